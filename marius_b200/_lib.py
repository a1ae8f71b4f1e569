"""ctypes binding of include/marius_b200.h (libmarius_b200.so).  There is NO fallback: if the CUDA library is
missing this module raises at import, and every call raises MariusB200Error on a non-zero status."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmarius_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: build it with `python -m marius_b200.build` (the hot path has no CPU/eager fallback)")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


class MariusB200Error(RuntimeError):
    """Non-zero mb_status.  MB_ERR_INVALID mirrors the reference's std::runtime_error on bad shapes (storage.cpp:607-610,652-655)."""

    def __init__(self, status: int, msg: str):
        super().__init__(f"marius_b200 status {status}: {msg}")
        self.status = status


class mb_batch(C.Structure):
    _fields_ = [("decoder", C.c_int), ("U", C.c_int64), ("d", C.c_int64), ("B", C.c_int64), ("R", C.c_int64), ("C", C.c_int), ("N", C.c_int),
                ("edges", C.c_void_p), ("edge_cols", C.c_int), ("dst_negs", C.c_void_p), ("src_negs", C.c_void_p), ("rel", C.c_void_p),
                ("inv_rel", C.c_void_p)]


class mb_shards(C.Structure):
    _fields_ = [("tables", C.c_void_p * 8), ("states", C.c_void_p * 8), ("world", C.c_int), ("rows_per_rank", C.c_int64), ("rank", C.c_int),
                ("exchange", C.c_void_p * 8), ("exchange_rows", C.c_int64), ("single_process", C.c_int)]


_vp, _i64, _i32, _f = C.c_void_p, C.c_int64, C.c_int, C.c_float
_SIGS = {
    "mb_create": [_i32, C.POINTER(_vp)],
    "mb_gather_rows": [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp],
    "mb_scatter_add_rows": [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp],
    "mb_scatter_put_rows": [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp],
    "mb_global_to_local_map": [_vp, _i64, _i64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i32, _vp],
    "mb_adagrad_deltas": [_vp, _vp, _i64, _i64, _i64, _f, _vp, _vp, _vp],
    "mb_adagrad_update_rows": [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _f, _vp],
    "mb_map_tensors": [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp],
    "mb_sample_negatives": [_i64, _i32, _i32, _f, _vp, _i64, _i32, _i32, C.c_uint64, C.c_uint32, _vp, _vp],
    "mb_edge_sample": [_vp, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "mb_reduce_rows_by_key": [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp],
    "mb_decoder_forward": [_vp, C.POINTER(mb_batch), _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp],
    "mb_apply_score_filter": [_vp, _i64, _i64, _i64, _vp, _i64, _vp],
    "mb_compute_ranks": [_vp, _vp, _i64, _i64, _i64, _vp, _vp],
    "mb_evaluate_batch": [_vp, C.POINTER(mb_batch), _vp, _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "mb_decoder_backward": [_vp, C.POINTER(mb_batch), _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mb_train_batch": [_vp, C.POINTER(mb_batch), _vp, _i64, _vp, _i64, _f, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mb_train_step": [_vp, C.POINTER(mb_batch), _vp, _vp, _i64, _i64, _vp, _f, _i32, _i32, _vp, _vp, _vp, _vp],
    "mb_train_step_sharded": [_vp, C.POINTER(mb_batch), C.POINTER(mb_shards), _i64, _vp, _f, _i32, _i32, _vp, _vp, _vp, _vp],
    "mb_train_step_sharded_host": [_vp, C.POINTER(mb_batch), C.POINTER(mb_shards), _i64, _vp, _f, _i32, _i32, _vp, _vp, _vp, _vp],
    "mb_train_step_host": [_vp, C.POINTER(mb_batch), _vp, _vp, _i64, _i64, _vp, _f, _i32, _i32, _vp, _vp, _vp, _vp],
    "mb_train_step_host_async": [_vp, C.POINTER(mb_batch), _vp, _vp, _i64, _i64, _vp, _f, _i32, _i32, _vp, _vp, C.POINTER(C.c_int), _vp],
    "mb_train_step_sharded_host_async": [_vp, C.POINTER(mb_batch), C.POINTER(mb_shards), _i64, _vp, _f, _i32, _i32, _vp, _vp, C.POINTER(C.c_int), _vp],
    "mb_train_step_host_wait": [_vp, _i32, C.POINTER(C.c_float)],
    "mb_train_step_edges_host_async": [_vp, _i32, _vp, _i64, _i64, _i32, _i32, C.c_uint64, C.c_uint32, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _f, _i32, _i32,
                                       _vp, _vp, C.POINTER(C.c_int), _vp],
    "mb_dense_adagrad_step": [_vp, _vp, _vp, _i64, _f, _f, _vp],
    "mb_profile_enable": [_vp, _i32],
    "mb_graph_enable": [_vp, _i32],
    "mb_enable_peer_access": [_vp, _i32],
    "mb_ipc_export": [_vp, _vp, C.POINTER(C.c_int64)],
    "mb_ipc_import": [_vp, _vp, _i64, C.POINTER(_vp)],
    "mb_profile_read": [_vp, C.POINTER(C.c_float), C.POINTER(C.c_int)],
    "mb_profile_timeline": [_vp, _i32, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)],
    "mb_filter_sort_edges": [_vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp],
    "mb_compute_filter": [_vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp, _i64, _vp, _vp],
    "mb_evaluate_all_nodes": [_vp, _i32, _vp, _i64, _i64, _i64, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp],
    "mb_shard_error": [_vp, C.POINTER(mb_shards), C.POINTER(C.c_int)],
    "mb_debug_wait_log": [C.POINTER(C.c_uint64), _i32],
    "mb_debug_gemm": [_vp, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
}
EXPORTS = list(_SIGS) + ["mb_profile_num_stages", "mb_profile_stage_name", "mb_destroy", "mb_last_error", "mb_version", "mb_launch_count", "mb_build_info", "mb_workspace_bytes", "mb_shard_exchange_bytes"]
for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = C.c_int
lib.mb_destroy.argtypes = [_vp]
lib.mb_destroy.restype = None
lib.mb_last_error.restype = C.c_char_p
lib.mb_version.restype = C.c_int
lib.mb_launch_count.restype = C.c_uint64
lib.mb_build_info.restype = C.c_char_p
lib.mb_profile_num_stages.restype = C.c_int
lib.mb_profile_stage_name.argtypes = [C.c_int]
lib.mb_profile_stage_name.restype = C.c_char_p
lib.mb_workspace_bytes.argtypes = [_vp]
lib.mb_workspace_bytes.restype = C.c_size_t
lib.mb_shard_exchange_bytes.argtypes = [C.c_int, C.c_int64, C.c_int64]
lib.mb_shard_exchange_bytes.restype = C.c_int64


def check(status: int) -> None:
    if status != 0:
        raise MariusB200Error(status, lib.mb_last_error().decode())


def launch_count() -> int:
    return int(lib.mb_launch_count())
