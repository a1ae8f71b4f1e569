"""oracle/ref_lib.py -- ctypes binding to oracle/_ref/libmarius_ref.so (the UNMODIFIED Marius reference C++,
built by oracle/Makefile from /root/reference, plus oracle/ref_driver.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py, bench.py's cpu_baseline /
``--impl reference`` legs and __graft_entry__.smoke().  Never by marius_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmarius_ref.so")

_lib = None

_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int64)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _fp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_F)


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_I)


def lib():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (libtorch must be loaded before the reference library resolves its symbols)

        # RTLD_LOCAL: the reference defines classes (Model, DistMult, PartitionBuffer ...) whose names the product's host adapters keep; its
        # symbols must never be visible to marius_b200/lib/_host*.so when both are loaded in one test process
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_num_threads.restype = C.c_int
        _lib.ref_set_num_threads.argtypes = [C.c_int]
        _lib.ref_inmemory_index_read.argtypes = [_F, C.c_int64, C.c_int64, _I, C.c_int64, _F]
        _lib.ref_inmemory_index_read_bad_rank.argtypes = [_F, C.c_int64, C.c_int64]
        _lib.ref_inmemory_index_add.argtypes = [_F, C.c_int64, C.c_int64, _I, C.c_int64, _F, C.c_int64, C.c_int64]
        _lib.ref_partition_buffer_exercise.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64, _I, C.c_int, _I, C.c_int64,
                                                       _F, _F, _F, _I, _I, _I, _I]
        _lib.ref_map_tensors.argtypes = [_I, C.c_int64, _I, _I]
        _lib.ref_map_tensors.restype = C.c_int64
        _lib.ref_train_batch.argtypes = [C.c_int, C.c_int, C.c_int, _F, _F, _F, _F, C.c_int64, _I, C.c_int64, _I, _I, C.c_int, C.c_int,
                                         C.c_float, C.c_int] + [_F] * 10
        _lib.ref_train_loop.argtypes = [C.c_int, C.c_int, C.c_int, _F, _F, C.c_int64, C.c_int, _I, _I, _I, C.c_int64, _I, _I, C.c_int, C.c_int,
                                        C.c_float, C.c_int, C.c_int]
        _lib.ref_train_loop.restype = C.c_double
        _lib.ref_evaluate_batch.argtypes = [C.c_int, C.c_int, C.c_int, _F, _F, _F, C.c_int64, _I, C.c_int64, _I, _I, C.c_int, C.c_int, _I, C.c_int64, _I,
                                            C.c_int64, _I, _I, _F, _F, _F, _F]
        _lib.ref_ranking_metrics.argtypes = [_I, C.c_int64, C.POINTER(C.c_double)]
        _lib.ref_compute_filter.argtypes = [_I, C.c_int64, C.c_int, _I, C.c_int, C.c_int, C.c_int, _I, C.c_int64, C.c_int64, _I, C.c_int64]
        _lib.ref_compute_filter.restype = C.c_int64
    return _lib


# decoder ids used by ref_driver.cpp::make_model (0 DistMult, 1 ComplEx, 2 TransE); oracle kinds are
# DOT=0, DISTMULT=1, COMPLEX=2 -> map oracle kind -> driver id (DOT has no reference decoder class with
# relations; it is DistMult with 2-column edges in the reference, handled by callers).
_KIND_TO_DRIVER = {1: 0, 2: 1}


def num_threads() -> int:
    return lib().ref_num_threads()


def set_num_threads(n: int) -> None:
    lib().ref_set_num_threads(n)


def index_read(table: np.ndarray, idx: np.ndarray) -> np.ndarray:
    out = np.empty((idx.shape[0], table.shape[1]), dtype=np.float32)
    rc = lib().ref_inmemory_index_read(_fp(table), table.shape[0], table.shape[1], _ip(idx), idx.shape[0], _fp(out))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return out


def index_read_bad_rank_throws(table: np.ndarray) -> bool:
    return lib().ref_inmemory_index_read_bad_rank(_fp(table), table.shape[0], table.shape[1]) == 1


def index_add(table: np.ndarray, idx: np.ndarray, vals: np.ndarray) -> None:
    rc = lib().ref_inmemory_index_add(_fp(table), table.shape[0], table.shape[1], _ip(idx), idx.shape[0], _fp(vals), vals.shape[0], vals.shape[1])
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())


def partition_buffer_exercise(filename, capacity, num_partitions, partition_size, d, total, states, idx, add_vals=None):
    states = np.ascontiguousarray(states, dtype=np.int64)
    ns = states.shape[0]
    n = idx.shape[0]
    read_out = np.empty((n, d), dtype=np.float32)
    read_after = np.empty((n, d), dtype=np.float32) if add_vals is not None else None
    map_cur = np.empty(total, dtype=np.int64)
    map_next = np.empty(total, dtype=np.int64) if ns > 1 else None
    admits = np.empty(max(ns - 1, 1), dtype=np.int64)
    evicts = np.empty(max(ns - 1, 1), dtype=np.int64)
    rc = lib().ref_partition_buffer_exercise(filename.encode(), capacity, num_partitions, partition_size, d, total, _ip(states), ns, _ip(idx), n,
                                             _fp(read_out), _fp(add_vals), _fp(read_after), _ip(map_cur), _ip(map_next), _ip(admits), _ip(evicts))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return dict(read=read_out, read_after_add=read_after, map_current=map_cur, map_next=map_next, admits=admits[:ns - 1], evicts=evicts[:ns - 1])


def map_tensors(all_ids: np.ndarray):
    uniq = np.empty_like(all_ids)
    mapped = np.empty_like(all_ids)
    u = lib().ref_map_tensors(_ip(all_ids), all_ids.shape[0], _ip(uniq), _ip(mapped))
    return uniq[:u].copy(), mapped


def train_batch(kind, emb, state, edges, rel, inv_rel, dst_negs, src_negs, lr, reduction):
    """Runs the reference Model::forward_lp + Model::train_batch.  Returns a dict of numpy arrays."""
    U, d = emb.shape
    B = edges.shape[0]
    Cc, N = dst_negs.shape
    Bp = Cc * int(np.ceil(B / Cc))
    R = rel.shape[0]
    inverse = inv_rel is not None and src_negs is not None
    out = dict(pos=np.zeros(Bp, np.float32), neg=np.zeros((Bp, N), np.float32), loss=np.zeros(1, np.float32), grad=np.zeros((U, d), np.float32),
               delta_e=np.zeros((U, d), np.float32), delta_s=np.zeros((U, d), np.float32), rel_grad=np.zeros((R, d), np.float32))
    if inverse:
        out.update(inv_pos=np.zeros(Bp, np.float32), inv_neg=np.zeros((Bp, N), np.float32), inv_rel_grad=np.zeros((R, d), np.float32))
    rc = lib().ref_train_batch(_KIND_TO_DRIVER[kind], d, R, _fp(rel), _fp(inv_rel) if inverse else None, _fp(emb), _fp(state), U, _ip(edges), B,
                               _ip(dst_negs), _ip(src_negs) if inverse else None, Cc, N, lr, reduction, _fp(out["pos"]), _fp(out["neg"]),
                               _fp(out.get("inv_pos")), _fp(out.get("inv_neg")), _fp(out["loss"]), _fp(out["grad"]), _fp(out["delta_e"]),
                               _fp(out["delta_s"]), _fp(out["rel_grad"]), _fp(out.get("inv_rel_grad")))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return out


def train_loop(kind, d, num_rel, table, state_table, uniq, uniq_off, edges, dst_negs, src_negs, lr, reduction, warmup_batches=1) -> float:
    """Timed synchronous CPU hot loop of the reference; returns seconds for len(uniq_off)-1 batches."""
    nb = uniq_off.shape[0] - 1
    B = edges.shape[1]
    Cc, N = dst_negs.shape[1], dst_negs.shape[2]
    t = lib().ref_train_loop(_KIND_TO_DRIVER[kind], d, num_rel, _fp(table), _fp(state_table), table.shape[0], nb, _ip(uniq), _ip(uniq_off),
                             _ip(edges), B, _ip(dst_negs), _ip(src_negs), Cc, N, lr, reduction, warmup_batches)
    if t < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return t


def evaluate_batch(kind, emb, edges, rel, inv_rel, dst_negs, src_negs, dst_filter=None, src_filter=None):
    """Runs the reference Model::evaluate_batch (forward_lp + score filters + LinkPredictionReporter::computeRanks)."""
    U, d = emb.shape
    B = edges.shape[0]
    Cc, N = dst_negs.shape
    Bp = Cc * int(np.ceil(B / Cc))
    inverse = inv_rel is not None and src_negs is not None
    out = dict(ranks=np.zeros(Bp, np.int64), pos=np.zeros(Bp, np.float32), neg=np.zeros((Bp, N), np.float32))
    if inverse:
        out.update(inv_ranks=np.zeros(Bp, np.int64), inv_pos=np.zeros(Bp, np.float32), inv_neg=np.zeros((Bp, N), np.float32))
    df = np.ascontiguousarray(dst_filter, dtype=np.int64) if dst_filter is not None else None
    sf = np.ascontiguousarray(src_filter, dtype=np.int64) if src_filter is not None else None
    rc = lib().ref_evaluate_batch(_KIND_TO_DRIVER[kind], d, rel.shape[0], _fp(rel), _fp(inv_rel) if inverse else None, _fp(emb), U, _ip(edges), B,
                                  _ip(dst_negs), _ip(src_negs) if inverse else None, Cc, N, _ip(df), 0 if df is None else df.shape[0], _ip(sf),
                                  0 if sf is None else sf.shape[0], _ip(out["ranks"]), _ip(out.get("inv_ranks")), _fp(out["pos"]), _fp(out["neg"]),
                                  _fp(out.get("inv_pos")), _fp(out.get("inv_neg")))
    if rc:
        raise RuntimeError(lib().ref_last_error().decode())
    return out


def ranking_metrics(ranks: np.ndarray) -> dict:
    out = (C.c_double * 5)()
    r = np.ascontiguousarray(ranks, dtype=np.int64)
    if lib().ref_ranking_metrics(_ip(r), r.shape[0], out):
        raise RuntimeError(lib().ref_last_error().decode())
    return {"mean_rank": out[0], "mrr": out[1], "hits@1": out[2], "hits@3": out[3], "hits@10": out[4]}


def compute_filter(edges: np.ndarray, corruption_nodes: np.ndarray, inverse: bool, graph_edges: Optional[np.ndarray] = None, num_nodes: int = 0):
    """Runs the reference compute_filter_corruption_cpu (global filter against graph_edges, or local filter against the batch)."""
    e = np.ascontiguousarray(edges, dtype=np.int64)
    negs = np.ascontiguousarray(corruption_nodes, dtype=np.int64)
    g = np.ascontiguousarray(graph_edges, dtype=np.int64) if graph_edges is not None else None
    cap = max(1, e.shape[0] * negs.shape[1] if g is None else 4 * (g.shape[0] + e.shape[0]))
    out = np.zeros((cap, 2), np.int64)
    n = lib().ref_compute_filter(_ip(e), e.shape[0], e.shape[1], _ip(negs), negs.shape[0], negs.shape[1], int(bool(inverse)), _ip(g),
                                 0 if g is None else g.shape[0], int(num_nodes), _ip(out), cap)
    if n < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    if n > cap:
        raise RuntimeError("filter buffer too small")
    return out[:n].copy()
