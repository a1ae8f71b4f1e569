"""marius_b200.ops -- thin Python entry points over the C ABI (include/marius_b200.h).

torch is used only for device memory and streams (tensors in, tensors out); every function below is one
C-ABI call into the hand-written sm_100a kernels.  There is no eager / CPU fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Optional

import torch

from . import _lib
from ._lib import MariusB200Error, check, lib, mb_batch, mb_shards

DOT, DISTMULT, COMPLEX = 0, 1, 2
REDUCTION_MEAN, REDUCTION_SUM = 0, 1
PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2

_INVALID = 1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise MariusB200Error(_INVALID, "marius_b200 ops need CUDA tensors (the hot path has no CPU fallback)")


def _rowmajor(t: torch.Tensor, name: str) -> int:
    """leading dimension (elements) of a 2-D tensor whose rows are contiguous"""
    if t.dim() != 2:
        raise MariusB200Error(_INVALID, f"{name} must be 2-D with contiguous rows")
    if t.numel() == 0:
        return max(t.size(1), 1)
    if t.size(1) > 1 and t.stride(1) != 1:
        raise MariusB200Error(_INVALID, f"{name} must be 2-D with contiguous rows")
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), t.stride(0))


def _check_indices(idx: torch.Tensor):
    # storage.cpp:607-610 / buffer.cpp:442-445: indices must be 1-D (std::runtime_error otherwise)
    if idx.dim() != 1 or idx.dtype != torch.int64:
        raise MariusB200Error(_INVALID, "indices must be a 1-D int64 tensor")
    if not idx.is_contiguous():
        raise MariusB200Error(_INVALID, "indices must be contiguous")


class Context:
    """Owns the device workspace of one in-flight batch (mb_context)."""

    def __init__(self, device: int = 0):
        self.device = int(device)
        h = C.c_void_p()
        check(lib.mb_create(self.device, C.byref(h)))
        self._h = h

    @property
    def handle(self):
        return self._h

    def workspace_bytes(self) -> int:
        return int(lib.mb_workspace_bytes(self._h))

    def enable_peer_access(self, peer_device: int) -> None:
        check(lib.mb_enable_peer_access(self._h, int(peer_device)))

    def graph(self, on: bool) -> None:
        """enable / disable CUDA-graph replay of the fused step for this context"""
        check(lib.mb_graph_enable(self._h, int(bool(on))))

    def profile(self, on) -> None:
        """False/0 off, True/1 stage timing (side streams folded into the caller's stream), 2 timeline (streams stay concurrent)"""
        check(lib.mb_profile_enable(self._h, int(on)))

    def profile_timeline(self, cap: int = 4096) -> list:
        """[(stage name, start ms, end ms)] of the stage launches recorded since the last read (profile(2)); synchronises the device."""
        st = (C.c_int * cap)()
        a = (C.c_float * cap)()
        b = (C.c_float * cap)()
        n = C.c_int(0)
        check(lib.mb_profile_timeline(self._h, cap, st, a, b, C.byref(n)))
        return [(lib.mb_profile_stage_name(st[i]).decode(), float(a[i]), float(b[i])) for i in range(n.value)]

    def profile_read(self) -> dict:
        """{stage name: (total ms, launches)} accumulated since the last read (synchronises the device)."""
        n = lib.mb_profile_num_stages()
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        check(lib.mb_profile_read(self._h, ms, cnt))
        return {lib.mb_profile_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def close(self):
        if getattr(self, "_h", None):
            lib.mb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if sys.is_finalizing():  # interpreter shutdown: leave device resources to the driver
                return
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------------
def gather_rows(table: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Storage::indexRead on a device-resident table (storage.cpp:613-614, buffer.cpp:441-455)."""
    _need_cuda(table, idx)
    _check_indices(idx)
    ld = _rowmajor(table, "table")
    n, d = idx.numel(), table.size(1)
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=table.device)
    check(lib.mb_gather_rows(_ptr(table), table.size(0), ld, d, _ptr(idx), n, _ptr(out), _rowmajor(out, "out") if n > 0 else d, _stream()))
    return out


def _check_values(table, idx, vals):
    # storage.cpp:652-655: values defined, idx 1-D, idx.size(0) == values.size(0), values.size(1) == dim1
    if vals is None or idx.dim() != 1 or vals.dim() != 2 or idx.size(0) != vals.size(0) or table.size(1) != vals.size(1):
        raise MariusB200Error(_INVALID, "indexAdd: values undefined or shape mismatch")


def scatter_add_rows(table: torch.Tensor, idx: torch.Tensor, vals: torch.Tensor) -> None:
    """Storage::indexAdd on a device-resident table, unique ids (storage.cpp:656-657, buffer.cpp:459-480)."""
    _need_cuda(table, idx, vals)
    _check_values(table, idx, vals)
    _check_indices(idx)
    check(lib.mb_scatter_add_rows(_ptr(table), table.size(0), _rowmajor(table, "table"), table.size(1), _ptr(idx), idx.numel(), _ptr(vals),
                                  _rowmajor(vals, "values"), _stream()))


def scatter_put_rows(table: torch.Tensor, idx: torch.Tensor, vals: torch.Tensor) -> None:
    """Storage::indexPut (storage.cpp:675-696)."""
    _need_cuda(table, idx, vals)
    _check_values(table, idx, vals)
    _check_indices(idx)
    check(lib.mb_scatter_put_rows(_ptr(table), table.size(0), _rowmajor(table, "table"), table.size(1), _ptr(idx), idx.numel(), _ptr(vals),
                                  _rowmajor(vals, "values"), _stream()))


def global_to_local_map(total_rows: int, partition_size: int, partition_ids, buffer_slots, device) -> torch.Tensor:
    """PartitionBuffer::getGlobalToLocalMap (buffer.cpp:581-633)."""
    n = len(partition_ids)
    out = torch.empty(total_rows, dtype=torch.int64, device=device)
    pa = (C.c_int32 * max(n, 1))(*[int(x) for x in partition_ids])
    sa = (C.c_int32 * max(n, 1))(*[int(x) for x in buffer_slots])
    check(lib.mb_global_to_local_map(_ptr(out), total_rows, partition_size, pa, sa, n, _stream()))
    return out


def adagrad_deltas(grad: torch.Tensor, state: torch.Tensor, lr: float):
    """Batch::accumulateGradients (batch.cpp:62-79): returns (node_gradients_ = delta_e, node_state_update_ = delta_s)."""
    _need_cuda(grad, state)
    g2 = grad.reshape(-1, grad.size(-1)) if grad.dim() > 1 else grad.reshape(1, -1)
    s2 = state.reshape(g2.shape)
    g2, s2 = g2.contiguous(), s2.contiguous()
    de, ds = torch.empty_like(g2), torch.empty_like(g2)
    check(lib.mb_adagrad_deltas(_ptr(g2), _ptr(s2), g2.size(0), g2.size(1), g2.size(1), float(lr), _ptr(de), _ptr(ds), _stream()))
    return de.reshape(grad.shape), ds.reshape(grad.shape)


def adagrad_update_rows(table: torch.Tensor, state_table: torch.Tensor, idx: torch.Tensor, grad: torch.Tensor, lr: float) -> None:
    """accumulateGradients + indexAdd(embeddings) + indexAdd(state), fused (dataloader.cpp:550-557)."""
    _need_cuda(table, state_table, idx, grad)
    _check_values(table, idx, grad)
    _check_indices(idx)
    ld = _rowmajor(table, "table")
    if _rowmajor(state_table, "state_table") != ld or state_table.shape != table.shape:
        raise MariusB200Error(_INVALID, "state table must have the table's shape and stride")
    check(lib.mb_adagrad_update_rows(_ptr(table), _ptr(state_table), table.size(0), ld, table.size(1), _ptr(idx), idx.numel(), _ptr(grad),
                                     _rowmajor(grad, "grad"), float(lr), _stream()))


def dense_adagrad_step(param: torch.Tensor, state_sum: torch.Tensor, grad: torch.Tensor, lr: float, eps: float = 1e-10) -> None:
    """AdagradOptimizer::step (optim.cpp:114-145) on one dense parameter."""
    _need_cuda(param, state_sum, grad)
    if not (param.is_contiguous() and state_sum.is_contiguous() and grad.is_contiguous()):
        raise MariusB200Error(_INVALID, "dense parameters must be contiguous")
    check(lib.mb_dense_adagrad_step(_ptr(param), _ptr(state_sum), _ptr(grad), param.numel(), float(lr), float(eps), _stream()))


def map_tensors(ctx: Context, all_ids: torch.Tensor, max_id: Optional[int] = None):
    """map_tensors (util.cpp:180-205): (sorted unique ids, position of every input id)."""
    _need_cuda(all_ids)
    if all_ids.dim() != 1:
        raise MariusB200Error(_INVALID, "Input tensors must be 1D")  # util.cpp:182-185
    all_ids = all_ids.contiguous()
    n = all_ids.numel()
    if max_id is None:
        max_id = int(all_ids.max().item()) if n else 0
    uniq = torch.empty(n, dtype=torch.int64, device=all_ids.device)
    mapped = torch.empty(n, dtype=torch.int64, device=all_ids.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=all_ids.device)
    check(lib.mb_map_tensors(ctx.handle, _ptr(all_ids), n, int(max_id), _ptr(uniq), _ptr(mapped), _ptr(cnt), _stream()))
    return uniq[: int(cnt.item())], mapped


def sample_negatives(num_nodes: int, C_: int, N: int, seed: int, batch_index: int, inverse: bool, device, degree_fraction: float = 0.0,
                     edges: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CorruptNodeNegativeSampler::getNegatives (negative.cpp:328-366) on the device: [C, N] int64 ids, a pure function of
    (seed, batch_index, inverse, position)."""
    _need_cuda(edges)
    out = torch.empty((C_, N), dtype=torch.int64, device=device)
    B, cols = (edges.size(0), edges.size(1)) if edges is not None else (0, 3)
    if edges is not None:
        edges = edges.contiguous()
    check(lib.mb_sample_negatives(int(num_nodes), int(C_), int(N), float(degree_fraction), _ptr(edges), B, cols, int(bool(inverse)), int(seed),
                                  int(batch_index), _ptr(out), _stream()))
    return out


def edge_sample(ctx: Context, edges: torch.Tensor, src_negs: Optional[torch.Tensor], dst_negs: torch.Tensor, max_id: int):
    """DataLoader::edgeSample (dataloader.cpp:389-471) on the device: global edges / negatives -> (unique ids padded with -1 to
    capacity, num_unique [1] int64 device, local edges, local src_negs or None, local dst_negs)."""
    _need_cuda(edges, src_negs, dst_negs)
    if edges.dim() != 2 or edges.size(1) not in (2, 3):
        raise MariusB200Error(_INVALID, "Edge list must be a 3 or 2 column tensor")
    edges, dst_negs = edges.contiguous(), dst_negs.contiguous()
    src_negs = src_negs.contiguous() if src_negs is not None else None
    B, cols = edges.shape
    Cc, N = dst_negs.shape
    n = 2 * B + Cc * N * (2 if src_negs is not None else 1)
    dev = edges.device
    uniq = torch.empty(n, dtype=torch.int64, device=dev)
    num = torch.zeros(1, dtype=torch.int64, device=dev)
    e_loc = torch.empty_like(edges)
    s_loc = torch.empty_like(src_negs) if src_negs is not None else None
    d_loc = torch.empty_like(dst_negs)
    check(lib.mb_edge_sample(ctx.handle, _ptr(edges), B, cols, _ptr(src_negs), _ptr(dst_negs), Cc, N, int(max_id), _ptr(uniq), _ptr(e_loc), _ptr(s_loc),
                             _ptr(d_loc), _ptr(num), _stream()))
    return uniq, num, e_loc, s_loc, d_loc


def reduce_rows_by_key(ctx: Context, ids: torch.Tensor, rows: torch.Tensor, max_id: Optional[int] = None, padded: bool = False):
    """(sorted unique ids, per-id sum of rows): the owner-side merge of gradient rows received from several ranks.
    padded=True returns the full-length buffers (unique ids padded with -1, zero rows) without any host synchronisation."""
    _need_cuda(ids, rows)
    _check_indices(ids)
    n = ids.numel()
    if rows.dim() != 2 or rows.size(0) != n:
        raise MariusB200Error(_INVALID, "rows must be [len(ids), d]")
    rows = rows.contiguous()
    if max_id is None:
        max_id = int(ids.max().item()) if n else 0
    uniq = torch.empty(n, dtype=torch.int64, device=ids.device)
    out = torch.empty_like(rows)
    cnt = torch.zeros(1, dtype=torch.int64, device=ids.device)
    check(lib.mb_reduce_rows_by_key(ctx.handle, _ptr(ids), _ptr(rows), n, rows.size(1), int(max_id), _ptr(uniq), _ptr(out), _ptr(cnt), _stream()))
    if padded:
        return uniq, out
    u = int(cnt.item())
    return uniq[:u], out[:u]


# ---------------------------------------------------------------------------------------------------------------
def _make_batch(kind, U, d, edges, rel, inv_rel, dst_negs, src_negs, host=False):
    if edges.dim() != 2 or edges.size(1) not in (2, 3):
        raise MariusB200Error(_INVALID, "Edge list must be a 3 or 2 column tensor")  # decoder_methods.cpp:66-72
    if dst_negs is None or dst_negs.dim() != 2:
        raise MariusB200Error(_INVALID, "negative index mapping must be [num_chunks, num_negatives]")
    keep = [edges.contiguous(), dst_negs.contiguous(), None if src_negs is None else src_negs.contiguous(),
            None if rel is None else rel.contiguous(), None if inv_rel is None else inv_rel.contiguous()]
    b = mb_batch()
    b.decoder = int(kind)
    b.U, b.d, b.B = int(U), int(d), int(edges.size(0))
    b.R = int(rel.size(0)) if rel is not None else 0
    b.C, b.N = int(dst_negs.size(0)), int(dst_negs.size(1))
    b.edges = keep[0].data_ptr()
    b.edge_cols = int(edges.size(1))
    b.dst_negs = keep[1].data_ptr()
    b.src_negs = keep[2].data_ptr() if keep[2] is not None else None
    b.rel = keep[3].data_ptr() if keep[3] is not None else None
    b.inv_rel = keep[4].data_ptr() if keep[4] is not None else None
    return b, keep


def padded_rows(B: int, C_: int) -> int:
    return C_ * ((B + C_ - 1) // C_)


def decoder_forward(ctx: Context, kind: int, emb: torch.Tensor, edges, rel, inv_rel, dst_negs, src_negs, precision: int = PREC_BF16X3):
    """Model::forward_lp / node_corrupt_forward (decoder_methods.cpp:57-114): (pos, neg, inv_pos, inv_neg)."""
    _need_cuda(emb, edges, rel, inv_rel, dst_negs, src_negs)
    if emb is None:
        raise MariusB200Error(_INVALID, "UndefinedTensor")
    b, keep = _make_batch(kind, emb.size(0), emb.size(1), edges, rel, inv_rel, dst_negs, src_negs)
    Bp = padded_rows(b.B, b.C)
    inverse = b.inv_rel is not None and b.src_negs is not None and kind != DOT and b.edge_cols == 3
    dev = emb.device
    pos = torch.empty(Bp, dtype=torch.float32, device=dev)
    neg = torch.empty((Bp, b.N), dtype=torch.float32, device=dev)
    inv_pos = torch.empty(Bp, dtype=torch.float32, device=dev) if inverse else None
    inv_neg = torch.empty((Bp, b.N), dtype=torch.float32, device=dev) if inverse else None
    check(lib.mb_decoder_forward(ctx.handle, C.byref(b), _ptr(emb), _rowmajor(emb, "node_embeddings"), int(precision), _ptr(pos), _ptr(neg),
                                 _ptr(inv_pos), _ptr(inv_neg), _stream()))
    return pos, neg, inv_pos, inv_neg


def _check_filter(f: Optional[torch.Tensor]):
    if f is None:
        return None, 0
    if f.dim() != 2 or f.size(1) != 2 or f.dtype != torch.int64:
        raise MariusB200Error(_INVALID, "score filter must be an int64 [F, 2] tensor of (row, column) pairs")
    f = f.contiguous()
    return f, int(f.size(0))


def apply_score_filter(scores: torch.Tensor, filt: Optional[torch.Tensor]) -> torch.Tensor:
    """apply_score_filter (negative.cpp:306-311): scores[filter[:,0], filter[:,1]] = -1e9, in place."""
    _need_cuda(scores, filt)
    f, F = _check_filter(filt)
    if F:
        check(lib.mb_apply_score_filter(_ptr(scores), scores.size(0), scores.size(1), _rowmajor(scores, "scores"), _ptr(f), F, _stream()))
    return scores


def compute_ranks(pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
    """LinkPredictionReporter::computeRanks (reporting.cpp:56-58): (neg >= pos.unsqueeze(1)).sum(1) + 1, int64."""
    _need_cuda(pos, neg)
    if pos.dim() != 1 or neg.dim() != 2 or pos.size(0) != neg.size(0):
        raise MariusB200Error(_INVALID, "computeRanks: pos [rows], neg [rows, N]")
    ranks = torch.empty(pos.size(0), dtype=torch.int64, device=pos.device)
    check(lib.mb_compute_ranks(_ptr(pos.contiguous()), _ptr(neg), neg.size(0), neg.size(1), _rowmajor(neg, "neg_scores"), _ptr(ranks), _stream()))
    return ranks


def evaluate_batch(ctx: Context, kind: int, emb: torch.Tensor, edges, rel, inv_rel, dst_negs, src_negs, dst_filter=None, src_filter=None,
                   precision: int = PREC_BF16X3):
    """Model::evaluate_batch (model.cpp:335-349): scores, score filters and ranks of both corruption sides in one call.
    Returns (ranks [Bp], inv_ranks [Bp] or None, pos [Bp], inv_pos [Bp] or None)."""
    _need_cuda(emb, edges, rel, inv_rel, dst_negs, src_negs, dst_filter, src_filter)
    if emb is None:
        raise MariusB200Error(_INVALID, "UndefinedTensor")
    b, keep = _make_batch(kind, emb.size(0), emb.size(1), edges, rel, inv_rel, dst_negs, src_negs)
    Bp = padded_rows(b.B, b.C)
    inverse = b.inv_rel is not None and b.src_negs is not None and kind != DOT and b.edge_cols == 3
    dev = emb.device
    df, Fd = _check_filter(dst_filter)
    sf, Fs = _check_filter(src_filter)
    ranks = torch.empty(Bp, dtype=torch.int64, device=dev)
    pos = torch.empty(Bp, dtype=torch.float32, device=dev)
    inv_ranks = torch.empty(Bp, dtype=torch.int64, device=dev) if inverse else None
    inv_pos = torch.empty(Bp, dtype=torch.float32, device=dev) if inverse else None
    check(lib.mb_evaluate_batch(ctx.handle, C.byref(b), _ptr(emb), _rowmajor(emb, "node_embeddings"), int(precision), _ptr(df), Fd, _ptr(sf), Fs,
                                _ptr(ranks), _ptr(inv_ranks), _ptr(pos), _ptr(inv_pos), _stream()))
    return ranks, inv_ranks, pos, inv_pos


def filter_sort_edges(ctx: Context, graph_edges: torch.Tensor, inverse: bool, max_id: Optional[int] = None) -> torch.Tensor:
    """The graph's edges sorted stably by the endpoint a corruption keeps (done once per evaluation; negative.cpp:62-112)."""
    _need_cuda(graph_edges)
    g = graph_edges.contiguous()
    if max_id is None:
        max_id = int(g.max().item()) if g.numel() else 0
    out = torch.empty_like(g)
    check(lib.mb_filter_sort_edges(ctx.handle, _ptr(g), g.size(0), g.size(1), int(bool(inverse)), int(max_id), _ptr(out), _stream()))
    return out


def compute_filter(ctx: Context, sorted_edges: torch.Tensor, batch_edges: torch.Tensor, inverse: bool, cap: Optional[int] = None) -> torch.Tensor:
    """compute_filter_corruption against the whole graph with all nodes as negatives (negative.cpp:152-163): [F, 2] (row, node id)."""
    _need_cuda(sorted_edges, batch_edges)
    e = batch_edges.contiguous()
    cnt = torch.zeros(1, dtype=torch.int64, device=e.device)
    cap = int(cap) if cap is not None else max(1024, 64 * e.size(0))
    while True:
        out = torch.empty((cap, 2), dtype=torch.int64, device=e.device)
        check(lib.mb_compute_filter(ctx.handle, _ptr(sorted_edges), sorted_edges.size(0), sorted_edges.size(1), int(bool(inverse)), _ptr(e), e.size(0),
                                    _ptr(out), cap, _ptr(cnt), _stream()))
        n = int(cnt.item())
        if n <= cap:
            return out[:n]
        cap = n


def evaluate_all_nodes(ctx: Context, kind: int, table: torch.Tensor, edges: torch.Tensor, rel, inv_rel, dst_filter=None, src_filter=None,
                       precision: int = PREC_BF16X3, tile_rows: int = 32768):
    """Model::evaluate_batch with every node as a negative (filtered evaluation), streamed over the table in tiles: (ranks, inv_ranks, pos, inv_pos)."""
    _need_cuda(table, edges, rel, inv_rel, dst_filter, src_filter)
    e = edges.contiguous()
    B = e.size(0)
    inverse = inv_rel is not None and rel is not None and kind != DOT and e.size(1) == 3
    df, Fd = _check_filter(dst_filter)
    sf, Fs = _check_filter(src_filter)
    dev = table.device
    ranks = torch.empty(B, dtype=torch.int64, device=dev)
    inv_ranks = torch.empty(B, dtype=torch.int64, device=dev) if inverse else None
    pos = torch.empty(B, dtype=torch.float32, device=dev)
    inv_pos = torch.empty(B, dtype=torch.float32, device=dev) if inverse else None
    check(lib.mb_evaluate_all_nodes(ctx.handle, int(kind), _ptr(table), table.size(0), _rowmajor(table, "table"), table.size(1), _ptr(e), B, e.size(1), _ptr(rel),
                                    _ptr(inv_rel) if inverse else None, rel.size(0) if rel is not None else 0, _ptr(df), Fd, _ptr(sf), Fs, int(precision),
                                    int(tile_rows), _ptr(ranks), _ptr(inv_ranks), _ptr(pos), _ptr(inv_pos), _stream()))
    return ranks, inv_ranks, pos, inv_pos


def ranking_metrics(ranks: torch.Tensor, ks=(1, 3, 10)) -> dict:
    """MeanRankMetric / MeanReciprocalRankMetric / HitskMetric (reporting.cpp:17-31) on a rank vector."""
    out = {"mean_rank": float(ranks.to(torch.float64).mean().item()), "mrr": float(ranks.to(torch.float32).reciprocal().mean().item())}
    for k in ks:
        out[f"hits@{k}"] = float((ranks <= k).sum().item()) / ranks.size(0)
    return out


def train_batch(ctx: Context, kind: int, emb, state, edges, rel, inv_rel, dst_negs, src_negs, lr: float, reduction: int = REDUCTION_SUM,
                precision: int = PREC_BF16X3, want_grad: bool = True):
    """Model::train_batch (model.cpp:290-333) on batch-local tensors.  Returns a dict: loss, grad, delta_e, delta_s, rel_grad, inv_rel_grad."""
    _need_cuda(emb, state, edges, rel, inv_rel, dst_negs, src_negs)
    b, keep = _make_batch(kind, emb.size(0), emb.size(1), edges, rel, inv_rel, dst_negs, src_negs)
    dev = emb.device
    U, d = emb.shape
    out = dict(loss=torch.empty(1, dtype=torch.float32, device=dev))
    out["grad"] = torch.empty((U, d), dtype=torch.float32, device=dev) if want_grad else None
    if state is not None:
        out["delta_e"] = torch.empty((U, d), dtype=torch.float32, device=dev)
        out["delta_s"] = torch.empty((U, d), dtype=torch.float32, device=dev)
    has_rel = rel is not None and kind != DOT and b.edge_cols == 3
    inverse = has_rel and inv_rel is not None and src_negs is not None
    out["rel_grad"] = torch.empty_like(rel) if has_rel else None
    out["inv_rel_grad"] = torch.empty_like(inv_rel) if inverse else None
    check(lib.mb_train_batch(ctx.handle, C.byref(b), _ptr(emb), _rowmajor(emb, "node_embeddings"), _ptr(state),
                             _rowmajor(state, "state") if state is not None else 0, float(lr), int(reduction), int(precision), _ptr(out["loss"]),
                             _ptr(out["grad"]), _ptr(out.get("delta_e")), _ptr(out.get("delta_s")), _ptr(out["rel_grad"]),
                             _ptr(out["inv_rel_grad"]), _stream()))
    return out


def train_step(ctx: Context, kind: int, table, state_table, unique_ids, edges, rel, inv_rel, dst_negs, src_negs, lr: float,
               reduction: int = REDUCTION_SUM, precision: int = PREC_BF16X3, loss=None, rel_grad=None, inv_rel_grad=None):
    """gather -> train_batch -> fused Adagrad scatter on a device-resident table (trainer.cpp:106-138 with DEVICE_MEMORY embeddings)."""
    _need_cuda(table, state_table, unique_ids, edges, rel, inv_rel, dst_negs, src_negs)
    _check_indices(unique_ids)
    b, keep = _make_batch(kind, unique_ids.numel(), table.size(1), edges, rel, inv_rel, dst_negs, src_negs)
    if loss is None:
        loss = torch.empty(1, dtype=torch.float32, device=table.device)
    check(lib.mb_train_step(ctx.handle, C.byref(b), _ptr(table), _ptr(state_table), table.size(0), _rowmajor(table, "table"), _ptr(unique_ids),
                            float(lr), int(reduction), int(precision), _ptr(loss), _ptr(rel_grad), _ptr(inv_rel_grad), _stream()))
    return loss


def train_step_host(ctx: Context, kind: int, table, state_table, unique_ids_h, edges_h, rel, inv_rel, dst_negs_h, src_negs_h, lr: float,
                    reduction: int = REDUCTION_SUM, precision: int = PREC_BF16X3, rel_grad=None, inv_rel_grad=None) -> float:
    """Same step with HOST (ideally pinned) index tensors: H2D of the batch indices and D2H of the loss are inside the call."""
    _need_cuda(table, state_table, rel, inv_rel)
    for t in (unique_ids_h, edges_h, dst_negs_h, src_negs_h):
        if t is not None and t.is_cuda:
            raise MariusB200Error(_INVALID, "train_step_host takes host index tensors")
    b, keep = _make_batch(kind, unique_ids_h.numel(), table.size(1), edges_h, rel, inv_rel, dst_negs_h, src_negs_h, host=True)
    loss = C.c_float(0.0)
    uid = unique_ids_h.contiguous()
    check(lib.mb_train_step_host(ctx.handle, C.byref(b), _ptr(table), _ptr(state_table), table.size(0), _rowmajor(table, "table"),
                                 C.c_void_p(uid.data_ptr()), float(lr), int(reduction), int(precision), C.cast(C.pointer(loss), C.c_void_p), _ptr(rel_grad),
                                 _ptr(inv_rel_grad), _stream()))
    return float(loss.value)


def train_step_host_async(ctx: Context, kind: int, table, state_table, unique_ids_h, edges_h, rel, inv_rel, dst_negs_h, src_negs_h, lr: float,
                          reduction: int = REDUCTION_SUM, precision: int = PREC_BF16X3, rel_grad=None, inv_rel_grad=None):
    """train_step_host without the final wait: returns (ticket, keep-alive) once the step is enqueued.  Pass the ticket to
    train_step_host_wait to get the loss; hold on to the second value (the host tensors the copies read from) until then.  At most
    two steps may be in flight."""
    _need_cuda(table, state_table, rel, inv_rel)
    for t in (unique_ids_h, edges_h, dst_negs_h, src_negs_h):
        if t is not None and t.is_cuda:
            raise MariusB200Error(_INVALID, "train_step_host takes host index tensors")
    b, keep = _make_batch(kind, unique_ids_h.numel(), table.size(1), edges_h, rel, inv_rel, dst_negs_h, src_negs_h, host=True)
    uid = unique_ids_h.contiguous()
    ticket = C.c_int(0)
    check(lib.mb_train_step_host_async(ctx.handle, C.byref(b), _ptr(table), _ptr(state_table), table.size(0), _rowmajor(table, "table"),
                                       C.c_void_p(uid.data_ptr()), float(lr), int(reduction), int(precision), _ptr(rel_grad), _ptr(inv_rel_grad),
                                       C.byref(ticket), _stream()))
    return int(ticket.value), (keep, uid)


def train_step_edges_host_async(ctx: Context, kind: int, table, state_table, edges_h, num_nodes: int, C_: int, N: int, seed: int, batch_index: int, rel, inv_rel,
                                lr: float, reduction: int = REDUCTION_SUM, precision: int = PREC_BF16X3, rel_grad=None, inv_rel_grad=None):
    """The step from RAW positive edges ([B,3] int64 host tensor, global node ids): negatives and the unique-id mapping are produced on
    the device.  Returns (ticket, keep-alive); train_step_host_wait(ctx, ticket) returns the loss."""
    _need_cuda(table, state_table, rel, inv_rel)
    if edges_h.is_cuda or edges_h.dim() != 2 or edges_h.size(1) != 3 or edges_h.dtype != torch.int64:
        raise MariusB200Error(_INVALID, "edges must be a host int64 [B, 3] tensor")
    e = edges_h.contiguous()
    ticket = C.c_int(0)
    check(lib.mb_train_step_edges_host_async(ctx.handle, int(kind), C.c_void_p(e.data_ptr()), e.size(0), int(num_nodes), int(C_), int(N), int(seed),
                                             int(batch_index), _ptr(rel), _ptr(inv_rel), rel.size(0), table.size(1), _ptr(table), _ptr(state_table),
                                             table.size(0), _rowmajor(table, "table"), float(lr), int(reduction), int(precision), _ptr(rel_grad),
                                             _ptr(inv_rel_grad), C.byref(ticket), _stream()))
    return int(ticket.value), (e,)


def train_step_host_wait(ctx: Context, ticket: int) -> float:
    loss = C.c_float(0.0)
    check(lib.mb_train_step_host_wait(ctx.handle, int(ticket), C.byref(loss)))
    return float(loss.value)


def train_step_sharded_host_async(ctx: Context, kind: int, shards, ld: int, d: int, unique_ids_h, edges_h, rel, inv_rel, dst_negs_h, src_negs_h, lr: float,
                                  reduction: int = REDUCTION_SUM, precision: int = PREC_BF16X3, rel_grad=None, inv_rel_grad=None):
    _need_cuda(rel, inv_rel)
    b, keep = _make_batch(kind, unique_ids_h.numel(), d, edges_h, rel, inv_rel, dst_negs_h, src_negs_h, host=True)
    uid = unique_ids_h.contiguous()
    ticket = C.c_int(0)
    check(lib.mb_train_step_sharded_host_async(ctx.handle, C.byref(b), C.byref(shards), int(ld), C.c_void_p(uid.data_ptr()), float(lr), int(reduction),
                                               int(precision), _ptr(rel_grad), _ptr(inv_rel_grad), C.byref(ticket), _stream()))
    return int(ticket.value), (keep, uid)


def ipc_export(t: torch.Tensor):
    """(64-byte CUDA IPC handle, byte offset) of a device tensor's storage -- picklable, for torch.distributed.all_gather_object."""
    _need_cuda(t)
    h = C.create_string_buffer(64)
    off = C.c_int64(0)
    check(lib.mb_ipc_export(_ptr(t), h, C.byref(off)))
    return bytes(h.raw), int(off.value)


def ipc_import(ctx: Context, handle: bytes, offset: int) -> int:
    """device pointer (int) of a peer's tensor opened in this process with this context's GPU current"""
    out = C.c_void_p()
    check(lib.mb_ipc_import(ctx.handle, C.create_string_buffer(handle, 64), int(offset), C.byref(out)))
    return int(out.value)


def shard_exchange_bytes(world: int, exchange_rows: int, d: int) -> int:
    """bytes of one rank's exchange area (barrier flags + one inbox of `exchange_rows` gradient rows per sender)"""
    n = int(lib.mb_shard_exchange_bytes(int(world), int(exchange_rows), int(d)))
    if n < 0:
        raise MariusB200Error(_INVALID, "bad exchange dimensions")
    return n


def make_shards_raw(table_ptrs, state_ptrs, rows_per_rank: int, rank: int = 0, exchange_ptrs=None, exchange_rows: int = 0, single_process: bool = False):
    sh = mb_shards()
    sh.rank = int(rank)
    for i, (t, s_) in enumerate(zip(table_ptrs, state_ptrs)):
        sh.tables[i] = int(t)
        sh.states[i] = int(s_)
    sh.world = len(table_ptrs)
    sh.rows_per_rank = int(rows_per_rank)
    if exchange_ptrs is not None:
        for i, e in enumerate(exchange_ptrs):
            sh.exchange[i] = int(e)
    sh.exchange_rows = int(exchange_rows)
    sh.single_process = int(bool(single_process))
    return sh


def make_shards(tables, states, rows_per_rank: int, rank: int = 0, exchange_rows: int = 1 << 14):
    """mb_shards over per-owner table / state tensors that all live in THIS process (one GPU holding several shards: tests, or a
    single-process driver): the exchange areas are allocated here and the call serves every owner (single_process).  The returned
    structure keeps the exchange tensors alive through its `_keep` attribute."""
    if len(tables) != len(states) or not (1 <= len(tables) <= 8):
        raise MariusB200Error(_INVALID, "1..8 shards")
    sh = mb_shards()
    keep = []
    for i, (t, s_) in enumerate(zip(tables, states)):
        _need_cuda(t, s_)
        if t.shape != tables[0].shape or s_.shape != t.shape or _rowmajor(t, "table") != _rowmajor(tables[0], "table"):
            raise MariusB200Error(_INVALID, "all shards must have the same shape and stride")
        sh.tables[i] = t.data_ptr()
        sh.states[i] = s_.data_ptr()
    sh.world = len(tables)
    sh.rows_per_rank = int(rows_per_rank)
    sh.rank = int(rank)
    if sh.world > 1:
        nbytes = shard_exchange_bytes(sh.world, exchange_rows, tables[0].size(1))
        for i in range(sh.world):
            ex = torch.zeros(nbytes, dtype=torch.uint8, device=tables[0].device)
            keep.append(ex)
            sh.exchange[i] = ex.data_ptr()
        sh.exchange_rows = int(exchange_rows)
    sh.single_process = 1
    sh._keep = keep
    return sh


def shard_error(ctx: Context, shards) -> int:
    """non-zero once a cross-rank barrier of the sharded step timed out (sticky)"""
    e = C.c_int(0)
    check(lib.mb_shard_error(ctx.handle, C.byref(shards), C.byref(e)))
    return int(e.value)


def train_step_sharded(ctx: Context, kind: int, shards, ld: int, d: int, unique_ids, edges, rel, inv_rel, dst_negs, src_negs, lr: float,
                       reduction: int = REDUCTION_SUM, precision: int = PREC_BF16X3, loss=None, rel_grad=None, inv_rel_grad=None):
    """mb_train_step on a table sharded over peer GPUs (global unique ids; remote rows over NVLink loads/stores)."""
    _need_cuda(unique_ids, edges, rel, inv_rel, dst_negs, src_negs)
    _check_indices(unique_ids)
    b, keep = _make_batch(kind, unique_ids.numel(), d, edges, rel, inv_rel, dst_negs, src_negs)
    if loss is None:
        loss = torch.empty(1, dtype=torch.float32, device=unique_ids.device)
    check(lib.mb_train_step_sharded(ctx.handle, C.byref(b), C.byref(shards), int(ld), _ptr(unique_ids), float(lr), int(reduction), int(precision),
                                    _ptr(loss), _ptr(rel_grad), _ptr(inv_rel_grad), _stream()))
    return loss


def train_step_sharded_host(ctx: Context, kind: int, shards, ld: int, d: int, unique_ids_h, edges_h, rel, inv_rel, dst_negs_h, src_negs_h, lr: float,
                            reduction: int = REDUCTION_SUM, precision: int = PREC_BF16X3, rel_grad=None, inv_rel_grad=None) -> float:
    _need_cuda(rel, inv_rel)
    b, keep = _make_batch(kind, unique_ids_h.numel(), d, edges_h, rel, inv_rel, dst_negs_h, src_negs_h, host=True)
    loss = C.c_float(0.0)
    uid = unique_ids_h.contiguous()
    check(lib.mb_train_step_sharded_host(ctx.handle, C.byref(b), C.byref(shards), int(ld), C.c_void_p(uid.data_ptr()), float(lr), int(reduction),
                                         int(precision), C.cast(C.pointer(loss), C.c_void_p), _ptr(rel_grad), _ptr(inv_rel_grad), _stream()))
    return float(loss.value)


def debug_gemm(ctx: Context, A: torch.Tensor, a_mn: bool, B: torch.Tensor, b_mn: bool, precision: int = PREC_BF16X3, block_n: int = 256):
    """Diagnostic: batched D = A . B over K through the contraction kernels (see mb_debug_gemm).  block_n 2 / 3: the backward kernels
    (A converted in the kernel; tensor-memory A / shared-memory A); 4 / 5: both backward problems in one grouped launch (identity / exp)."""
    _need_cuda(A, B)
    A, B = A.contiguous(), B.contiguous()
    batches = A.size(0)
    M, K = (A.size(2), A.size(1)) if a_mn else (A.size(1), A.size(2))
    N = B.size(2) if b_mn else B.size(1)
    if block_n in (4, 5):  # both backward problems from one square matrix: B holds [2 * batches][K][N], D gets [2 * batches][M][N]
        batches = A.size(0)
        D = torch.empty((2 * batches, M, N), dtype=torch.float32, device=A.device)
    else:
        D = torch.empty((batches, M, N), dtype=torch.float32, device=A.device)
    check(lib.mb_debug_gemm(ctx.handle, _ptr(A), int(a_mn), _ptr(B), int(b_mn), _ptr(D), M, N, K, batches, int(precision), int(block_n), _stream()))
    return D


def debug_wait_log():
    """Diagnostics (MB_TC_WAITLOG=1): [(block, thread, barrier shared address, parity)] of the bounded waits that gave up (see mb_debug_wait_log)."""
    import ctypes as C

    buf = (C.c_uint64 * 1000)()
    n = lib.mb_debug_wait_log(buf, 1000)
    return [(int(buf[2 * i] >> 32), int(buf[2 * i] & 0xFFFFFFFF), int(buf[2 * i + 1] >> 32), int(buf[2 * i + 1] & 0xFFFFFFFF)) for i in range(n)]
