"""oracle/marius_oracle.py -- CPU restatement (numpy) of Marius's per-batch embedding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``marius_b200/`` may import this module; it is the
checker for ``tests/``, ``__graft_entry__.smoke()`` and (as ``cpu_baseline`` "port") ``bench.py``.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function below against
  * the reference's own known-answer vectors (DistMult ``[12.5, -3.75, -0.25]``,
    ``test/python/bindings/integration/test_nn.py:148-160``; the Adagrad rule,
    ``test_data.py:34-47``; the PartitionBuffer global->local map, ``test/cpp/unit/test_buffer.cpp:310-318``),
  * ``tests/golden/*.npz`` -- outputs of the *unmodified* reference C++ (``oracle/_ref/libmarius_ref.so``,
    built by ``oracle/Makefile`` from ``/root/reference``) produced by ``tests/golden/make_golden.py``,
  * and, when ``oracle/_ref`` is present, the reference library directly on fresh seeded inputs.

All citations are relative to ``/root/reference/src/cpp``.  ids are int64, values fp32, row-major.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

F32 = np.float32

# decoder kinds (mirrors DecoderType, include/configuration/options.h:60, restricted to DotCompare models)
DOT = 0        # no relation operator (2-column edges), DotCompare          comparators.cpp:62-72
DISTMULT = 1   # HadamardOperator + DotCompare                              distmult.cpp:7-19
COMPLEX = 2    # ComplexHadamardOperator + DotCompare                       complex.cpp:7-19

REDUCTION_MEAN = 0  # LossReduction::MEAN  options.h:24
REDUCTION_SUM = 1   # LossReduction::SUM


# ----------------------------------------------------------------------------------------------
# storage: gather / scatter-add                                            (a3,a4,a16,a17)
# ----------------------------------------------------------------------------------------------
def _check_indices(indices: np.ndarray) -> None:
    # storage.cpp:607-610 / buffer.cpp:442-445 : indices must be 1-D else std::runtime_error
    if indices.ndim != 1:
        raise RuntimeError("indices must be 1-dimensional")


def index_read(table: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """InMemory::indexRead (storage/storage.cpp:606-649) == PartitionBuffer::indexRead
    (storage/buffer.cpp:441-455): ``out[i,:] = table[indices[i],:]`` -- a pure copy, bit-exact."""
    _check_indices(indices)
    return np.ascontiguousarray(table[indices])


def index_add(table: np.ndarray, indices: np.ndarray, values: np.ndarray) -> None:
    """InMemory::indexAdd (storage.cpp:651-673) == PartitionBuffer::indexAdd (buffer.cpp:459-480):
    ``table[indices[i], j] += values[i, j]`` in place.  Indices are unique within one call
    (buffer.cpp:459), so the fp32 result does not depend on order."""
    if values is None or indices.ndim != 1 or indices.shape[0] != values.shape[0] or table.shape[1] != values.shape[1]:
        raise RuntimeError("indexAdd: shape mismatch")  # storage.cpp:652-655
    # unique ids => fancy += is exact; np.add.at keeps reference semantics should a caller pass duplicates
    np.add.at(table, indices, values.astype(table.dtype, copy=False))


def global_to_local_map(total_embeddings: int, partition_size: int, buffer_state, buffer_slots=None) -> np.ndarray:
    """PartitionBuffer::getGlobalToLocalMap(true) (buffer.cpp:581-602): ``map[g] = slot*partition_size + (g - p*partition_size)``
    for every partition ``p`` resident in slot ``slot``; -1 elsewhere.  ``buffer_slots[i]`` is the slot of
    ``buffer_state[i]`` (defaults to i, as after PartitionBuffer::load, buffer.cpp:386-395)."""
    m = -np.ones(total_embeddings, dtype=np.int64)
    if buffer_slots is None:
        buffer_slots = list(range(len(buffer_state)))
    for p, slot in zip(buffer_state, buffer_slots):
        lo = int(p) * partition_size
        hi = min(lo + partition_size, total_embeddings)  # last partition may be short (buffer.cpp:347-350)
        m[lo:hi] = np.arange(slot * partition_size, slot * partition_size + (hi - lo), dtype=np.int64)
    return m


def map_tensors(all_ids: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """map_tensors (common/util.cpp:180-205): ``torch::_unique2(cat, sorted=true, return_inverse=true)``.
    Returns (sorted unique ids, position of every input id in that list)."""
    uniq, inv = np.unique(all_ids, return_inverse=True)
    return uniq.astype(np.int64), inv.astype(np.int64).reshape(all_ids.shape)


# ----------------------------------------------------------------------------------------------
# decoder pieces                                                           (a7-a10)
# ----------------------------------------------------------------------------------------------
def apply_relation(kind: int, embs: np.ndarray, rels: Optional[np.ndarray]) -> np.ndarray:
    """HadamardOperator / ComplexHadamardOperator (nn/decoders/edge/relation_operators.cpp:7-35)."""
    if kind == DOT or rels is None:
        return embs
    if kind == DISTMULT:
        return embs * rels
    if kind == COMPLEX:
        d = embs.shape[1]
        h = d // 2
        er, ei = embs[:, :h], embs[:, h:]
        rr, ri = rels[:, :h], rels[:, h:]
        out = np.zeros_like(embs)
        out[:, :h] = er * rr - ei * ri
        out[:, h:] = er * ri + ei * rr
        return out
    raise ValueError(kind)


def apply_relation_backward(kind: int, g: np.ndarray, embs: np.ndarray, rels: Optional[np.ndarray]):
    """Gradient of apply_relation w.r.t. (embs, rels) for upstream gradient ``g``."""
    if kind == DOT or rels is None:
        return g, None
    if kind == DISTMULT:
        return g * rels, g * embs
    if kind == COMPLEX:
        h = embs.shape[1] // 2
        er, ei = embs[:, :h], embs[:, h:]
        rr, ri = rels[:, :h], rels[:, h:]
        gr, gi = g[:, :h], g[:, h:]
        de = np.empty_like(embs)
        dr = np.empty_like(embs)
        de[:, :h] = gr * rr + gi * ri
        de[:, h:] = -gr * ri + gi * rr
        dr[:, :h] = gr * er + gi * ei
        dr[:, h:] = -gr * ei + gi * er
        return de, dr
    raise ValueError(kind)


def pad_and_reshape(x: np.ndarray, num_chunks: int) -> np.ndarray:
    """pad_and_reshape (comparators.cpp:7-20): zero-pad rows to C*ceil(B/C) and view as [C, ceil(B/C), d]."""
    b = x.shape[0]
    per = int(math.ceil(b / num_chunks))
    if per * num_chunks != b:
        x = np.concatenate([x, np.zeros((per * num_chunks - b, x.shape[1]), dtype=x.dtype)], axis=0)
    return x.reshape(num_chunks, per, x.shape[1])


def dot_compare(src: np.ndarray, dst: np.ndarray, acc=F32) -> np.ndarray:
    """DotCompare::operator() (comparators.cpp:62-72)."""
    if src.shape == dst.shape:
        return (src.astype(acc) * dst.astype(acc)).sum(-1).astype(F32)
    s = pad_and_reshape(src, dst.shape[0])
    out = np.matmul(s.astype(acc), np.transpose(dst.astype(acc), (0, 2, 1)))
    return out.reshape(-1, out.shape[-1]).astype(F32)


@dataclass
class Scores:
    pos: np.ndarray                 # [Bp]
    neg: np.ndarray                 # [Bp, N]
    inv_pos: Optional[np.ndarray]   # [Bp]
    inv_neg: Optional[np.ndarray]   # [Bp, N]


def node_corrupt_forward(kind, emb, edges, rel, inv_rel, dst_negs, src_negs, acc=F32) -> Scores:
    """node_corrupt_forward (nn/decoders/edge/decoder_methods.cpp:57-114).  ``emb`` [U,d] batch-local rows,
    ``edges`` [B,3] (or [B,2]) local ids, ``dst_negs``/``src_negs`` [C,N] local ids."""
    src = emb[edges[:, 0]]
    dst = emb[edges[:, -1]]
    has_rel = edges.shape[1] == 3 and kind != DOT
    r = rel[edges[:, 1]] if has_rel else None
    a = apply_relation(kind, src, r)
    pos = dot_compare(a, dst, acc)
    neg = dot_compare(a, emb[dst_negs.reshape(-1)].reshape(dst_negs.shape[0], dst_negs.shape[1], -1), acc)
    inv_pos = inv_neg = None
    if has_rel and inv_rel is not None and src_negs is not None:
        ri = inv_rel[edges[:, 1]]
        b = apply_relation(kind, dst, ri)
        inv_pos = dot_compare(b, src, acc)
        inv_neg = dot_compare(b, emb[src_negs.reshape(-1)].reshape(src_negs.shape[0], src_negs.shape[1], -1), acc)
    if pos.shape[0] != neg.shape[0]:  # decoder_methods.cpp:103-111
        extra = neg.shape[0] - pos.shape[0]
        pos = np.concatenate([pos, np.zeros(extra, dtype=F32)])
        if inv_pos is not None:
            inv_pos = np.concatenate([inv_pos, np.zeros(extra, dtype=F32)])
    return Scores(pos, neg, inv_pos, inv_neg)


# ----------------------------------------------------------------------------------------------
# batch assembly: negative sampling + edgeSample                           
# ----------------------------------------------------------------------------------------------
def philox4x32_10_u64(seed: int, index: np.ndarray, stream: int, batch: int) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11; the generator cuRAND / libtorch CUDA use): key = seed, counter = (index lo, index hi,
    stream, batch); returns the first two output words as one uint64 per index.  This is the definition of the device sampler's
    stream (marius_b200/csrc/sample_kernels.cu), restated with numpy integer arithmetic."""
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    idx = index.astype(np.uint64)
    c0, c1 = idx & MASK, idx >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(stream))
    c3 = np.full_like(c0, np.uint64(batch))
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2  # 32x32 -> 64-bit products
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ k0, p1 & MASK, (p0 >> np.uint64(32)) ^ c3 ^ k1, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0 | (c1 << np.uint64(32))


def sample_negatives(num_nodes: int, C: int, N: int, seed: int, batch_index: int, inverse: bool, degree_fraction: float = 0.0,
                     edges: Optional[np.ndarray] = None) -> np.ndarray:
    """CorruptNodeNegativeSampler::getNegatives (data/samplers/negative.cpp:328-366): per chunk, ``(int)(N * degree_fraction)``
    endpoints of random batch edges (batch_sample, negative.cpp:7-19: source column when ``inverse``, else destination) followed by
    uniform ids in [0, num_nodes) -- ``random % range`` like libtorch's randint, with Philox instead of libtorch's generator."""
    num_batch = int(np.float32(N) * np.float32(degree_fraction))
    r = philox4x32_10_u64(seed, np.arange(C * N, dtype=np.uint64), 1 if inverse else 0, batch_index)
    j = np.arange(C * N) % max(N, 1)
    out = (r % np.uint64(num_nodes)).astype(np.int64)
    if num_batch > 0:
        col = 0 if inverse else edges.shape[1] - 1
        deg = edges[(r % np.uint64(edges.shape[0])).astype(np.int64), col]
        out = np.where(j < num_batch, deg, out)
    return out.reshape(C, N)


def edge_sample(edges: np.ndarray, src_negs: Optional[np.ndarray], dst_negs: np.ndarray):
    """DataLoader::edgeSample without a neighbour sampler (data/dataloader.cpp:389-471): map_tensors over
    cat(src, dst, src_negs, dst_negs), edges rewritten to (mapped src, rel, mapped dst), negatives keep their [C,N] shape.
    Returns (unique ids, local edges, local src_negs or None, local dst_negs)."""
    B = edges.shape[0]
    parts = [edges[:, 0], edges[:, -1]] + ([src_negs.reshape(-1)] if src_negs is not None else []) + [dst_negs.reshape(-1)]
    uniq, mapped = map_tensors(np.concatenate(parts))
    local = edges.copy()
    local[:, 0], local[:, -1] = mapped[:B], mapped[B:2 * B]
    off = 2 * B
    s_loc = None
    if src_negs is not None:
        s_loc = mapped[off:off + src_negs.size].reshape(src_negs.shape)
        off += src_negs.size
    return uniq, local, s_loc, mapped[off:].reshape(dst_negs.shape)


# ----------------------------------------------------------------------------------------------
# evaluation: score filter, ranks, ranking metrics                         
# ----------------------------------------------------------------------------------------------
def apply_score_filter(scores: np.ndarray, filt: Optional[np.ndarray]) -> np.ndarray:
    """apply_score_filter (data/samplers/negative.cpp:306-311): ``scores[filter[:,0], filter[:,1]] = -1e9`` in place
    (called on the negative scores of each side by Model::forward_lp, nn/model.cpp:279-285)."""
    if filt is not None and filt.shape[0] > 0:
        scores[filt[:, 0], filt[:, 1]] = F32(-1e9)
    return scores


def compute_filter_corruption(edges: np.ndarray, corruption_nodes: np.ndarray, inverse: bool, graph_edges: Optional[np.ndarray] = None) -> np.ndarray:
    """compute_filter_corruption_cpu (data/samplers/negative.cpp:62-195): the [F,2] (row, column) pairs whose negative score must be
    masked because the corrupted triple is a true edge.

    ``graph_edges`` given  -> GLOBAL filter (``filtered`` evaluation, negative.cpp:152-163): the negatives are all nodes
        (``arange(num_nodes)``), so for batch edge i every graph edge with the same kept endpoint (source, or destination when
        ``inverse``) and the same relation contributes (i, its corrupted endpoint's node id).  Order: by batch edge, then in the
        order of the graph's edge list sorted (stably) by the kept endpoint.
    ``graph_edges`` None   -> LOCAL filter against the batch itself (negative.cpp:164-182): (i, j) if replacing the corrupted
        endpoint of edge i by negative j of i's chunk gives another edge of the batch (first match per negative)."""
    has_rel = edges.shape[1] == 3
    tup, cor = ((edges.shape[1] - 1), 0) if inverse else (0, edges.shape[1] - 1)
    nodes = edges[:, tup]
    pool = graph_edges if graph_edges is not None else edges
    pool = pool[np.argsort(pool[:, tup], kind="stable")]
    keys = pool[:, tup]
    starts, ends = np.searchsorted(keys, nodes, side="left"), np.searchsorted(keys, nodes + 1, side="left")
    num_chunks = corruption_nodes.shape[0]
    chunk_size = int(math.ceil(edges.shape[0] / num_chunks))
    out = []
    for i in range(edges.shape[0]):
        cand = pool[starts[i]:ends[i]]
        if has_rel:
            cand = cand[cand[:, 1] == edges[i, 1]]
        if graph_edges is not None:
            out.extend((i, int(c)) for c in cand[:, cor])
        else:
            present = set(int(c) for c in cand[:, cor])
            out.extend((i, j) for j, n in enumerate(corruption_nodes[i // chunk_size]) if int(n) in present)
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def compute_ranks(pos: np.ndarray, neg: np.ndarray) -> np.ndarray:
    """LinkPredictionReporter::computeRanks (reporting/reporting.cpp:56-58): ``(neg >= pos.unsqueeze(1)).sum(1) + 1`` as int64.
    Padding rows (pos = 0, all-zero scores) get rank N + 1, exactly as in the reference."""
    return (neg >= pos[:, None]).sum(axis=1).astype(np.int64) + 1


def ranking_metrics(ranks: np.ndarray, ks=(1, 3, 10)) -> dict:
    """MeanRankMetric / MeanReciprocalRankMetric / HitskMetric (reporting/reporting.cpp:17-31): mean rank in float64, MRR as the
    float32 mean of float32 reciprocals, Hits@k = #(rank <= k) / n in float64."""
    out = {"mean_rank": float(ranks.astype(np.float64).mean()), "mrr": float((F32(1.0) / ranks.astype(F32)).mean(dtype=F32))}
    for k in ks:
        out[f"hits@{k}"] = float((ranks <= k).sum()) / ranks.shape[0]
    return out


def evaluate_batch(kind, emb, edges, rel, inv_rel, dst_negs, src_negs, dst_filter=None, src_filter=None, acc=F32):
    """Model::evaluate_batch (nn/model.cpp:335-349): forward_lp with the batch's filters applied, then one computeRanks per side
    (the reporter receives the dst-corruption ranks first, then the src-corruption ranks).  Returns (ranks, inv_ranks or None, Scores)."""
    sc = node_corrupt_forward(kind, emb, edges, rel, inv_rel, dst_negs, src_negs, acc)
    apply_score_filter(sc.neg, dst_filter)
    ranks = compute_ranks(sc.pos, sc.neg)
    inv_ranks = None
    if sc.inv_neg is not None:
        apply_score_filter(sc.inv_neg, src_filter)
        inv_ranks = compute_ranks(sc.inv_pos, sc.inv_neg)
    return ranks, inv_ranks, sc


# ----------------------------------------------------------------------------------------------
# loss                                                                     (a11)
# ----------------------------------------------------------------------------------------------
def softmax_ce(pos: np.ndarray, neg: np.ndarray, reduction: int, acc=F32):
    """SoftmaxCrossEntropy (nn/loss.cpp:50-67): CE([pos_i, logsumexp_j neg_ij], label 0).
    Returns (loss, dloss/dpos [Bp], dloss/dneg [Bp,N])."""
    p = pos.astype(acc)
    n = neg.astype(acc)
    m = np.maximum(p, n.max(axis=1))
    z = m + np.log(np.exp(p - m) + np.exp(n - m[:, None]).sum(axis=1))  # log(e^pos + sum e^neg)
    li = z - p
    w = acc(1.0) if reduction == REDUCTION_SUM else acc(1.0 / pos.shape[0])
    loss = (li.sum() * w).astype(F32)
    gpos = ((np.exp(p - z) - 1.0) * w).astype(F32)
    gneg = (np.exp(n - z[:, None]) * w).astype(F32)
    return loss, gpos, gneg


# ----------------------------------------------------------------------------------------------
# sparse Adagrad                                                           (a13)
# ----------------------------------------------------------------------------------------------
def accumulate_gradients(grad: np.ndarray, state: np.ndarray, lr: float):
    """Batch::accumulateGradients (data/batch.cpp:62-79), fp32 op-for-op:
    ``ds = g^2 ; s = s + ds ; de = -lr * (g / (sqrt(s) + 1e-10))``.  Returns (delta_e, delta_s)."""
    g = grad.astype(F32)
    ds = g * g
    s = state.astype(F32) + ds
    de = F32(-lr) * (g / (np.sqrt(s) + F32(1e-10)))
    return de.astype(F32), ds.astype(F32)


# ----------------------------------------------------------------------------------------------
# one training batch = a6..a13                                             model.cpp:290-333
# ----------------------------------------------------------------------------------------------
@dataclass
class TrainBatchResult:
    scores: Scores
    loss: np.float32
    grad: np.ndarray          # [U,d]  dLoss/d node_embeddings_
    delta_e: np.ndarray       # [U,d]  node_gradients_ after accumulateGradients
    delta_s: np.ndarray       # [U,d]  node_state_update_
    rel_grad: Optional[np.ndarray]      # [R,d]
    inv_rel_grad: Optional[np.ndarray]  # [R,d]


def _one_side_backward(kind, emb, a, e_head, r, e_tail, head_ids, tail_ids, neg_ids, gpos, gneg, grad, rel_ids, rel_grad, acc):
    """Backward of one corruption side: scores = <a, e_tail>, neg = pad(a) . emb[neg_ids]^T with a = op(e_head, r)."""
    B, d = a.shape
    C, N = neg_ids.shape
    a_pad = pad_and_reshape(a, C).astype(acc)                  # [C,Bc,d]
    neg_e = emb[neg_ids.reshape(-1)].reshape(C, N, d).astype(acc)
    g3 = gneg.reshape(C, -1, N).astype(acc)                    # [C,Bc,N]
    da = np.matmul(g3, neg_e).reshape(-1, d)[:B]               # bmm backward wrt src (pad rows dropped)
    dneg = np.matmul(np.transpose(g3, (0, 2, 1)), a_pad)       # [C,N,d]
    da = da + gpos[:B, None].astype(acc) * e_tail.astype(acc)
    dtail = gpos[:B, None].astype(acc) * a.astype(acc)
    dhead, dr = apply_relation_backward(kind, da.astype(F32), e_head, r)
    np.add.at(grad, head_ids, dhead.astype(F32))
    np.add.at(grad, tail_ids, dtail.astype(F32))
    np.add.at(grad, neg_ids.reshape(-1), dneg.reshape(-1, d).astype(F32))
    if dr is not None:
        np.add.at(rel_grad, rel_ids, dr.astype(F32))


def train_batch(kind, emb, state, edges, rel, inv_rel, dst_negs, src_negs, lr, reduction=REDUCTION_SUM, acc=F32) -> TrainBatchResult:
    """Model::train_batch for link prediction (nn/model.cpp:290-333): forward_lp -> SoftmaxCE both sides
    (model.cpp:309-312) -> backward -> Batch::accumulateGradients.  The dense relation optimizer step is
    not part of this function (relation gradients are returned)."""
    emb = emb.astype(F32)
    sc = node_corrupt_forward(kind, emb, edges, rel, inv_rel, dst_negs, src_negs, acc)
    U, d = emb.shape
    B = edges.shape[0]
    has_rel = edges.shape[1] == 3 and kind != DOT
    grad = np.zeros((U, d), dtype=F32)
    rel_grad = np.zeros_like(rel) if has_rel else None
    inv_rel_grad = None

    src = emb[edges[:, 0]]
    dst = emb[edges[:, -1]]
    r = rel[edges[:, 1]] if has_rel else None
    a = apply_relation(kind, src, r)

    loss, gpos, gneg = softmax_ce(sc.pos, sc.neg, reduction, acc)
    _one_side_backward(kind, emb, a, src, r, dst, edges[:, 0], edges[:, -1], dst_negs, gpos, gneg, grad,
                       edges[:, 1] if has_rel else None, rel_grad, acc)
    if sc.inv_neg is not None:
        ri = inv_rel[edges[:, 1]]
        b = apply_relation(kind, dst, ri)
        inv_rel_grad = np.zeros_like(inv_rel)
        l2, gpos2, gneg2 = softmax_ce(sc.inv_pos, sc.inv_neg, reduction, acc)
        loss = F32(l2 + loss)  # model.cpp:312  loss = lhs_loss + rhs_loss
        _one_side_backward(kind, emb, b, dst, ri, src, edges[:, -1], edges[:, 0], src_negs, gpos2, gneg2, grad,
                           edges[:, 1], inv_rel_grad, acc)

    de, ds = accumulate_gradients(grad, state, lr)
    return TrainBatchResult(sc, F32(loss), grad, de, ds, rel_grad, inv_rel_grad)


def train_step_on_table(kind, table, state_table, unique_ids, edges, rel, inv_rel, dst_negs, src_negs, lr,
                        reduction=REDUCTION_SUM, acc=F32) -> TrainBatchResult:
    """The synchronous trainer's per-batch sequence (pipeline/trainer.cpp:106-138):
    gather emb + state (dataloader.cpp:505-548) -> train_batch -> indexAdd x2 (dataloader.cpp:550-564).
    Mutates ``table`` / ``state_table`` in place."""
    emb = index_read(table, unique_ids)
    st = index_read(state_table, unique_ids)
    res = train_batch(kind, emb, st, edges, rel, inv_rel, dst_negs, src_negs, lr, reduction, acc)
    index_add(table, unique_ids, res.delta_e)
    index_add(state_table, unique_ids, res.delta_s)
    return res


# ----------------------------------------------------------------------------------------------
# synthetic batch construction the way DataLoader::edgeSample does         dataloader.cpp:389-471
# ----------------------------------------------------------------------------------------------
def make_batch(rng: np.random.Generator, num_nodes: int, num_rel: int, B: int, C: int, N: int, with_rel=True):
    """Uniform edges + uniform negatives over ``num_nodes`` (negative.cpp:342), then map_tensors over
    cat(src, dst, src_negs, dst_negs) (dataloader.cpp:399-409,447-461).
    Returns (unique_ids [U], edges_local [B,3|2], dst_negs_local [C,N], src_negs_local [C,N])."""
    src = rng.integers(0, num_nodes, size=B, dtype=np.int64)
    dst = rng.integers(0, num_nodes, size=B, dtype=np.int64)
    relid = rng.integers(0, max(num_rel, 1), size=B, dtype=np.int64)
    src_negs = rng.integers(0, num_nodes, size=(C, N), dtype=np.int64)
    dst_negs = rng.integers(0, num_nodes, size=(C, N), dtype=np.int64)
    all_ids = np.concatenate([src, dst, src_negs.reshape(-1), dst_negs.reshape(-1)])
    uniq, inv = map_tensors(all_ids)
    s_l, d_l = inv[:B], inv[B:2 * B]
    sn_l = inv[2 * B:2 * B + C * N].reshape(C, N)
    dn_l = inv[2 * B + C * N:].reshape(C, N)
    edges = np.stack([s_l, relid, d_l], axis=1) if with_rel else np.stack([s_l, d_l], axis=1)
    return uniq, np.ascontiguousarray(edges), np.ascontiguousarray(dn_l), np.ascontiguousarray(sn_l)
