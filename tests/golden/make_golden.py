"""tests/golden/make_golden.py -- regenerate the golden fixtures from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   make -C oracle && python tests/golden/make_golden.py
Every array in tests/golden/*.npz is an input to, or an output of, the reference C++ itself
(oracle/_ref/libmarius_ref.so = the reference's TUs compiled in place + oracle/ref_driver.cpp):
  InMemory::indexRead/indexAdd, PartitionBuffer::{indexRead,indexAdd,getGlobalToLocalMap,getNextAdmit,getNextEvict},
  map_tensors, Model::forward_lp, Model::train_batch, Model::evaluate_batch (+ LinkPredictionReporter::computeRanks, ranking metrics), compute_filter_corruption_cpu.
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import marius_oracle as O  # noqa: E402
from oracle import ref_lib as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def train_case(name, kind, B, C, N, d, num_nodes, num_rel, reduction, seed, lr=0.1, emb_scale=0.5):
    rng = np.random.default_rng(seed)
    uniq, edges, dn, sn = O.make_batch(rng, num_nodes, num_rel, B, C, N)
    U = len(uniq)
    emb = rng.uniform(-emb_scale, emb_scale, (U, d)).astype(np.float32)
    state = rng.uniform(0, 0.1, (U, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (num_rel, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (num_rel, d)).astype(np.float32)
    ref = R.train_batch(kind, emb, state, edges, rel, inv_rel, dn, sn, lr, reduction)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind=kind, B=B, C=C, N=N, d=d, lr=np.float32(lr), reduction=reduction, uniq=uniq, edges=edges,
                        dst_negs=dn, src_negs=sn, emb=emb, state=state, rel=rel, inv_rel=inv_rel, **{"ref_" + k: v for k, v in ref.items()})
    print(name, "U", U)


def eval_case(name, kind, B, C, N, d, num_nodes, num_rel, seed, n_filter, all_nodes=False, emb_scale=0.5):
    """Model::evaluate_batch: scores -> score filters -> ranks.  all_nodes: the filtered-evaluation shape (negative.cpp:321-325,355:
    one chunk whose negatives are every node, arange(num_nodes))."""
    rng = np.random.default_rng(seed)
    if all_nodes:
        U, C, N = num_nodes, 1, num_nodes
        edges = np.stack([rng.integers(0, U, B), rng.integers(0, num_rel, B), rng.integers(0, U, B)], axis=1).astype(np.int64)
        dn = np.arange(U, dtype=np.int64).reshape(1, U)
        sn = dn.copy()
    else:
        uniq, edges, dn, sn = O.make_batch(rng, num_nodes, num_rel, B, C, N)
        U = len(uniq)
    emb = rng.uniform(-emb_scale, emb_scale, (U, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (num_rel, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (num_rel, d)).astype(np.float32)
    Bp = C * int(np.ceil(B / C))
    mk = lambda: np.unique(np.stack([rng.integers(0, Bp, n_filter), rng.integers(0, N, n_filter)], axis=1).astype(np.int64), axis=0)
    dst_filter, src_filter = (mk(), mk()) if n_filter > 0 else (None, None)
    if all_nodes:  # the true destination / source of every edge is always filtered (it is among the negatives)
        dst_filter = np.unique(np.concatenate([dst_filter, np.stack([np.arange(B), edges[:, 2]], axis=1)]), axis=0)
        src_filter = np.unique(np.concatenate([src_filter, np.stack([np.arange(B), edges[:, 0]], axis=1)]), axis=0)
    ref = R.evaluate_batch(kind, emb, edges, rel, inv_rel, dn, sn, dst_filter, src_filter)
    metrics = R.ranking_metrics(np.concatenate([ref["ranks"], ref["inv_ranks"]]))
    empty = np.zeros((0, 2), np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind=kind, B=B, C=C, N=N, d=d, edges=edges, dst_negs=dn, src_negs=sn, emb=emb, rel=rel,
                        inv_rel=inv_rel, dst_filter=empty if dst_filter is None else dst_filter, src_filter=empty if src_filter is None else src_filter,
                        metrics=np.array([metrics[k] for k in ("mean_rank", "mrr", "hits@1", "hits@3", "hits@10")]),
                        **{"ref_" + k: v for k, v in ref.items()})
    print(name, "U", U, metrics)


def filtered_eval_case(name, kind, num_nodes, num_rel, num_graph_edges, B, d, seed):
    """The reference's filtered evaluation end to end: global score filters built by compute_filter_corruption_cpu from the graph
    (negative.cpp:62-195), negatives = all nodes (negative.cpp:355), Model::evaluate_batch -> ranks; plus the local (in-batch) filter of
    the same edges against sampled negatives."""
    rng = np.random.default_rng(seed)
    graph = np.unique(np.stack([rng.integers(0, num_nodes, num_graph_edges), rng.integers(0, num_rel, num_graph_edges),
                                rng.integers(0, num_nodes, num_graph_edges)], axis=1).astype(np.int64), axis=0)
    graph = graph[rng.permutation(len(graph))]
    edges = np.ascontiguousarray(graph[rng.choice(len(graph), B, replace=False)])
    all_nodes = np.arange(num_nodes, dtype=np.int64).reshape(1, num_nodes)
    dst_filter = R.compute_filter(edges, all_nodes, False, graph_edges=graph, num_nodes=num_nodes)
    src_filter = R.compute_filter(edges, all_nodes, True, graph_edges=graph, num_nodes=num_nodes)
    local_negs = rng.integers(0, num_nodes, (4, 50)).astype(np.int64)
    local_dst = R.compute_filter(edges, local_negs, False)
    local_src = R.compute_filter(edges, local_negs, True)
    emb = rng.uniform(-0.5, 0.5, (num_nodes, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (num_rel, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (num_rel, d)).astype(np.float32)
    ref = R.evaluate_batch(kind, emb, edges, rel, inv_rel, all_nodes, all_nodes.copy(), dst_filter, src_filter)
    metrics = R.ranking_metrics(np.concatenate([ref["ranks"], ref["inv_ranks"]]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind=kind, B=B, C=1, N=num_nodes, d=d, graph=graph, edges=edges, dst_negs=all_nodes,
                        src_negs=all_nodes, emb=emb, rel=rel, inv_rel=inv_rel, dst_filter=dst_filter, src_filter=src_filter, local_negs=local_negs,
                        local_dst_filter=local_dst, local_src_filter=local_src,
                        metrics=np.array([metrics[k] for k in ("mean_rank", "mrr", "hits@1", "hits@3", "hits@10")]),
                        **{"ref_" + k: v for k, v in ref.items()})
    print(name, "graph edges", len(graph), "filter sizes", dst_filter.shape[0], src_filter.shape[0], local_dst.shape[0], local_src.shape[0], metrics)


def storage_case():
    rng = np.random.default_rng(7)
    table = rng.standard_normal((257, 24)).astype(np.float32)
    idx = rng.integers(0, 257, size=100, dtype=np.int64)          # duplicates allowed for reads
    read = R.index_read(table, idx)
    uidx = rng.permutation(257)[:90].astype(np.int64)             # unique for adds (buffer.cpp:459)
    vals = rng.standard_normal((90, 24)).astype(np.float32)
    after = table.copy()
    R.index_add(after, uidx, vals)
    throws = R.index_read_bad_rank_throws(table)
    all_ids = rng.integers(0, 50, size=200, dtype=np.int64)
    uniq, mapped = R.map_tensors(all_ids)
    np.savez_compressed(os.path.join(OUT, "storage.npz"), table=table, idx=idx, read=read, uidx=uidx, vals=vals, after=after, bad_rank_throws=throws,
                        all_ids=all_ids, uniq=uniq, mapped=mapped)
    print("storage ok, bad rank throws:", throws)


def buffer_case():
    # test/cpp/unit/test_buffer.cpp:20-75 : 45 rows, 5 partitions of 10 (last has 5), capacity 2, and the
    # 10-state ordering of TestPartitionBufferOrdering (:241-259); embedding_size reduced 10000 -> 16.
    rng = np.random.default_rng(11)
    total, nparts, psize, d, cap = 45, 5, 10, 16, 2
    table = rng.standard_normal((total, d)).astype(np.float32)
    states = np.array([[0, 1], [0, 2], [0, 3], [0, 4], [1, 4], [1, 3], [1, 2], [3, 2], [4, 2], [4, 3]], dtype=np.int64)
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "emb.bin")
        table.tofile(fn)
        idx = rng.permutation(20)[:12].astype(np.int64)           # buffer-local ids, unique
        vals = rng.standard_normal((12, d)).astype(np.float32)
        res = R.partition_buffer_exercise(fn, cap, nparts, psize, d, total, states, idx, vals)
        file_after = np.fromfile(fn, dtype=np.float32).reshape(total, d)
    np.savez_compressed(os.path.join(OUT, "partition_buffer.npz"), table=table, states=states, idx=idx, vals=vals, file_after=file_after,
                        total=total, nparts=nparts, psize=psize, d=d, cap=cap, **res)
    print("buffer admits", res["admits"], "evicts", res["evicts"])


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first: make -C oracle"
    storage_case()
    buffer_case()
    train_case("train_distmult_pad", O.DISTMULT, 7, 3, 5, 8, 40, 3, O.REDUCTION_SUM, 1)
    train_case("train_complex_pad", O.COMPLEX, 7, 3, 5, 8, 40, 3, O.REDUCTION_SUM, 2)
    train_case("train_distmult_mean", O.DISTMULT, 64, 2, 32, 16, 300, 5, O.REDUCTION_MEAN, 3)
    train_case("train_complex_mid", O.COMPLEX, 96, 3, 64, 48, 2000, 7, O.REDUCTION_SUM, 4)
    train_case("train_distmult_dup", O.DISTMULT, 128, 4, 64, 32, 60, 4, O.REDUCTION_SUM, 5)   # heavy id collisions
    eval_case("eval_distmult_pad", O.DISTMULT, 7, 3, 5, 8, 40, 3, 21, 0)                      # padded rows get rank N + 1
    eval_case("eval_complex_filter", O.COMPLEX, 96, 2, 64, 32, 500, 7, 22, 200)               # sampled negatives + score filters
    eval_case("eval_distmult_all", O.DISTMULT, 50, 1, 0, 16, 120, 5, 23, 60, all_nodes=True)  # filtered evaluation against all nodes
    filtered_eval_case("eval_filtered_graph", O.COMPLEX, 320, 6, 4000, 64, 32, 24)              # global + local filters from a graph
    train_case("train_complex_d100", O.COMPLEX, 200, 2, 100, 100, 14541, 237, O.REDUCTION_SUM, 6, emb_scale=0.1)  # FB15k-237-sized
