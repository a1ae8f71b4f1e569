"""Pins the numpy oracle (oracle/marius_oracle.py) to the reference:
(1) the reference tests' own known-answer vectors, (2) golden fixtures produced by the unmodified
reference C++ (tests/golden/make_golden.py), (3) the reference library itself when oracle/_ref is built."""
import glob
import os

import numpy as np
import pytest

from oracle import marius_oracle as O
from oracle import ref_lib as R


def rel_err(a, b, tau=1e-6):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), tau))


# ---- (1) reference known-answer vectors -------------------------------------------------------
def test_distmult_known_answer():
    # test/python/bindings/integration/test_nn.py:15-25,148-160
    emb = np.array([[1.5, 2.5], [2.5, 3.5], [4.25, 1.0], [-1.0, 0.5]], np.float32)
    edges = np.array([[0, 0, 1], [2, 0, 3], [3, 1, 0]], np.int64)
    rel = np.ones((2, 2), np.float32)  # DistMult::reset, distmult.cpp:21-28
    negs = np.array([[2, 0], [0, 1], [1, 0]], np.int64)  # test_nn.py:170
    sc = O.node_corrupt_forward(O.DISTMULT, emb, edges, rel, None, negs, None)
    assert np.array_equal(sc.pos, np.array([12.5, -3.75, -0.25], np.float32))
    assert sc.neg.shape == (3, 2)


def test_adagrad_known_answer():
    # test/python/bindings/integration/test_data.py:34-47
    g = np.array([0.5, -1.0], np.float32)
    de, ds = O.accumulate_gradients(g, np.zeros(2, np.float32), 1.0)
    assert np.array_equal(ds, g * g)
    expected = np.float32(-1.0) * (g / (np.sqrt(ds) + np.float32(1e-10)))
    assert np.array_equal(de, expected)


def test_global_to_local_map_known_answer():
    # test/cpp/unit/test_buffer.cpp:310-318
    m = O.global_to_local_map(45, 10, [0, 1])
    exp = -np.ones(45, np.int64)
    exp[:20] = np.arange(20)
    assert np.array_equal(m, exp)
    # after the first swap (admit 2, evict 1): partition 2 takes partition 1's slot
    exp[10:20] = -1
    exp[20:30] = np.arange(10, 20)
    assert np.array_equal(O.global_to_local_map(45, 10, [0, 2], [0, 1]), exp)


def test_index_errors():
    t = np.zeros((4, 3), np.float32)
    with pytest.raises(RuntimeError):  # storage.cpp:607-610, test_buffer.cpp:282
        O.index_read(t, np.zeros((2, 2), np.int64))
    with pytest.raises(RuntimeError):  # storage.cpp:652-655, test_buffer.cpp:294-296
        O.index_add(t, np.zeros(2, np.int64), np.zeros((3, 3), np.float32))
    with pytest.raises(RuntimeError):
        O.index_add(t, np.zeros(2, np.int64), np.zeros((2, 4), np.float32))


def test_pad_and_reshape():
    x = np.arange(14, dtype=np.float32).reshape(7, 2)
    y = O.pad_and_reshape(x, 3)  # comparators.cpp:7-20 : 7 rows, 3 chunks -> [3,3,2] with 2 zero rows
    assert y.shape == (3, 3, 2)
    assert np.array_equal(y.reshape(9, 2)[:7], x) and not y.reshape(9, 2)[7:].any()


# ---- (2) golden fixtures from the reference C++ ------------------------------------------------
def test_golden_storage(golden_dir):
    g = np.load(os.path.join(golden_dir, "storage.npz"))
    assert np.array_equal(O.index_read(g["table"], g["idx"]), g["read"])
    t = g["table"].copy()
    O.index_add(t, g["uidx"], g["vals"])
    assert np.array_equal(t, g["after"])
    assert bool(g["bad_rank_throws"])
    u, m = O.map_tensors(g["all_ids"])
    assert np.array_equal(u, g["uniq"]) and np.array_equal(m, g["mapped"])


def test_golden_partition_buffer(golden_dir):
    g = np.load(os.path.join(golden_dir, "partition_buffer.npz"))
    total, psize = int(g["total"]), int(g["psize"])
    states = g["states"]
    assert np.array_equal(O.global_to_local_map(total, psize, states[0]), g["map_current"])
    # next state: admitted partition takes the evicted partition's slot (buffer.cpp:603-630)
    assert np.array_equal(O.global_to_local_map(total, psize, [0, 2], [0, 1]), g["map_next"])
    # buffer-local reads == rows of the resident partitions
    cur = g["map_current"]
    local_to_global = {int(l): gl for gl, l in enumerate(cur) if l >= 0}
    rows = np.array([local_to_global[int(i)] for i in g["idx"]])
    assert np.array_equal(g["table"][rows], g["read"])
    assert np.array_equal(g["table"][rows] + g["vals"], g["read_after_add"])
    # write-back on unload(true): file rows updated exactly
    exp = g["table"].copy()
    exp[rows] += g["vals"]
    assert np.array_equal(exp, g["file_after"])
    # swap order of test_buffer.cpp:241-259
    assert list(g["admits"]) == [2, 3, 4, 1, 3, 2, 3, 4, 3]
    assert list(g["evicts"]) == [1, 2, 3, 0, 4, 3, 1, 3, 2]


TRAIN_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "train_*.npz")))


@pytest.mark.parametrize("name", TRAIN_CASES)
@pytest.mark.parametrize("acc", [np.float32, np.float64])
def test_golden_train_batch(golden_dir, name, acc):
    g = np.load(os.path.join(golden_dir, name))
    res = O.train_batch(int(g["kind"]), g["emb"], g["state"], g["edges"], g["rel"], g["inv_rel"], g["dst_negs"], g["src_negs"], float(g["lr"]),
                        int(g["reduction"]), acc=acc)
    tol = 2e-5
    assert rel_err(res.scores.pos, g["ref_pos"]) < tol
    assert rel_err(res.scores.neg, g["ref_neg"]) < tol
    assert rel_err(res.scores.inv_pos, g["ref_inv_pos"]) < tol
    assert rel_err(res.scores.inv_neg, g["ref_inv_neg"]) < tol
    assert abs(float(res.loss) - float(g["ref_loss"][0])) <= tol * abs(float(g["ref_loss"][0]))
    assert rel_err(res.grad, g["ref_grad"]) < tol
    assert rel_err(res.delta_e, g["ref_delta_e"]) < 1e-4
    assert rel_err(res.delta_s, g["ref_delta_s"]) < tol
    assert rel_err(res.rel_grad, g["ref_rel_grad"]) < tol
    assert rel_err(res.inv_rel_grad, g["ref_inv_rel_grad"]) < tol
    # Adagrad rule given the reference's own gradient (batch.cpp:62-79): delta_s is bit-exact; delta_e is
    # within 1 ulp -- libtorch 2.11's AVX512 vectorised sqrt is not correctly rounded (~0.6 % of inputs are
    # 1 ulp off IEEE sqrt), so bit-equality with the reference binary is a property of its math library,
    # not of the algorithm.  The oracle (and the CUDA path) use IEEE round-to-nearest sqrt / div.
    de, ds = O.accumulate_gradients(g["ref_grad"], g["state"], float(g["lr"]))
    assert np.array_equal(ds, g["ref_delta_s"])
    ulp = np.abs(de.view(np.int32).astype(np.int64) - g["ref_delta_e"].view(np.int32).astype(np.int64))
    assert ulp.max() <= 4 and (ulp != 0).mean() < 0.02


# ---- (3) the reference library itself on fresh inputs ------------------------------------------
@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (make -C oracle)")
@pytest.mark.parametrize("kind", [O.DISTMULT, O.COMPLEX])
def test_against_reference_library(kind):
    rng = np.random.default_rng(123 + kind)
    B, C, N, d = 50, 4, 40, 20
    uniq, edges, dn, sn = O.make_batch(rng, 500, 6, B, C, N)
    U = len(uniq)
    emb = rng.uniform(-0.5, 0.5, (U, d)).astype(np.float32)
    state = rng.uniform(0, 0.2, (U, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (6, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (6, d)).astype(np.float32)
    ref = R.train_batch(kind, emb, state, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM)
    res = O.train_batch(kind, emb, state, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM)
    for a, b in [(res.scores.neg, ref["neg"]), (res.scores.inv_neg, ref["inv_neg"]), (res.grad, ref["grad"]), (res.delta_s, ref["delta_s"]),
                 (res.rel_grad, ref["rel_grad"])]:
        assert rel_err(a, b) < 2e-5
    # gather / scatter-add are bit exact
    table = rng.standard_normal((300, 12)).astype(np.float32)
    idx = rng.integers(0, 300, 64, dtype=np.int64)
    assert np.array_equal(O.index_read(table, idx), R.index_read(table, idx))


# ---- (4) evaluation path: score filter, ranks, ranking metrics -----------------
EVAL_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "eval_*.npz")))


def ranks_match_up_to_ties(ranks, ref_ranks, pos, neg, tol=1e-5):
    """Ranks are integers: two implementations whose scores differ in the last bits may only disagree on a row by the number of
    negatives that tie with the positive within `tol` (relative to the largest score)."""
    scale = max(float(np.abs(neg[neg > -1e8]).max()), 1e-12)
    near = (np.abs(neg - pos[:, None]) <= tol * scale).sum(axis=1)
    return bool(np.all(np.abs(ranks - ref_ranks) <= near))


def test_eval_known_answers():
    # reporting.cpp:56-58 on a hand-checked case: ties count against the positive (>=), rank is 1-based
    pos = np.array([1.0, 0.5, -2.0], np.float32)
    neg = np.array([[0.9, 1.0, 1.1, -5.0], [0.4, 0.3, 0.2, 0.1], [0.0, 0.0, 0.0, 0.0]], np.float32)
    assert O.compute_ranks(pos, neg).tolist() == [3, 1, 5]
    # negative.cpp:306-311: filtered entries become -1e9 and so never outrank the positive
    f = np.array([[0, 2], [0, 1], [2, 0]], np.int64)
    assert O.compute_ranks(pos, O.apply_score_filter(neg.copy(), f)).tolist() == [1, 1, 4]
    m = O.ranking_metrics(np.array([1, 2, 4, 10, 11], np.int64))
    assert m["mean_rank"] == pytest.approx(5.6) and m["hits@1"] == 0.2 and m["hits@3"] == 0.4 and m["hits@10"] == 0.8
    assert m["mrr"] == pytest.approx((1 + 0.5 + 0.25 + 0.1 + 1 / 11) / 5, rel=1e-6)


@pytest.mark.parametrize("name", EVAL_CASES)
def test_golden_evaluate_batch(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    ranks, inv_ranks, sc = O.evaluate_batch(int(g["kind"]), g["emb"], g["edges"], g["rel"], g["inv_rel"], g["dst_negs"], g["src_negs"], g["dst_filter"],
                                            g["src_filter"])
    assert rel_err(sc.neg, g["ref_neg"]) < 2e-5 and rel_err(sc.inv_neg, g["ref_inv_neg"]) < 2e-5
    # the rank rule itself is exact on the reference's own scores
    assert np.array_equal(O.compute_ranks(g["ref_pos"], g["ref_neg"]), g["ref_ranks"])
    assert np.array_equal(O.compute_ranks(g["ref_inv_pos"], g["ref_inv_neg"]), g["ref_inv_ranks"])
    assert ranks_match_up_to_ties(ranks, g["ref_ranks"], g["ref_pos"], g["ref_neg"])
    assert ranks_match_up_to_ties(inv_ranks, g["ref_inv_ranks"], g["ref_inv_pos"], g["ref_inv_neg"])
    m = O.ranking_metrics(np.concatenate([g["ref_ranks"], g["ref_inv_ranks"]]))
    ref_m = g["metrics"]
    assert m["mean_rank"] == pytest.approx(ref_m[0], rel=1e-12) and m["mrr"] == pytest.approx(ref_m[1], rel=1e-6)
    assert (m["hits@1"], m["hits@3"], m["hits@10"]) == (ref_m[2], ref_m[3], ref_m[4])


# ---- (5) batch assembly: negative sampling + edgeSample ---------------------------
def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10 (first two output words); counter = (index lo, index hi, stream, batch), key = seed
    kat = [((0, 0, 0, 0), 0xE169C58D6627E8D5), ((0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), 0x41C83B0E408F276D),
           ((0x299F31D0A4093822, 0x85A308D3243F6A88, 0x13198A2E, 0x03707344), 0x94FDCCEBD16CFE09)]
    for (seed, index, stream, batch), want in kat:
        assert int(O.philox4x32_10_u64(seed, np.array([index], dtype=np.uint64), stream, batch)[0]) == want


def test_sample_negatives_contract():
    # shapes / ranges of negative.cpp:328-366 and test/python/bindings/integration/test_data.py:108-162
    neg = O.sample_negatives(1000, 4, 250, seed=3, batch_index=0, inverse=False)
    assert neg.shape == (4, 250) and neg.dtype == np.int64 and neg.min() >= 0 and neg.max() < 1000
    assert not np.array_equal(neg, O.sample_negatives(1000, 4, 250, seed=3, batch_index=1, inverse=False))      # a new batch draws new ids
    assert not np.array_equal(neg, O.sample_negatives(1000, 4, 250, seed=3, batch_index=0, inverse=True))       # the two sides are independent
    assert np.array_equal(neg, O.sample_negatives(1000, 4, 250, seed=3, batch_index=0, inverse=False))          # stateless
    big = O.sample_negatives(50, 1, 200000, seed=9, batch_index=0, inverse=False).reshape(-1)
    counts = np.bincount(big, minlength=50)
    chi2 = float(((counts - 4000.0) ** 2 / 4000.0).sum())
    assert chi2 < 100  # 49 degrees of freedom: p(chi2 > 100) ~ 2e-5
    # degree-based fraction: the first (int)(N * f) ids of every chunk are endpoints of batch edges (negative.cpp:7-19,334,347)
    rng = np.random.default_rng(0)
    edges = np.stack([rng.integers(5000, 5010, 64), rng.integers(0, 3, 64), rng.integers(7000, 7010, 64)], axis=1).astype(np.int64)
    d = O.sample_negatives(1000, 3, 100, seed=1, batch_index=2, inverse=False, degree_fraction=0.5, edges=edges)
    assert np.all((d[:, :50] >= 7000) & (d[:, :50] < 7010)) and np.all(d[:, 50:] < 1000)
    s = O.sample_negatives(1000, 3, 100, seed=1, batch_index=2, inverse=True, degree_fraction=0.25, edges=edges)
    assert np.all((s[:, :25] >= 5000) & (s[:, :25] < 5010)) and np.all(s[:, 25:] < 1000)


def test_edge_sample_matches_map_tensors_golden(golden_dir):
    # edgeSample = map_tensors over cat(src, dst, src_negs, dst_negs) (dataloader.cpp:398-461); map_tensors is pinned by the golden
    rng = np.random.default_rng(4)
    edges = np.stack([rng.integers(0, 300, 40), rng.integers(0, 5, 40), rng.integers(0, 300, 40)], axis=1).astype(np.int64)
    sn, dn = rng.integers(0, 300, (2, 30)).astype(np.int64), rng.integers(0, 300, (2, 30)).astype(np.int64)
    uniq, local, s_loc, d_loc = O.edge_sample(edges, sn, dn)
    assert np.array_equal(uniq, np.unique(np.concatenate([edges[:, 0], edges[:, 2], sn.reshape(-1), dn.reshape(-1)])))
    assert np.array_equal(uniq[local[:, 0]], edges[:, 0]) and np.array_equal(uniq[local[:, 2]], edges[:, 2]) and np.array_equal(local[:, 1], edges[:, 1])
    assert np.array_equal(uniq[s_loc], sn) and np.array_equal(uniq[d_loc], dn)
    if R.available():
        all_ids = np.concatenate([edges[:, 0], edges[:, 2], sn.reshape(-1), dn.reshape(-1)])
        ru, rm = R.map_tensors(all_ids)
        assert np.array_equal(ru, uniq) and np.array_equal(rm[:40], local[:, 0]) and np.array_equal(rm[80:140].reshape(2, 30), s_loc)
    u2, l2, s2, d2 = O.edge_sample(edges[:, [0, 2]], None, dn)  # 2-column edges, no inverse side
    assert s2 is None and l2.shape == (40, 2) and np.array_equal(u2[d2], dn)


# ---- (6) score-filter construction (negative.cpp:62-195) -----------------------------------------------
def test_filter_construction_known_answers():
    # graph: 0 -r0-> 1, 0 -r0-> 2, 0 -r1-> 3, 4 -r0-> 1 ; batch edge (0, r0, 1)
    graph = np.array([[0, 0, 1], [0, 0, 2], [0, 1, 3], [4, 0, 1]], np.int64)
    edges = np.array([[0, 0, 1]], np.int64)
    all_nodes = np.arange(5, dtype=np.int64).reshape(1, 5)
    # corrupting the destination: every true (0, r0, *) destination is masked, including the positive's own
    assert O.compute_filter_corruption(edges, all_nodes, False, graph_edges=graph).tolist() == [[0, 1], [0, 2]]
    # corrupting the source: every true (*, r0, 1) source
    assert O.compute_filter_corruption(edges, all_nodes, True, graph_edges=graph).tolist() == [[0, 0], [0, 4]]
    # local filter: negatives [2, 3, 1] of the only chunk against the batch {(0,r0,1), (0,r0,2)} -> columns of nodes 2 and 1
    batch = np.array([[0, 0, 1], [0, 0, 2]], np.int64)
    assert O.compute_filter_corruption(batch, np.array([[2, 3, 1]], np.int64), False).tolist() == [[0, 0], [0, 2], [1, 0], [1, 2]]


def test_golden_filter_construction_and_filtered_eval(golden_dir):
    g = np.load(os.path.join(golden_dir, "eval_filtered_graph.npz"))
    all_nodes = g["dst_negs"]
    # integer work: bit-exact, including the order of the pairs
    assert np.array_equal(O.compute_filter_corruption(g["edges"], all_nodes, False, graph_edges=g["graph"]), g["dst_filter"])
    assert np.array_equal(O.compute_filter_corruption(g["edges"], all_nodes, True, graph_edges=g["graph"]), g["src_filter"])
    assert np.array_equal(O.compute_filter_corruption(g["edges"], g["local_negs"], False), g["local_dst_filter"])
    assert np.array_equal(O.compute_filter_corruption(g["edges"], g["local_negs"], True), g["local_src_filter"])
    # the positive's own endpoint is always among the filtered columns of its row (it is a true edge of the graph)
    df = set(map(tuple, g["dst_filter"].tolist()))
    assert all((i, int(e[2])) in df for i, e in enumerate(g["edges"]))
    # filtered evaluation end to end with the oracle-built filters
    ranks, inv_ranks, sc = O.evaluate_batch(int(g["kind"]), g["emb"], g["edges"], g["rel"], g["inv_rel"], all_nodes, g["src_negs"],
                                            O.compute_filter_corruption(g["edges"], all_nodes, False, graph_edges=g["graph"]),
                                            O.compute_filter_corruption(g["edges"], all_nodes, True, graph_edges=g["graph"]))
    assert ranks_match_up_to_ties(ranks, g["ref_ranks"], g["ref_pos"], g["ref_neg"])
    assert ranks_match_up_to_ties(inv_ranks, g["ref_inv_ranks"], g["ref_inv_pos"], g["ref_inv_neg"])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (make -C oracle)")
def test_filter_construction_against_reference_library():
    rng = np.random.default_rng(8)
    num_nodes = 40
    graph = np.unique(np.stack([rng.integers(0, num_nodes, 400), rng.integers(0, 3, 400), rng.integers(0, num_nodes, 400)], axis=1).astype(np.int64), axis=0)
    graph = graph[rng.permutation(len(graph))]
    edges = graph[rng.choice(len(graph), 30, replace=False)]
    all_nodes = np.arange(num_nodes, dtype=np.int64).reshape(1, -1)
    negs = rng.integers(0, num_nodes, (3, 25)).astype(np.int64)
    for inverse in (False, True):
        assert np.array_equal(O.compute_filter_corruption(edges, all_nodes, inverse, graph_edges=graph),
                              R.compute_filter(edges, all_nodes, inverse, graph_edges=graph, num_nodes=num_nodes))
        assert np.array_equal(O.compute_filter_corruption(edges[:, [0, 2]], all_nodes, inverse, graph_edges=graph[:, [0, 2]]),
                              R.compute_filter(edges[:, [0, 2]], all_nodes, inverse, graph_edges=graph[:, [0, 2]], num_nodes=num_nodes))
        assert np.array_equal(O.compute_filter_corruption(edges, negs, inverse), R.compute_filter(edges, negs, inverse))


# ---- (7) size-independent properties of the integer / index work (the same properties the GPU tests check at full size) ----------
@pytest.mark.parametrize("seed", range(5))
def test_properties_ranks_filters_assembly(seed):
    rng = np.random.default_rng(1000 + seed)
    rows, N = int(rng.integers(1, 60)), int(rng.integers(1, 90))
    neg = np.round(rng.standard_normal((rows, N)).astype(np.float32), 1)  # heavy ties
    pos = np.round(rng.standard_normal(rows).astype(np.float32), 1)
    ranks = O.compute_ranks(pos, neg)
    assert ranks.min() >= 1 and ranks.max() <= N + 1
    assert np.array_equal(ranks, O.compute_ranks(pos, neg[:, rng.permutation(N)]))            # order of the negatives is irrelevant
    assert np.all(O.compute_ranks(pos + 1.0, neg) <= ranks)                                    # a better positive never ranks worse
    F = int(rng.integers(0, rows * N))
    filt = np.stack([rng.integers(0, rows, F), rng.integers(0, N, F)], axis=1).astype(np.int64)
    once = O.apply_score_filter(neg.copy(), filt)
    assert np.array_equal(once, O.apply_score_filter(once.copy(), filt))                       # idempotent
    assert np.all(O.compute_ranks(pos, once) <= ranks)                                         # filtering never worsens a rank
    full = np.stack([np.repeat(np.arange(rows), N), np.tile(np.arange(N), rows)], axis=1)
    assert np.all(O.compute_ranks(pos, O.apply_score_filter(neg.copy(), full)) == 1)           # everything filtered -> rank 1
    # batch assembly round trip: local ids index the sorted unique list back to the global ids, for both edge widths
    num_nodes, B, C, Nn = int(rng.integers(2, 500)), int(rng.integers(1, 200)), int(rng.integers(1, 4)), int(rng.integers(1, 50))
    edges = np.stack([rng.integers(0, num_nodes, B), rng.integers(0, 5, B), rng.integers(0, num_nodes, B)], axis=1).astype(np.int64)
    sn, dn = O.sample_negatives(num_nodes, C, Nn, seed, 0, True), O.sample_negatives(num_nodes, C, Nn, seed, 0, False)
    uniq, local, s_loc, d_loc = O.edge_sample(edges, sn, dn)
    assert np.all(np.diff(uniq) > 0) and uniq.min() >= 0 and uniq.max() < num_nodes
    assert np.array_equal(uniq[local[:, 0]], edges[:, 0]) and np.array_equal(uniq[local[:, 2]], edges[:, 2]) and np.array_equal(local[:, 1], edges[:, 1])
    assert np.array_equal(uniq[s_loc], sn) and np.array_equal(uniq[d_loc], dn)
    assert len(uniq) == len(set(edges[:, 0]) | set(edges[:, 2]) | set(sn.reshape(-1)) | set(dn.reshape(-1)))
    u2, l2, s2, d2 = O.edge_sample(edges[:, [0, 2]], None, dn)
    assert s2 is None and np.array_equal(u2[l2[:, 0]], edges[:, 0]) and np.array_equal(u2[l2[:, 1]], edges[:, 2]) and np.array_equal(u2[d2], dn)
