"""GPU parity of the decoder / training step through the C ABI against the oracle and the reference-generated golden fixtures.
Tolerance: 1e-4 relative (north_star) with rel(a,b) = max|a-b| / max(max|b|, tau), tau = 1e-6; indices / gathers bit-exact."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import marius_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    from marius_b200 import ops as o

    return o


@pytest.fixture(scope="module")
def ctx(ops):
    return ops.Context(0)


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_err(a, b, tau=1e-6):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), tau))


def test_distmult_known_answer(ops, ctx):
    # test/python/bindings/integration/test_nn.py:15-25,148-160 : exact [12.5, -3.75, -0.25]
    emb = torch.tensor([[1.5, 2.5], [2.5, 3.5], [4.25, 1.0], [-1.0, 0.5]], device="cuda")
    edges = torch.tensor([[0, 0, 1], [2, 0, 3], [3, 1, 0]], device="cuda")
    rel = torch.ones(2, 2, device="cuda")
    negs = torch.tensor([[2, 0], [0, 1], [1, 0]], device="cuda")
    pos, neg, inv_pos, inv_neg = ops.decoder_forward(ctx, ops.DISTMULT, emb, edges, rel, None, negs, None)
    assert torch.equal(pos, torch.tensor([12.5, -3.75, -0.25], device="cuda"))
    assert inv_pos is None and inv_neg is None
    ref = O.node_corrupt_forward(O.DISTMULT, emb.cpu().numpy(), edges.cpu().numpy(), rel.cpu().numpy(), None, negs.cpu().numpy(), None)
    assert np.array_equal(neg.cpu().numpy(), ref.neg)


TRAIN_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "train_*.npz")))


@pytest.mark.parametrize("name", TRAIN_CASES)
@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_golden_train_batch(ops, ctx, golden_dir, name, prec):
    """Against outputs of the unmodified reference C++ (Model::forward_lp + Model::train_batch)."""
    g = np.load(os.path.join(golden_dir, name))
    p = ops.PREC_FP32 if prec == "fp32" else ops.PREC_BF16X3
    kind = int(g["kind"])
    emb, state, edges, rel, inv_rel, dn, sn = (dev(g[k]) for k in ("emb", "state", "edges", "rel", "inv_rel", "dst_negs", "src_negs"))
    pos, neg, ipos, ineg = ops.decoder_forward(ctx, kind, emb, edges, rel, inv_rel, dn, sn, p)
    assert rel_err(pos, g["ref_pos"]) < TOL and rel_err(neg, g["ref_neg"]) < TOL
    assert rel_err(ipos, g["ref_inv_pos"]) < TOL and rel_err(ineg, g["ref_inv_neg"]) < TOL
    out = ops.train_batch(ctx, kind, emb, state, edges, rel, inv_rel, dn, sn, float(g["lr"]), int(g["reduction"]), p)
    assert abs(float(out["loss"].item()) - float(g["ref_loss"][0])) <= TOL * abs(float(g["ref_loss"][0]))
    assert rel_err(out["grad"], g["ref_grad"]) < TOL
    assert rel_err(out["delta_e"], g["ref_delta_e"]) < TOL
    assert rel_err(out["delta_s"], g["ref_delta_s"]) < TOL
    assert rel_err(out["rel_grad"], g["ref_rel_grad"]) < TOL
    assert rel_err(out["inv_rel_grad"], g["ref_inv_rel_grad"]) < TOL


def _random_problem(seed, kind, B, C, N, d, num_nodes, R, scale=0.3, with_rel=True):
    rng = np.random.default_rng(seed)
    uniq, edges, dn, sn = O.make_batch(rng, num_nodes, R, B, C, N, with_rel=with_rel)
    U = len(uniq)
    emb = rng.uniform(-scale, scale, (U, d)).astype(np.float32)
    state = rng.uniform(0, 0.05, (U, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    return uniq, edges, dn, sn, emb, state, rel, inv_rel


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("kind,B,C,N,d", [(1, 100, 1, 64, 32), (2, 333, 4, 200, 64), (2, 1000, 1, 1000, 400), (1, 2000, 2, 1000, 400),
                                          (1, 17, 5, 24, 16), (2, 130, 3, 136, 72), (2, 96, 2, 72, 432)])
def test_train_batch_vs_oracle(ops, ctx, prec, kind, B, C, N, d):
    p = ops.PREC_FP32 if prec == "fp32" else ops.PREC_BF16X3
    uniq, edges, dn, sn, emb, state, rel, inv_rel = _random_problem(B + d, kind, B, C, N, d, 50000, 11)
    ref = O.train_batch(kind, emb, state, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM, acc=np.float64)
    out = ops.train_batch(ctx, kind, dev(emb), dev(state), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, ops.REDUCTION_SUM, p)
    pos, neg, ipos, ineg = ops.decoder_forward(ctx, kind, dev(emb), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), p)
    assert rel_err(pos, ref.scores.pos) < TOL and rel_err(neg, ref.scores.neg) < TOL
    assert rel_err(ipos, ref.scores.inv_pos) < TOL and rel_err(ineg, ref.scores.inv_neg) < TOL
    assert abs(float(out["loss"].item()) - float(ref.loss)) <= TOL * abs(float(ref.loss))
    assert rel_err(out["grad"], ref.grad) < TOL
    assert rel_err(out["delta_e"], ref.delta_e) < TOL and rel_err(out["delta_s"], ref.delta_s) < TOL
    assert rel_err(out["rel_grad"], ref.rel_grad) < TOL and rel_err(out["inv_rel_grad"], ref.inv_rel_grad) < TOL


def test_no_inverse_and_dot_decoder(ops, ctx):
    uniq, edges, dn, sn, emb, state, rel, inv_rel = _random_problem(7, 1, 200, 2, 64, 32, 3000, 5)
    # use_inverse_relations = false (decoder_methods.cpp:90)
    ref = O.train_batch(O.DISTMULT, emb, state, edges, rel, None, dn, None, 0.1, O.REDUCTION_SUM, acc=np.float64)
    for p in (ops.PREC_FP32, ops.PREC_BF16X3):
        out = ops.train_batch(ctx, ops.DISTMULT, dev(emb), dev(state), dev(edges), dev(rel), None, dev(dn), None, 0.1, ops.REDUCTION_SUM, p)
        assert out["inv_rel_grad"] is None
        assert rel_err(out["grad"], ref.grad) < TOL and rel_err(out["rel_grad"], ref.rel_grad) < TOL
        assert abs(float(out["loss"].item()) - float(ref.loss)) <= TOL * abs(float(ref.loss))
    # 2-column edges: no relations at all (decoder_methods.cpp:99-101)
    e2 = np.ascontiguousarray(edges[:, [0, 2]])
    ref = O.train_batch(O.DOT, emb, state, e2, None, None, dn, None, 0.1, O.REDUCTION_SUM, acc=np.float64)
    out = ops.train_batch(ctx, ops.DOT, dev(emb), dev(state), dev(e2), None, None, dev(dn), None, 0.1, ops.REDUCTION_SUM, ops.PREC_BF16X3)
    assert out["rel_grad"] is None
    assert rel_err(out["grad"], ref.grad) < TOL and rel_err(out["delta_e"], ref.delta_e) < TOL


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("kind", [1, 2])
def test_fused_train_step_on_table(ops, ctx, prec, kind):
    """gather -> train_batch -> Adagrad scatter on a device-resident table == the reference's per-batch sequence (trainer.cpp:106-138)."""
    p = ops.PREC_FP32 if prec == "fp32" else ops.PREC_BF16X3
    rng = np.random.default_rng(99)
    num_nodes, R, B, C, N, d = 20000, 9, 512, 2, 256, 128
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    state = rng.uniform(0, 0.01, (num_nodes, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    t, st = dev(table), dev(state)
    rg, irg = torch.empty(R, d, device="cuda"), torch.empty(R, d, device="cuda")
    for step in range(3):
        uniq, edges, dn, sn = O.make_batch(rng, num_nodes, R, B, C, N)
        loss = ops.train_step(ctx, kind, t, st, dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, ops.REDUCTION_SUM, p,
                              rel_grad=rg, inv_rel_grad=irg)
        res = O.train_step_on_table(kind, table, state, uniq, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM, acc=np.float64)
        assert abs(float(loss.item()) - float(res.loss)) <= TOL * abs(float(res.loss))
        assert rel_err(rg, res.rel_grad) < TOL and rel_err(irg, res.inv_rel_grad) < TOL
        # untouched rows are bit-identical, touched rows within tolerance
        got_t, got_s = t.cpu().numpy(), st.cpu().numpy()
        mask = np.ones(num_nodes, bool)
        mask[uniq] = False
        assert np.array_equal(got_t[mask], table[mask]) and np.array_equal(got_s[mask], state[mask])
        assert rel_err(got_t[uniq] - table[uniq] + res.delta_e, res.delta_e) < 2 * TOL  # the applied update
        assert rel_err(got_s, state) < TOL
        # continue from the device result so errors do not compound in the comparison
        table, state = got_t.copy(), got_s.copy()


def test_host_buffer_step_matches_device_step(ops, ctx):
    rng = np.random.default_rng(5)
    num_nodes, R, B, C, N, d = 5000, 4, 256, 1, 128, 64
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    uniq, edges, dn, sn = O.make_batch(rng, num_nodes, R, B, C, N)
    t1, s1 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    t2, s2 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    l1 = ops.train_step(ctx, ops.COMPLEX, t1, s1, dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    l2 = ops.train_step_host(ctx, ops.COMPLEX, t2, s2, pin(uniq), pin(edges), dev(rel), dev(inv_rel), pin(dn), pin(sn), 0.1)
    assert float(l1.item()) == l2
    assert torch.equal(t1, t2) and torch.equal(s1, s2)  # deterministic: same kernels, same order


def test_determinism_with_heavy_duplicates(ops, ctx):
    """60 nodes, 128 edges x 4 chunks x 64 negatives: every node id collides many times; two runs are bit-identical."""
    uniq, edges, dn, sn, emb, state, rel, inv_rel = _random_problem(5, 1, 128, 4, 64, 32, 60, 4)
    args = (dev(emb), dev(state), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1)
    a = ops.train_batch(ctx, ops.DISTMULT, *args)
    b = ops.train_batch(ctx, ops.DISTMULT, *args)
    for k in ("grad", "delta_e", "delta_s", "rel_grad", "inv_rel_grad", "loss"):
        assert torch.equal(a[k], b[k])


def test_mean_reduction_and_padding(ops, ctx):
    # B % C != 0: padded rows add log(1+N) each to the loss and divide the MEAN by C*ceil(B/C) (hard part 2)
    uniq, edges, dn, sn, emb, state, rel, inv_rel = _random_problem(21, 2, 50, 4, 40, 24, 500, 3)
    for red in (O.REDUCTION_MEAN, O.REDUCTION_SUM):
        ref = O.train_batch(O.COMPLEX, emb, state, edges, rel, inv_rel, dn, sn, 0.1, red, acc=np.float64)
        out = ops.train_batch(ctx, ops.COMPLEX, dev(emb), dev(state), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, red, ops.PREC_BF16X3)
        assert abs(float(out["loss"].item()) - float(ref.loss)) <= TOL * abs(float(ref.loss))
        assert rel_err(out["grad"], ref.grad) < TOL


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("block_n", [1024])
def test_contraction_kernel_layouts(ops, ctx, a_mn, b_mn, block_n):
    """tcgen05 GEMM, every operand-major combination the path uses, ragged M/N/K, vs fp64 matmul."""
    torch.manual_seed(1)
    for (bt, M, N, K) in [(1, 128, 256, 64), (3, 1000, 1000, 400), (2, 1000, 400, 1000), (2, 200, 72, 136), (1, 8, 16, 8)]:
        A = torch.randn(bt, M, K, device="cuda")
        Bm = torch.randn(bt, N, K, device="cuda")
        ref = torch.matmul(A.double(), Bm.double().transpose(1, 2))
        Ain = A.transpose(1, 2).contiguous() if a_mn else A
        Bin = Bm.transpose(1, 2).contiguous() if b_mn else Bm
        D = ops.debug_gemm(ctx, Ain, a_mn, Bin, b_mn, ops.PREC_BF16X3, block_n)
        err = float((D.double() - ref).abs().max() / ref.abs().max())
        assert err < 3e-5, (a_mn, b_mn, block_n, bt, M, N, K, err)
        D1 = ops.debug_gemm(ctx, Ain, a_mn, Bin, b_mn, ops.PREC_BF16, block_n)
        err1 = float((D1.double() - ref).abs().max() / ref.abs().max())
        assert err1 < 2e-2
        D0 = ops.debug_gemm(ctx, Ain, a_mn, Bin, b_mn, ops.PREC_FP32, block_n)
        assert float((D0.double() - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("which", [2, 3])
@pytest.mark.parametrize("a_mn", [False, True])
def test_backward_contraction_kernels_vs_fp64(ops, ctx, which, a_mn):
    """The backward contraction kernels (A operand converted inside the kernel from an fp32 matrix; which = 2: through tensor memory,
    one full-width tile per 256-row block; 3: through shared memory), identity conversion, ragged shapes, vs an fp64 matmul."""
    torch.manual_seed(3)
    for (bt, M, N, K) in [(1, 256, 64, 32), (1, 256, 416, 64), (3, 232, 104, 40), (2, 1000, 400, 1000), (2, 300, 224, 72), (2, 520, 8, 200), (3, 77, 232, 1000),
                          (2, 200, 424, 136)]:  # N = 424 > 416 accumulator columns: falls back to the shared-memory-A kernel
        if a_mn and M % 8:
            continue  # (TMA: the inner extent of the fp32 matrix must be a multiple of 8)
        A = torch.randn(bt, K, M, device="cuda") if a_mn else torch.randn(bt, M, K, device="cuda")
        Bm = torch.randn(bt, K, N, device="cuda")
        ref = (A.double().transpose(1, 2) if a_mn else A.double()) @ Bm.double()
        D = ops.debug_gemm(ctx, A, a_mn, Bm, True, ops.PREC_BF16X3, which)
        err = float((D.double() - ref).abs().max() / ref.abs().max())
        assert err < 3e-5, (which, a_mn, bt, M, N, K, err)


def test_single_bf16_product_mode(ops, ctx):
    """MB_PREC_BF16 (one bf16 product, never the default: it misses the 1e-4 bar) through the backward kernels and the whole batch:
    the error is what bf16 operands give (~1e-3 .. 1e-2), not garbage."""
    torch.manual_seed(5)
    for which in (2, 3):
        for a_mn in (False, True):
            A = torch.randn(2, 1000, 1000, device="cuda")
            Bm = torch.randn(2, 1000, 400, device="cuda")
            ref = (A.double().transpose(1, 2) if a_mn else A.double()) @ Bm.double()
            D = ops.debug_gemm(ctx, A, a_mn, Bm, True, ops.PREC_BF16, which)
            err = float((D.double() - ref).abs().max() / ref.abs().max())
            assert 1e-5 < err < 2e-2, (which, a_mn, err)
    uniq, edges, dn, sn, emb, state, rel, inv_rel = _random_problem(77, 2, 333, 4, 200, 64, 50000, 11)
    ref = O.train_batch(2, emb, state, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM, acc=np.float64)
    out = ops.train_batch(ctx, 2, dev(emb), dev(state), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, ops.REDUCTION_SUM, ops.PREC_BF16)
    assert abs(float(out["loss"].item()) - float(ref.loss)) <= 2e-2 * abs(float(ref.loss))
    assert rel_err(out["grad"], ref.grad) < 5e-2


@pytest.mark.parametrize("which", [4, 5])
def test_backward_contractions_back_to_back(ops, ctx, which, monkeypatch):
    """Both backward problems in one grouped launch, several tiles per CTA pair, launched back to back without a synchronisation in
    between (how they run inside the step's CUDA graph): guards the raw-tile ring against waiting on a barrier two phases ahead, which
    showed up only under this timing as corrupted tiles / a trapped arrive.  which = 4: G = S, 5: G = exp(S)."""
    torch.manual_seed(4)
    bt, M, N = 60, 1000, 400
    S = torch.randn(bt, M, M, device="cuda") * 0.5
    Bm = torch.randn(2 * bt, M, N, device="cuda")
    f = S.double() if which == 4 else torch.exp(S.double())
    ref = torch.cat([f @ Bm[:bt].double(), f.transpose(1, 2) @ Bm[bt:].double()], 0)
    for ts in ("1", "0"):
        monkeypatch.setenv("MB_CONV_TS", ts)
        for _ in range(10):
            D = ops.debug_gemm(ctx, S, False, Bm, True, ops.PREC_BF16X3, which)
        torch.cuda.synchronize()
        err = float((D.double() - ref).abs().max() / ref.abs().max())
        assert err < 3e-5, (which, ts, err)


def test_full_shape_property_checks(ops, ctx):
    """BASELINE configs[1] shape (ComplEx d=400, 1000 negatives, B=10000 / 10 chunks): tensor-core path vs the fp32 SIMT path
    of the same library, plus size-independent properties: the loss gradient rows sum to 0 over (pos, negs) and
    zero relation => zero scores."""
    uniq, edges, dn, sn, emb, state, rel, inv_rel = _random_problem(1234, 2, 10000, 10, 1000, 400, 10**8, 1000, scale=0.1)
    a = ops.train_batch(ctx, ops.COMPLEX, dev(emb), dev(state), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, ops.REDUCTION_SUM,
                        ops.PREC_BF16X3)
    b = ops.train_batch(ctx, ops.COMPLEX, dev(emb), dev(state), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, ops.REDUCTION_SUM,
                        ops.PREC_FP32)
    for k in ("grad", "delta_e", "delta_s", "rel_grad", "inv_rel_grad"):
        assert rel_err(a[k], b[k].cpu().numpy()) < TOL, k
    assert abs(float(a["loss"].item()) - float(b["loss"].item())) <= TOL * abs(float(b["loss"].item()))
    zero = torch.zeros_like(dev(rel))
    pos, neg, ipos, ineg = ops.decoder_forward(ctx, ops.COMPLEX, dev(emb), dev(edges), zero, zero, dev(dn), dev(sn))
    assert float(pos.abs().max()) == 0.0 and float(neg.abs().max()) == 0.0 and float(ineg.abs().max()) == 0.0


def test_sharded_table_single_rank_matches_fused_step(ops, ctx):
    """marius_b200.dist.ShardedTable with world_size 1 (all three exchanges degenerate to copies) == the fused mb_train_step:
    exercises OpsBackend (gather, train_batch, reduce_rows_by_key, adagrad_update_rows) on the GPU."""
    from marius_b200.dist import OpsBackend, ShardedTable

    rng = np.random.default_rng(17)
    num_nodes, R, B, C, N, d = 8000, 5, 256, 2, 128, 64
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    uniq, edges, dn, sn = O.make_batch(rng, num_nodes, R, B, C, N)
    t1, s1 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    t2, s2 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    rg = torch.empty(R, d, device="cuda")
    l1 = ops.train_step(ctx, ops.COMPLEX, t1, s1, dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, rel_grad=rg)
    st = ShardedTable(num_nodes, OpsBackend(t2, s2, ctx))
    out = st.train_step(ops.COMPLEX, dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, ops.REDUCTION_SUM)
    assert float(l1.item()) == float(out["loss"].item())
    assert torch.equal(t1, t2) and torch.equal(s1, s2)
    assert torch.equal(rg, out["rel_grad"])
    # reduce_rows_by_key (owner-side merge primitive) with real duplicates
    ids = torch.tensor([5, 3, 5, 9, 3, 5], device="cuda")
    rows = torch.arange(6 * 8, device="cuda", dtype=torch.float32).reshape(6, 8)
    u, s = ops.reduce_rows_by_key(ctx, ids, rows)
    assert u.tolist() == [3, 5, 9]
    assert torch.equal(s, torch.stack([rows[1] + rows[4], rows[0] + rows[2] + rows[5], rows[3]]))


@pytest.mark.parametrize("host", [False, True])
def test_graph_replay_matches_eager(ops, host):
    """The CUDA-graph replay of the fused step (varying index tensors and unique counts, fixed output pointers) is bit-identical to
    plain stream launches."""
    rng = np.random.default_rng(31)
    num_nodes, R, B, C, N, d = 30000, 6, 512, 2, 256, 64
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    batches = [O.make_batch(rng, num_nodes, R, B, C, N) for _ in range(5)]
    assert len({len(b[0]) for b in batches}) > 1  # the unique-row count really varies
    results = []
    for graph_on in (False, True):
        c = ops.Context(0)
        c.graph(graph_on)
        t, st = dev(table), torch.zeros(num_nodes, d, device="cuda")
        rl, irl = dev(rel), dev(inv_rel)
        loss = torch.zeros(1, device="cuda")
        rg, irg = torch.zeros(R, d, device="cuda"), torch.zeros(R, d, device="cuda")
        losses = []
        for (u, e, dn, sn) in batches:
            if host:
                pin = lambda a: torch.from_numpy(a).pin_memory()
                losses.append(ops.train_step_host(c, ops.COMPLEX, t, st, pin(u), pin(e), rl, irl, pin(dn), pin(sn), 0.1, rel_grad=rg, inv_rel_grad=irg))
            else:
                ops.train_step(c, ops.COMPLEX, t, st, dev(u), dev(e), rl, irl, dev(dn), dev(sn), 0.1, loss=loss, rel_grad=rg, inv_rel_grad=irg)
                losses.append(float(loss.item()))
        torch.cuda.synchronize()
        results.append((t.clone(), st.clone(), rg.clone(), losses))
    assert torch.equal(results[0][0], results[1][0]) and torch.equal(results[0][1], results[1][1])
    assert torch.equal(results[0][2], results[1][2]) and results[0][3] == results[1][3]


def test_sharded_entry_point_with_one_shard_equals_fused_step(ops, ctx):
    """mb_train_step_sharded with world == 1 is the plain fused step (same kernels, pointer table of one)."""
    rng = np.random.default_rng(23)
    num_nodes, R, B, C, N, d = 6000, 4, 200, 2, 96, 48
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    uniq, edges, dn, sn = O.make_batch(rng, num_nodes, R, B, C, N)
    t1, s1 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    t2, s2 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    l1 = ops.train_step(ctx, ops.COMPLEX, t1, s1, dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1)
    sh = ops.make_shards([t2], [s2], num_nodes)
    l2 = ops.train_step_sharded(ctx, ops.COMPLEX, sh, t2.stride(0), d, dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1)
    assert float(l1.item()) == float(l2.item())
    assert torch.equal(t1, t2) and torch.equal(s1, s2)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_step_with_local_shards_vs_oracle(ops, ctx, world):
    """Shard arithmetic of mb_train_step_sharded without any IPC: the `world` shards are separate tensors on ONE GPU; global unique ids
    address rows across them.  Must equal the oracle on the concatenated table (and the unsharded fused step)."""
    rng = np.random.default_rng(40 + world)
    rows, R, B, C, N, d = 3000, 4, 200, 2, 96, 48
    total = rows * world
    full = rng.uniform(-0.3, 0.3, (total, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    uniq, edges, dn, sn = O.make_batch(rng, total, R, B, C, N)
    tables = [dev(full[r * rows:(r + 1) * rows].copy()) for r in range(world)]
    states = [torch.zeros(rows, d, device="cuda") for _ in range(world)]
    sh = ops.make_shards(tables, states, rows)
    rg = torch.empty(R, d, device="cuda")
    for step in range(3):  # step 0 eager, step 1 captures the graph, step 2 replays it
        loss = ops.train_step_sharded(ctx, ops.COMPLEX, sh, tables[0].stride(0), d, dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1,
                                      loss=torch.zeros(1, device="cuda") if step == 0 else loss, rel_grad=rg)
    exp_t, exp_s = full.copy(), np.zeros_like(full)
    for step in range(3):
        res = O.train_step_on_table(O.COMPLEX, exp_t, exp_s, uniq, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM, acc=np.float64)
    got_t = np.concatenate([t.cpu().numpy() for t in tables])
    got_s = np.concatenate([s.cpu().numpy() for s in states])
    assert rel_err(got_t, exp_t) < 3e-4 and rel_err(got_s, exp_s) < 3e-4
    assert abs(float(loss.item()) - float(res.loss)) < 3e-4 * abs(float(res.loss))


def test_async_host_steps_match_sync_host_steps(ops):
    """mb_train_step_host_async / _wait with one step in flight (batch i+1 enqueued before the loss of batch i is read) gives exactly
    the tables and losses of the synchronous host entry point: same kernels, same stream order, two loss slots."""
    rng = np.random.default_rng(21)
    num_nodes, R, B, C, N, d = 6000, 5, 512, 2, 256, 64
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    rel, inv_rel = dev(rng.uniform(-1, 1, (R, d)).astype(np.float32)), dev(rng.uniform(-1, 1, (R, d)).astype(np.float32))
    batches = [tuple(torch.from_numpy(x).pin_memory() for x in O.make_batch(rng, num_nodes, R, B, C, N)) for _ in range(6)]
    ctx_a, ctx_b = ops.Context(0), ops.Context(0)
    t1, s1 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    t2, s2 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    sync_losses = [ops.train_step_host(ctx_a, ops.COMPLEX, t1, s1, u, e, rel, inv_rel, dn, sn, 0.1) for (u, e, dn, sn) in batches]
    async_losses, prev, keep = [], None, []
    for (u, e, dn, sn) in batches:
        cur = ops.train_step_host_async(ctx_b, ops.COMPLEX, t2, s2, u, e, rel, inv_rel, dn, sn, 0.1)
        keep.append(cur[1])
        if prev is not None:
            async_losses.append(ops.train_step_host_wait(ctx_b, prev[0]))
        prev = cur
    async_losses.append(ops.train_step_host_wait(ctx_b, prev[0]))
    assert async_losses == sync_losses
    assert torch.equal(t1, t2) and torch.equal(s1, s2)
    assert {ops.train_step_host_async(ctx_b, ops.COMPLEX, t2, s2, *batches[0][:2], rel, inv_rel, *batches[0][2:], 0.1)[0] for _ in range(1)} <= {0, 1}
    torch.cuda.synchronize()


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_wide_rows_take_the_general_path(ops, ctx, prec):
    """d > 512 is outside the vectorised row kernels (4 chunks of 16 B per lane): the general-d kernels take over, same results."""
    p = ops.PREC_FP32 if prec == "fp32" else ops.PREC_BF16X3
    rng = np.random.default_rng(77)
    num_nodes, R, B, C, N, d = 3000, 4, 96, 2, 64, 520
    table = rng.uniform(-0.2, 0.2, (num_nodes, d)).astype(np.float32)
    rel, inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32), rng.uniform(-1, 1, (R, d)).astype(np.float32)
    uniq, edges, dn, sn = O.make_batch(rng, num_nodes, R, B, C, N)
    t, st = dev(table), torch.zeros(num_nodes, d, device="cuda")
    loss = ops.train_step(ctx, ops.DISTMULT, t, st, dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn), 0.1, ops.REDUCTION_SUM, p)
    ref_t, ref_s = table.copy(), np.zeros_like(table)
    res = O.train_step_on_table(O.DISTMULT, ref_t, ref_s, uniq, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM, acc=np.float64)
    assert rel_err(t, ref_t) < TOL and rel_err(st, ref_s) < TOL
    assert abs(float(loss.item()) - float(res.loss)) <= TOL * abs(float(res.loss))


def test_bench_shape_parity_vs_reference(ops, ctx):
    """The headline configuration itself (B = 50 000, C = 50, N = 1000, ComplEx d = 400, 2*10^6-row table) through the C ABI against
    oracle/_ref (the reference C++): unique ids / gathered rows bit-exact, scores, loss, deltas and updated rows within 1e-4.  Same
    checker bench.py prints as its `parity` block."""
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench
    from oracle import ref_lib as R

    if not R.available():
        pytest.skip("oracle/_ref not built")
    res = bench.parity_single(ops, ctx, torch.device("cuda", 0), ops.PREC_BF16X3, 50000, rows=2_000_000)
    assert res["checked"] and res["unique_ids_bit_exact"] and res["gathered_rows_bit_exact"] and res["untouched_rows_bit_exact"], res
    assert res["fused_step_equals_batch_step_bit_exact"], res
    assert res["max_err"] <= TOL, res
    assert res["ok"], res
