"""GPU parity of batch assembly: the device negative sampler and DataLoader::edgeSample through the C ABI.
Indices are bit-exact against the oracle; the sampler's stream is Philox4x32-10 (pinned by the Random123 known-answer vectors in
tests/test_oracle.py), its distribution is checked the way the reference's sampler tests check theirs (shape, range, uniformity)."""
import numpy as np
import pytest
import torch

from oracle import marius_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from marius_b200 import ops as o

    return o


@pytest.fixture(scope="module")
def ctx(ops):
    return ops.Context(0)


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("num_nodes,C,N", [(1000, 4, 250), (14541, 10, 500), (100_000_000, 50, 1000), (3, 1, 1), ((1 << 40) + 12345, 2, 333)])
@pytest.mark.parametrize("inverse", [False, True])
def test_sampler_bit_exact(ops, num_nodes, C, N, inverse):
    for seed, batch in ((0, 0), (0xDEADBEEFCAFEF00D, 7), (12345, 4_000_000_000)):
        got = ops.sample_negatives(num_nodes, C, N, seed, batch, inverse, "cuda").cpu().numpy()
        assert np.array_equal(got, O.sample_negatives(num_nodes, C, N, seed, batch, inverse))
        assert got.shape == (C, N) and got.min() >= 0 and got.max() < num_nodes


def test_sampler_degree_fraction_and_errors(ops):
    from marius_b200 import MariusB200Error

    rng = np.random.default_rng(0)
    edges = np.stack([rng.integers(5000, 5100, 777), rng.integers(0, 3, 777), rng.integers(7000, 7100, 777)], axis=1).astype(np.int64)
    for inverse, f in ((False, 0.5), (True, 0.25), (False, 1.0)):
        got = ops.sample_negatives(1000, 3, 100, 1, 2, inverse, "cuda", degree_fraction=f, edges=dev(edges)).cpu().numpy()
        assert np.array_equal(got, O.sample_negatives(1000, 3, 100, 1, 2, inverse, degree_fraction=f, edges=edges))
    got2 = ops.sample_negatives(1000, 3, 100, 1, 2, False, "cuda", degree_fraction=0.5, edges=dev(edges[:, [0, 2]])).cpu().numpy()
    assert np.array_equal(got2, O.sample_negatives(1000, 3, 100, 1, 2, False, degree_fraction=0.5, edges=edges[:, [0, 2]]))
    with pytest.raises(MariusB200Error):
        ops.sample_negatives(1000, 3, 100, 1, 2, False, "cuda", degree_fraction=0.5)  # needs the batch's edges
    with pytest.raises(MariusB200Error):
        ops.sample_negatives(0, 3, 100, 1, 2, False, "cuda")


def test_sampler_uniformity_full_size(ops):
    # BASELINE shape: 50 chunks x 1000 negatives over 1e8 nodes, 20 batches = 1e6 draws bucketed into 100 bins
    draws = torch.cat([ops.sample_negatives(100_000_000, 50, 1000, 42, b, False, "cuda").reshape(-1) for b in range(20)])
    counts = torch.bincount(draws // 1_000_000, minlength=100).double()
    chi2 = float(((counts - 10000.0) ** 2 / 10000.0).sum())
    assert chi2 < 180  # 99 degrees of freedom: p(chi2 > 180) ~ 1e-6
    assert draws.unique().numel() > 0.99 * draws.numel()  # 1e6 draws from 1e8 ids: ~0.5 % collide


@pytest.mark.parametrize("B,C,N,num_nodes,cols,inverse", [(40, 2, 30, 300, 3, True), (1000, 4, 250, 5000, 3, True), (333, 3, 111, 50, 2, False),
                                                           (50000, 50, 1000, 100_000_000, 3, True), (1, 1, 1, 1, 3, True)])
def test_edge_sample_bit_exact(ops, ctx, B, C, N, num_nodes, cols, inverse):
    rng = np.random.default_rng(B + N)
    edges = np.stack([rng.integers(0, num_nodes, B), rng.integers(0, 7, B), rng.integers(0, num_nodes, B)], axis=1).astype(np.int64)
    if cols == 2:
        edges = np.ascontiguousarray(edges[:, [0, 2]])
    dn = O.sample_negatives(num_nodes, C, N, 5, 0, False)
    sn = O.sample_negatives(num_nodes, C, N, 5, 0, True) if inverse else None
    uniq, num, e_loc, s_loc, d_loc = ops.edge_sample(ctx, dev(edges), dev(sn), dev(dn), num_nodes - 1)
    ru, rl, rs, rd = O.edge_sample(edges, sn, dn)
    U = int(num.item())
    assert U == len(ru) and np.array_equal(uniq[:U].cpu().numpy(), ru) and bool((uniq[U:] == -1).all())
    assert np.array_equal(e_loc.cpu().numpy(), rl) and np.array_equal(d_loc.cpu().numpy(), rd)
    if inverse:
        assert np.array_equal(s_loc.cpu().numpy(), rs)
    else:
        assert s_loc is None


def test_sample_map_train_equals_oracle(ops, ctx):
    """raw global edges -> device sampler -> device edgeSample -> fused step with U = capacity (no host round trip for the unique
    count) gives the table the oracle gets from the same edges."""
    rng = np.random.default_rng(11)
    num_nodes, R, B, C, N, d = 3000, 5, 512, 2, 256, 64
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    rel, inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32), rng.uniform(-1, 1, (R, d)).astype(np.float32)
    edges = np.stack([rng.integers(0, num_nodes, B), rng.integers(0, R, B), rng.integers(0, num_nodes, B)], axis=1).astype(np.int64)
    e = dev(edges)
    sn = ops.sample_negatives(num_nodes, C, N, 77, 3, True, "cuda", degree_fraction=0.5, edges=e)
    dn = ops.sample_negatives(num_nodes, C, N, 77, 3, False, "cuda", degree_fraction=0.5, edges=e)
    uniq, num, e_loc, s_loc, d_loc = ops.edge_sample(ctx, e, sn, dn, num_nodes - 1)
    tab, st = dev(table), torch.zeros(num_nodes, d, device="cuda")
    # unique ids padded with -1 up to capacity: the step is launched without reading the unique count back
    loss = ops.train_step(ctx, ops.COMPLEX, tab, st, uniq, e_loc, dev(rel), dev(inv_rel), d_loc, s_loc, 0.1, ops.REDUCTION_SUM, ops.PREC_FP32)
    ru, rl, rs, rd = O.edge_sample(edges, sn.cpu().numpy(), dn.cpu().numpy())
    ref_tab, ref_st = table.copy(), np.zeros_like(table)
    res = O.train_step_on_table(O.COMPLEX, ref_tab, ref_st, ru, rl, rel, inv_rel, rd, rs, 0.1, O.REDUCTION_SUM, acc=np.float64)
    err = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
    assert err(tab.cpu().numpy(), ref_tab) < 1e-4 and err(st.cpu().numpy(), ref_st) < 1e-4
    assert abs(float(loss.item()) - float(res.loss)) < 1e-4 * abs(float(res.loss))


def test_step_from_raw_edges_equals_the_assembled_step(ops):
    """mb_train_step_edges_host_async (raw global edges from the host; negatives + unique mapping on the device) leaves exactly the
    tables / losses of: sampler (oracle, bit-exact with the device sampler) -> map_tensors (oracle) -> mb_train_step."""
    rng = np.random.default_rng(5)
    num_nodes, R, B, C, N, d = 20000, 6, 1024, 2, 512, 64
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    rel, inv_rel = dev(rng.uniform(-1, 1, (R, d)).astype(np.float32)), dev(rng.uniform(-1, 1, (R, d)).astype(np.float32))
    t1, s1 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    t2, s2 = dev(table), torch.zeros(num_nodes, d, device="cuda")
    ctx_a, ctx_b = ops.Context(0), ops.Context(0)
    seed = 99
    losses_a, losses_b = [], []
    for step in range(4):
        edges = np.stack([rng.integers(0, num_nodes, B), rng.integers(0, R, B), rng.integers(0, num_nodes, B)], axis=1).astype(np.int64)
        e_h = torch.from_numpy(edges).pin_memory()
        ticket, keep = ops.train_step_edges_host_async(ctx_a, ops.COMPLEX, t1, s1, e_h, num_nodes, C, N, seed, step, rel, inv_rel, 0.1)
        losses_a.append(ops.train_step_host_wait(ctx_a, ticket))
        dn = O.sample_negatives(num_nodes, C, N, seed, step, False)
        sn = O.sample_negatives(num_nodes, C, N, seed, step, True)
        uniq, inv = O.map_tensors(np.concatenate([edges[:, 0], edges[:, 2], sn.reshape(-1), dn.reshape(-1)]))
        e_loc = np.ascontiguousarray(np.stack([inv[:B], edges[:, 1], inv[B:2 * B]], axis=1))
        sn_loc = np.ascontiguousarray(inv[2 * B:2 * B + C * N].reshape(C, N))
        dn_loc = np.ascontiguousarray(inv[2 * B + C * N:].reshape(C, N))
        loss = ops.train_step(ctx_b, ops.COMPLEX, t2, s2, dev(uniq), dev(e_loc), rel, inv_rel, dev(dn_loc), dev(sn_loc), 0.1)
        losses_b.append(float(loss.item()))
    torch.cuda.synchronize()
    assert losses_a == losses_b
    assert torch.equal(t1, t2) and torch.equal(s1, s2)
