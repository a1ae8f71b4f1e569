"""GPU parity of the storage path (gather / scatter-add / Adagrad) through the C ABI -- bit-exact against the oracle and the
reference-generated golden fixtures.  Mirrors test/cpp/unit/test_buffer.cpp:275-297 and test_storage.cpp:260-292."""
import os

import numpy as np
import pytest
import torch

from oracle import marius_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from marius_b200 import ops as o

    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_golden_index_read_add(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "storage.npz"))
    table = dev(g["table"])
    out = ops.gather_rows(table, dev(g["idx"]))
    assert np.array_equal(out.cpu().numpy(), g["read"])  # == reference InMemory::indexRead
    ops.scatter_add_rows(table, dev(g["uidx"]), dev(g["vals"]))
    assert np.array_equal(table.cpu().numpy(), g["after"])  # == reference InMemory::indexAdd


@pytest.mark.parametrize("d", [1, 2, 3, 10, 100, 400, 404, 1000])
@pytest.mark.parametrize("n", [0, 1, 33, 4000])
def test_gather_scatter_bit_exact(ops, d, n):
    rng = np.random.default_rng(d * 1000 + n)
    rows = 5000
    table = rng.standard_normal((rows, d)).astype(np.float32)
    idx = rng.integers(0, rows, n, dtype=np.int64)
    t = dev(table)
    got = ops.gather_rows(t, dev(idx))
    assert got.shape == (n, d)
    assert np.array_equal(got.cpu().numpy(), O.index_read(table, idx))
    uidx = rng.permutation(rows)[:n].astype(np.int64)
    vals = rng.standard_normal((n, d)).astype(np.float32)
    ops.scatter_add_rows(t, dev(uidx), dev(vals))
    ref = table.copy()
    O.index_add(ref, uidx, vals)
    assert np.array_equal(t.cpu().numpy(), ref)
    ops.scatter_put_rows(t, dev(uidx), dev(vals))
    ref[uidx] = vals
    assert np.array_equal(t.cpu().numpy(), ref)


def test_gather_strided_table_and_views(ops):
    # a column-slice view: ld > d (EmbeddingLayer::forward narrows, embedding.cpp:17)
    rng = np.random.default_rng(5)
    full = dev(rng.standard_normal((300, 48)).astype(np.float32))
    view = full[:, 8:40]
    idx = dev(rng.integers(0, 300, 77, dtype=np.int64))
    assert torch.equal(ops.gather_rows(view, idx), view.index_select(0, idx))


def test_errors(ops):
    from marius_b200 import MariusB200Error

    t = torch.zeros(45, 16, device="cuda")
    with pytest.raises(MariusB200Error):  # test_buffer.cpp:282
        ops.gather_rows(t, torch.zeros((10, 10), dtype=torch.int64, device="cuda"))
    ids = torch.arange(5, device="cuda")
    with pytest.raises(MariusB200Error):  # test_buffer.cpp:294-296
        ops.scatter_add_rows(t, ids, torch.zeros(6, 16, device="cuda"))
    with pytest.raises(MariusB200Error):
        ops.scatter_add_rows(t, ids, torch.zeros(5, 17, device="cuda"))
    with pytest.raises(MariusB200Error):
        ops.scatter_add_rows(t, torch.zeros((10, 10), dtype=torch.int64, device="cuda"), torch.zeros(5, 16, device="cuda"))


def test_adagrad_known_answer(ops):
    # test/python/bindings/integration/test_data.py:34-47
    g = torch.tensor([0.5, -1.0], device="cuda")
    de, ds = ops.adagrad_deltas(g, torch.zeros(2, device="cuda"), 1.0)
    assert torch.equal(ds, g.pow(2))
    expected = -1.0 * (g / (ds.sqrt().add_(1e-10)))
    assert torch.equal(de, expected)


@pytest.mark.parametrize("d", [2, 100, 400])
def test_adagrad_bit_exact_vs_oracle(ops, d):
    rng = np.random.default_rng(d)
    n = 3000
    g = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    s = rng.uniform(0, 0.1, (n, d)).astype(np.float32)
    s[::7] = 0
    de, ds = ops.adagrad_deltas(dev(g), dev(s), 0.1)
    rde, rds = O.accumulate_gradients(g, s, 0.1)
    assert np.array_equal(ds.cpu().numpy(), rds)
    assert np.array_equal(de.cpu().numpy(), rde)  # IEEE sqrt/div, no fma contraction: bit-exact with the oracle
    # fused update == accumulateGradients + indexAdd x2
    rows = 10000
    table = rng.standard_normal((rows, d)).astype(np.float32)
    state = rng.uniform(0, 0.1, (rows, d)).astype(np.float32)
    idx = rng.permutation(rows)[:n].astype(np.int64)
    t, st = dev(table), dev(state)
    ops.adagrad_update_rows(t, st, dev(idx), dev(g), 0.1)
    de2, ds2 = O.accumulate_gradients(g, state[idx], 0.1)
    O.index_add(table, idx, de2)
    O.index_add(state, idx, ds2)
    assert np.array_equal(t.cpu().numpy(), table)
    assert np.array_equal(st.cpu().numpy(), state)


def test_golden_adagrad(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "train_complex_d100.npz"))
    de, ds = ops.adagrad_deltas(dev(g["ref_grad"]), dev(g["state"]), float(g["lr"]))
    assert np.array_equal(ds.cpu().numpy(), g["ref_delta_s"])
    ulp = np.abs(de.cpu().numpy().view(np.int32).astype(np.int64) - g["ref_delta_e"].view(np.int32).astype(np.int64))
    assert ulp.max() <= 4  # libtorch's AVX512 sqrt is not correctly rounded (tests/test_oracle.py)


def test_global_to_local_map(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "partition_buffer.npz"))
    m = ops.global_to_local_map(int(g["total"]), int(g["psize"]), [0, 1], [0, 1], "cuda")
    assert np.array_equal(m.cpu().numpy(), g["map_current"])  # test_buffer.cpp:310-318
    m2 = ops.global_to_local_map(int(g["total"]), int(g["psize"]), [0, 2], [0, 1], "cuda")
    assert np.array_equal(m2.cpu().numpy(), g["map_next"])


def test_map_tensors(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "storage.npz"))
    ctx = ops.Context(0)
    u, m = ops.map_tensors(ctx, dev(g["all_ids"]))
    assert np.array_equal(u.cpu().numpy(), g["uniq"]) and np.array_equal(m.cpu().numpy(), g["mapped"])
    rng = np.random.default_rng(3)
    for n, hi in [(1, 10), (5000, 300), (100000, 10**9), (70000, 2**40)]:
        ids = rng.integers(0, hi, n, dtype=np.int64)
        u, m = ops.map_tensors(ctx, dev(ids))
        ru, rm = O.map_tensors(ids)
        assert np.array_equal(u.cpu().numpy(), ru) and np.array_equal(m.cpu().numpy(), rm)


def test_large_roundtrip_properties(ops):
    """Full-size property checks (d=400): gather(scatter_put(x)) == x ; scatter_add is linear ; update idempotence of zero grad."""
    torch.manual_seed(0)
    rows, d, n = 200000, 400, 40000
    table = torch.randn(rows, d, device="cuda")
    idx = torch.randperm(rows, device="cuda")[:n]
    vals = torch.randn(n, d, device="cuda")
    before = ops.gather_rows(table, idx)
    ops.scatter_add_rows(table, idx, vals)
    ops.scatter_add_rows(table, idx, -vals)
    assert torch.allclose(ops.gather_rows(table, idx), before, atol=1e-6)
    ops.scatter_put_rows(table, idx, vals)
    assert torch.equal(ops.gather_rows(table, idx), vals)
    st = torch.rand(rows, d, device="cuda")
    t0, s0 = table.clone(), st.clone()
    ops.adagrad_update_rows(table, st, idx, torch.zeros(n, d, device="cuda"), 0.1)
    assert torch.equal(table, t0) and torch.equal(st, s0)
