"""The C++/libtorch adapters, exercised the way the reference's own tests exercise the originals:
test/python/bindings/integration/test_nn.py (forward_lp / train_batch), test_data.py (Batch.accumulateGradients),
test/cpp/unit/test_buffer.cpp + test_storage.cpp (PartitionBuffer / InMemory indexRead / indexAdd / swap order / maps)."""
import os

import numpy as np
import pytest
import torch

from oracle import marius_oracle as O

pytestmark = pytest.mark.gpu
CUDA = torch.device("cuda", 0)


@pytest.fixture(scope="module")
def host():
    from marius_b200 import host as h

    return h


# ---- test_nn.py ------------------------------------------------------------------------------------------------
node_embeddings = torch.tensor([[1.5, 2.5], [2.5, 3.5], [4.25, 1.0], [-1.0, 0.5]])
batch_edges = torch.tensor([[0, 0, 1], [2, 0, 3], [3, 1, 0]])


def get_test_model_lp(host, mode="infer", inverse=False):
    decoder = host.nn.decoders.edge.DistMult(num_relations=2, embedding_dim=2, use_inverse_relations=inverse, device=CUDA, mode=mode)
    loss = host.nn.SoftmaxCrossEntropy(reduction="sum")
    return host.nn.Model(decoder, loss, CUDA)


def test_forward_lp(host):
    # test_nn.py:148-160
    model = get_test_model_lp(host)
    batch = host.data.Batch(False)
    batch.node_embeddings = node_embeddings.to(CUDA)
    batch.edges = batch_edges.to(CUDA)
    scores, _, _, _ = model.forward_lp(batch=batch, train=False)
    assert torch.all(torch.eq(scores.cpu(), torch.tensor([12.5, -3.75, -0.25]))).item() is True


def test_train_batch(host):
    # test_nn.py:162-172, with the result checked against the oracle
    model = get_test_model_lp(host, mode="train")
    batch = host.data.Batch(True)
    batch.node_embeddings = node_embeddings.to(CUDA)
    batch.node_embeddings_state = torch.zeros_like(node_embeddings).to(CUDA)
    batch.edges = batch_edges.to(CUDA)
    negs = torch.tensor([[2, 0], [0, 1], [1, 0]])
    batch.dst_neg_indices_mapping = negs.to(CUDA)
    model.train_batch(batch, True)
    ref = O.train_batch(O.DISTMULT, node_embeddings.numpy(), np.zeros((4, 2), np.float32), batch_edges.numpy(), np.ones((2, 2), np.float32), None,
                        negs.numpy(), None, 0.1, O.REDUCTION_SUM)
    assert batch.node_embeddings_state is None
    assert np.allclose(batch.node_gradients.cpu().numpy(), ref.delta_e, rtol=1e-5, atol=1e-7)
    assert np.allclose(batch.node_state_update.cpu().numpy(), ref.delta_s, rtol=1e-5, atol=1e-9)
    # dense relation step happened (Adagrad on ones): relations moved where the gradient is non-zero
    assert not torch.equal(model.decoder.relations.detach().cpu(), torch.ones(2, 2))


@pytest.mark.parametrize("decoder_name,kind", [("DistMult", 1), ("ComplEx", 2)])
def test_generic_autograd_path_matches_fused(host, decoder_name, kind):
    """loss.backward() through the fused decoder autograd::Function (any libtorch loss) == the fully fused SoftmaxCE path == oracle."""
    rng = np.random.default_rng(3)
    uniq, edges, dn, sn = O.make_batch(rng, 2000, 5, 96, 2, 64)
    U, d = len(uniq), 32
    emb = rng.uniform(-0.4, 0.4, (U, d)).astype(np.float32)
    state = rng.uniform(0, 0.05, (U, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (5, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (5, d)).astype(np.float32)
    ref = O.train_batch(kind, emb, state, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM, acc=np.float64)

    def make():
        dec = getattr(host.nn.decoders.edge, decoder_name)(num_relations=5, embedding_dim=d, use_inverse_relations=True, device=CUDA, mode="train")
        with torch.no_grad():
            dec.relations.copy_(torch.from_numpy(rel))
            dec.inverse_relations.copy_(torch.from_numpy(inv_rel))
        m = host.nn.Model(dec, host.nn.SoftmaxCrossEntropy(reduction="sum"), CUDA)
        b = host.data.Batch(True)
        b.node_embeddings = torch.from_numpy(emb).to(CUDA)
        b.node_embeddings_state = torch.from_numpy(state).to(CUDA)
        b.edges = torch.from_numpy(edges).to(CUDA)
        b.dst_neg_indices_mapping = torch.from_numpy(dn).to(CUDA)
        b.src_neg_indices_mapping = torch.from_numpy(sn).to(CUDA)
        return m, b

    # (1) fused path
    m, b = make()
    m.train_batch(b, False)
    rel_err = lambda a, r: float(np.abs(a.cpu().numpy() - r).max() / max(np.abs(r).max(), 1e-6))
    assert rel_err(b.node_gradients, ref.delta_e) < 1e-4 and rel_err(b.node_state_update, ref.delta_s) < 1e-4
    assert rel_err(m.decoder.relations.grad, ref.rel_grad) < 1e-4 and rel_err(m.decoder.inverse_relations.grad, ref.inv_rel_grad) < 1e-4
    # (2) generic autograd path: scores -> libtorch loss -> backward through mb_decoder_backward
    m2, b2 = make()
    b2.node_embeddings.requires_grad_()
    pos, neg, ipos, ineg = m2.forward_lp(b2, True)
    assert rel_err(neg.detach(), ref.scores.neg) < 1e-4 and rel_err(ineg.detach(), ref.scores.inv_neg) < 1e-4
    loss = m2.loss_function(ipos, ineg, True) + m2.loss_function(pos, neg, True)
    assert abs(float(loss.item()) - float(ref.loss)) < 1e-4 * abs(float(ref.loss))
    loss.backward()
    assert rel_err(b2.node_embeddings.grad, ref.grad) < 1e-4
    assert rel_err(m2.decoder.relations.grad, ref.rel_grad) < 1e-4
    b2.accumulateGradients(0.1)
    assert rel_err(b2.node_gradients, ref.delta_e) < 1e-4


# ---- test_data.py ----------------------------------------------------------------------------------------------
def test_batch_construction_and_accumulate_gradients(host):
    b1 = host.data.Batch(train=False)
    assert b1.node_embeddings is None and b1.train is False and b1.device_id == -1
    # test_data.py:34-47
    b = host.data.Batch(train=True)
    b.node_embeddings = torch.tensor([2.0, 4.0], device=CUDA)
    b.node_embeddings.grad = torch.tensor([0.5, -1.0], device=CUDA)
    b.node_embeddings_state = torch.tensor([0.0, 0.0], device=CUDA)
    b.accumulateGradients(learning_rate=1.0)
    assert b.node_embeddings_state is None
    assert torch.all(torch.eq(b.node_state_update, b.node_embeddings.grad.pow(2))).item() is True
    expected = -1.0 * (b.node_embeddings.grad / (b.node_state_update.sqrt().add_(1e-10)))
    assert torch.all(torch.eq(b.node_gradients, expected)).item() is True
    b.clear()
    assert b.node_embeddings is None and b.node_gradients is None


# ---- test_storage.cpp / test_buffer.cpp ------------------------------------------------------------------------
def test_inmemory_index_read_add_put(host, tmp_path):
    # test_storage.cpp:260-316
    rand = torch.randn(200, 24)
    fn = str(tmp_path / "emb.bin")
    st = host.storage.InMemory(fn, rand, CUDA)
    st.load()
    idx = torch.randint(200, (50,))
    assert torch.equal(st.indexRead(idx).cpu(), rand.index_select(0, idx))
    with pytest.raises(RuntimeError):
        st.indexRead(torch.randint(100, (10, 10)))
    uidx = torch.randperm(200)[:60]
    vals = torch.randn(60, 24)
    st.indexAdd(uidx, vals)
    exp = rand.clone().index_add_(0, uidx, vals)
    assert torch.equal(st.indexRead(uidx).cpu(), exp.index_select(0, uidx))
    with pytest.raises(RuntimeError):
        st.indexAdd(uidx, torch.randn(61, 24))
    with pytest.raises(RuntimeError):
        st.indexAdd(uidx, torch.randn(60, 25))
    st.indexPut(uidx, vals)
    assert torch.equal(st.indexRead(uidx).cpu(), vals)
    assert torch.equal(st.range(10, 5).cpu(), st.indexRead(torch.arange(10, 15)).cpu())
    # flat fp32 row-major file format on unload(write=true) (constants.h:39-42)
    st.unload(True)
    disk = torch.from_numpy(np.fromfile(fn, dtype=np.float32).reshape(200, 24))
    exp[uidx] = vals
    assert torch.equal(disk, exp)


@pytest.fixture(params=[False, True], ids=["sync", "prefetching"])
def prefetching(request):
    """both swap engines: synchronous (buffer.cpp:635-683 without prefetching) and LookaheadBlock / AsyncWriteBlock (buffer.cpp:118-322)"""
    return request.param


class TestPartitionBuffer:
    # test_buffer.cpp:20-75 fixture: 45 rows, 5 partitions of 10 (last 5), capacity 2, ordering of TestPartitionBufferOrdering
    total, nparts, psize, d, cap = 45, 5, 10, 16, 2
    states = [[0, 1], [0, 2], [0, 3], [0, 4], [1, 4], [1, 3], [1, 2], [3, 2], [4, 2], [4, 3]]

    @pytest.fixture(autouse=True)
    def _mode(self, prefetching):
        self.prefetching = prefetching

    def make(self, host, tmp_path):
        rand = torch.randn(self.total, self.d)
        fn = str(tmp_path / "pb.bin")
        rand.numpy().tofile(fn)
        pb = host.storage.PartitionBuffer(self.cap, self.nparts, 1, self.psize, self.d, self.total, fn, self.prefetching, CUDA)
        pb.setBufferOrdering([torch.tensor(s) for s in self.states])
        pb.load()
        return pb, rand, fn

    def test_ordering(self, host, tmp_path):
        # test_buffer.cpp:241-259
        pb, _, _ = self.make(host, tmp_path)
        admits, evicts = [], []
        while pb.hasSwap():
            admits.append(pb.getNextAdmit()[0])
            evicts.append(pb.getNextEvict()[0])
            pb.performNextSwap()
        assert admits == [2, 3, 4, 1, 3, 2, 3, 4, 3] and evicts == [1, 2, 3, 0, 4, 3, 1, 3, 2]
        assert pb.hasSwap() is False

    def test_index_read_add(self, host, tmp_path, golden_dir):
        # test_buffer.cpp:275-297
        pb, rand, _ = self.make(host, tmp_path)
        idx = pb.getRandomIds(20)
        assert torch.equal(rand.index_select(0, idx), pb.indexRead(idx).cpu())
        with pytest.raises(RuntimeError):
            pb.indexRead(torch.randint(1000, (10, 10)))
        uidx = torch.unique(pb.getRandomIds(1000))
        vals = torch.randint(1000, (uidx.size(0), self.d)).float()
        upd = rand.clone().index_add_(0, uidx, vals).index_select(0, uidx)
        pb.indexAdd(uidx, vals)
        assert torch.equal(upd, pb.indexRead(uidx).cpu())
        with pytest.raises(RuntimeError):
            pb.indexAdd(uidx, torch.zeros(uidx.size(0) + 1, self.d))
        with pytest.raises(RuntimeError):
            pb.indexAdd(uidx, torch.zeros(uidx.size(0), self.d + 1))
        with pytest.raises(RuntimeError):
            pb.indexAdd(torch.randint(1000, (10, 10)), vals)

    def test_global_map_and_sync(self, host, tmp_path):
        # test_buffer.cpp:299-318
        pb, rand, fn = self.make(host, tmp_path)
        exp = -torch.ones(self.total, dtype=torch.int64)
        exp[:20] = torch.arange(20)
        assert torch.equal(exp, pb.getGlobalToLocalMap(True).cpu())
        exp[10:20] = -1
        exp[20:30] = torch.arange(10, 20)
        assert torch.equal(exp, pb.getGlobalToLocalMap(False).cpu())
        # updates survive eviction: add to partition 1 rows, swap it out, read the file
        ids = torch.arange(10, 20)
        vals = torch.ones(10, self.d)
        pb.indexAdd(ids, vals)
        pb.performNextSwap()  # evicts partition 1
        pb.sync()  # (with prefetching the write-back is asynchronous: sync() returns once it has reached the file)
        disk = torch.from_numpy(np.fromfile(fn, dtype=np.float32).reshape(self.total, self.d))
        assert torch.equal(disk[10:20], rand[10:20] + 1)
        # partition 2 now lives in slot 1: buffer-local row 10 is global row 20
        assert torch.equal(pb.indexRead(torch.tensor([10])).cpu(), rand[20:21])
        pb.unload(True)
        disk = torch.from_numpy(np.fromfile(fn, dtype=np.float32).reshape(self.total, self.d))
        assert torch.equal(disk[20:30], rand[20:30]) and torch.equal(disk[:10], rand[:10])

    def test_against_reference_golden(self, host, tmp_path, golden_dir):
        g = np.load(os.path.join(golden_dir, "partition_buffer.npz"))
        fn = str(tmp_path / "pbg.bin")
        g["table"].tofile(fn)
        pb = host.storage.PartitionBuffer(int(g["cap"]), int(g["nparts"]), 1, int(g["psize"]), int(g["d"]), int(g["total"]), fn, self.prefetching, CUDA)
        pb.setBufferOrdering([torch.from_numpy(s) for s in g["states"]])
        pb.load()
        idx = torch.from_numpy(g["idx"])
        assert np.array_equal(pb.indexRead(idx).cpu().numpy(), g["read"])
        assert np.array_equal(pb.getGlobalToLocalMap(True).cpu().numpy(), g["map_current"])
        assert np.array_equal(pb.getGlobalToLocalMap(False).cpu().numpy(), g["map_next"])
        pb.indexAdd(idx, torch.from_numpy(g["vals"]))
        assert np.array_equal(pb.indexRead(idx).cpu().numpy(), g["read_after_add"])
        while pb.hasSwap():
            pb.performNextSwap()
        pb.unload(True)
        assert np.array_equal(np.fromfile(fn, dtype=np.float32).reshape(g["table"].shape), g["file_after"])


    def test_updates_survive_a_full_epoch_of_swaps(self, host, tmp_path):
        """Every buffer state: add a known value to random resident rows, swap.  After the epoch the file equals the host-side replay
        exactly, whatever the swap engine (asynchronous write-backs / lookahead reads must never lose or reorder an update)."""
        pb, rand, fn = self.make(host, tmp_path)
        exp = rand.clone()
        g = torch.Generator().manual_seed(5)
        si = 0
        while True:
            state = self.states[si]
            local = torch.unique(torch.randint(0, 2 * self.psize, (12,), generator=g))
            # buffer-local row -> global row through the current map
            m = pb.getGlobalToLocalMap(True).cpu()
            inv = {int(m[gid]): gid for gid in range(self.total) if int(m[gid]) >= 0}
            local = torch.tensor([l for l in local.tolist() if l in inv])
            vals = torch.randint(1, 50, (local.numel(), self.d), generator=g).float()
            pb.indexAdd(local, vals)
            for l, v in zip(local.tolist(), vals):
                exp[inv[l]] += v
            assert sorted(int(x) for x in pb.getBufferState()) == sorted(state)
            if not pb.hasSwap():
                break
            pb.performNextSwap()
            si += 1
        pb.unload(True)
        disk = torch.from_numpy(np.fromfile(fn, dtype=np.float32).reshape(self.total, self.d))
        assert torch.equal(disk, exp)


class TestPartitionBufferStorage:
    """storage.cpp:67-201 / test_storage.cpp:294-464: the Storage facade GraphModelStorage holds for buffered tables."""

    def test_facade(self, host, tmp_path, prefetching):
        total, d, nparts, cap = 45, 8, 5, 2
        rand = torch.randn(total, d)
        fn = str(tmp_path / "pbs.bin")
        st = host.storage.PartitionBufferStorage(fn, rand, nparts, cap, prefetching, 1, CUDA)  # appends the tensor to the file (storage.cpp:82-96)
        assert st.dim0_size == total and st.dim1_size == d
        assert np.array_equal(np.fromfile(fn, dtype=np.float32).reshape(total, d), rand.numpy())
        st.setBufferOrdering([torch.tensor(s) for s in TestPartitionBuffer.states])
        st.load()
        assert st.getNumInMemory() == cap * 9  # partition_size = ceil(45 / 5)
        idx = torch.arange(0, 18)
        assert torch.equal(st.indexRead(idx).cpu(), rand[:18])
        st.indexAdd(idx, torch.ones(18, d))
        assert torch.equal(st.indexRead(idx).cpu(), rand[:18] + 1)
        for bad in (lambda: st.range(0, 4), lambda: st.indexPut(idx, torch.ones(18, d)), lambda: st.rangePut(0, 4, torch.ones(4, d)),
                    lambda: st.shuffle(), lambda: st.sort(True)):
            with pytest.raises(RuntimeError):  # storage.cpp:178-201
                bad()
        assert st.hasSwap() and st.getNextAdmit() == [2] and st.getNextEvict() == [1]
        exp = -torch.ones(total, dtype=torch.int64)
        exp[:18] = torch.arange(18)
        assert torch.equal(st.getGlobalToLocalMap(True).cpu(), exp)
        st.performNextSwap()
        st.unload(True)
        disk = torch.from_numpy(np.fromfile(fn, dtype=np.float32).reshape(total, d))
        assert torch.equal(disk[:18], rand[:18] + 1) and torch.equal(disk[18:], rand[18:])


def test_train_batch_fused_on_device_tables(host):
    rng = np.random.default_rng(11)
    num_nodes, R, B, C, N, d = 3000, 4, 200, 2, 100, 48
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    state = np.zeros((num_nodes, d), np.float32)
    emb_st = host.storage.InMemory(torch.from_numpy(table).to(CUDA))
    state_st = host.storage.InMemory(torch.from_numpy(state).to(CUDA))
    dec = host.nn.decoders.edge.ComplEx(num_relations=R, embedding_dim=d, use_inverse_relations=True, device=CUDA, mode="train")
    model = host.nn.Model(dec, host.nn.SoftmaxCrossEntropy(reduction="sum"), CUDA)
    rel = dec.relations.detach().cpu().numpy().copy()
    inv = dec.inverse_relations.detach().cpu().numpy().copy()
    uniq, edges, dn, sn = O.make_batch(rng, num_nodes, R, B, C, N)
    b = host.data.Batch(True)
    b.unique_node_indices = torch.from_numpy(uniq)
    b.edges = torch.from_numpy(edges)
    b.dst_neg_indices_mapping = torch.from_numpy(dn)
    b.src_neg_indices_mapping = torch.from_numpy(sn)
    loss = model.train_batch_fused(b, emb_st, state_st, False)
    res = O.train_step_on_table(O.COMPLEX, table, state, uniq, edges, rel, inv, dn, sn, 0.1, O.REDUCTION_SUM, acc=np.float64)
    assert abs(loss - float(res.loss)) < 1e-4 * abs(float(res.loss))
    got = emb_st.data.cpu().numpy()
    assert np.abs(got - table).max() / np.abs(table).max() < 1e-4
    assert np.abs(state_st.data.cpu().numpy() - state).max() / max(np.abs(state).max(), 1e-6) < 1e-4


# ---- evaluation: Model::evaluate_batch + LinkPredictionReporter (model.cpp:335-349, reporting.cpp:44-95) --------------------------
@pytest.mark.parametrize("name", ["eval_distmult_pad.npz", "eval_complex_filter.npz", "eval_distmult_all.npz"])
def test_evaluate_batch_reporter(host, name):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    kind, d, R = int(g["kind"]), int(g["d"]), g["rel"].shape[0]
    cls = host.nn.decoders.edge.DistMult if kind == O.DISTMULT else host.nn.decoders.edge.ComplEx
    dec = cls(num_relations=R, embedding_dim=d, use_inverse_relations=True, device=CUDA, mode="train")
    with torch.no_grad():
        dec.relations.copy_(torch.from_numpy(g["rel"]))
        dec.inverse_relations.copy_(torch.from_numpy(g["inv_rel"]))
    model = host.nn.Model(dec, host.nn.SoftmaxCrossEntropy(reduction="sum"), CUDA)
    rep = host.report.LinkPredictionReporter()
    for m in (host.report.MeanRank(), host.report.MeanReciprocalRank(), host.report.Hitsk(1), host.report.Hitsk(3), host.report.Hitsk(10)):
        rep.add_metric(m)
    model.reporter = rep
    batch = host.data.Batch(False)
    batch.node_embeddings = torch.from_numpy(g["emb"]).to(CUDA)
    batch.edges = torch.from_numpy(g["edges"]).to(CUDA)
    batch.dst_neg_indices_mapping = torch.from_numpy(g["dst_negs"]).to(CUDA)
    batch.src_neg_indices_mapping = torch.from_numpy(g["src_negs"]).to(CUDA)
    if g["dst_filter"].shape[0]:
        batch.dst_neg_filter = torch.from_numpy(g["dst_filter"]).to(CUDA)
        batch.src_neg_filter = torch.from_numpy(g["src_filter"]).to(CUDA)
    model.evaluate_batch(batch)
    assert len(rep.per_batch_ranks) == 2  # dst-corruption ranks first, then src-corruption ranks (model.cpp:343-348)
    ranks, inv_ranks = rep.per_batch_ranks[0].cpu().numpy(), rep.per_batch_ranks[1].cpu().numpy()
    # the generic route of the adapter (forward_lp -> add_result -> compute_ranks) gives the same integers
    pos, neg, inv_pos, inv_neg = model.forward_lp(batch, True)
    assert np.array_equal(rep.compute_ranks(pos, neg).cpu().numpy(), ranks) and np.array_equal(rep.compute_ranks(inv_pos, inv_neg).cpu().numpy(), inv_ranks)
    assert np.array_equal(ranks, O.compute_ranks(pos.detach().cpu().numpy(), neg.detach().cpu().numpy()))

    def close(r, ref, p, n):  # equal to the reference's ranks up to near-ties
        scale = np.abs(n[n > -1e8]).max()
        return bool(np.all(np.abs(r - ref) <= (np.abs(n - p[:, None]) <= 2e-4 * scale).sum(axis=1)))

    assert close(ranks, g["ref_ranks"], g["ref_pos"], g["ref_neg"]) and close(inv_ranks, g["ref_inv_ranks"], g["ref_inv_pos"], g["ref_inv_neg"])
    text = rep.report()
    assert f"Link Prediction: {2 * len(ranks)} edges evaluated" in text and "MRR: " in text and "Hits@10: " in text
    assert rep.all_ranks.device.type == "cpu" and rep.all_ranks.numel() == 2 * len(ranks) and len(rep.per_batch_ranks) == 0
    mrr = float(text.split("MRR: ")[1].split("\n")[0])
    assert mrr == pytest.approx(float(g["metrics"][1]), rel=0.02)


def test_evaluate_batch_needs_reporter(host):
    model = get_test_model_lp(host, mode="train")
    batch = host.data.Batch(False)
    batch.node_embeddings = node_embeddings.to(CUDA)
    batch.edges = batch_edges.to(CUDA)
    batch.dst_neg_indices_mapping = torch.tensor([[2, 0], [0, 1], [1, 0]]).to(CUDA)
    with pytest.raises(RuntimeError):
        model.evaluate_batch(batch)


def test_compute_worker_gpu_stage(host):
    """ComputeWorkerGPU adapter (pipeline_gpu.cpp:33-104 for device-resident tables): batches pushed into its loaded-batches queue are
    trained by its worker thread (load -> train -> update fused) in order, and come out of the update-batches queue with their loss;
    tables equal the same batches applied by direct Model::train_batch_fused calls."""
    rng = np.random.default_rng(17)
    num_nodes, R, B, C, N, d = 4000, 4, 256, 2, 128, 64
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    batches = [O.make_batch(rng, num_nodes, R, B, C, N) for _ in range(6)]

    def mk():
        emb = host.storage.InMemory(torch.from_numpy(table).to(CUDA))
        st = host.storage.InMemory(torch.zeros(num_nodes, d).to(CUDA))
        dec = host.nn.decoders.edge.ComplEx(num_relations=R, embedding_dim=d, use_inverse_relations=True, device=CUDA, mode="train")
        return emb, st, host.nn.Model(dec, host.nn.SoftmaxCrossEntropy(reduction="sum"), CUDA)

    def mkbatch(b):
        uniq, edges, dn, sn = b
        x = host.data.Batch(True)
        x.unique_node_indices = torch.from_numpy(uniq)
        x.edges = torch.from_numpy(edges)
        x.dst_neg_indices_mapping = torch.from_numpy(dn)
        x.src_neg_indices_mapping = torch.from_numpy(sn)
        return x

    e1, s1, m1 = mk()
    direct = [m1.train_batch_fused(mkbatch(b), e1, s1, True) for b in batches]
    e2, s2, m2 = mk()
    w = host.pipeline.ComputeWorkerGPU(m2, e2, s2, 2)
    w.start()
    for b in batches:
        w.push(mkbatch(b))
    got = [w.pop_finished().loss for _ in batches]
    w.stop()
    assert w.error == "" and w.batches_processed == len(batches) and w.edges_processed == len(batches) * B
    assert got == direct
    assert torch.equal(e1.data, e2.data) and torch.equal(s1.data, s2.data)
