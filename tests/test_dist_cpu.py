"""Multi-rank row exchange (marius_b200.dist.ShardedTable) on CPU: world_size 2, gloo, numpy-oracle backend.
The routing logic (owner bucketing, the three all-to-alls, owner-side merge of duplicate rows, Adagrad at the owner) must
reproduce a single-process computation exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """CPU stand-in for OpsBackend built on the oracle (tests only)."""

    def __init__(self, table, state):
        from oracle import marius_oracle as O

        self.O, self.table, self.state = O, table, state

    def gather(self, local_rows, out=None):
        rows = torch.from_numpy(self.O.index_read(self.table, local_rows.numpy()))
        if out is not None:
            out.copy_(rows)
            return out
        return rows

    def train_batch(self, kind, emb, edges, rel, inv_rel, dn, sn, reduction):
        r = self.O.train_batch(kind, emb.numpy(), np.zeros_like(emb.numpy()), edges.numpy(), rel.numpy(), inv_rel.numpy(), dn.numpy(), sn.numpy(), 0.0,
                               reduction)
        return dict(loss=torch.tensor([float(r.loss)]), grad=torch.from_numpy(r.grad), rel_grad=torch.from_numpy(r.rel_grad),
                    inv_rel_grad=torch.from_numpy(r.inv_rel_grad))

    def update(self, local_rows, grads, lr):
        idx = local_rows.numpy()
        de, ds = self.O.accumulate_gradients(grads.numpy(), self.state[idx], lr)
        self.O.index_add(self.table, idx, de)
        self.O.index_add(self.state, idx, ds)


def _problem(world, rows_per_rank, d, R, B, C, N):
    from oracle import marius_oracle as O

    rng = np.random.default_rng(2024)
    total = world * rows_per_rank
    table = rng.uniform(-0.4, 0.4, (total, d)).astype(np.float32)
    state = rng.uniform(0, 0.05, (total, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    batches = [O.make_batch(rng, total, R, B, C, N) for _ in range(world)]  # heavy overlap between ranks: total is small
    return table, state, rel, inv_rel, batches


def _worker(rank, world, port, rows_per_rank, d, R, B, C, N, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from marius_b200.dist import ShardedTable, owner_bounds
    from oracle import marius_oracle as O

    table, state, rel, inv_rel, batches = _problem(world, rows_per_rank, d, R, B, C, N)
    lo, hi = rank * rows_per_rank, (rank + 1) * rows_per_rank
    shard, sstate = table[lo:hi].copy(), state[lo:hi].copy()
    st = ShardedTable(rows_per_rank, OracleBackend(shard, sstate))
    uniq, edges, dn, sn = batches[rank]
    # bucketing sanity: slices partition the sorted unique list by owner
    b = owner_bounds(torch.from_numpy(uniq), rows_per_rank, world).tolist()
    for j in range(world):
        seg = uniq[b[j]:b[j + 1]]
        assert ((seg // rows_per_rank) == j).all()
    out = st.train_step(O.DISTMULT, torch.from_numpy(uniq), torch.from_numpy(edges), torch.from_numpy(rel), torch.from_numpy(inv_rel),
                        torch.from_numpy(dn), torch.from_numpy(sn), 0.1, O.REDUCTION_SUM)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), shard=shard, state=sstate, rel_grad=out["rel_grad"].numpy(), loss=out["loss"].numpy(),
             remote=st.last_remote_rows)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_sharded_train_step_world2(tmp_path):
    from oracle import marius_oracle as O

    world, rows_per_rank, d, R, B, C, N = 2, 150, 16, 3, 40, 2, 24
    mp.spawn(_worker, args=(world, _free_port(), rows_per_rank, d, R, B, C, N, str(tmp_path)), nprocs=world, join=True)
    # single-process reference: both batches read the same pre-update table; every owner then applies the contributions rank by
    # rank, each as one sparse Adagrad update (Batch::accumulateGradients + indexAdd x2)
    table, state, rel, inv_rel, batches = _problem(world, rows_per_rank, d, R, B, C, N)
    total = world * rows_per_rank
    gsum = np.zeros((total, d), np.float32)
    touched = np.zeros(total, bool)
    rel_grad = np.zeros_like(rel)
    losses = []
    per_owner = [[] for _ in range(world)]
    for r, (uniq, edges, dn, sn) in enumerate(batches):
        res = O.train_batch(O.DISTMULT, table[uniq], np.zeros((len(uniq), d), np.float32), edges, rel, inv_rel, dn, sn, 0.0, O.REDUCTION_SUM)
        losses.append(float(res.loss))
        rel_grad += res.rel_grad
        for o in range(world):
            m = (uniq // rows_per_rank) == o
            per_owner[o].append((uniq[m], res.grad[m]))
    exp_table, exp_state = table.copy(), state.copy()
    for o in range(world):
        for ids, g in per_owner[o]:  # sender rank order
            de, ds = O.accumulate_gradients(g, exp_state[ids], 0.1)
            exp_table[ids] += de
            exp_state[ids] += ds
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        lo, hi = r * rows_per_rank, (r + 1) * rows_per_rank
        assert np.array_equal(got["shard"], exp_table[lo:hi])
        assert np.array_equal(got["state"], exp_state[lo:hi])
        assert np.allclose(got["rel_grad"], rel_grad, rtol=1e-6, atol=1e-7)  # all-reduced
        assert abs(float(got["loss"][0]) - losses[r]) < 1e-4 * abs(losses[r])
        assert int(got["remote"]) > 0  # rows really crossed the rank boundary
