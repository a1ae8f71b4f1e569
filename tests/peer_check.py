"""tests/peer_check.py -- 2+ rank check of the peer-memory sharded step (run under torchrun, one rank per GPU).
Rank r builds a batch over the WHOLE global id space restricted to ids == r (mod world): the batches of different ranks touch
disjoint rows, so there are no cross-rank races and the final tables must equal the oracle applying every batch once."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from marius_b200 import ops
    from marius_b200.dist import PeerShardedTable
    from oracle import marius_oracle as O

    rows, d, R, B, C, N = 4096, 64, 5, 256, 2, 128
    total = rows * world
    rng = np.random.default_rng(7)
    full = rng.uniform(-0.3, 0.3, (total, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    batches = []
    for r in range(world):
        brng = np.random.default_rng(100 + r)
        pick = lambda size: (brng.integers(0, total // world, size=size, dtype=np.int64) * world + r)  # ids == r (mod world): all partitions
        src, dst, relid = pick(B), pick(B), brng.integers(0, R, size=B, dtype=np.int64)
        sn, dn = pick((C, N)), pick((C, N))
        uniq, inv = O.map_tensors(np.concatenate([src, dst, sn.reshape(-1), dn.reshape(-1)]))
        edges = np.ascontiguousarray(np.stack([inv[:B], relid, inv[B:2 * B]], axis=1))
        batches.append((uniq, edges, np.ascontiguousarray(inv[2 * B + C * N:].reshape(C, N)), np.ascontiguousarray(inv[2 * B:2 * B + C * N].reshape(C, N))))
    table = torch.from_numpy(full[rank * rows:(rank + 1) * rows].copy()).to(dev)
    state = torch.zeros(rows, d, device=dev)
    ctx = ops.Context(local)
    pst = PeerShardedTable(table, state, ctx)
    t = lambda a: torch.from_numpy(a).to(dev)
    uniq, edges, dn, sn = batches[rank]
    rg, irg = torch.empty(R, d, device=dev), torch.empty(R, d, device=dev)
    loss = pst.train_step(ops.COMPLEX, t(uniq), t(edges), t(rel), t(inv_rel), t(dn), t(sn), 0.1, ops.REDUCTION_SUM, rel_grad=rg, inv_rel_grad=irg)
    torch.cuda.synchronize()
    dist.barrier()
    # oracle: every batch applied once to the full table (disjoint rows => order-free)
    exp_t, exp_s = full.copy(), np.zeros_like(full)
    ref_loss = None
    for r, (u, e, dnn, snn) in enumerate(batches):
        res = O.train_step_on_table(O.COMPLEX, exp_t, exp_s, u, e, rel, inv_rel, dnn, snn, 0.1, O.REDUCTION_SUM, acc=np.float64)
        if r == rank:
            ref_loss = float(res.loss)
    got_t, got_s = table.cpu().numpy(), state.cpu().numpy()
    lo, hi = rank * rows, (rank + 1) * rows
    err = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-9))
    et, es = err(got_t, exp_t[lo:hi]), err(got_s, exp_s[lo:hi])
    touched_remote = int(((uniq // rows) != rank).sum())
    ok = et < 1e-4 and es < 1e-4 and abs(float(loss.item()) - ref_loss) < 1e-4 * abs(ref_loss) and touched_remote > 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"PEER_CHECK {'OK' if int(flag.item()) == 1 else 'FAIL'} world={world} table_err={et:.2e} state_err={es:.2e} remote_rows={touched_remote}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
