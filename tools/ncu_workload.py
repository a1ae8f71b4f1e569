"""tools/ncu_workload.py -- the workload the committed ncu captures are taken on (profiles/r1_*).
One pass = the three standalone storage ops (gather / scatter-add / fused Adagrad row update on U = 40 000 random rows of a
2e6 x 400 table) followed by eager (MB_GRAPH=0) fused training steps of the bench shape (ComplEx d=400, 1000 negatives, batch 10 000).
The table is 3.2 GB (>> 126 MB of L2, random rows) but small enough that ncu's per-kernel save/restore stays cheap.
Usage: MB_GRAPH=0 python tools/ncu_workload.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from marius_b200 import ops

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
d, rows, B, C, N, R = 400, 2_000_000, 10_000, 10, 1000, 1000
g = torch.Generator(device="cuda").manual_seed(0)
table = (torch.rand(rows, d, device="cuda", generator=g) - 0.5) * 0.2
state = torch.zeros(rows, d, device="cuda")
rel = torch.rand(R, d, device="cuda", generator=g) - 0.5
inv_rel = torch.rand(R, d, device="cuda", generator=g) - 0.5
ctx = ops.Context(0)
ctx.graph(False)
rng = np.random.default_rng(0)


def batch():
    ids = rng.choice(rows, size=2 * B + 2 * C * N, replace=False)
    uniq = np.sort(ids).astype(np.int64)
    U = uniq.shape[0]
    edges = np.stack([rng.integers(0, U, B), rng.integers(0, R, B), rng.integers(0, U, B)], 1).astype(np.int64)
    dn = rng.integers(0, U, (C, N)).astype(np.int64)
    sn = rng.integers(0, U, (C, N)).astype(np.int64)
    return [torch.from_numpy(a).cuda() for a in (uniq, edges, dn, sn)]


for it in range(2):  # standalone storage ops (InMemory / PartitionBuffer indexRead, indexAdd, and the fused Adagrad update)
    uniq = batch()[0]
    got = ops.gather_rows(table, uniq)
    ops.scatter_add_rows(table, uniq, got * 1e-3)
    ops.adagrad_update_rows(table, state, uniq, got, 0.1)
loss = torch.zeros(1, device="cuda")
rg, irg = torch.empty(R, d, device="cuda"), torch.empty(R, d, device="cuda")
for it in range(steps):
    uniq, edges, dn, sn = batch()
    ops.train_step(ctx, ops.COMPLEX, table, state, uniq, edges, rel, inv_rel, dn, sn, 0.1, ops.REDUCTION_SUM, ops.PREC_BF16X3, loss=loss, rel_grad=rg,
                   inv_rel_grad=irg)
torch.cuda.synchronize()
print("loss", float(loss.item()))
