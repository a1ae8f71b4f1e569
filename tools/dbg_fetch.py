"""tools/dbg_fetch.py -- bench-shape sharded step with all shards on ONE GPU (single_process shards): exercises the remote-row fetch,
the inbox shipping and the owner-side apply at B = 50 000 without a multi-GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from marius_b200 import ops
world, rows, d, B = 2, 1_000_000, 400, 50000
C, N, R = 50, 1000, 1000
dev = torch.device("cuda", 0)
tables = [torch.empty(rows, d, device=dev).uniform_(-0.1, 0.1) for _ in range(world)]
states = [torch.zeros(rows, d, device=dev) for _ in range(world)]
sh = ops.make_shards(tables, states, rows, rank=0, exchange_rows=2 * B + 2 * C * N)
ctx = ops.Context(0)
rng = np.random.default_rng(0)
rel = torch.ones(R, d, device=dev); inv = rel.clone()
rg, irg = torch.empty_like(rel), torch.empty_like(rel)
loss = torch.zeros(1, device=dev)
batches, _ = bench.make_sharded_batches(rng, rows, 0, world, 6, B)
for i, b in enumerate(batches):
    u, e, dn, sn = (torch.from_numpy(x).to(dev) for x in b)
    ops.train_step_sharded(ctx, ops.COMPLEX, sh, d, d, u, e, rel, inv, dn, sn, 0.1, loss=loss, rel_grad=rg, inv_rel_grad=irg)
    torch.cuda.synchronize()
    print("step", i, "loss", float(loss.item()), flush=True)
print("DBG_FETCH OK")
