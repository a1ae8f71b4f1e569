"""tools/timeline.py -- device timeline of one overlapped step (CUDA events on the step's own streams, eager launches).
    python tools/timeline.py [--batch 20000] [--steps 3]
Prints, for the last step, every stage launch with its start / end in us from the step's first kernel."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from marius_b200 import ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--nodes", type=int, default=4_000_000)
    args = ap.parse_args()
    D, NEG, R, B = 400, 1000, 1000, args.batch
    C = B // 1000
    dev = torch.device("cuda:0")
    table = torch.empty((args.nodes, D), device=dev).uniform_(-0.1, 0.1)
    state = torch.zeros_like(table)
    rels = torch.zeros((2, R, D), device=dev)
    rels[:, :, : D // 2] = 1.0
    rg, irg = torch.empty(R, D, device=dev), torch.empty(R, D, device=dev)
    loss = torch.zeros(1, device=dev)
    ctx = ops.Context(0)
    rng = np.random.default_rng(0)
    host_batches, _ = bench.make_batches(rng, args.nodes, args.steps + 2, B)
    batches = [tuple(torch.from_numpy(x).to(dev) for x in b) for b in host_batches]
    for i in range(2):
        u, e, dn, sn = batches[i]
        ops.train_step(ctx, ops.COMPLEX, table, state, u, e, rels[0], rels[1], dn, sn, 0.1, ops.REDUCTION_SUM, ops.PREC_BF16X3, loss=loss, rel_grad=rg, inv_rel_grad=irg)
    torch.cuda.synchronize()
    ctx.profile(2)
    for i in range(2, 2 + args.steps):
        u, e, dn, sn = batches[i]
        ops.train_step(ctx, ops.COMPLEX, table, state, u, e, rels[0], rels[1], dn, sn, 0.1, ops.REDUCTION_SUM, ops.PREC_BF16X3, loss=loss, rel_grad=rg, inv_rel_grad=irg)
    tl = ctx.profile_timeline()
    ctx.profile(0)
    per = len(tl) // args.steps
    last = tl[-per:]
    t0 = min(a for _, a, _ in last)
    print(f"B={B}: {per} stage launches per step; last step spans {1e3 * (max(b for _, _, b in last) - t0):.1f} us")
    for name, a, b in sorted(last, key=lambda x: x[1]):
        print(f"  {1e3 * (a - t0):8.1f} -> {1e3 * (b - t0):8.1f} us  ({1e3 * (b - a):7.1f})  {name}")


if __name__ == "__main__":
    main()
