import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from marius_b200 import ops
from oracle import marius_oracle as O
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def run(world, steps, graph, kind=ops.COMPLEX, d=48):
    rng = np.random.default_rng(40 + world)
    rows, R, B, C, N = 3000, 4, 200, 2, 96
    total = rows * world
    full = rng.uniform(-0.3, 0.3, (total, d)).astype(np.float32)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32); inv_rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    uniq, edges, dn, sn = O.make_batch(rng, total, R, B, C, N)
    ctx = ops.Context(0); ctx.graph(graph)
    tables = [dev(full[r * rows:(r + 1) * rows].copy()) for r in range(world)]
    states = [torch.zeros(rows, d, device="cuda") for _ in range(world)]
    sh = ops.make_shards(tables, states, rows)
    rg = torch.empty(R, d, device="cuda"); loss = torch.zeros(1, device="cuda")
    keep = []
    for step in range(steps):
        args = (dev(uniq), dev(edges), dev(rel), dev(inv_rel), dev(dn), dev(sn)); keep.append(args)
        ops.train_step_sharded(ctx, kind, sh, tables[0].stride(0), d, *args, 0.1, loss=loss, rel_grad=rg)
    torch.cuda.synchronize()
    exp_t, exp_s = full.copy(), np.zeros_like(full)
    for step in range(steps):
        res = O.train_step_on_table(kind, exp_t, exp_s, uniq, edges, rel, inv_rel, dn, sn, 0.1, O.REDUCTION_SUM, acc=np.float64)
    got_t = np.concatenate([t.cpu().numpy() for t in tables]); got_s = np.concatenate([s.cpu().numpy() for s in states])
    nan_rows = np.unique(np.argwhere(np.isnan(got_t))[:, 0])
    err = np.nanmax(np.abs(got_t - exp_t)) / np.abs(exp_t).max()
    print(f"world={world} steps={steps} graph={graph} kind={kind}: nan_rows={len(nan_rows)} first={nan_rows[:6]} err(excl nan)={err:.2e} loss={float(loss.item()):.4f} ref={float(res.loss):.4f} "
          f"nan_in_uniq={np.isin(nan_rows, uniq).all() if len(nan_rows) else None} owners={np.unique(nan_rows // rows) if len(nan_rows) else None}", flush=True)
for world in (1, 2):
    for steps, graph in ((1, False), (3, False), (3, True)):
        run(world, steps, graph)
run(2, 1, False, kind=ops.DISTMULT)
run(2, 1, False, d=128)
