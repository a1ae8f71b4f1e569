"""tools/gpu_probe.py -- first-contact GPU probe: each tcgen05 GEMM variant runs in its own subprocess (a device trap
poisons the CUDA context) under a timeout; then micro-timings of the HBM kernels and of the full step.
Usage on the GPU box:  python tools/gpu_probe.py [gemm|time|all]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from marius_b200 import ops
a_mn, b_mn, block_n, bt, M, N, K, prec = %s
ctx = ops.Context(0)
torch.manual_seed(0)
A = torch.randn(bt, M, K, device="cuda"); B = torch.randn(bt, N, K, device="cuda")
ref = torch.matmul(A.double(), B.double().transpose(1, 2))
Ain = A.transpose(1, 2).contiguous() if a_mn else A
Bin = B.transpose(1, 2).contiguous() if b_mn else B
D = ops.debug_gemm(ctx, Ain, a_mn, Bin, b_mn, prec, block_n)
torch.cuda.synchronize()
err = float((D.double() - ref).abs().max() / ref.abs().max())
nan = int(torch.isnan(D).sum())
print("RESULT", err, nan)
'''


def run_gemm_probe():
    cases = []
    cfgs = [int(x) for x in os.environ.get('PROBE_CFGS', '256,2560,128,512,5120').split(',')]
    for block_n in cfgs:
        for (a_mn, b_mn) in ((False, False), (False, True), (True, True), (True, False)):
            for shape in ((1, 128, 256, 64), (1, 128, 256, 128), (2, 1000, 1000, 400), (2, 1000, 400, 1000), (1, 200, 72, 136)):
                for prec in (2, 1):
                    cases.append((a_mn, b_mn, block_n) + shape + (prec,))
    for c in cases:
        code = CHILD % (ROOT, repr(c))
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
            out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            msg = out[0] if out else ("FAIL rc=%d %s" % (r.returncode, (r.stderr.strip().splitlines() or ["?"])[-1][:200]))
        except subprocess.TimeoutExpired:
            msg = "TIMEOUT"
        print("GEMM a_mn=%d b_mn=%d bn=%d bt=%d M=%d N=%d K=%d prec=%d -> %s  (%.1fs)" % (c + (msg, time.time() - t0)), flush=True)


def timeit(fn, iters=20, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_timing():
    import numpy as np
    import torch

    from marius_b200 import ops
    from oracle import marius_oracle as O

    ctx = ops.Context(0)
    dev = "cuda"
    d = 400
    rows = 20_000_000
    table = torch.empty(rows, d, device=dev).uniform_(-0.1, 0.1)
    state = torch.zeros(rows, d, device=dev)
    print("table GB", table.numel() * 4 / 1e9, flush=True)
    for n in (4000, 40000, 200000):
        idx = torch.randperm(rows, device=dev)[:n]
        out = torch.empty(n, d, device=dev)
        g = torch.randn(n, d, device=dev) * 0.01
        ms = timeit(lambda: ops.gather_rows(table, idx, out))
        print(f"gather n={n}: {ms*1e3:.1f} us  {8*n*d/ms/1e6:.1f} GB/s (algorithmic 8nd)", flush=True)
        ms = timeit(lambda: ops.scatter_add_rows(table, idx, g))
        print(f"scatter_add n={n}: {ms*1e3:.1f} us  {12*n*d/ms/1e6:.1f} GB/s (12nd)", flush=True)
        ms = timeit(lambda: ops.adagrad_update_rows(table, state, idx, g, 0.1))
        print(f"adagrad_update n={n}: {ms*1e3:.1f} us  {20*n*d/ms/1e6:.1f} GB/s (20nd)", flush=True)
    rng = np.random.default_rng(0)
    R = 1000
    rel = torch.ones(R, d, device=dev)
    inv_rel = torch.ones(R, d, device=dev)
    rg, irg = torch.empty(R, d, device=dev), torch.empty(R, d, device=dev)
    for (B, C) in ((1000, 1), (10000, 10), (50000, 50)):
        uniq, edges, dn, sn = O.make_batch(rng, rows, R, B, C, 1000)
        U = len(uniq)
        tu, te, tdn, tsn = (torch.from_numpy(x).to(dev) for x in (uniq, edges, dn, sn))
        for prec, name in ((ops.PREC_BF16X3, "bf16x3"), (ops.PREC_BF16, "bf16"), (ops.PREC_FP32, "fp32")):
            if prec == ops.PREC_FP32 and B > 10000:
                continue
            fn = lambda: ops.train_step(ctx, ops.COMPLEX, table, state, tu, te, rel, inv_rel, tdn, tsn, 0.1, ops.REDUCTION_SUM, prec, rel_grad=rg,
                                        inv_rel_grad=irg)
            ms = timeit(fn, iters=10 if prec != ops.PREC_FP32 else 3)
            print(f"train_step B={B} C={C} U={U} {name}: {ms:.3f} ms  {B/ms*1e3/1e6:.2f} M edges/s  alg-HBM {16*U*d/ms/1e6:.0f} GB/s  "
                  f"alg-TF {12*B*1000*d/ms/1e9:.1f}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gemm", "all"):
        run_gemm_probe()
    if what in ("time", "all"):
        run_timing()
