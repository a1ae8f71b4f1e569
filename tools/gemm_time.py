"""tools/gemm_time.py -- time the three contractions of the named shape in isolation through mb_debug_gemm."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from marius_b200 import ops

ctx = ops.Context(0)
cfg = int(os.environ.get("MB_TC_CFG", "256"))
bt, Bc, N, d = 20, 1000, 1000, 400
shapes = {"scores A.NegT (K,K)": (False, False, Bc, N, d), "dA G.Neg (K,MN)": (False, True, Bc, d, N), "dNeg GT.A (MN,MN)": (True, True, N, d, Bc)}
only = os.environ.get('GT_ONLY')
for name, (a_mn, b_mn, M, Nn, K) in shapes.items():
    if only and not name.startswith(only):
        continue
    A = torch.randn(bt, K, M, device="cuda") if a_mn else torch.randn(bt, M, K, device="cuda")
    B = torch.randn(bt, K, Nn, device="cuda") if b_mn else torch.randn(bt, Nn, K, device="cuda")
    for prec, pn in ((ops.PREC_BF16X3, "bf16x3"),):
        f = lambda: ops.debug_gemm(ctx, A, a_mn, B, b_mn, prec, cfg)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * bt * M * Nn * K
        print(f"cfg {cfg} dbg {os.environ.get('MB_TC_DEBUG','0')} {name:22s} {pn:7s} {ms*1e3:8.1f} us (incl. hi/lo split + output alloc)  {fl/ms/1e9:8.1f} TF/s algorithmic", flush=True)
