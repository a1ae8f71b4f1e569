"""tests/link/run_link_test.py -- drives oracle/_ref/liblink_test.so (tests/link/link_test.cpp) in its own process: the reference's
GraphModelStorage + DataLoader::loadGPUParameters / updateEmbeddings over a Storage subclass that calls the C ABI, on cuda:0, against the
numpy oracle applying the same batches.  Prints LINK_TEST OK / FAIL."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch  # noqa: F401  (libtorch first)

from oracle import marius_oracle as O


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblink_test.so"), mode=C.RTLD_LOCAL)
    lib.link_last_error.restype = C.c_char_p
    F, I = C.POINTER(C.c_float), C.POINTER(C.c_int64)
    lib.link_train_loop.argtypes = [F, F, C.c_int64, C.c_int, C.c_int, F, F, C.c_int, I, I, I, C.c_int64, I, I, C.c_int, C.c_int, C.c_float, F]
    fp = lambda a: a.ctypes.data_as(F)
    ip = lambda a: a.ctypes.data_as(I)
    rng = np.random.default_rng(3)
    num_nodes, R, B, Cc, N, d, steps = 6000, 5, 512, 2, 256, 64, 3
    table = rng.uniform(-0.3, 0.3, (num_nodes, d)).astype(np.float32)
    state = rng.uniform(0.01, 0.1, (num_nodes, d)).astype(np.float32)  # (a trained table: at state 0 the first Adagrad step is -lr * sign(g),
    #                                                                       which is not a continuous function of the gradient)
    rel = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    inv = rng.uniform(-1, 1, (R, d)).astype(np.float32)
    batches = [O.make_batch(rng, num_nodes, R, B, Cc, N) for _ in range(steps)]
    uniq = np.concatenate([b[0] for b in batches])
    off = np.zeros(steps + 1, np.int64)
    off[1:] = np.cumsum([len(b[0]) for b in batches])
    edges = np.ascontiguousarray(np.stack([b[1] for b in batches]))
    dn = np.ascontiguousarray(np.stack([b[2] for b in batches]))
    sn = np.ascontiguousarray(np.stack([b[3] for b in batches]))
    got_t, got_s = table.copy(), state.copy()
    losses = np.zeros(steps, np.float32)
    rc = lib.link_train_loop(fp(got_t), fp(got_s), num_nodes, d, R, fp(rel), fp(inv), steps, ip(uniq), ip(off), ip(edges), B, ip(dn), ip(sn), Cc, N, 0.1, fp(losses))
    if rc != 0:
        print("LINK_TEST FAIL:", lib.link_last_error().decode())
        return 1
    exp_t, exp_s = table.copy(), state.copy()
    exp_losses = []
    for (u, e, dnn, snn) in batches:
        res = O.train_step_on_table(O.COMPLEX, exp_t, exp_s, u, e, rel, inv, dnn, snn, 0.1, O.REDUCTION_SUM, acc=np.float64)
        exp_losses.append(float(res.loss))
    err = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
    et, es = err(got_t, exp_t), err(got_s, exp_s)
    el = max(abs(float(a) - b) / abs(b) for a, b in zip(losses, exp_losses))
    conv = lib.link_error_conventions()
    ok = et < 1e-4 and es < 1e-4 and el < 1e-4 and conv == 3
    print(f"LINK_TEST {'OK' if ok else 'FAIL'} table_err={et:.2e} state_err={es:.2e} loss_err={el:.2e} runtime_errors_through_the_reference_facade={conv}/3")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
