#!/usr/bin/env python
"""bench.py -- edges/sec through the Marius per-batch embedding hot path (gather + score fwd/bwd + Adagrad update).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref) on the host cores

A "step" is one pass of the hot path over one batch of synthetic edges (default 50 000 positives): ComplEx, d=400, 1000 negatives
per chunk of 1000 positives, both-side corruption, SoftmaxCE-SUM, sparse Adagrad lr 0.1 (BASELINE.json configs[1] shape; the
table is sized to fit one B200 WITH its Adagrad state).  Negative sampling and unique-id mapping are
inputs: batches are pre-built and excluded from every timed region, for both arms.

One JSON line on stdout (rank 0).  `value` = edges/s with the batch indices already resident in HBM; `e2e` = the same
step through the host-buffer C-ABI call (pinned host index tensors in, loss out, every step); `roofline` = the dominant
kernel of the step timed live with CUDA events on the launching stream; `cpu_baseline` = the reference CPU path on a
bounded sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "edges/sec (gather+score+update) at d=400, 1000 negs"
UNIT = "edges/s"
D, NEG, CHUNK = 400, 1000, 1000  # embedding dim, negatives per chunk, positives per chunk
NUM_REL = 1000
LR = 0.1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return dict(hbm_gbs=float(j["hbm_gbs"]), bf16_tflops=float(j["bf16_tflops"]),
                        bf16_tflops_sustained=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), source="measured")
        except Exception:
            pass
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def synthetic_batch(rng, src, dst, relid, src_negs, dst_negs):
    """Batch assembly of the synthetic inputs on the host (numpy): unique-id mapping over cat(src, dst, src_negs, dst_negs), the
    order DataLoader::edgeSample uses (dataloader.cpp:398-461).  Returns (unique ids, local edges [B,3], local dst_negs, local src_negs)."""
    B, (C, N) = src.shape[0], src_negs.shape
    uniq, inv = np.unique(np.concatenate([src, dst, src_negs.reshape(-1), dst_negs.reshape(-1)]), return_inverse=True)
    uniq, inv = uniq.astype(np.int64), inv.astype(np.int64).reshape(-1)
    edges = np.ascontiguousarray(np.stack([inv[:B], relid, inv[B:2 * B]], axis=1))
    return uniq, edges, np.ascontiguousarray(inv[2 * B + C * N:].reshape(C, N)), np.ascontiguousarray(inv[2 * B:2 * B + C * N].reshape(C, N))


def make_batches(rng, num_nodes, n_batches, B):
    """Uniform edges and uniform negatives over the whole table (negative.cpp:342)."""
    C = max(B // CHUNK, 1)
    out = []
    for _ in range(n_batches):
        src = rng.integers(0, num_nodes, size=B, dtype=np.int64)
        dst = rng.integers(0, num_nodes, size=B, dtype=np.int64)
        relid = rng.integers(0, NUM_REL, size=B, dtype=np.int64)
        sn = rng.integers(0, num_nodes, size=(C, NEG), dtype=np.int64)
        dn = rng.integers(0, num_nodes, size=(C, NEG), dtype=np.int64)
        out.append(synthetic_batch(rng, src, dst, relid, sn, dn))
    return out, C


def make_sharded_batches(rng, rows_per_rank, rank, world, n_batches, B):
    """Bucket-wise routing: sources and both negative pools come from the rank's own partition, destinations are
    uniform over ALL partitions, so (world-1)/world of the destination rows -- ~22 % of a batch's unique rows at 8 GPUs -- cross
    NVLink.  Returns [(unique GLOBAL ids sorted, edges local, dst_negs local, src_negs local)]."""
    C = max(B // CHUNK, 1)
    lo = rank * rows_per_rank
    out = []
    for _ in range(n_batches):
        src = rng.integers(lo, lo + rows_per_rank, size=B, dtype=np.int64)
        dst = rng.integers(0, world * rows_per_rank, size=B, dtype=np.int64)
        relid = rng.integers(0, NUM_REL, size=B, dtype=np.int64)
        sn = rng.integers(lo, lo + rows_per_rank, size=(C, NEG), dtype=np.int64)
        dn = rng.integers(lo, lo + rows_per_rank, size=(C, NEG), dtype=np.int64)
        out.append(synthetic_batch(rng, src, dst, relid, sn, dn))
    return out, C


# ------------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  NVML is polled from a thread
    every few ms (the nvidia-smi CLI needs ~0.2 s to start, longer than a short timed region); nvidia-smi -lms is the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("hw_power_brake", 0x80))

    def __init__(self, gpu_index, period_s=0.004):
        self.gpu = gpu_index
        self.period = period_s
        self.proc = None
        self.lines = []
        self.sm, self.reasons, self.mx = [], set(), None
        self.nvml = None
        self._stop = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except TypeError:
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.h = self._nvml_handle()
            self.mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for name, bit in self.BITS:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=2)
            if not self.sm:
                return dict(sm_mhz=None, sm_max_mhz=self.mx, reasons=["no samples"])
            sm = sorted(self.sm)
            return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.mx, reasons=sorted(self.reasons), samples=len(sm), source="nvml")
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")


# ------------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(steps, warmup, B, num_nodes, seed=0):
    """Times the reference's CPU hot loop (InMemory::indexRead x2 -> Model::train_batch -> indexAdd x2) on the host cores.
    Uses oracle/_ref (the unmodified reference C++) when it was built, else the numpy oracle port."""
    from oracle import marius_oracle as O
    from oracle import ref_lib as R

    rng = np.random.default_rng(seed)
    table = rng.uniform(-0.1, 0.1, (num_nodes, D)).astype(np.float32)
    state = np.zeros((num_nodes, D), np.float32)
    batches, C = make_batches(rng, num_nodes, max(steps, 1), B)
    if R.available():
        # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must use all host cores whatever launched it
        R.set_num_threads(os.cpu_count() or 1)
        uniq = np.concatenate([b[0] for b in batches])
        off = np.zeros(len(batches) + 1, np.int64)
        off[1:] = np.cumsum([len(b[0]) for b in batches])
        edges = np.ascontiguousarray(np.stack([b[1] for b in batches]))
        dn = np.ascontiguousarray(np.stack([b[2] for b in batches]))
        sn = np.ascontiguousarray(np.stack([b[3] for b in batches]))
        cores = R.num_threads()
        secs = R.train_loop(O.COMPLEX, D, NUM_REL, table, state, uniq, off, edges, dn, sn, LR, O.REDUCTION_SUM, warmup_batches=min(warmup, len(batches)))
        kind = "reference"
    else:
        rel = np.zeros((NUM_REL, D), np.float32)
        rel[:, : D // 2] = 1
        inv = rel.copy()
        t0 = time.perf_counter()
        for (u, e, dnn, snn) in batches:
            O.train_step_on_table(O.COMPLEX, table, state, u, e, rel, inv, dnn, snn, LR, O.REDUCTION_SUM)
        secs = time.perf_counter() - t0
        cores = os.cpu_count() or 1
        kind = "port"
    n = len(batches)
    return dict(value=n * B / secs, ms_per_step=secs / n * 1e3, cores=cores, kind=kind,
                sample=f"{n} batches x {B} edges (C={C}, N={NEG}, d={D}, ComplEx), host table {num_nodes} rows", C=C)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B = args.ref_batch
    # bounded sample: the CPU path needs ~0.7 s (16 cores) to ~2.5 s (8 cores) per 50 000-edge batch, so at most --ref-max-steps batches are
    # timed (after at most 2 warm-up batches) whatever K / W ask for; throughput per batch is steady, the line says how many were timed
    steps, warmup = max(1, min(args.steps, args.ref_max_steps)), min(args.warmup, 2)
    res = cpu_reference_run(steps, warmup, B, args.ref_nodes)
    line = dict(metric=METRIC, value=res["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=res["ms_per_step"],
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=f"ComplEx d={D}, {NEG} negatives, batch {B} ({res['C']} chunks), both-side corruption, SoftmaxCE-SUM, Adagrad; "
                                     f"reference CPU path (InMemory host table {args.ref_nodes} rows)", timing="wall clock, host only"),
                cpu_baseline=dict(value=res["value"], unit=UNIT, cores=res["cores"], kind=res["kind"], sample=res["sample"]),
                e2e=dict(value=res["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------ the other shapes SURVEY.md 8d names
def make_batches_shape(rng, num_nodes, n_batches, B, C, N, R):
    out = []
    for _ in range(n_batches):
        src = rng.integers(0, num_nodes, size=B, dtype=np.int64)
        dst = rng.integers(0, num_nodes, size=B, dtype=np.int64)
        relid = rng.integers(0, R, size=B, dtype=np.int64)
        sn = rng.integers(0, num_nodes, size=(C, N), dtype=np.int64)
        dn = rng.integers(0, num_nodes, size=(C, N), dtype=np.int64)
        out.append(synthetic_batch(rng, src, dst, relid, sn, dn))
    return out


def cpu_reference_shape(kind, d, R, batches, num_nodes, seed=3):
    """The reference's CPU hot loop on the given batches (bounded sample); edges/s, or None when oracle/_ref is not built."""
    from oracle import marius_oracle as O
    from oracle import ref_lib as Rf

    if not Rf.available():
        return None
    Rf.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(seed)
    table = rng.uniform(-0.1, 0.1, (num_nodes, d)).astype(np.float32)
    state = np.zeros((num_nodes, d), np.float32)
    uniq = np.concatenate([b[0] for b in batches])
    off = np.zeros(len(batches) + 1, np.int64)
    off[1:] = np.cumsum([len(b[0]) for b in batches])
    edges = np.ascontiguousarray(np.stack([b[1] for b in batches]))
    dn = np.ascontiguousarray(np.stack([b[2] for b in batches]))
    sn = np.ascontiguousarray(np.stack([b[3] for b in batches]))
    okind = {"distmult": O.DISTMULT, "complex": O.COMPLEX}[kind]
    secs = Rf.train_loop(okind, d, R, table, state, uniq, off, edges, dn, sn, LR, O.REDUCTION_SUM, warmup_batches=1)
    B = edges.shape[1]
    return dict(value=len(batches) * B / secs, unit=UNIT, cores=Rf.num_threads(), kind="reference", sample=f"{len(batches)} batches x {B} edges, host table {num_nodes} rows")


EXTRA_SHAPES = [
    # name, decoder, nodes (None = the bench table), R, d, B, C, N, cpu batches
    ("FB15k-237 sizes: DistMult d=100, 100 negatives, batch 1000 (BASELINE configs[0])", "distmult", 14541, 237, 100, 1000, 1, 100, 20),
    ("DistMult d=400, 1000 negatives, batch 50000", "distmult", None, NUM_REL, D, 50000, 50, NEG, 2),
    ("ComplEx d=400, 1000 negatives, batch 1000 (Marius default batch)", "complex", None, NUM_REL, D, 1000, 1, NEG, 20),
    ("ComplEx d=400, 1000 negatives, batch 10000", "complex", None, NUM_REL, D, 10000, 10, NEG, 6),
]


def run_extra_shapes(ops, ctx, dev, prec, table, state, steps, cpu_nodes, with_cpu):
    """Device-timed edges/s of the fused step at the other shapes (indices resident in HBM, fresh random rows of a table >> L2 every
    step -- the FB15k-237-sized table fits L2, as it does in the reference), each with its own bounded CPU-reference number."""
    import torch

    out = []
    for name, kind, nodes, R, d, B, C, N, cpu_batches in EXTRA_SHAPES:
        rng = np.random.default_rng(77)
        if nodes is None:
            t, st, n_nodes = table, state, table.size(0)
        else:
            n_nodes = nodes
            t = torch.empty((n_nodes, d), dtype=torch.float32, device=dev).uniform_(-0.1, 0.1)
            st = torch.zeros((n_nodes, d), dtype=torch.float32, device=dev)
        okind = {"distmult": ops.DISTMULT, "complex": ops.COMPLEX}[kind]
        rel = torch.ones((R, d), device=dev)
        if kind == "complex":
            rel[:, d // 2:] = 0
        inv_rel = rel.clone()
        rg, irg = torch.empty_like(rel), torch.empty_like(rel)
        loss = torch.zeros(1, device=dev)
        W = 3
        batches = make_batches_shape(rng, n_nodes, W + steps, B, C, N, R)
        res = [tuple(torch.from_numpy(x).to(dev) for x in b) for b in batches]
        for i in range(W):
            ops.train_step(ctx, okind, t, st, *res[i][:2], rel, inv_rel, res[i][2], res[i][3], LR, ops.REDUCTION_SUM, prec, loss=loss, rel_grad=rg, inv_rel_grad=irg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(W, W + steps):
            ops.train_step(ctx, okind, t, st, *res[i][:2], rel, inv_rel, res[i][2], res[i][3], LR, ops.REDUCTION_SUM, prec, loss=loss, rel_grad=rg, inv_rel_grad=irg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        entry = dict(workload=name, value=B / (ms * 1e-3), unit=UNIT, ms_per_step=ms, steps=steps, unique_rows_per_step=float(np.mean([len(b[0]) for b in batches])))
        if with_cpu:
            try:
                cb = make_batches_shape(np.random.default_rng(78), min(cpu_nodes, n_nodes), cpu_batches, B, C, N, R)
                entry["cpu_baseline"] = cpu_reference_shape(kind, d, R, cb, min(cpu_nodes, n_nodes))
            except Exception as ex:
                entry["cpu_baseline"] = dict(value=None, kind="unavailable", sample=str(ex)[:200])
        out.append(entry)
        del res
    return out


# ------------------------------------------------------------------------------------------------------ buffered table (BASELINE configs[2] shape)
def run_buffered(ops, dev, prec, num_partitions=16, capacity=8, partition_rows=250_000, edges_per_bucket=200_000, B=50_000, seed=5, modes=(True, False)):
    """One epoch over a table that does NOT fit the buffer: `num_partitions` partitions in a backing file, `capacity` of them resident in
    an HBM slab (marius_b200.host.PartitionBuffer, embeddings + Adagrad state), BETA ordering (marius_b200.ordering), DistMult d=400,
    1000 negatives drawn from the resident partitions, `edges_per_bucket` synthetic edges per edge bucket in batches of B through the
    raw-edge step.  edges/s over the whole epoch INCLUDING every swap; with the asynchronous swap engine (LookaheadBlock /
    AsyncWriteBlock) and with synchronous swaps."""
    import shutil
    import tempfile

    import torch

    from marius_b200 import host, ordering

    d, C, N, R = D, max(B // CHUNK, 1), NEG, NUM_REL
    total = num_partitions * partition_rows
    need = 2 * total * d * 4
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need + (8 << 30) else tempfile.gettempdir()
    if shutil.disk_usage(base).free < need + (2 << 30):
        return dict(skipped=f"not enough space under {base} for {need / 1e9:.1f} GB of partition files")
    tmp = tempfile.mkdtemp(prefix="mb_buffered_", dir=base)
    out = dict(workload=f"BASELINE configs[2] shape, scaled: DistMult d={d}, {N} negatives, {num_partitions} partitions x {partition_rows} rows "
                        f"({total * d * 4 / 1e9:.1f} GB embeddings + as much Adagrad state in backing files under {base}), buffer capacity {capacity} "
                        f"partitions in HBM, BETA ordering, {edges_per_bucket} edges per edge bucket in batches of {B}", unit=UNIT)
    try:
        rng = np.random.default_rng(seed)
        block = rng.uniform(-0.1, 0.1, (partition_rows, d)).astype(np.float32)
        zeros = np.zeros((partition_rows, d), np.float32)
        f_emb, f_state = os.path.join(tmp, "embeddings.bin"), os.path.join(tmp, "embeddings_state.bin")
        with open(f_emb, "wb") as fe, open(f_state, "wb") as fs:
            for _ in range(num_partitions):
                fe.write(block.tobytes())
                fs.write(zeros.tobytes())
        del block, zeros
        states, buckets = ordering.beta_ordering(num_partitions, capacity, seed)
        out["buffer_states"] = len(states)
        out["swaps"] = len(states) - 1
        n_batches = max(1, edges_per_bucket // B)
        edges_total = sum(len(b) for b in buckets) * n_batches * B
        rel = torch.ones((R, d), device=dev)
        inv_rel = rel.clone()
        rels = torch.stack([rel, inv_rel])
        rel_states, rel_grads = torch.zeros_like(rels), torch.empty_like(rels)
        for prefetching in modes:
            ctx = ops.Context(dev.index)
            emb = host.storage.PartitionBuffer(capacity, num_partitions, 1, partition_rows, d, total, f_emb, prefetching, dev)
            st = host.storage.PartitionBuffer(capacity, num_partitions, 1, partition_rows, d, total, f_state, prefetching, dev)
            order = [torch.tensor(s_) for s_ in states]
            emb.setBufferOrdering(order)
            st.setBufferOrdering(order)
            emb.load()
            st.load()
            torch.cuda.synchronize()
            brng = np.random.default_rng(seed + 1)
            bi = 0
            prev = None
            swap_s = 0.0
            t0 = time.perf_counter()
            for si, bks in enumerate(buckets):
                m = emb.getGlobalToLocalMap(True)
                slot_of = {p_: int(m[p_ * partition_rows].item()) // partition_rows for p_ in states[si]}
                table, state = emb.bufferTensor(), st.bufferTensor()
                for (ps, pd) in bks:
                    for _ in range(n_batches):
                        e = np.empty((B, 3), np.int64)
                        e[:, 0] = slot_of[ps] * partition_rows + brng.integers(0, partition_rows, B)
                        e[:, 1] = brng.integers(0, R, B)
                        e[:, 2] = slot_of[pd] * partition_rows + brng.integers(0, partition_rows, B)
                        cur = ops.train_step_edges_host_async(ctx, ops.DISTMULT, table, state, torch.from_numpy(e), capacity * partition_rows, C, N, 99, bi, rels[0],
                                                              rels[1], LR, ops.REDUCTION_SUM, prec, rel_grad=rel_grads[0], inv_rel_grad=rel_grads[1])
                        ops.dense_adagrad_step(rels, rel_states, rel_grads, LR)
                        if prev is not None:
                            ops.train_step_host_wait(ctx, prev[0])
                        prev = cur
                        bi += 1
                if emb.hasSwap():
                    ts = time.perf_counter()
                    emb.performNextSwap()
                    st.performNextSwap()
                    swap_s += time.perf_counter() - ts
            if prev is not None:
                ops.train_step_host_wait(ctx, prev[0])
            torch.cuda.synchronize()
            secs = time.perf_counter() - t0
            emb.unload(True)
            st.unload(True)
            key = "async_swaps" if prefetching else "sync_swaps"
            out[key] = dict(value=edges_total / secs, seconds=secs, host_seconds_blocked_in_swaps=swap_s, edges=edges_total, batches=bi)
            del emb, st, ctx
        if "async_swaps" in out:
            out["value"] = out["async_swaps"]["value"]
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------------ parity (checker, never timed)
PARITY_TOL = 1e-4


def _errs(a, b):
    """(max|a-b| / max|b|,  max elementwise |a-b| / max(|b|, tau)) with tau = rms(b) -- the tau SURVEY.md 8d asks to be stated."""
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    if b.size == 0:
        return 0.0, 0.0
    diff = np.abs(a - b)
    tau = max(float(np.sqrt(np.mean(b * b))), 1e-30)
    return float(diff.max() / max(np.abs(b).max(), 1e-30)), float((diff / np.maximum(np.abs(b), tau)).max())


def parity_single(ops, ctx, dev, prec, B, rows=2_000_000, seed=11, rel_tables=None):
    """One bench-shape batch (B positives, C = B/1000 chunks, 1000 negatives, ComplEx d=400, both sides) on a `rows`-row table, through the
    C ABI (mb_edge_sample -> mb_gather_rows -> mb_decoder_forward -> mb_train_step) against oracle/_ref = the reference's own
    map_tensors / InMemory::indexRead / Model::forward_lp / Model::train_batch (model.cpp:290-333, storage.cpp:606-673).  Outside every timed region."""
    import torch

    from oracle import marius_oracle as O
    from oracle import ref_lib as R

    if not R.available():
        return dict(checked=False, why="oracle/_ref not built")
    R.set_num_threads(os.cpu_count() or 1)
    C = max(B // CHUNK, 1)
    rng = np.random.default_rng(seed)
    table_h = rng.uniform(-0.1, 0.1, (rows, D)).astype(np.float32)
    state_h = (rng.uniform(0, 1, (rows, D)) < 0.25).astype(np.float32) * rng.uniform(0, 0.5, (rows, D)).astype(np.float32)  # part of the rows already trained
    if rel_tables is None:
        rel_h = rng.uniform(-1, 1, (NUM_REL, D)).astype(np.float32)
        inv_h = rng.uniform(-1, 1, (NUM_REL, D)).astype(np.float32)
    else:
        rel_h, inv_h = rel_tables
    src = rng.integers(0, rows, size=B, dtype=np.int64)
    dst = rng.integers(0, rows, size=B, dtype=np.int64)
    relid = rng.integers(0, NUM_REL, size=B, dtype=np.int64)
    sn = rng.integers(0, rows, size=(C, NEG), dtype=np.int64)
    dn = rng.integers(0, rows, size=(C, NEG), dtype=np.int64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = dict(checked=True, tol=PARITY_TOL, shape=f"ComplEx d={D} B={B} C={C} N={NEG}, table {rows} rows", oracle="oracle/_ref (reference C++)",
               tau="elementwise error = |a-b| / max(|b|, rms(b)); global error = max|a-b| / max|b|")
    # (1) unique-id mapping: bit-exact
    raw_edges = np.ascontiguousarray(np.stack([src, relid, dst], axis=1))
    uq_d, num_d, e_loc, s_loc, d_loc = ops.edge_sample(ctx, t(raw_edges), t(sn), t(dn), rows)
    U = int(num_d.item())
    uq_ref, inv = R.map_tensors(np.concatenate([src, dst, sn.reshape(-1), dn.reshape(-1)]))
    e_ref = np.ascontiguousarray(np.stack([inv[:B], relid, inv[B:2 * B]], axis=1))
    sn_ref = np.ascontiguousarray(inv[2 * B:2 * B + C * NEG].reshape(C, NEG))
    dn_ref = np.ascontiguousarray(inv[2 * B + C * NEG:].reshape(C, NEG))
    out["unique_ids_bit_exact"] = bool(U == len(uq_ref) and np.array_equal(uq_d[:U].cpu().numpy(), uq_ref) and np.array_equal(e_loc.cpu().numpy(), e_ref)
                                       and np.array_equal(s_loc.cpu().numpy(), sn_ref) and np.array_equal(d_loc.cpu().numpy(), dn_ref))
    out["unique_rows"] = U
    uq = uq_d[:U].contiguous()
    # (2) gathered rows: bit-exact
    table, state = t(table_h), t(state_h)
    emb_d, st_d = ops.gather_rows(table, uq), ops.gather_rows(state, uq)
    emb_ref, st_ref = R.index_read(table_h, uq_ref), R.index_read(state_h, uq_ref)
    out["gathered_rows_bit_exact"] = bool(np.array_equal(emb_d.cpu().numpy(), emb_ref) and np.array_equal(st_d.cpu().numpy(), st_ref))
    # (3) scores, loss, deltas: the reference's Model::train_batch on the gathered rows
    ref = R.train_batch(O.COMPLEX, emb_ref, st_ref, e_ref, rel_h, inv_h, dn_ref, sn_ref, LR, O.REDUCTION_SUM)
    rel, inv_rel = t(rel_h), t(inv_h)
    pos, neg, ipos, ineg = ops.decoder_forward(ctx, ops.COMPLEX, emb_d, e_loc, rel, inv_rel, d_loc, s_loc, prec)
    for name, a, b in (("pos", pos, ref["pos"]), ("neg", neg, ref["neg"]), ("inv_pos", ipos, ref["inv_pos"]), ("inv_neg", ineg, ref["inv_neg"])):
        g, e = _errs(a.cpu().numpy(), b)
        out[f"{name}_err"] = g
        out[f"{name}_err_elementwise"] = e
    del pos, neg, ipos, ineg
    # (4) Model::train_batch on the gathered rows: loss, node gradients, Adagrad deltas, relation gradients
    tb = ops.train_batch(ctx, ops.COMPLEX, emb_d, st_d, e_loc, rel, inv_rel, d_loc, s_loc, LR, ops.REDUCTION_SUM, prec)
    torch.cuda.synchronize()
    out["loss_err"] = abs(float(tb["loss"].item()) - float(ref["loss"][0])) / abs(float(ref["loss"][0]))
    out["grad_err"], out["grad_err_elementwise"] = _errs(tb["grad"].cpu().numpy(), ref["grad"])
    out["delta_s_err"], out["delta_s_err_elementwise"] = _errs(tb["delta_s"].cpu().numpy(), ref["delta_s"])
    # delta_e = -lr * g / (sqrt(s + g^2) + 1e-10) is DISCONTINUOUS in g where the state is 0 (first Adagrad step of a row: -lr * sign(g)), so an
    # element whose reference gradient is within rounding noise of 0 can legitimately come out with the other sign.  Those elements
    # (|g_ref| < 1e-3 rms(g_ref), counted below) are compared through g and delta_s only; every other element must match.
    g_ref = ref["grad"]
    cond = np.abs(g_ref) >= 1e-3 * float(np.sqrt(np.mean(g_ref.astype(np.float64) ** 2)))
    de = tb["delta_e"].cpu().numpy()
    out["delta_e_err"], out["delta_e_err_elementwise"] = _errs(de[cond], ref["delta_e"][cond])
    out["delta_e_ill_conditioned_elements"] = int((~cond).sum())
    out["delta_e_sign_flips_among_them"] = int((np.sign(de[~cond]) != np.sign(ref["delta_e"][~cond])).sum())
    out["delta_e_elements"] = int(cond.size)
    out["rel_grad_err"] = _errs(tb["rel_grad"].cpu().numpy(), ref["rel_grad"])[0]
    out["inv_rel_grad_err"] = _errs(tb["inv_rel_grad"].cpu().numpy(), ref["inv_rel_grad"])[0]
    # (5) the fused step on the table (mb_train_step: gather fused away, Adagrad + scatter fused into the update kernel) must leave exactly
    # the rows `emb + delta_e`, `state + delta_s` of (4) behind, bit for bit, and report the same loss / relation gradients
    rg, irg = torch.empty_like(rel), torch.empty_like(inv_rel)
    loss = ops.train_step(ctx, ops.COMPLEX, table, state, uq, e_loc, rel, inv_rel, d_loc, s_loc, LR, ops.REDUCTION_SUM, prec, rel_grad=rg, inv_rel_grad=irg)
    torch.cuda.synchronize()
    new_e, new_s = ops.gather_rows(table, uq), ops.gather_rows(state, uq)
    out["fused_step_equals_batch_step_bit_exact"] = bool(torch.equal(new_e, emb_d + tb["delta_e"]) and torch.equal(new_s, st_d + tb["delta_s"])
                                                         and torch.equal(loss.reshape(-1), tb["loss"].reshape(-1)) and torch.equal(rg, tb["rel_grad"])
                                                         and torch.equal(irg, tb["inv_rel_grad"]))
    new_e, new_s = new_e.cpu().numpy(), new_s.cpu().numpy()
    out["updated_rows_err"], out["updated_rows_err_elementwise"] = _errs(new_e[cond], (emb_ref + ref["delta_e"])[cond])
    out["updated_state_err"], out["updated_state_err_elementwise"] = _errs(new_s, st_ref + ref["delta_s"])
    del tb
    # untouched rows stay bit-identical
    mask = np.ones(rows, bool)
    mask[uq_ref] = False
    probe = np.flatnonzero(mask)[:: max(1, rows // 4096)]
    out["untouched_rows_bit_exact"] = bool(np.array_equal(table[torch.from_numpy(probe).to(dev)].cpu().numpy(), table_h[probe]))
    keys = [k for k in out if k.endswith("_err")]
    out["max_err"] = max(out[k] for k in keys)
    out["ok"] = bool(out["unique_ids_bit_exact"] and out["gathered_rows_bit_exact"] and out["untouched_rows_bit_exact"]
                     and out["fused_step_equals_batch_step_bit_exact"] and out["max_err"] <= PARITY_TOL)
    del table, state
    return out


def parity_sharded(ops, ctx, dev, prec, rank, world, rows=65536, B=2000, seed=23, overlap=True):
    """N > 1: one step of the sharded path at d=400 / 1000 negatives per chunk inside the bench's process group, on separate small
    shards.  Every rank draws its batch over the WHOLE global id space (overlap=True: the batches of different ranks share rows), all
    ranks step once, and every rank compares ITS shard with the expectation built from oracle/_ref gradients of every rank's batch on
    the pre-step table, applied in the order the sharded step defines: the owner's own contribution first, then the
    senders' in rank order, each as one sparse-Adagrad step (batch.cpp:62-79)."""
    import torch
    import torch.distributed as dist

    from marius_b200.dist import PeerShardedTable
    from oracle import marius_oracle as O
    from oracle import ref_lib as R

    if not R.available():
        return dict(checked=False, why="oracle/_ref not built")
    R.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    C = max(B // CHUNK, 1)
    total = rows * world
    rng = np.random.default_rng(seed)
    full = rng.uniform(-0.1, 0.1, (total, D)).astype(np.float32)
    full_s = (rng.uniform(0, 1, (total, D)) < 0.25).astype(np.float32) * rng.uniform(0, 0.5, (total, D)).astype(np.float32)
    rel_h = rng.uniform(-1, 1, (NUM_REL, D)).astype(np.float32)
    inv_h = rng.uniform(-1, 1, (NUM_REL, D)).astype(np.float32)
    batches = []
    for r in range(world):
        brng = np.random.default_rng(seed + 100 + r)
        if overlap:
            pool = brng.integers(0, total, size=3 * B, dtype=np.int64)  # a small pool per rank + a pool shared by all ranks: plenty of shared rows
            shared = np.random.default_rng(seed + 99).integers(0, total, size=B, dtype=np.int64)
            pool = np.concatenate([pool, shared])
            pick = lambda size: pool[brng.integers(0, len(pool), size=size)]
        else:
            pick = lambda size: brng.integers(0, total // world, size=size, dtype=np.int64) * world + r
        src, dst, relid = pick(B), pick(B), brng.integers(0, NUM_REL, size=B, dtype=np.int64)
        sn, dn = pick((C, NEG)), pick((C, NEG))
        uniq, inv = O.map_tensors(np.concatenate([src, dst, sn.reshape(-1), dn.reshape(-1)]))
        edges = np.ascontiguousarray(np.stack([inv[:B], relid, inv[B:2 * B]], axis=1))
        batches.append((uniq, edges, np.ascontiguousarray(inv[2 * B + C * NEG:].reshape(C, NEG)), np.ascontiguousarray(inv[2 * B:2 * B + C * NEG].reshape(C, NEG))))
    lo, hi = rank * rows, (rank + 1) * rows
    table = torch.from_numpy(full[lo:hi].copy()).to(dev)
    state = torch.from_numpy(full_s[lo:hi].copy()).to(dev)
    pctx = ops.Context(dev.index)
    pst = PeerShardedTable(table, state, pctx, exchange_rows=2 * B + 2 * C * NEG)
    t = lambda a: torch.from_numpy(a).to(dev)
    uniq, edges, dn, sn = batches[rank]
    rel, inv_rel = t(rel_h), t(inv_h)
    rg, irg = torch.empty_like(rel), torch.empty_like(inv_rel)
    loss = pst.train_step(ops.COMPLEX, t(uniq), t(edges), rel, inv_rel, t(dn), t(sn), LR, ops.REDUCTION_SUM, prec, rel_grad=rg, inv_rel_grad=irg)
    torch.cuda.synchronize()
    dist.barrier()
    # expectation for MY shard: gradients of every rank's batch on the pre-step table (reference C++), contributions in the defined order
    exp_t, exp_s = full[lo:hi].copy(), full_s[lo:hi].copy()
    my_loss = None
    contrib = {}
    for r, (u, e, dnn, snn) in enumerate(batches):
        res = R.train_batch(O.COMPLEX, np.ascontiguousarray(full[u]), np.ascontiguousarray(full_s[u]), e, rel_h, inv_h, dnn, snn, LR, O.REDUCTION_SUM)
        if r == rank:
            my_loss = float(res["loss"][0])
        mine = (u >= lo) & (u < hi)
        contrib[r] = (u[mine] - lo, res["grad"][mine])
    order = [rank] + [r for r in range(world) if r != rank]
    seen = np.zeros(rows, np.int32)
    cond = np.ones((rows, D), bool)  # elements whose every contribution is well-conditioned (see parity_single: the Adagrad step is
    for r in order:                  # discontinuous in g where the state is 0, so |g| within rounding noise of 0 may flip the sign of delta_e)
        idx, g = contrib[r]
        seen[idx] += 1
        cond[idx] &= np.abs(g) >= 1e-3 * float(np.sqrt(np.mean(g.astype(np.float64) ** 2)) + 1e-30)
        s_new = exp_s[idx] + g * g                                 # batch.cpp:67-69
        exp_t[idx] = exp_t[idx] + (-LR * g / (np.sqrt(s_new) + np.float32(1e-10))).astype(np.float32)
        exp_s[idx] = s_new
    shared_rows = int((seen > 1).sum())
    got_t, got_s = table.cpu().numpy(), state.cpu().numpy()
    et, es = _errs(got_t[cond], exp_t[cond])[0], _errs(got_s, exp_s)[0]
    barrier_error = pst.error()
    el = abs(float(loss.item()) - my_loss) / abs(my_loss)
    remote = int(((uniq // rows) != rank).sum())
    vals = torch.tensor([et, es, el, float(remote), float(shared_rows), float(barrier_error)], device=dev, dtype=torch.float64)
    gathered = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(gathered, vals)
    per_rank = [dict(rank=i, table_err=float(g[0]), state_err=float(g[1]), loss_err=float(g[2]), remote_rows=int(g[3]), rows_updated_by_several_ranks=int(g[4]),
                     barrier_timeouts=int(g[5])) for i, g in enumerate(gathered)]
    mx = max(max(p["table_err"], p["state_err"], p["loss_err"]) for p in per_rank)
    del pst
    return dict(checked=True, tol=PARITY_TOL, shape=f"ComplEx d={D} B={B} C={C} N={NEG}, {world} shards x {rows} rows, batches over the whole id space"
                + (" (rows shared between ranks)" if overlap else " (disjoint rows per rank)"), oracle="oracle/_ref gradients + ordered owner-side Adagrad",
                per_rank=per_rank, max_err=mx,
                ok=bool(mx <= PARITY_TOL and all(p["remote_rows"] > 0 and p["barrier_timeouts"] == 0 for p in per_rank)))


# ------------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from marius_b200 import _lib, ops  # the measured arm imports nothing from oracle/

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the marius_b200 hot path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, K, W = args.batch, args.steps, args.warmup
    C = max(B // CHUNK, 1)
    prec = {"bf16x3": ops.PREC_BF16X3, "fp32": ops.PREC_FP32, "bf16": ops.PREC_BF16}[args.precision]
    # per-GPU shard of the node-partitioned table (weak scaling: rows per GPU fixed; ids below are shard-local rows)
    free, total = torch.cuda.mem_get_info()
    rows = args.nodes
    need = 2 * rows * D * 4 + (6 << 30)
    if need > free:
        rows = int((free - (6 << 30)) // (2 * D * 4))
    table = torch.empty((rows, D), dtype=torch.float32, device=dev).uniform_(-0.1, 0.1)
    state = torch.zeros((rows, D), dtype=torch.float32, device=dev)
    # the two relation tables (and their Adagrad sums / gradients) live back to back so that the dense optimizer is one launch
    rels = torch.zeros((2, NUM_REL, D), device=dev)
    rels[:, :, : D // 2] = 1.0  # ComplEx::reset (complex.cpp:21-30)
    rel, inv_rel = rels[0], rels[1]
    rel_states = torch.zeros_like(rels)
    rel_grads = torch.empty_like(rels)
    rg, irg = rel_grads[0], rel_grads[1]
    ctx = ops.Context(local)

    rng = np.random.default_rng(1000 + rank)
    n_b = K + W
    sharded = None
    peer = None
    if world > 1:
        from marius_b200.dist import OpsBackend, PeerShardedTable, ShardedTable

        host_batches, _ = make_sharded_batches(rng, rows, rank, world, n_b, B)
        if args.exchange == "peer":
            # peers' shards mapped over NVLink (CUDA IPC): the fused, graph-replayed step reads / updates remote rows directly
            peer = PeerShardedTable(table, state, ctx)
        else:
            cpu_group = dist.new_group(backend="gloo")
            sharded = ShardedTable(rows, OpsBackend(table, state, ctx, prec), cpu_group=cpu_group)
    else:
        host_batches, _ = make_batches(rng, rows, n_b, B)
    pinned = [tuple(torch.from_numpy(x).pin_memory() for x in b) for b in host_batches]
    resident = [tuple(t.to(dev) for t in b) for b in pinned]
    U_mean = float(np.mean([len(b[0]) for b in host_batches]))
    # rows of a batch that live on another rank (they cross NVLink: the embedding row in, the gradient row + its id out)
    remote_mean = float(np.mean([int(((b[0] // rows) != rank).sum()) for b in host_batches])) if world > 1 else 0.0
    # routing metadata (owner bucket sizes) depends only on the unique ids: prepared with the batches, like sampling + mapping
    routes = [sharded.make_plan(torch.from_numpy(b[0]), ids_device=r[0]) for b, r in zip(host_batches, resident)] if sharded is not None else None
    if sharded is not None:
        torch.cuda.synchronize()
    loss = torch.zeros(1, device=dev)

    dense_calls = [0]

    def dense_step():
        # Model::step(): the reference's dense Adagrad on the relation tables (optim.cpp:114-145), every batch on every rank.  The relation
        # gradients are all-reduced across ranks every `gpu_sync_interval` batches, the reference's own cadence (pipeline_gpu.cpp:52-78,
        # default 16: marius_config.py:674); between syncs the replicas of the (small, dense) relation tables drift exactly as they do there.
        dense_calls[0] += 1
        if world > 1 and dense_calls[0] % args.gpu_sync_interval == 0:
            dist.all_reduce(rel_grads)
        ops.dense_adagrad_step(rels, rel_states, rel_grads, LR)

    remote_rows = []

    def step_sharded(u, e, dn, sn, route):
        # ids -> rows -> gradients exchanged over NCCL all-to-all (marius_b200/dist.py); relation grads all-reduced inside
        out = sharded.train_step(ops.COMPLEX, u, e, rel, inv_rel, dn, sn, LR, ops.REDUCTION_SUM, route=route)
        rel_grads[0].copy_(out["rel_grad"])
        rel_grads[1].copy_(out["inv_rel_grad"])
        ops.dense_adagrad_step(rels, rel_states, rel_grads, LR)
        remote_rows.append(sharded.last_remote_rows)
        return out["loss"]

    def step_resident(i):
        u, e, dn, sn = resident[i]
        if sharded is not None:
            step_sharded(u, e, dn, sn, routes[i])
            return
        if peer is not None:
            peer.train_step(ops.COMPLEX, u, e, rel, inv_rel, dn, sn, LR, ops.REDUCTION_SUM, prec, loss=loss, rel_grad=rg, inv_rel_grad=irg)
            dense_step()
            return
        ops.train_step(ctx, ops.COMPLEX, table, state, u, e, rel, inv_rel, dn, sn, LR, ops.REDUCTION_SUM, prec, loss=loss, rel_grad=rg, inv_rel_grad=irg)
        dense_step()

    def step_host(i):
        u, e, dn, sn = pinned[i]
        if sharded is not None:
            l = step_sharded(*(t.to(dev, non_blocking=True) for t in (u, e, dn, sn)), routes[i])
            return float(l.item())  # D2H read of the step's loss
        if peer is not None:
            l = peer.train_step_host(ops.COMPLEX, u, e, rel, inv_rel, dn, sn, LR, ops.REDUCTION_SUM, prec, rel_grad=rg, inv_rel_grad=irg)
            dense_step()
            return l
        l = ops.train_step_host(ctx, ops.COMPLEX, table, state, u, e, rel, inv_rel, dn, sn, LR, ops.REDUCTION_SUM, prec, rel_grad=rg, inv_rel_grad=irg)
        dense_step()
        return l

    # ---- value: indices resident in HBM, device-timed, max over ranks
    for i in range(W):
        step_resident(i)
    if world > 1:
        # the first collective of a given size loads NCCL's kernel for it and sets its channels up (~15 ms, measured): with a sync interval
        # longer than the warm-up that one-time cost would otherwise land inside the timed steps
        dist.all_reduce(rel_grads)
    barrier()
    if world > 1:
        # The ranks leave the host barrier up to a few ms apart.  One more untimed step: its device-side flag barriers line the GPUs up, so
        # that every rank's start event marks (nearly) the same instant and the max over ranks is not inflated by host skew.
        step_resident(W - 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(K)] if world > 1 else []
    e0.record()
    for i in range(W, W + K):
        step_resident(i)
        if marks:
            marks[i - W].record()  # per-step device times (N > 1: shows which steps carry the relation all-reduce / rank skew)
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = world * K * B / (total_ms / 1e3)
    step_trace = None
    if marks:
        ts = [e0.elapsed_time(m) for m in marks]
        step_trace = [round(b_ - a_, 3) for a_, b_ in zip([0.0] + ts[:-1], ts)]

    # ---- e2e: host (pinned) index buffers in, loss out, every step; wall clock bracketed by barriers.  The steps go through the
    # asynchronous host entry point (mb_train_step_host_async): batch i+1 is enqueued -- its H2D copies queue behind batch i's kernels on
    # the same stream -- before the loss of batch i is read back, the shape of the reference's transfer / compute pipeline.  Every
    # step's inputs cross PCIe inside the timed region and every step's loss is read on the host.
    def step_host_async(i):
        u, e, dn, sn = pinned[i]
        if peer is not None:
            t = peer.train_step_host_async(ops.COMPLEX, u, e, rel, inv_rel, dn, sn, LR, ops.REDUCTION_SUM, prec, rel_grad=rg, inv_rel_grad=irg)
        else:
            t = ops.train_step_host_async(ctx, ops.COMPLEX, table, state, u, e, rel, inv_rel, dn, sn, LR, ops.REDUCTION_SUM, prec, rel_grad=rg, inv_rel_grad=irg)
        dense_step()
        return t

    for i in range(min(W, 2)):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    last_loss = 0.0
    if sharded is not None:
        for i in range(W, W + K):
            last_loss = step_host(i)
    else:
        prev = None
        for i in range(W, W + K):
            cur = step_host_async(i)
            if prev is not None:
                last_loss = ops.train_step_host_wait(ctx, prev[0])
            prev = cur
        last_loss = ops.train_step_host_wait(ctx, prev[0])
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * K * B / float(e2e_s.item())
    h2d = int(np.mean([sum(t.numel() * 8 for t in b) for b in pinned]))

    # ---- e2e from RAW edges (N = 1): the host hands over only the positive edges with global ids (24 B / edge); negatives are sampled and
    # the unique-id mapping is built on the device inside the call.  Reported next to `e2e` (whose inputs are the
    # pre-mapped index tensors the reference's Batch::to moves), with the sampling + mapping time on its own.
    e2e_raw = None
    if world == 1:
        raw_edges = []
        for _ in range(min(K, 20) + 2):
            raw_edges.append(torch.from_numpy(np.stack([rng.integers(0, rows, B), rng.integers(0, NUM_REL, B), rng.integers(0, rows, B)], axis=1).astype(np.int64)).pin_memory())
        def raw_step(i):
            return ops.train_step_edges_host_async(ctx, ops.COMPLEX, table, state, raw_edges[i], rows, C, NEG, 1234, i, rel, inv_rel, LR, ops.REDUCTION_SUM, prec,
                                                   rel_grad=rg, inv_rel_grad=irg)
        for i in range(2):
            ops.train_step_host_wait(ctx, raw_step(i)[0])
            dense_step()
        barrier()
        t0 = time.perf_counter()
        prev = None
        for i in range(2, len(raw_edges)):
            cur = raw_step(i)
            dense_step()
            if prev is not None:
                ops.train_step_host_wait(ctx, prev[0])
            prev = cur
        ops.train_step_host_wait(ctx, prev[0])
        barrier()
        raw_s = time.perf_counter() - t0
        e2e_raw = dict(value=(len(raw_edges) - 2) * B / raw_s, unit=UNIT, h2d_bytes_per_step=int(raw_edges[0].numel() * 8), d2h_bytes_per_step=4,
                       steps=len(raw_edges) - 2, note="positive edges only over PCIe; negative sampling (Philox) + unique-id mapping (radix sort) on the device inside the call")

    # ---- roofline: per-stage CUDA-event timing of the same steps (separate pass: events add launch gaps)
    ctx.profile(True)
    for i in range(W, W + K):
        step_resident(i)
    stages = ctx.profile_read()
    sample_ms = None
    if e2e_raw is not None:  # sampling + unique mapping, timed separately
        for i in range(2, min(6, len(raw_edges))):
            ops.train_step_host_wait(ctx, raw_step(i)[0])
        st2 = ctx.profile_read()
        if st2.get("negative_sampling+unique_mapping", (0, 0))[1] > 0:
            sample_ms = st2["negative_sampling+unique_mapping"][0] / st2["negative_sampling+unique_mapping"][1]
            e2e_raw["sampling_mapping_ms_per_step"] = sample_ms
    ctx.profile(False)
    pk = peaks()
    # per-STEP stage times (a stage may be several launches: the two corruption sides are pipelined on two streams)
    per_stage = {k: v[0] / K for k, v in stages.items() if v[1] > 0}
    stage_launches = {k: v[1] / K for k, v in stages.items() if v[1] > 0}
    if "gemm_dNeg" in per_stage:  # both backward contractions are reported as one stage
        per_stage["gemm_dA"] = per_stage.get("gemm_dA", 0.0) + per_stage.pop("gemm_dNeg")
        stage_launches["gemm_dA"] = stage_launches.get("gemm_dA", 0.0) + stage_launches.pop("gemm_dNeg")
    step_stage_ms = sum(v[0] for v in stages.values()) / K
    dom = max(per_stage, key=per_stage.get) if per_stage else None
    Bc = B // C
    # algorithmic fp32 flops per step: scores 2*sides*C*Bc*N*d ; backward dA + dNeg twice that
    flops = {"gemm_scores": 2 * 2 * C * Bc * NEG * D, "gemm_dA": 2 * (2 * 2 * C * Bc * D * NEG)}
    byts = {"gather_rows": 8 * U_mean * D, "segment_reduce+adagrad_update": 20 * U_mean * D}
    def ncu_traffic(kernel):
        # dram bytes per launch of the dominant kernel, from the committed ncu capture of this workload (profiles/r2_traffic.json)
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            if int(t.get("batch", -1)) == B and kernel in t and world == 1:
                return float(t[kernel]["dram_bytes_per_launch"])
        except Exception:
            pass
        return None

    roof = None
    if dom in flops:
        ach = flops[dom] / (per_stage[dom] * 1e-3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        roof = dict(bound="tensor", kernel=dom, achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak, traffic=ncu_traffic(dom),
                    launches_per_step=stage_launches.get(dom), issued_bf16_tflops=3 * ach, issued_frac=3 * ach / peak,
                    note=f"algorithmic fp32 flops (2MNK summed over the stage's launches in one step) / the stage's time per step; the kernel issues 3 bf16 "
                         f"products per fp32 product (bf16x3, needed for the 1e-4 parity bar), so frac <= 1/3 and issued_frac = 3 frac is the fraction of the peak the tensor cores actually deliver; peak = {pk['source']} sustained cuBLAS bf16")
    elif dom in byts:
        ach = byts[dom] / (per_stage[dom] * 1e-3) / 1e9
        roof = dict(bound="hbm", kernel=dom, achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"], traffic=None,
                    note=f"peak = {pk['source']} copy bandwidth")
    elif dom is not None:
        roof = dict(bound="hbm", kernel=dom, achieved=None, peak=pk["hbm_gbs"], unit="GB/s", frac=None, traffic=None, note="non-roofline stage dominant")
    step_ms = total_ms / K
    step_hbm = 16 * U_mean * D / (step_ms * 1e-3) / 1e9
    # the path is tensor-bound at fp32-equivalent accuracy: 12 B N d fp32 flops per step, each issued as 3 bf16 products
    tensor_ceiling = pk["bf16_tflops_sustained"] * 1e12 / (3 * 12 * NEG * D) * world  # edges/s if every other kernel hid under the contractions
    if roof is not None:
        roof["step_hbm_frac"] = step_hbm / pk["hbm_gbs"]
        roof["tensor_bound_ceiling_edges_per_s"] = tensor_ceiling
        roof["step_frac_of_tensor_ceiling"] = value / tensor_ceiling
        roof["hbm_bound_ceiling_edges_per_s"] = pk["hbm_gbs"] * 1e9 / (16 * U_mean * D / B) * world
        tj = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        except Exception:
            pass
        if tj is not None and int(tj.get("batch", -1)) == B and world == 1:
            roof["step_traffic_bytes"] = tj.get("step_dram_bytes")
            roof["step_traffic_over_algorithmic"] = (tj.get("step_dram_bytes") or 0) / (16 * U_mean * D) if tj.get("step_dram_bytes") else None

    # ---- CPU baseline (rank 0, N = 1): the reference path on a bounded sample on this box's host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(args.cpu_steps, 1, args.ref_batch, args.ref_nodes, seed=7)
            cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"])
        except Exception as ex:  # the baseline is reported, never required
            cpu = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="unavailable", sample=str(ex)[:200])

    # ---- the other shapes (N = 1): device-timed value + bounded CPU-reference number each
    extra = None
    if world == 1 and not args.no_extra_shapes:
        try:
            extra = run_extra_shapes(ops, ctx, dev, prec, table, state, min(K, 20), args.ref_nodes, not args.no_cpu_baseline)
        except Exception as ex:
            extra = [dict(workload="extra shapes failed", error=f"{type(ex).__name__}: {ex}"[:300])]

    # ---- buffered table: one BETA epoch over a partitioned table larger than its HBM buffer (swap engine included in the time)
    buffered = None
    if world == 1 and not args.no_buffered:
        try:
            buffered = run_buffered(ops, dev, prec, partition_rows=args.buffered_partition_rows, edges_per_bucket=args.buffered_edges_per_bucket, B=B)
        except Exception as ex:
            buffered = dict(error=f"{type(ex).__name__}: {ex}"[:300])

    # ---- parity (outside every timed region): the measured path against the reference's CPU path, at the bench shape (N = 1) or through the
    # sharded protocol inside this process group (N > 1)
    parity = None
    if not args.no_parity:
        try:
            if world == 1:
                parity = parity_single(ops, ctx, dev, prec, B, rows=min(args.ref_nodes, rows))
            else:
                parity = parity_sharded(ops, ctx, dev, prec, rank, world, overlap=not args.parity_disjoint)
        except Exception as ex:  # a failed check is reported as such, never hidden
            parity = dict(checked=False, ok=False, why=f"{type(ex).__name__}: {ex}"[:300])
            if world > 1:
                raise

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=step_ms, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=f"BASELINE configs[1] shape: ComplEx d={D}, {NEG} negatives/chunk, batch {B} = {C} chunks x {CHUNK}, both-side "
                                         f"corruption, SoftmaxCE-SUM, sparse Adagrad lr {LR}; table {rows} rows x {D} fp32 + Adagrad state resident per GPU "
                                         f"({2 * rows * D * 4 / 1e9:.0f} GB; 1e8 rows + state = 320 GB does not fit 180 GB)",
                                precision=args.precision,
                                parallelism=(f"table sharded by node partition over {world} GPUs; per batch: src + negatives local, dst uniform over all "
                                             f"partitions (~{(world - 1) / world * 25:.0f}% of a batch's rows are remote); "
                                             + ("remote rows fetched once per step over NVLink-mapped peer memory (CUDA IPC); the gradient row of every remote row is shipped to the owner's inbox and applied there (ordered sparse Adagrad, device-side flag barriers)"
                                                if peer is not None else
                                                f"rows / gradient rows exchanged by grouped NCCL send/recv, remote rows/step/rank {np.mean(remote_rows) if remote_rows else 0:.0f}")
                                             + f"; relation grads all-reduced (NCCL) every {args.gpu_sync_interval} batches (reference gpu_sync_interval)") if world > 1 else "single GPU, fused gather+score+update step",
                                l2="inputs larger than L2: every step gathers/updates a fresh uniform-random row set of a table >> 126 MB",
                                unique_rows_per_step=U_mean, step_hbm_gbs_algorithmic=step_hbm, step_hbm_frac=step_hbm / pk["hbm_gbs"],
                                **(dict(step_ms_trace_rank0=step_trace) if step_trace else {}),
                                **(dict(remote_rows_per_step_rank0=remote_mean, remote_fraction=remote_mean / U_mean,
                                        nvlink_gbs_per_gpu_each_direction=remote_mean * (D * 4 + 8) / (step_ms * 1e-3) / 1e9,
                                        nvlink_note="per remote row: 1 embedding row in (fetch) and 1 gradient row + id out (inbox); measured peer-copy peak 770 GB/s per direction")
                                   if world > 1 and peer is not None else {}),
                                stage_ms={k: round(v, 4) for k, v in per_stage.items()}, stage_sum_ms=step_stage_ms, last_loss=last_loss),
                    clocks=clocks, e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=4), gpu_launches=int(launches),
                    roofline=roof, cpu_baseline=cpu, parity=parity, e2e_raw_edges=e2e_raw, configs=extra, buffered=buffered, impl="marius_b200")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=50000,
                    help="positives per step, in chunks of 1000 (50 000 = the batch the Marius paper trains Freebase86m with; 10 000 runs ~25 %% slower "
                         "per edge because the contractions have too few tiles per SM, see DESIGN.md 6)")
    ap.add_argument("--nodes", type=int, default=40_000_000, help="table rows per GPU (with Adagrad state: 2 x rows x 1600 B)")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "fp32", "bf16"])
    ap.add_argument("--ref-nodes", type=int, default=2_000_000, help="host table rows for the CPU reference arm")
    ap.add_argument("--ref-batch", type=int, default=0, help="batch of the CPU reference arm (0 = --batch)")
    ap.add_argument("--ref-max-steps", type=int, default=24, help="--impl reference: upper bound on the timed CPU batches")
    ap.add_argument("--cpu-steps", type=int, default=3, help="batches of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-buffered", action="store_true", help="skip the buffered-table epoch (16 partitions, capacity 8, BETA ordering, swaps included)")
    ap.add_argument("--buffered-partition-rows", type=int, default=250_000)
    ap.add_argument("--buffered-edges-per-bucket", type=int, default=200_000)
    ap.add_argument("--no-extra-shapes", action="store_true", help="skip the other shapes (FB15k-237 sizes, DistMult d=400, batch 1000 / 10000)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (one bench-shape batch against the reference's CPU path, untimed)")
    ap.add_argument("--parity-disjoint", action="store_true", help="N > 1 parity: give every rank disjoint rows (no row updated by two ranks)")
    ap.add_argument("--gpu-sync-interval", type=int, default=16,
                    help="N > 1: all-reduce the dense relation gradients every this many batches (the reference's gpu_sync_interval, default 16)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = remote rows read / updated over NVLink-mapped peer memory inside the fused step; "
                         "'nccl' = rows and gradient rows exchanged with grouped NCCL send/recv (marius_b200/dist.py ShardedTable)")
    args = ap.parse_args()
    if args.ref_batch <= 0:
        args.ref_batch = args.batch
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
