"""tools/ncu_summary.py -- turn gpurun_out/*.ncu-rep and the launch-list CSV into the tracked summaries under profiles/.
Usage: python tools/ncu_summary.py <tag> <launches.csv> <report1.ncu-rep> [report2.ncu-rep ...]"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[h]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("mb::<unnamed>::", "").replace("mb::", "")[-70:]
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    return agg


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                d[w] = (r[hdr.index(w)], units[hdr.index(w)])
        res.append(d)
    return res


def main():
    tag, lcsv, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    out = [f"# ncu summary {tag}", "", "Source: `gpurun` on one B200 (`ncu --clock-control none`).  Launch list: `python bench.py --steps 2 --warmup 3 --nodes 2000000 "
           "--no-cpu-baseline` (the bench workload: ComplEx d=400, 1000 negatives, default batch; 2e6-row table so that ncu's replay save/restore stays "
           "cheap).  `--set full` captures: `*_rows*` = `MB_GRAPH=0 python tools/ncu_workload.py 3` (same shape at batch 10 000), `*50k*` = the bench command "
           "with `MB_GRAPH=0 --steps 1` at the default batch of 50 000.",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (the bench line has the real timings).", ""]
    if lcsv and os.path.exists(lcsv):
        agg = launches(lcsv)
        ours = {k: v for k, v in agg.items() if "at::" not in k and "ise_kernel" not in k and "uniform" not in k}
        tot = sum(v[1] for v in ours.values())
        out += ["## Launch list (`--metrics gpu__time_duration.sum`), our kernels only", "", "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
        for k, v in sorted(ours.items(), key=lambda x: -x[1][1]):
            out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
        out.append("")
    for rep in reps:
        out += [f"## `--set full` capture: {os.path.basename(rep)}", ""]
        for d in raw(rep):
            out.append(f"### `{d['kernel'][:150]}`")
            out.append("")
            out.append("| metric | value |")
            out.append("|---|---|")
            for w in WANT:
                if w in d:
                    out.append(f"| {w} | {d[w][0]} {d[w][1]} |")
            out.append("")
    path = os.path.join(ROOT, "profiles", f"{tag}.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
