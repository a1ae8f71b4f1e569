"""tools/ncu_step_table.py <launches.csv> <first_id> <last_id> [summary.json] -- per-kernel table (markdown) + traffic summary (json) of one bench step
from the long-format CSV of `ncu --replay-mode application --metrics ... --csv` (tools/gpu_r2_prof.sh)."""
import collections
import csv
import json
import sys

path, first, last = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
json_out = sys.argv[4] if len(sys.argv) > 4 else "/tmp/step_table.json"
rows = list(csv.reader(open(path)))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[h]
ki, idi, mn, mu, mv, st = (hdr.index(x) for x in ("Kernel Name", "ID", "Metric Name", "Metric Unit", "Metric Value", "Stream"))
data = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= mv:
        continue
    i = int(r[idi])
    if i < first or i > last:
        continue
    d = data.setdefault(i, {"name": r[ki].split("(")[0].replace("void ", "").replace("mb::<unnamed>::", "").replace("mb::", ""), "stream": r[st]})
    v = float(r[mv].replace(",", ""))
    u = r[mu]
    if r[mn] == "gpu__time_duration.sum":
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else (v * 1e6 if u == "s" else v))
    if r[mn].startswith("dram__bytes") or r[mn].startswith("lts__t_bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r[mn]] = v
streams = collections.Counter(d["stream"] for d in data.values())
main_stream = max(streams, key=lambda s: sum(x.get("gpu__time_duration.sum", 0) for x in data.values() if x["stream"] == s))
print("| kernel | stream | time (us) | DRAM read (MB) | DRAM write (MB) | GB/s | % of copy peak | warps active % | regs | tensor pipe % |")
print("|---|---|---|---|---|---|---|---|---|---|")
tot_us = tot_main = tot_b = 0.0
kernels = []
for i, d in data.items():
    us = d["gpu__time_duration.sum"]
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    gbs = (rd + wr) / us / 1e3
    tot_us += us
    tot_b += rd + wr
    main = d["stream"] == main_stream
    tot_main += us if main else 0
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    wa = d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
    regs = int(d.get("launch__registers_per_thread", 0))
    print(f"| `{d['name']}` | {'main' if main else 'side'} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {gbs / 65.40:.0f} | {wa:.1f} | {regs} | {tp:.1f} |")
    kernels.append(dict(kernel=d["name"], us=round(us, 1), dram_read_mb=round(rd / 1e6, 1), dram_write_mb=round(wr / 1e6, 1), gbs=round(gbs), warps_active_pct=round(wa, 1),
                        regs=regs, tensor_pct=round(tp, 1), grid=int(d.get("launch__grid_size", 0))))
print(f"\nSum: {tot_us:.0f} us of kernel time (main stream {tot_main:.0f} us), {tot_b / 1e9:.3f} GB of DRAM traffic")
json.dump(dict(step_dram_bytes=tot_b, step_kernel_us_serialised=tot_us, kernels=kernels), open(json_out, "w"), indent=1)
