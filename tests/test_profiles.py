"""The numbers bench.py reports from the committed ncu capture (roofline.traffic, step_traffic_bytes) are reproducible from the raw
launch list under profiles/ (no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_table_matches_the_raw_launch_list(tmp_path):
    csv_path = os.path.join(ROOT, "profiles", "r2_ncu_step.csv")
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_step_table.py"), csv_path, "155", "183", str(tmp_path / "step.json")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    step = json.load(open(tmp_path / "step.json"))
    assert len(step["kernels"]) == 29  # one step of the eager (MB_GRAPH=0) launch sequence
    assert abs(step["step_dram_bytes"] - traffic["step_dram_bytes"]) <= 1e-6 * traffic["step_dram_bytes"]
    by = {k["kernel"]: k for k in step["kernels"]}
    ts = by["gemm_tc_ts_kernel"]
    assert abs((ts["dram_read_mb"] + ts["dram_write_mb"]) * 1e6 - traffic["gemm_dA"]["dram_bytes_per_launch"]) <= 1e6
    # the share of the step the dominant kernel takes under ncu (serialised, cold) agrees with the bench's stage timers (overlapped, warm)
    main_us = sum(k["us"] for k in step["kernels"] if k["kernel"] in ("bulk::neg_rows_bulk_kernel<4>", "vec::edge_rows_kernel<2, 2>", "gemm_tc_group_kernel<0>",
                                                                        "loss_merge_kernel", "gemm_tc_ts_kernel", "vec::edge_backward_kernel<2, 2>",
                                                                        "vec::segment_reduce_kernel<2, 4>"))
    bench = json.loads([l for l in open(os.path.join(ROOT, "profiles", "r2_bench_n1_final.json")) if l.startswith("{")][0])
    st = bench["config"]["stage_ms"]
    crit = sum(st[k] for k in ("edge_prep+neg_gather", "gemm_scores", "loss_grad", "gemm_dA", "edge_backward", "segment_reduce+adagrad_update"))
    assert abs(ts["us"] / main_us - st["gemm_dA"] / crit) < 0.03
