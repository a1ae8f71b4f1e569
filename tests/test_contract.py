"""Repository contracts that need no GPU: the product never reaches into oracle/, and the reference arm of bench.py prints the
JSON line the driver expects."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_touches_the_oracle():
    offenders = []
    for base, _, files in os.walk(os.path.join(ROOT, "marius_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "oracle/" in text and f.endswith(".py"):
                    offenders.append(os.path.join(base, f))
    assert offenders == []
    # bench.py: only the CPU-reference leg imports oracle/ (inside cpu_reference_run)
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def run_ours("):src.index("def main(")]
    assert "oracle" not in body.replace("oracle/", "")


def test_reference_arm_json_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--ref-batch", "2000",
           "--ref-nodes", "20000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:] + out.stderr[-2000:]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "edges/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["metric"].startswith("edges/sec (gather+score+update) at d=400, 1000 negs")
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["steps"] == 2 and "workload" in j["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
