#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2f_bench.json') if l.startswith('{')][0])
print(j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['config']['stage_ms'])
print('parity', j['parity']['ok'], j['parity']['max_err'])
print('raw', j['e2e_raw_edges'])
for c in j['configs'] or []: print(c)
print('cpu', j['cpu_baseline'])
PY
tail -3 gpurun_out/r2f_bench.err
