#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -6 gpurun_out/r2e_pytest.log
for v in 1; do
MB_ROW_BULK=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r2e_bench_$v.json 2> gpurun_out/r2e_bench_$v.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/r2e_bench_$v.json') if l.startswith('{')][0])
print('bulk=$v', j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['config']['stage_ms'])
PY
done
