#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_host.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-shapes --no-buffered > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2j_bench.json') if l.startswith('{')][0])
print(j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['config']['stage_ms'], j['parity']['ok'], j['parity']['max_err'])
PY
tail -3 gpurun_out/r2j_bench.err
