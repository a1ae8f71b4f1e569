#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_link.py tests/test_gpu_host.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -12 gpurun_out/r2g_pytest.log
df -h /dev/shm /tmp | tail -3; free -g | head -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-extra-shapes > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2g_bench.json') if l.startswith('{')][0])
print(j['value']/1e6, j['ms_per_step'])
print('buffered', json.dumps(j['buffered'], indent=1))
PY
tail -3 gpurun_out/r2g_bench.err
