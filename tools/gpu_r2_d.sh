#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2d_bench.json') if l.startswith('{')][0])
print(j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['config']['stage_ms'], j['parity']['ok'], j['parity']['max_err'])
PY
tail -3 gpurun_out/r2d_bench.err
MB_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_group_kernel -s 5 -c 1 -o gpurun_out/r2d_gemm_bwd \
   python bench.py --steps 1 --warmup 3 --nodes 2000000 --no-cpu-baseline --no-parity > gpurun_out/r2d_ncu.out 2>&1; echo "ncu rc=$?"
