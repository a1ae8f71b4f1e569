#!/bin/bash
# validation pass: all GPU tests, the default bench line, ncu --set full of the backward contraction, eager timeline of one step
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/f1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f1_pytest.log
timeout 900 python bench.py > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/f1_bench.json') if l.startswith('{')][0])
    print('bench', j['value']/1e6, 'e2e', j['e2e']['value']/1e6, 'ms', j['ms_per_step'], j['config'].get('stage_ms'))
    print('parity', j['parity']['ok'], j['parity']['max_err'], 'roofline', j['roofline'], 'launches', j.get('gpu_launches'), 'clocks', j.get('clocks'))
    print('raw', j.get('e2e_raw_edges'), 'buffered', j.get('buffered'))
    print('configs', [(c.get('workload'), c.get('value')) for c in j.get('configs', [])])
    print('cpu', j.get('cpu_baseline'))
except Exception as e: print('no bench line', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_ts_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_ts python tools/ts_check.py group 100 1000 400 > gpurun_out/f1_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_ncu_ts.ncu-rep --page raw --csv > gpurun_out/r2_ncu_ts_raw.csv 2>/dev/null
ls -la gpurun_out/r2_ncu_ts.ncu-rep
timeout 300 python tools/timeline.py --batch 50000 --steps 3 > gpurun_out/f1_timeline.log 2>&1; tail -40 gpurun_out/f1_timeline.log
