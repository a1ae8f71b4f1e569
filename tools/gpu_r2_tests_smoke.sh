#!/bin/bash
# all GPU tests, smoke(), and the default bench line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/f3_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f3_pytest.log
python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')
" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/f3_bench.json') if l.startswith('{')][0])
    print('bench', j['value']/1e6, 'e2e', j['e2e']['value']/1e6, 'ms', j['ms_per_step'], j['config'].get('stage_ms'))
    print('parity', j['parity']['ok'], j['parity']['max_err'], 'launches', j.get('gpu_launches'), 'clocks', j.get('clocks'))
    r=j['roofline']; print('roofline', {k:r[k] for k in r if k!='note'})
    print('buffered', j['buffered']['async_swaps']['value'], j['buffered']['sync_swaps']['value'])
except Exception as e: print('no bench line', e)
PY
