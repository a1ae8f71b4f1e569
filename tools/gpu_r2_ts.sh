#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ts_check.py small > gpurun_out/ts_small.log 2>&1; rc=$?
cat gpurun_out/ts_small.log | tail -60
if [ $rc -ne 0 ]; then echo "TS small failed rc=$rc"; dmesg 2>/dev/null | tail -3; exit 0; fi
timeout 300 python tools/ts_check.py time > gpurun_out/ts_time.log 2>&1; cat gpurun_out/ts_time.log | tail
timeout 600 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q > gpurun_out/ts_pytest_dec.log 2>&1; rc=$?; tail -5 gpurun_out/ts_pytest_dec.log
if [ $rc -ne 0 ]; then echo "decoder tests failed"; exit 0; fi
for ts in 1 0; do
MB_CONV_TS=$ts timeout 600 python bench.py --steps 20 --warmup 5 --no-extra-shapes --no-buffered > gpurun_out/ts_bench_$ts.json 2> gpurun_out/ts_bench_$ts.err; echo "bench ts=$ts rc=$?"
done
python - <<'PY'
import json
for f in ['ts_bench_1','ts_bench_0']:
    try:
        j=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0])
        print(f, j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['config'].get('stage_ms'))
        if j.get('parity'): print('parity', j['parity']['ok'], j['parity']['max_err'])
    except Exception as e: print(f, 'no line', e)
PY
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/ts_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/ts_pytest_all.log
