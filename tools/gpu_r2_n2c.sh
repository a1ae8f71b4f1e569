#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $RUN --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2n2c_bench.json 2> gpurun_out/r2n2c_bench.err; echo "bench rc=$?"
timeout 600 $RUN --nproc-per-node 2 --master-port 29562 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2n2c_ref.json 2> gpurun_out/r2n2c_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ['r2n2c_bench','r2n2c_ref']:
    try:
        j=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0])
        print(f, j['value']/1e6, j['e2e']['value']/1e6, j.get('ms_per_step'), j['config'].get('stage_ms') if isinstance(j.get('config'),dict) else None, j.get('cpu_baseline'))
        if j.get('parity'): print('parity', j['parity']['ok'], j['parity']['max_err'])
        if 'step_ms_trace_rank0' in j.get('config',{}): print(j['config']['step_ms_trace_rank0'])
    except Exception as e: print(f, 'no line', e)
PY
tail -3 gpurun_out/r2n2c_bench.err
