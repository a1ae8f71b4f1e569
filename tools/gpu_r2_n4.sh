#!/bin/bash
# 4-GPU pass: bench at N=4 (peer exchange) with its parity block
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $RUN --nproc-per-node 4 --master-port 29571 bench.py --gpus 4 --steps 40 --warmup 5 > gpurun_out/r2n4c_bench.json 2> gpurun_out/r2n4c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/r2n4c_bench.json') if l.startswith('{')][0])
    print('n4', j['value']/1e6, j['e2e']['value']/1e6, j.get('ms_per_step'), j['config'].get('stage_ms'))
    print('parity', j['parity']['ok'], j['parity']['max_err'], j['config'].get('step_ms_trace_rank0'))
    print({k: j['config'].get(k) for k in ('remote_fraction','nvlink_gbs_per_gpu_each_direction')}, j.get('scaling_info'))
except Exception as e: print('no line', e)
PY
tail -3 gpurun_out/r2n4c_bench.err
