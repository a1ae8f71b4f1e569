#!/bin/bash
# 2-GPU pass: multi-GPU tests (peer protocol with shared rows), bench at N=2 (peer exchange) with its parity block, NCCL exchange once
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2n2_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n2_pytest.log
tail -12 gpurun_out/r2n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 \
   > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/r2n2_bench.json') if l.startswith('{')][0])
    print(j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['config'].get('stage_ms'), j['parity'])
    print({k:v for k,v in j['config'].items() if 'nvlink' in k or 'remote' in k})
except Exception as e: print('no bench line', e)
PY
tail -5 gpurun_out/r2n2_bench.err
