#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_link.py tests/test_gpu_decoder.py -m gpu -x -q > gpurun_out/r2n2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n2b_pytest.log
tail -6 gpurun_out/r2n2b_pytest.log
for fb in 1 0; do
MB_FETCH_BULK=$fb timeout 900 $RUN --nproc-per-node 2 --master-port 2955$fb bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n2b_bench_$fb.json 2> gpurun_out/r2n2b_bench_$fb.err; echo "bench rc=$?"
done
CUDA_LAUNCH_BLOCKING=1 timeout 600 $RUN --nproc-per-node 2 --master-port 29557 bench.py --gpus 2 --steps 4 --warmup 3 --exchange nccl --no-parity > gpurun_out/r2n2b_nccl.json 2> gpurun_out/r2n2b_nccl.err; echo "nccl rc=$?"
python - <<'PY'
import json
for f in ['r2n2b_bench_1','r2n2b_bench_0','r2n2b_nccl']:
    try:
        j=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0])
        print(f, j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['config'].get('stage_ms'))
        if j.get('parity'): print('parity', j['parity']['ok'], j['parity']['max_err'])
    except Exception as e: print(f, 'no line', e)
PY
grep -i "error\|what()" gpurun_out/r2n2b_nccl.err | head -5
