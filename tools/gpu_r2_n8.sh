#!/bin/bash
# 8-GPU pass: peer protocol check, bench at N=8 / 4 / 2 (peer exchange) with parity, NCCL exchange once
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $RUN --nproc-per-node 8 --master-port 29541 tests/peer_check.py > gpurun_out/r2n8_peer_check.log 2>&1; echo "peer_check rc=$?"
grep "PEER_CHECK" gpurun_out/r2n8_peer_check.log | cut -c1-160
timeout 900 $RUN --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err; echo "bench8 rc=$?"
timeout 900 $RUN --nproc-per-node 4 --master-port 29543 bench.py --gpus 4 --steps 40 --warmup 5 --no-parity > gpurun_out/r2n4_bench.json 2> gpurun_out/r2n4_bench.err; echo "bench4 rc=$?"
timeout 900 $RUN --nproc-per-node 2 --master-port 29545 bench.py --gpus 2 --steps 40 --warmup 5 --no-parity > gpurun_out/r2n2_bench.json 2> gpurun_out/r2n2_bench.err; echo "bench2 rc=$?"
timeout 240 $RUN --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --steps 6 --warmup 3 --exchange nccl --no-parity > gpurun_out/r2n8_bench_nccl.json 2> gpurun_out/r2n8_bench_nccl.err; echo "bench8 nccl rc=$?"
python - <<'PY'
import json
for f in ['r2n8_bench','r2n4_bench','r2n2_bench','r2n8_bench_nccl']:
    try:
        j=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0])
        print(f, round(j['value']/1e6,2), round(j['e2e']['value']/1e6,2), round(j['ms_per_step'],4), j['config'].get('stage_ms'))
        print('   ', {k:v for k,v in j['config'].items() if 'nvlink_gbs' in k or 'remote_f' in k}, 'trace', j['config'].get('step_ms_trace_rank0'))
        if j.get('parity'): print('   parity', j['parity']['ok'], j['parity']['max_err'])
    except Exception as e: print(f, 'no line', e)
PY
