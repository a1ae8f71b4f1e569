#!/bin/bash
# 8-GPU pass: peer protocol check, bench at N=8 (peer exchange) with its parity block, NCCL exchange once, N=4 peer
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2n8_gpus.txt
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $RUN --nproc-per-node 8 --master-port 29541 tests/peer_check.py > gpurun_out/r2n8_peer_check.log 2>&1; echo "peer_check rc=$?"
grep PEER_CHECK gpurun_out/r2n8_peer_check.log | cut -c1-400
timeout 900 $RUN --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err; echo "bench8 rc=$?"
timeout 900 $RUN --nproc-per-node 4 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 5 --no-parity > gpurun_out/r2n4_bench.json 2> gpurun_out/r2n4_bench.err; echo "bench4 rc=$?"
timeout 900 $RUN --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 --exchange nccl --no-parity > gpurun_out/r2n8_bench_nccl.json 2> gpurun_out/r2n8_bench_nccl.err; echo "bench8 nccl rc=$?"
python - <<'PY'
import json
for f in ['r2n8_bench','r2n4_bench','r2n8_bench_nccl']:
    try:
        j=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0])
        print(f, j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step'], j['config'].get('stage_ms'))
        print({k:v for k,v in j['config'].items() if 'nvlink_gbs' in k or 'remote' in k})
        if j.get('parity'): print('parity', j['parity']['ok'], j['parity']['max_err'], [ (p['table_err'],p['barrier_timeouts']) for p in j['parity']['per_rank']])
    except Exception as e: print(f, 'no line', e)
PY
tail -3 gpurun_out/r2n8_bench.err
