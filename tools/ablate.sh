#!/bin/bash
# GEMM ablation: stage timings of the bench step with pieces of the grouped tcgen05 kernel switched off (MB_TC_DEBUG bits).
for dbg in ${DBGS:-0 8 1 2 4 10}; do
  MB_TC_DEBUG=$dbg timeout 120 python bench.py --steps 20 --warmup 3 --nodes 2000000 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{\"metric'):
        j=json.loads(l); s=j['config']['stage_ms']; print('dbg', $dbg, 'scores', s.get('gemm_scores'), 'bwd', s.get('gemm_dA'), 'step_ms', round(j['ms_per_step'],4))
"
done
