#!/bin/bash
# A/B of an environment switch on the bench step: tools/ab.sh VAR "v1 v2 ..." [bench args]
var=$1; vals=$2; shift 2
for v in $vals; do
  env $var=$v timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('$var=$v', 'value', round(j['value']/1e6,2), 'e2e', round(j['e2e']['value']/1e6,2), 'ms', round(j['ms_per_step'],4), j['config']['stage_ms'])
"
done
