#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" 2>gpurun_out/c23_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f  e2e %.0f  e2e ms %.2f  %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['api'][:60]))
"; }
B="python bench.py --steps 5 --warmup 3 --no-reads-leg --no-cpu-baseline"
for m in 37888 60000 110000 170000; do echo "== IT_MIN $m"; run GACT_IT_MIN=$m $B; done
echo "== IT_MIN 60000 chunk 192Ki"; run GACT_IT_MIN=60000 $B --chunk 196608
echo "== IT_MIN 110000 chunk 320Ki"; run GACT_IT_MIN=110000 $B --chunk 327680
