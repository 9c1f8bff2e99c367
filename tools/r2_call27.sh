#!/bin/bash
run() { env "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f  e2e %.0f  e2e ms %.2f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))
"; }
B="python bench.py --steps 5 --warmup 3 --no-reads-leg --no-cpu-baseline"
echo "== default"; run $B
echo "== GACT_NARROW=0"; run GACT_NARROW=0 $B
echo "== narrow on IT stream"; run GACT_NARROW_STREAM=1 $B
echo "== carveout 64"; run GACT_CARVEOUT=64 $B
echo "== carveout 64 + narrow stream"; run GACT_CARVEOUT=64 GACT_NARROW_STREAM=1 $B
