#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" 2>gpurun_out/c22_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f  e2e %.0f  e2e ms %.2f  %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['api'][:60]))
"; }
B="python bench.py --steps 5 --warmup 3 --no-reads-leg --no-cpu-baseline"
for cw in 1 2 4; do echo "== cta warps $cw"; run GACT_IT_CTA_WARPS=$cw $B; done
echo "== cta warps 1 chunk 128Ki"; run GACT_IT_CTA_WARPS=1 $B --chunk 131072
for cw in 1 4; do GACT_IT_CTA_WARPS=$cw python tools/ncu_tile_driver.py 524288 | tail -n 1 | cut -c1-90;  GACT_IT_CTA_WARPS=$cw FULL_FRAC=1 FIRST_FRAC=0 python tools/ncu_tile_driver.py 524288 | tail -n 1 | cut -c1-90; done
timeout 600 python -m pytest tests/test_tiles_gpu.py -x -q -m gpu 2>&1 | tail -n 2
