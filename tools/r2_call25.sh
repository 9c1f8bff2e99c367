#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiles_gpu.py tests/test_hazards.py -x -q -m gpu > gpurun_out/c25_tests.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/c25_tests.log
GACT_LIB=darwin-gpu_b200/libgact_b200_check.so timeout 600 python tools/sanitize_run.py 2>&1 | tail -n 3
for v in "GACT_NARROW=0" ""; do echo "== $v"; env $v python tools/ncu_tile_driver.py 524288 | tail -n 1 | cut -c1-120; done
python bench.py --steps 5 --warmup 3 --no-reads-leg --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f  e2e %.0f  launches %d' % (d['value'], d['e2e']['value'], d['gpu_launches']))
"
