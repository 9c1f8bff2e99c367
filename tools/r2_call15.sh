#!/bin/bash
# inter-task traceback rework: parity, then timing of the variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiles_gpu.py -x -q -m gpu > gpurun_out/c15_tests.log 2>&1; echo "tests rc=$?"
tail -n 3 gpurun_out/c15_tests.log
for v in "" "GACT_IT_QS=0" "GACT_IT_BAND=24" "GACT_IT_BAND=28"; do
  echo "== $v"; env $v python tools/ncu_tile_driver.py 524288 2>&1 | tail -n 1
done
B="python bench.py --steps 5 --warmup 3 --no-reads-leg --no-cpu-baseline"
$B > gpurun_out/c15_bench.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/c15_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('bench value', d['value'], 'e2e', d['e2e']['value'], 'routing', d.get('tile_routing'))
PY
