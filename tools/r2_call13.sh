#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-reads-leg --no-cpu-baseline --steps 5 --warmup 3"
for cfg in "4 65536" "4 262144" "3 262144" "2 262144"; do
  set -- $cfg
  echo "== GACT_IT_CTAS=$1 chunk=$2" >> gpurun_out/c13_bench.log
  GACT_IT_CTAS=$1 timeout 600 $B --chunk $2 2>gpurun_out/c13_err.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('gcups', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e_ms', round(d['e2e']['ms_per_step'],2))" >> gpurun_out/c13_bench.log 2>&1
done
(time timeout 1200 python -m pytest tests/test_tiles_gpu.py tests/test_hazards.py -m gpu -q -x) > gpurun_out/c13_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c13_pytest.log
echo done
