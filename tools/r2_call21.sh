#!/bin/bash
mkdir -p gpurun_out
run() { env GACT_HOST_TRACE=1 "$@" 2>gpurun_out/c21_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f  e2e %.0f  e2e ms %.2f  %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['api'][:60]))
"; grep HOST_TRACE gpurun_out/c21_err.log; }
B="python bench.py --steps 5 --warmup 3 --no-reads-leg --no-cpu-baseline"
echo "== default"; run $B
echo "== threads 4"; run env GACT_HOST_THREADS=4 $B
echo "== threads 1"; run env GACT_HOST_THREADS=1 $B
echo "== chunk 128Ki"; run $B --chunk 131072
echo "== chunk 512Ki"; run $B --chunk 524288
timeout 600 python -m pytest tests/test_tiles_gpu.py -x -q -m gpu 2>&1 | tail -n 2
