#!/bin/bash
mkdir -p gpurun_out
D=$PWD/darwin-gpu_b200
GACT_LIB=$D/libgact_b200_check.so python tools/sanitize_run.py > gpurun_out/c6_bounds_check.log 2>&1
for v in prof nopf u2 u4; do
  echo "== variant $v" >> gpurun_out/c6_latency_variants.log
  GACT_CHAIN_MODE=1 GACT_LIB=$D/libgact_b200_$v.so python tools/chain_latency.py 30 1 592 >> gpurun_out/c6_latency_variants.log 2>&1
  GACT_CHAIN_MODE=3 GACT_LIB=$D/libgact_b200_$v.so python tools/chain_latency.py 30 4736 >> gpurun_out/c6_latency_variants.log 2>&1
done
for v in prof u2; do
  echo "== variant $v" >> gpurun_out/c6_bench_variants.log
  GACT_LIB=$D/libgact_b200_$v.so python bench.py --no-reads-leg --no-cpu-baseline --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('gcups', d['value'], 'e2e', d['e2e']['value'])" >> gpurun_out/c6_bench_variants.log 2>&1
done
python -m pytest tests/test_tiles_gpu.py tests/test_extend_gpu.py tests/test_hazards.py -m gpu -q -x > gpurun_out/c6_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c6_pytest.log
echo done
