#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/c3_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c3_pytest.log
P=$PWD/darwin-gpu_b200/libgact_b200_prof.so
for m in 1 2 3; do
  GACT_CHAIN_MODE=$m GACT_LIB=$P python tools/chain_latency.py 30 1 148 592 1184 2368 > gpurun_out/c3_latency_mode$m.log 2>&1
done
rm -f gpurun_out/c3_chain_profile.log
for mb in 6 12 25 50; do
  for m in 0 1 2 3 4; do
    GACT_CHAIN_MODE=$m python tools/chain_profile.py $mb >> gpurun_out/c3_chain_profile.log 2>&1
  done
done
python bench.py --no-reads-leg --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo done
