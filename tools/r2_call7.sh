#!/bin/bash
mkdir -p gpurun_out
D=$PWD/darwin-gpu_b200
(time python -m pytest tests -m gpu -q) > gpurun_out/c7_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c7_pytest.log
for v in prof u8; do
  echo "== variant $v" >> gpurun_out/c7_latency_variants.log
  GACT_CHAIN_MODE=1 GACT_LIB=$D/libgact_b200_$v.so python tools/chain_latency.py 30 1 592 >> gpurun_out/c7_latency_variants.log 2>&1
done
for mb in 6 12 25 50; do python tools/chain_profile.py $mb; done > gpurun_out/c7_chain_profile.log 2>&1
python tools/e2e_sweep.py 1 50 "" > gpurun_out/c7_e2e_1gpu.log 2>&1
python tools/e2e_sweep.py 1 6.25 "" > gpurun_out/c7_e2e_1gpu_shard8.log 2>&1
python bench.py --no-reads-leg --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
echo done
