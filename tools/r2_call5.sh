#!/bin/bash
mkdir -p gpurun_out
P=$PWD/darwin-gpu_b200/libgact_b200_prof.so
(time python -m pytest tests/test_extend_gpu.py tests/test_tiles_gpu.py -m gpu -q -x) > gpurun_out/c5_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c5_pytest.log
GACT_CHAIN_MODE=1 GACT_LIB=$P python tools/chain_latency.py 30 1 592 > gpurun_out/c5_latency_mode1.log 2>&1
for mb in 6 12 25 50; do python tools/chain_profile.py $mb; done > gpurun_out/c5_chain_profile.log 2>&1
python tools/e2e_sweep.py 1 50 "DARWIN_BATCH_READS=0/" > gpurun_out/c5_e2e_1gpu.log 2>&1
python tools/e2e_sweep.py 1 6.25 "" > gpurun_out/c5_e2e_1gpu_shard8.log 2>&1
echo done
