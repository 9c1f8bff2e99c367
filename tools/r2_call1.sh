#!/bin/bash
# GPU call 1 of round 2: baseline tests, chain latency anatomy, sanitizer logs, full-batch parity.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_gpu.txt
(time python -m pytest tests -m gpu -x -q) > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c1_pytest.log
GACT_LIB=$PWD/darwin-gpu_b200/libgact_b200_prof.so python tools/chain_latency.py 30 1 148 592 1184 2368 > gpurun_out/c1_latency_lat.log 2>&1
GACT_CHAIN_THROUGHPUT=1 GACT_LIB=$PWD/darwin-gpu_b200/libgact_b200_prof.so python tools/chain_latency.py 30 1 148 592 1184 2368 4736 > gpurun_out/c1_latency_thr.log 2>&1
python tools/chain_profile.py 6 > gpurun_out/c1_chain_profile_6mb.log 2>&1
python tools/chain_profile.py 50 > gpurun_out/c1_chain_profile_50mb.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/c1_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/c1_racecheck.log 2>&1
python tools/full_batch_parity.py 1048576 gpurun_out/c1_full_batch_parity.json > gpurun_out/c1_full_batch.log 2>&1
echo done
