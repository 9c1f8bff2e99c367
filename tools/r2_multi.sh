#!/bin/bash
# usage: r2_multi.sh N  -- multi-GPU checks on an N-GPU box
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m${N}_gpus.txt
python -m pytest tests/test_e2e_gpu.py -m gpu -q -k "multi_gpu" > gpurun_out/m${N}_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/m${N}_pytest.log
L="1"; for g in 2 4 8; do if [ $g -le $N ]; then L="$L,$g"; fi; done
python tools/e2e_sweep.py $L 50 "DARWIN_MULTIPROC=1/DARWIN_MULTIPROC=0" > gpurun_out/m${N}_e2e_strong.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo done
