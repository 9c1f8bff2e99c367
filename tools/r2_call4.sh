#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_e2e_gpu.py tests/test_abi.py tests/test_dsoft_gpu.py -m gpu -q) > gpurun_out/c4_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c4_pytest.log
python tools/ref_gpu_bench.py 65536 > gpurun_out/c4_ref_gpu.json 2> gpurun_out/c4_ref_gpu.err
(time python bench.py --steps 5 --warmup 3) > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/c4_bench_ref.json 2> gpurun_out/c4_bench_ref.err
echo done
