#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_e2e_gpu.py tests/test_dsoft_gpu.py -m gpu -q) > gpurun_out/c9_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c9_pytest.log
python tools/e2e_sweep.py 1 50 "DARWIN_TRACE=1" > gpurun_out/c9_e2e_trace.log 2>&1
python tools/e2e_sweep.py 1 6.25 "DARWIN_TRACE=1" > gpurun_out/c9_e2e_trace_shard8.log 2>&1
python bench.py --config 1 --cpu-reads 100 > gpurun_out/c9_config1.json 2> gpurun_out/c9_config1.err
echo done
