#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_tiles_gpu.py -m gpu -q -x -k "inter_task or large_batch or golden or microbatch_default") > gpurun_out/c10_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c10_pytest.log
timeout 600 python bench.py --no-reads-leg --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
GACT_IT=0 timeout 600 python bench.py --no-reads-leg --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c10_bench_noit.json 2> gpurun_out/c10_bench_noit.err
for w in 16 24 48; do GACT_IT_BAND=$w timeout 600 python bench.py --no-reads-leg --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/c10_bench_w$w.json 2> gpurun_out/c10_bench_w$w.err; done
echo done
