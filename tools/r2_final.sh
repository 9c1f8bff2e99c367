#!/bin/bash
# driver-like sequence on one GPU: tests, smoke, bench (both arms)
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/f_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
echo "smoke rc $?" >> gpurun_out/f_smoke.log
(time python bench.py) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
(time python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
echo done
