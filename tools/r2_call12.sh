#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_tiles_gpu.py -m gpu -q -x -k "inter_task or large_batch or exceptions") > gpurun_out/c12_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/c12_pytest.log
timeout 600 python bench.py --no-reads-leg --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_plain_it2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gact_tile_it_kernel -s 1 -c 1 -o gpurun_out/r2_it2 -f python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_it2.log 2>&1
echo done
