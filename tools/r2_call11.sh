#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_plain_it.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gact_tile_it_kernel -s 1 -c 1 -o gpurun_out/r2_it -f python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_it.log 2>&1
ls -la gpurun_out/r2_it.ncu-rep
echo done
