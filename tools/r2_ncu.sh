#!/bin/bash
# ncu evidence of the round (one call): launch list of the bench command, full captures of the two tile kernels and of both chain kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-reads-leg --no-cpu-baseline"
$B > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_plain_tile.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gact_tile_it_kernel -s 1 -c 1 -o gpurun_out/r2_it -f python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_it.log 2>&1
python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_plain_tile2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*gact_tile_s16h_kernel<\(int\)10.*' -s 2 -c 1 -o gpurun_out/r2_wavefront -f python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_wavefront.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*gact_tile_s16h_kernel<\(int\)5.*' -s 1 -c 1 -o gpurun_out/r2_narrow -f python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_narrow.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
