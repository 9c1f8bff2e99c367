#!/bin/bash
# the two wavefront-kernel captures of tools/r2_ncu.sh (demangled names carry the template arguments as "<(int)10, (int)16, (bool)1>")
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*gact_tile_s16h_kernel<\(int\)10.*' -s 2 -c 1 -o gpurun_out/r2_wavefront -f python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_wavefront.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:.*gact_tile_s16h_kernel<\(int\)5.*' -s 1 -c 1 -o gpurun_out/r2_narrow -f python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_narrow.log 2>&1
tail -n 2 gpurun_out/ncu_wavefront.log gpurun_out/ncu_narrow.log
ls -la gpurun_out/r2_wavefront.ncu-rep gpurun_out/r2_narrow.ncu-rep
