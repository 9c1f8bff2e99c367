#!/bin/bash
for lib in va vb; do
  L=darwin-gpu_b200/libgact_b200_$lib.so
  GACT_LIB=$L FULL_FRAC=1 FIRST_FRAC=0 ncu --set full --clock-control none --import-source on -k regex:gact_tile_it_kernel -s 1 -c 1 -o gpurun_out/r2_it_$lib -f python tools/ncu_tile_driver.py 524288 > gpurun_out/ncu_it_$lib.log 2>&1
  tail -n 2 gpurun_out/ncu_it_$lib.log | cut -c1-100
done
