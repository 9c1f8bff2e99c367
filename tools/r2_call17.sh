#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tiles_gpu.py -x -q -m gpu > gpurun_out/c17_tests.log 2>&1; echo "tests rc=$?"
tail -n 3 gpurun_out/c17_tests.log
for lib in prev ""; do
  L=darwin-gpu_b200/libgact_b200${lib:+_$lib}.so
  for w in "FULL_FRAC=1 FIRST_FRAC=0" "FULL_FRAC=0.82"; do
    echo "== $L $w"; env GACT_LIB=$L $w python tools/ncu_tile_driver.py 524288 2>&1 | tail -n 1
  done
done
