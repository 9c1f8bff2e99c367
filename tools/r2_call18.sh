#!/bin/bash
for lib in prev va vf vg vh ""; do
  L=darwin-gpu_b200/libgact_b200${lib:+_$lib}.so
  echo "== $L"; env GACT_LIB=$L FULL_FRAC=1 FIRST_FRAC=0 python tools/ncu_tile_driver.py 524288 2>&1 | tail -n 1 | cut -c1-80
  env GACT_LIB=$L python tools/ncu_tile_driver.py 524288 2>&1 | tail -n 1 | cut -c1-80
done
