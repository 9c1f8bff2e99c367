#!/bin/bash
# inter-task kernel alone (all tiles full, none first), previous vs current build, and band widths
for lib in prev ""; do
  L=darwin-gpu_b200/libgact_b200${lib:+_$lib}.so
  for w in "FULL_FRAC=1 FIRST_FRAC=0" "FULL_FRAC=0.82"; do
    echo "== $L $w"; env GACT_LIB=$L $w python tools/ncu_tile_driver.py 524288 2>&1 | tail -n 1
  done
done
for v in "GACT_IT_QS=0" "GACT_IT_BAND=24" "GACT_IT_BAND=20" "GACT_IT_BAND=48" "GACT_IT_CTAS=3"; do
  echo "== $v IT only"; env $v FULL_FRAC=1 FIRST_FRAC=0 python tools/ncu_tile_driver.py 524288 2>&1 | tail -n 1
done
