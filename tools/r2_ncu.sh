#!/bin/bash
# ncu evidence of the round (one call): launch list of the bench command, full captures of the tile kernel and of both chain kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-reads-leg --no-cpu-baseline"
$B > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_tile_driver.py > gpurun_out/ncu_plain_tile.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gact_tile_s16h_kernel -s 1 -c 1 -o gpurun_out/r2_tile -f python tools/ncu_tile_driver.py > gpurun_out/ncu_tile.log 2>&1
python tools/chain_profile.py 6 > gpurun_out/ncu_plain_chain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gact_chain_s16h_kernel -s 2 -c 1 -o gpurun_out/r2_chain_latency -f python tools/chain_profile.py 6 > gpurun_out/ncu_chain6.log 2>&1
python tools/chain_profile.py 50 > gpurun_out/ncu_plain_chain50.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gact_chain_s16h_kernel -s 2 -c 1 -o gpurun_out/r2_chain_throughput -f python tools/chain_profile.py 50 > gpurun_out/ncu_chain50.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
