#!/bin/bash
mkdir -p gpurun_out
python tools/e2e_sweep.py 1 50 "DARWIN_TRACE=1/DARWIN_TRACE=1,DARWIN_BATCH_READS=0" > gpurun_out/c8_e2e_trace.log 2>&1
python tools/e2e_sweep.py 1 6.25 "DARWIN_TRACE=1" > gpurun_out/c8_e2e_trace_shard8.log 2>&1
python bench.py --config 5 --cpu-reads 100 > gpurun_out/c8_config5.json 2> gpurun_out/c8_config5.err
python bench.py --config 1 --cpu-reads 100 > gpurun_out/c8_config1.json 2> gpurun_out/c8_config1.err
echo done
