#!/bin/bash
mkdir -p gpurun_out
N=$1
lscpu | grep -E "^CPU\(s\)|NUMA|Socket|Model name" > gpurun_out/c24_lscpu.txt
nvidia-smi topo -m > gpurun_out/c24_topo.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-reads-leg --no-cpu-baseline"
BENCH_NUMA_BIND=0 $T > gpurun_out/c24_nobind.json 2> gpurun_out/c24_nobind.err
$T > gpurun_out/c24_bind.json 2> gpurun_out/c24_bind.err
GACT_HOST_THREADS=4 $T > gpurun_out/c24_bind_t4.json 2> gpurun_out/c24_bind_t4.err
python - <<'PY'
import json
for f in ('nobind','bind','bind_t4'):
    for l in open(f'gpurun_out/c24_{f}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'value %.0f e2e %.0f e2e_ms %.2f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']), d.get('host_binding'))
PY
cat gpurun_out/c24_lscpu.txt
