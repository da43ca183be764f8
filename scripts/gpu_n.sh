#!/bin/bash
# usage: gpu_n.sh N
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --workload c3 > gpurun_out/bench_r1f_n${N}_c3.log 2>&1
echo "exit $?"; tail -1 gpurun_out/bench_r1f_n${N}_c3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],2), round(d['value']), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['value']), d['picked_marker'])" || tail -20 gpurun_out/bench_r1f_n${N}_c3.log
