#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 3 --warmup 3 --workload ${W:-c5} --no-e2e --no-cpu > gpurun_out/bench_r1f_n${N}_${W:-c5}.log 2>&1
echo "exit $?"; tail -1 gpurun_out/bench_r1f_n${N}_${W:-c5}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['ms_per_step'],2), round(d['value']), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['picked_marker']); print({k:(round(v['achieved'],1), round(v['frac'],3)) for k,v in d['rooflines'].items()})" || tail -20 gpurun_out/bench_r1f_n${N}_${W:-c5}.log
