#!/bin/bash
EAGLE_BENCH_DEBUG=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/r1g_bench_c3_n8.log 2>gpurun_out/r1g_bench_c3_n8.err; echo rc=$?
grep "stage_ms" gpurun_out/r1g_bench_c3_n8.err | head -3; tail -1 gpurun_out/r1g_bench_c3_n8.log | cut -c1-200
