#!/bin/bash
# N=2 bench at c3 (multi-rank path after the allocator / bench changes)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r1g_bench_c3_n2.log 2>gpurun_out/r1g_bench_c3_n2.err; echo rc=$?
tail -c 600 gpurun_out/r1g_bench_c3_n2.log | cut -c1-400; tail -3 gpurun_out/r1g_bench_c3_n2.err
