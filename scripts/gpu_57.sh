#!/bin/bash
EAGLE_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e 2>&1 | grep "stage_ms" | tail -4
