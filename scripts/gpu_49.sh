#!/bin/bash
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --search > gpurun_out/bench49_search.log 2>gpurun_out/bench49_search.err; echo rc=$?
tail -3 gpurun_out/bench49_search.err
