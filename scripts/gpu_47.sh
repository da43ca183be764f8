#!/bin/bash
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --search > gpurun_out/bench47_search.log 2>gpurun_out/bench47_search.err; echo rc=$?
tail -c 3000 gpurun_out/bench47_search.log | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['forward_search'], indent=1)); print(d['ms_per_step'])"
tail -3 gpurun_out/bench47_search.err
