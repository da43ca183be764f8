#!/bin/bash
timeout 300 python -m pytest tests/test_am.py tests/test_gpu_algebra.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['forward_search']; print(f['iterations'], f['all_picked_1based'], f['seconds'])"
