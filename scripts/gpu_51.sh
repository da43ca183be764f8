#!/bin/bash
for i in 1 2; do timeout 600 python bench.py --no-search --no-cpu --e2e-steps 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('e2e ms', e['ms_per_step'], 'packed', e['from_packed_container']['ms_per_step'], 'step', d['ms_per_step'])"; done
