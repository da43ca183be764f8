#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -6
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-100
python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench35_c3.log; python -c "
import json; d=json.loads(open('gpurun_out/bench35_c3.log').read()); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}); print({k:(round(v['achieved'],1), round(v['frac'],3)) for k,v in d['rooflines'].items()}, d['clocks'], d['picked_marker'], 'e2e', round(d['e2e']['ms_per_step'],1))"
