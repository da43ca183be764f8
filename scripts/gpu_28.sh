#!/bin/bash
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "config5 or preproducts" 2>&1 | tail -5
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}); print({k:(round(v['achieved'],1), round(v['frac'],3), v.get('timed_alone_frac')) for k,v in d['rooflines'].items()})"
