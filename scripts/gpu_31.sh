#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['roofline']['kernel_ms'], d['gpu_launches'], d['picked_marker'])"
python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['roofline']['kernel_ms'], d['picked_marker'])"
