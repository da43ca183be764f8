#!/bin/bash
timeout 600 python -m pytest tests/test_packed.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "pinned or mmt or store" 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), 'e2e', d['e2e']['ms_per_step'], d['e2e']['abi_stage_ms']); print(d['e2e']['from_packed_container']['ms_per_step'], d['picked_marker'])"
