#!/bin/bash
for pair in 0 1; do
EAGLE_SI_PAIR=$pair python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pair=$pair', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['roofline']['kernel_ms'], d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['ms_per_step'],1), d['e2e']['abi_stage_ms'])"
done
