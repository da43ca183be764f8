#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_c3.py > gpurun_out/prof26_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'decode_ascii|syrk_i8|scan_i8_kernel|prep_i8_kernel|transpose' -o gpurun_out/prof_r1f_c3 -f python scripts/prof_c3.py > gpurun_out/prof26_ncu.log 2>&1
tail -2 gpurun_out/prof26_plain.log gpurun_out/prof26_ncu.log
python bench.py --workload c3 --steps 3 --warmup 3 > gpurun_out/bench26_c3.log 2>&1; echo bench exit $?; tail -c 1500 gpurun_out/bench26_c3.log
