#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary2.txt; tail -4 gpurun_out/$name.log; }
rm -f gpurun_out/summary2.txt
TMO=600 run t2_parity python -m pytest tests/test_gpu_parity.py -q -k "scan or reduced or forward" -p no:cacheprovider
TMO=300 run prof_plain python scripts/prof_kernels.py
if grep -q "^ok" gpurun_out/prof_plain.log; then
TMO=900 run prof_ncu ncu --set full --clock-control none --import-source on -k regex:"decode_ascii|syrk_i8|scan_f64" -c 6 -f -o gpurun_out/prof_r1a python scripts/prof_kernels.py
fi
TMO=600 run bench_c2 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu
TMO=600 run launches_c2 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu
TMO=900 run bench_c3 python bench.py --workload c3 --steps 3 --warmup 3
cat gpurun_out/summary2.txt
