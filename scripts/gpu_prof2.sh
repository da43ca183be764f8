#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summaryp2.txt; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summaryp2.txt
export PROF_LSCAN=75776
TMO=300 run profd_plain python scripts/prof_kernels.py
TMO=900 run profd_ncu ncu --set full --clock-control none --import-source on -k regex:"scan_i8_kernel|decode_ascii|syrk_i8" -c 6 -f -o gpurun_out/prof_r1d python scripts/prof_kernels.py
unset PROF_LSCAN
TMO=600 run benchd_c2 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu
TMO=600 run launchesd_c2 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1d_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu
cat gpurun_out/summaryp2.txt
