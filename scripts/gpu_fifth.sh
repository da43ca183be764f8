#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary5.txt; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary5.txt
TMO=600 run t5_scan python -m pytest tests/test_gpu_parity.py tests/test_gpu_device.py -q -k "scan or forward or shards or config2 or reduced" -p no:cacheprovider
TMO=300 run prof5_c2 python scripts/prof_kernels.py
TMO=900 run prof5_ncu ncu --set full --clock-control none --import-source on -k regex:"scan_f64" -c 2 -f -o gpurun_out/prof_r1c python scripts/prof_kernels.py
export PROF_N=10000 PROF_L=1000000 PROF_LSCAN=18944
TMO=900 run prof5c3_ncu ncu --set full --clock-control none -k regex:"syrk_i8" -s 1 -c 1 -f -o gpurun_out/prof_r1c_syrk_c3 python scripts/prof_kernels.py
unset PROF_N PROF_L PROF_LSCAN
TMO=600 run bench5_c2 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu
cat gpurun_out/summary5.txt
