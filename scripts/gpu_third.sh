#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary3.txt; tail -4 gpurun_out/$name.log; }
rm -f gpurun_out/summary3.txt
TMO=600 run t3_dev python -m pytest tests/test_gpu_device.py -q -p no:cacheprovider
TMO=600 run t3_parity python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider
TMO=300 run prof3_plain python scripts/prof_kernels.py
TMO=900 run prof3_ncu ncu --set full --clock-control none --import-source on -k regex:"decode_ascii|scan_f64" -c 4 -f -o gpurun_out/prof_r1b python scripts/prof_kernels.py
export PROF_N=10000 PROF_L=1000000 PROF_LSCAN=18944
TMO=300 run prof3c3_plain python scripts/prof_kernels.py
TMO=900 run prof3c3_ncu ncu --set full --clock-control none -k regex:"syrk_i8" -s 1 -c 1 -f -o gpurun_out/prof_r1b_syrk_c3 python scripts/prof_kernels.py
unset PROF_N PROF_L PROF_LSCAN
TMO=600 run bench3_c2 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu
cat gpurun_out/summary3.txt
