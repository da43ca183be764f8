#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary4.txt; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary4.txt
TAILN=30 TMO=120 run dmma_bench scripts/microbench/dmma_bench
TMO=600 run t4_mmt python -m pytest tests/test_gpu_parity.py tests/test_gpu_device.py -q -k "mmt or config2 or shards" -p no:cacheprovider
export PROF_N=10000 PROF_L=1000000 PROF_LSCAN=18944
TMO=300 run prof4c3_flow python scripts/prof_kernels.py
EAGLE_SYRK_FLOWCTL=0 TMO=300 run prof4c3_noflow python scripts/prof_kernels.py
unset PROF_N PROF_L PROF_LSCAN
TMO=300 run prof4c2_flow python scripts/prof_kernels.py
cat gpurun_out/summary4.txt
