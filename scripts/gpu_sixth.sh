#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary6.txt; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary6.txt
TMO=900 run t6_all python -m pytest tests -q -m gpu -p no:cacheprovider
TMO=300 run prof6_c2 python scripts/prof_kernels.py
export PROF_N=10000 PROF_L=1000000 PROF_LSCAN=37888
TMO=300 run prof6_c3 python scripts/prof_kernels.py
unset PROF_N PROF_L PROF_LSCAN
TMO=600 run bench6_c2 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu
cat gpurun_out/summary6.txt
