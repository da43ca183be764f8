#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary12.txt; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary12.txt
TAILN=25 TMO=900 run t12_all python -m pytest tests -q -m gpu -p no:cacheprovider
TMO=600 run bench12_c2 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu
TMO=900 run bench12_c3 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu
cat gpurun_out/summary12.txt
