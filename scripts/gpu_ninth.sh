#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary9.txt; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f gpurun_out/summary9.txt
TAILN=30 TMO=900 run t9_scan python -m pytest tests -q -m gpu -p no:cacheprovider -k "scan or forward or shards" -x
TAILN=6 TMO=600 run t9_all python -m pytest tests -q -m gpu -p no:cacheprovider
EAGLE_SCAN_MODE=i8 TMO=600 run bench9_c2_i8 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu
cat gpurun_out/summary9.txt
