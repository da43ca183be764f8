#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary16.txt
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a $S; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f $S
TAILN=30 TMO=300 run t16_prep python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "preproducts"
TAILN=25 TMO=600 run t16_all python -m pytest tests -q -m gpu -p no:cacheprovider -x
TAILN=3 TMO=600 run bench16_c3 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu
cat $S
