#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary20.txt
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a $S; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f $S
TAILN=25 TMO=900 run t20_all python -m pytest tests -q -m gpu -p no:cacheprovider
TAILN=3 TMO=300 run smoke20 python __graft_entry__.py smoke
TAILN=2 TMO=600 run bench20_c2 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu
TAILN=2 TMO=900 run bench20_c3 python bench.py --workload c3 --steps 3 --warmup 3
cat $S
