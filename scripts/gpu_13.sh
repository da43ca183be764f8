#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary13.txt
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a $S; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f $S
TAILN=25 TMO=600 run t13_scan python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "scan or forward or store_shards or config2"
TAILN=8 TMO=300 SW_SHAPES=37x4,24x6,30x5 run sweep13 python scripts/scan_shape_sweep.py
cat $S
