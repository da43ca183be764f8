#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary14.txt
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a $S; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f $S
TAILN=25 TMO=300 run t14_scan python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "scan or forward or store_shards or config2"
TAILN=12 TMO=300 SW_SHAPES=18x4,37x2,12x6,24x3,9x8,36x4 run sweep14 python scripts/scan_shape_sweep.py
TAILN=4 TMO=200 EAGLE_SI_PAIR=0 SW_SHAPES=37x4 run sweep14_single python scripts/scan_shape_sweep.py
cat $S
