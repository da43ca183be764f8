#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary23.txt
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a $S; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f $S
TAILN=5 TMO=300 run t23_scan python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "scan or forward or store_shards"
echo c2; SW_N=2000 SW_L=500000 SW_SHAPES=37x4 timeout 300 python scripts/scan_shape_sweep.py 2>&1 | tail -1
echo c2 pair; EAGLE_SI_PAIR=1 SW_N=2000 SW_L=500000 SW_SHAPES=18x4 timeout 300 python scripts/scan_shape_sweep.py 2>&1 | tail -1
echo c3; SW_SHAPES=37x4 timeout 300 python scripts/scan_shape_sweep.py 2>&1 | tail -1
echo c3 pair; EAGLE_SI_PAIR=1 SW_SHAPES=18x4 timeout 300 python scripts/scan_shape_sweep.py 2>&1 | tail -1
