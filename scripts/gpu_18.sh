#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary18.txt
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a $S; tail -${TAILN:-4} gpurun_out/$name.log; }
rm -f $S
TAILN=10 TMO=300 run t18_prep python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "preproducts"
for cfg in "4 3" "8 2" "2 4" "16 2" "8 4" "4 6"; do set -- $cfg
  echo "--- phase $1 lag $2"; EAGLE_PREP_PHASE=$1 EAGLE_PREP_LAG=$2 SW_REP=10 SW_MODES=i8 timeout 200 python scripts/prof_prep.py 2>&1 | tail -1
done
echo "--- no flow"; EAGLE_PREP_FLOWCTL=0 SW_REP=10 SW_MODES=i8 timeout 200 python scripts/prof_prep.py 2>&1 | tail -1
cat $S
