#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second
SW_REP=1 SW_MODES=i8 python scripts/prof_prep.py > gpurun_out/prep19_plain.log 2>&1 || exit 1
for cfg in "4 3" "16 2" "8 4" "4 6"; do set -- $cfg
  EAGLE_PREP_PHASE=$1 EAGLE_PREP_LAG=$2 SW_REP=1 SW_MODES=i8 ncu --metrics $M --clock-control none -k regex:prep_i8 -s 2 -c 2 --csv --log-file gpurun_out/prep19_$1_$2.csv python scripts/prof_prep.py > /dev/null 2>&1
done
echo done
