#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_prep.py > gpurun_out/prep17_plain.log 2>&1 && \
SW_REP=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prep17_launches.csv python scripts/prof_prep.py > gpurun_out/prep17_ncu1.log 2>&1
SW_REP=1 SW_MODES=i8 ncu --set full --clock-control none --import-source on -k regex:prep_i8 -c 2 -o gpurun_out/prof_r1f_prep python scripts/prof_prep.py > gpurun_out/prep17_ncu2.log 2>&1
cat gpurun_out/prep17_plain.log; tail -3 gpurun_out/prep17_ncu2.log
