#!/bin/bash
# round-1g measurement pass: tests, smoke, bench (both arms, with the forward search), ncu launch list + full captures
mkdir -p gpurun_out
S=gpurun_out/summary_r1g.txt
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "$name exit $?" | tee -a $S; tail -${TAILN:-2} gpurun_out/$name.log | cut -c1-300; }
rm -f $S
TAILN=4 TMO=900 run r1g_tests python -m pytest tests -q -m gpu -p no:cacheprovider
TMO=300 run r1g_smoke python __graft_entry__.py smoke
TMO=600 run r1g_bench_c3_reference python bench.py --impl reference --workload c3 --steps 2 --warmup 1
TMO=600 run r1g_bench_c2_n1 python bench.py --workload c2 --steps 5 --warmup 3
TMO=900 run r1g_bench_c3_n1 python bench.py --steps 3 --warmup 3 --search-host
python bench.py --steps 1 --warmup 1 --no-cpu --no-search > gpurun_out/r1g_launch_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1g_launches_bench_c3.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-search > gpurun_out/r1g_launch_ncu.log 2>&1
echo "launch list exit $?" | tee -a $S
python scripts/prof_c3.py > gpurun_out/r1g_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'decode_kb|syrk_i8|scan_i8_kernel|prep_i8_kernel|transpose_kb128|gemv_i8' -o gpurun_out/prof_r1g_c3 -f python scripts/prof_c3.py > gpurun_out/r1g_prof_ncu.log 2>&1
echo "ncu full exit $?" | tee -a $S
SW_N=4000 python scripts/ingest_bench.py > gpurun_out/r1g_ingest_plain.log 2>&1 && \
SW_N=4000 ncu --set full --clock-control none --import-source on -k regex:'tok_|encode_ascii' -c 12 -o gpurun_out/prof_r1g_ingest -f python scripts/ingest_bench.py > gpurun_out/r1g_ingest_ncu.log 2>&1
echo "ncu ingest exit $?" | tee -a $S
python scripts/ingest_bench.py > gpurun_out/r1g_ingest_bench.log 2>&1; echo "ingest bench exit $?" | tee -a $S
cat $S
