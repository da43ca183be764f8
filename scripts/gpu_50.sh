#!/bin/bash
# full default bench lines for the record (c3 with search + host route, c2), reference arm
timeout 1200 python bench.py --search-host > gpurun_out/r1g_bench_c3_n1.log 2>gpurun_out/r1g_bench_c3_n1.err; echo c3 rc=$?
timeout 600 python bench.py --workload c2 > gpurun_out/r1g_bench_c2_n1.log 2>gpurun_out/r1g_bench_c2_n1.err; echo c2 rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1g_bench_c3_reference.log 2>&1; echo ref rc=$?
tail -2 gpurun_out/r1g_bench_c3_n1.err
