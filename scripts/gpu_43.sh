#!/bin/bash
SW_N=2000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tok_emit -c 2 -o gpurun_out/prof_ingest -f python scripts/ingest_bench.py > gpurun_out/prof_ingest.log 2>&1
tail -3 gpurun_out/prof_ingest.log
