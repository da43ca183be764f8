#!/bin/bash
SW_N=4000 ncu --set full --clock-control none --import-source on -k regex:'tok_emit|encode_ascii' -c 3 -o gpurun_out/prof_r1g_ingest2 -f python scripts/ingest_bench.py > gpurun_out/r1g_ingest_ncu2.log 2>&1; echo rc=$?
