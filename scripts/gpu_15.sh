#!/bin/bash
mkdir -p gpurun_out
TMO=400 timeout 400 python scripts/scan_sustained.py > gpurun_out/sustained15.log 2>&1; echo "exit $?"; cat gpurun_out/sustained15.log | tail -12
