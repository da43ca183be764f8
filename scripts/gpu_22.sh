#!/bin/bash
mkdir -p gpurun_out
for fc in 1 0; do for pair in 0 1; do echo "flow=$fc pair=$pair"; EAGLE_SCAN_FLOWCTL=$fc EAGLE_SI_PAIR=$pair SW_N=2000 SW_L=500000 SW_SHAPES=37x4,18x4 timeout 300 python scripts/scan_shape_sweep.py 2>&1 | tail -2; done; done
