#!/bin/bash
mkdir -p gpurun_out
SW_N=2000 SW_L=500000 SW_SHAPES=37x4,16x9,74x2,148x1,24x6,8x18 timeout 300 python scripts/scan_shape_sweep.py 2>&1 | tail -8
