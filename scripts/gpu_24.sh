#!/bin/bash
echo c2; SW_N=2000 SW_L=500000 SW_SHAPES=37x4 timeout 300 python scripts/scan_shape_sweep.py 2>&1 | tail -5
echo c3; SW_SHAPES=37x4 timeout 300 python scripts/scan_shape_sweep.py 2>&1 | tail -5
