#!/bin/bash
timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "decode or kblocked or readblock or mmt" 2>&1 | tail -3
timeout 200 python scripts/decode_bench.py 2>&1 | tail -6
SW_N=2000 SW_L=500000 timeout 200 python scripts/decode_bench.py 2>&1 | tail -6
