#!/bin/bash
timeout 900 python -m pytest tests/test_ingest.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "reshape or recycled or createMt" 2>&1 | tail -15
