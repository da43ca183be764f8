#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_ingest.py -q -m gpu -p no:cacheprovider -x -k "path_cache or createMt or reshape_and or recycled or demo" 2>&1 | tail -5
