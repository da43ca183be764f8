#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_algebra.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -25
