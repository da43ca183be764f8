#!/bin/bash
timeout 600 python -m pytest tests/test_ingest.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -5
timeout 600 python scripts/ingest_bench.py 2>&1 | tail -8
