#!/bin/bash
timeout 600 python -m pytest tests/test_ingest.py -q -m gpu -p no:cacheprovider -x -k fuzz 2>&1 | tail -25
