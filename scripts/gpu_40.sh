#!/bin/bash
# ingest tests (tokeniser, encoder, createMt) under compute-sanitizer for the small cases, then plain
timeout 600 python -m pytest tests/test_ingest.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15
