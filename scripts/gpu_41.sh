#!/bin/bash
timeout 600 python scripts/ingest_bench.py 2>&1 | tail -8
