#!/bin/bash
# memcheck of the ingest kernels on the small cases, then the whole GPU suite
EAGLE_INGEST_PIECE_BYTES=3000 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ingest.py -q -m gpu -p no:cacheprovider -x -k "matches_oracle and None or createMt" > gpurun_out/memcheck45.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck45.log
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t45_all.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/t45_all.log
