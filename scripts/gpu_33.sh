#!/bin/bash
for sr in 12 4 24 37 74 148; do echo "srows $sr"; EAGLE_PREP_SROWS=$sr SW_REP=5 SW_MODES=i8 timeout 200 python scripts/prof_prep.py 2>&1 | tail -1; done
