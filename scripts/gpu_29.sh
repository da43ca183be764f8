#!/bin/bash
SW_WORLD=2 timeout 200 python scripts/prof_prep_cols.py 2>&1 | tail -4
SW_WORLD=8 timeout 200 python scripts/prof_prep_cols.py 2>&1 | tail -8
