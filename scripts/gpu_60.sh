#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -6
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-120
