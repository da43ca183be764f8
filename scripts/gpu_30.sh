#!/bin/bash
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "config3" 2>&1 | tail -15
