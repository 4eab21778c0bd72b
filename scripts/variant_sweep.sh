#!/bin/bash
# Development probe: cfg3 round-1 profile for a list of nvcc flag sets (each a quoted string).
for flags in "$@"; do python scripts/variant_perf.py $flags 2>&1 | tail -1; done
