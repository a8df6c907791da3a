#!/bin/bash
# Build with several compile-time configurations of the MSD window kernel and time C5 (100k atoms x 5000 frames).
for cfg in "$@"; do
    echo "=== $cfg"
    AMOFB_NVCC_FLAGS="$cfg" python amof_b200/build.py --force > /dev/null 2>&1 || { echo "build failed"; continue; }
    python bench.py --workload c5 --atoms 100000 --frames 5000 --steps 2 2>&1 | tail -1 | cut -c1-140
done
