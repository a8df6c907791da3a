#!/bin/bash
# Build with several compile-time configurations of the bond-angle kernel and time C4.
for cfg in "$@"; do
    echo "=== $cfg"
    AMOFB_NVCC_FLAGS="$cfg" python amof_b200/build.py --force > /dev/null 2>&1 || { echo "build failed"; continue; }
    python tools/profile_bad.py 1000 3 | tail -1
done
