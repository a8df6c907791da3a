#!/bin/bash
# Build the library with several compile-time configurations of the tiled pair kernel and time each on C2 frames.
# usage (on the GPU box): bash tools/sweep_pair.sh "<flags 1>" "<flags 2>" ...
for cfg in "$@"; do
    echo "=== $cfg"
    AMOFB_NVCC_FLAGS="$cfg" python amof_b200/build.py --force > /dev/null 2>&1 || { echo "build failed"; continue; }
    python tools/profile_pair.py c2 214 2 | tail -1
done
