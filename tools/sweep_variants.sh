#!/bin/bash
# Time prebuilt library variants (tools/build_variants.sh) on the GPU box: usage: bash tools/sweep_variants.sh <workload> <frames> tag...
wl=$1; fr=$2; shift 2
for tag in "$@"; do
    echo "=== $tag"
    AMOFB_LIB=experiments/build/libamofb_$tag.so timeout 300 python tools/profile_pair.py $wl $fr 3 | tail -1
done
