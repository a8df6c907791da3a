#!/bin/bash
# Build libamofb.so with several compile-time configurations HERE (nvcc cross-compiles without a GPU) so that one GPU
# call can time them all: experiments/build/libamofb_<tag>.so, picked up by tools/profile_*.py through AMOFB_LIB.
# usage: bash tools/build_variants.sh tag1 "<nvcc flags 1>" tag2 "<nvcc flags 2>" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p experiments/build
cp amof_b200/libamofb.so /tmp/libamofb_keep.so 2>/dev/null || true
while [ $# -ge 2 ]; do
    tag=$1; flags=$2; shift 2
    echo "=== $tag: $flags"
    AMOFB_NVCC_FLAGS="$flags" python amof_b200/build.py --force > /dev/null
    cp amof_b200/libamofb.so experiments/build/libamofb_$tag.so
done
python amof_b200/build.py --force > /dev/null
