#!/bin/bash
mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_pair.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
echo "== dynamic"; python tools/profile_pair.py c2 214 3 | tail -1
echo "== static"; AMOFB_LIB=experiments/build/libamofb_static.so python tools/profile_pair.py c2 214 3 | tail -1
done
echo "== dynamic c3"; python tools/profile_pair.py c3 20 3 | tail -1
echo "== static c3"; AMOFB_LIB=experiments/build/libamofb_static.so python tools/profile_pair.py c3 20 3 | tail -1
