#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_msd.py -m gpu -x -q 2>&1 | tail -2
for v in mkb8 nkb4 nkb6 nkb8; do echo "== $v"; AMOFB_LIB=experiments/build/libamofb_$v.so python tools/profile_msd.py 100000 5000 2 2>&1 | tail -1; done | tee $O/sweep2.log
