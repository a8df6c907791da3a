#!/bin/bash
mkdir -p gpurun_out/r2p
R=gpurun_out/r2p; rm -f $R/msd.log
timeout 900 python -m pytest tests/test_gpu_msd.py tests/test_gpu_classes.py -m gpu -x -q > $R/pytest.log 2>&1
tail -3 $R/pytest.log
run() { echo "== $1" >> $R/msd.log; shift; env "$@" timeout 300 python tools/profile_msd.py 100000 5000 3 2>&1 | tail -1 >> $R/msd.log; }
run "warp scan" A=1
run "serial scan" AMOFB_LIB=experiments/build/libamofb_serialscan.so
cat $R/msd.log
