#!/bin/bash
mkdir -p gpurun_out/r2p
R=gpurun_out/r2p; rm -f $R/msd.log
timeout 900 python -m pytest tests/test_gpu_msd.py -m gpu -x -q > $R/pytest.log 2>&1
tail -3 $R/pytest.log
run() { echo "== $1" >> $R/msd.log; shift; env "$@" timeout 300 python tools/profile_msd.py 100000 5000 3 2>&1 | tail -1 >> $R/msd.log; }
run "reg commit ilp4" A=1
run "reg commit ilp8" AMOFB_LIB=experiments/build/libamofb_ilp8.so
run "reg commit ilp2" AMOFB_LIB=experiments/build/libamofb_ilp2.so
run "block commit" AMOFB_MSD_NO_COLUMN_COMMIT=1
cat $R/msd.log
