#!/bin/bash
mkdir -p gpurun_out/r2p
R=gpurun_out/r2p; rm -f $R/msd.log
run() { echo "== $1" >> $R/msd.log; shift; env "$@" timeout 300 python tools/profile_msd.py 100000 5000 3 2>&1 | tail -1 >> $R/msd.log; }
run "A=32 T=192" AMOFB_LIB=experiments/build/libamofb_a32.so
run "A=16 T=96" AMOFB_LIB=experiments/build/libamofb_a16.so
run "A=32 T=96" AMOFB_LIB=experiments/build/libamofb_a32t96.so
run "A=16 T=192" AMOFB_LIB=experiments/build/libamofb_a16t192.so
cat $R/msd.log
