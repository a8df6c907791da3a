#!/bin/bash
mkdir -p gpurun_out/r2r
R=gpurun_out/r2r; rm -f $R/msd.log
run() { echo "== $1" >> $R/msd.log; shift; env "$@" 2>&1 | tail -1 >> $R/msd.log; }
run "100k x 512 block" AMOFB_MSD_NO_COLUMN_COMMIT=1 timeout 300 python tools/profile_msd.py 100000 512 3
run "100k x 512 reg" timeout 300 python tools/profile_msd.py 100000 512 3
run "100k x 2048 block" AMOFB_MSD_NO_COLUMN_COMMIT=1 timeout 300 python tools/profile_msd.py 100000 2048 3
run "100k x 2048 reg" timeout 300 python tools/profile_msd.py 100000 2048 3
run "100k x 5000 block" AMOFB_MSD_NO_COLUMN_COMMIT=1 timeout 300 python tools/profile_msd.py 100000 5000 3
run "979200 x 512 block" AMOFB_MSD_NO_COLUMN_COMMIT=1 timeout 300 python tools/profile_msd.py 979200 512 3
run "979200 x 512 reg" timeout 300 python tools/profile_msd.py 979200 512 3
cat $R/msd.log
