#!/bin/bash
mkdir -p gpurun_out/r2t
R=gpurun_out/r2t
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $R/lists.csv python tools/profile_bad.py 900 2 > $R/lists.log 2>&1
AMOFB_BAD_ONE_LIST=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $R/one.csv python tools/profile_bad.py 900 2 > $R/one.log 2>&1
AMOFB_BAD_ONE_LIST=1 AMOFB_BAD_NO_CENTRE_LIST=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $R/old.csv python tools/profile_bad.py 900 2 > $R/old.log 2>&1
tail -1 $R/lists.log $R/one.log $R/old.log
