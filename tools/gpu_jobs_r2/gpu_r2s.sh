#!/bin/bash
mkdir -p gpurun_out/r2s
R=gpurun_out/r2s; rm -f $R/*.log
timeout 900 python -m pytest tests/test_gpu_bad.py tests/test_gpu_classes.py tests/test_gpu_guard.py -m gpu -x -q > $R/pytest.log 2>&1
tail -3 $R/pytest.log
run() { echo "== $1" >> $R/bad.log; shift; env "$@" timeout 300 python tools/profile_bad.py 2000 3 2>&1 | tail -1 >> $R/bad.log; }
run "lists per species" A=1
run "one list" AMOFB_BAD_ONE_LIST=1
cat $R/bad.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $R/lists.csv python tools/profile_bad.py 900 2 > $R/lists_ncu.log 2>&1
AMOFB_BAD_ONE_LIST=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 20 -c 60 --csv --log-file $R/one.csv python tools/profile_bad.py 900 2 > $R/one_ncu.log 2>&1
