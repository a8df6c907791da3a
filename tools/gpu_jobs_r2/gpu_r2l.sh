#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
python tools/profile_bad.py 2000 3 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file $O/launches_c4.csv python tools/profile_bad.py 1000 2 > /dev/null 2>&1
python tools/launch_shares.py $O/launches_c4.csv 2>/dev/null | head -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/profile_cn.py c2 1000 2>&1 | tail -2
