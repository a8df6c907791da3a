#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bad.py tests/test_gpu_classes.py tests/test_gpu_neigh.py -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for w in 1.0 1.25 1.5 2.0 2.5; do echo "== widen $w"; AMOFB_BAD_CELL_WIDEN=$w python tools/profile_bad.py 1000 3 | tail -1; done | tee $O/sweep.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file $O/launches_c4.csv python tools/profile_bad.py 500 2 > /dev/null 2>&1
python tools/launch_shares.py $O/launches_c4.csv 2>/dev/null | head -12
