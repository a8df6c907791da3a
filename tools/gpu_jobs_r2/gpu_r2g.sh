#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_msd.py tests/test_gpu_classes.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python tools/profile_msd.py 100000 5000 3 > $O/msd_100k.log 2>&1; cat $O/msd_100k.log
timeout 900 python tools/profile_msd.py 979200 5000 3 > $O/msd_full.log 2>&1; cat $O/msd_full.log
