#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_msd.py tests/test_gpu_classes.py tests/test_gpu_bad.py -m gpu -x -q > $O/pytest.log 2>&1; tail -15 $O/pytest.log
timeout 600 python tools/profile_msd.py 100000 5000 2 > $O/msd_100k.log 2>&1; cat $O/msd_100k.log
timeout 600 python tools/profile_msd.py 100000 5000 2 legacy > $O/msd_100k_legacy.log 2>&1; cat $O/msd_100k_legacy.log
