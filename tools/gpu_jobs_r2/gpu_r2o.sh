#!/bin/bash
mkdir -p gpurun_out/r2o
R=gpurun_out/r2o
timeout 900 python -m pytest tests/test_gpu_msd.py -m gpu -x -q > $R/pytest.log 2>&1
tail -5 $R/pytest.log
for kb in 6 8 10; do
  echo "== wide KB=$kb" >> $R/msd.log
  AMOFB_MSD_WIDE_KB=$kb timeout 300 python tools/profile_msd.py 100000 5000 3 >> $R/msd.log 2>&1
done
grep -v "^\[amofb msd\]" $R/msd.log | tail -30
