#!/bin/bash
# round 2: wide MSD window kernel, guard test, per-frame take, native parser
mkdir -p gpurun_out/r2n
R=gpurun_out/r2n
timeout 900 python -m pytest tests/test_gpu_msd.py tests/test_gpu_guard.py tests/test_gpu_classes.py -m gpu -x -q > $R/pytest.log 2>&1
tail -5 $R/pytest.log
for kb in 6 8 10; do
  echo "== wide KB=$kb" >> $R/msd.log
  AMOFB_MSD_DEBUG=1 AMOFB_MSD_WIDE_KB=$kb timeout 300 python tools/profile_msd.py 100000 5000 2 >> $R/msd.log 2>&1
done
echo "== narrow" >> $R/msd.log
AMOFB_MSD_NO_WIDE=1 timeout 300 python tools/profile_msd.py 100000 5000 2 >> $R/msd.log 2>&1
grep -v "^\[amofb msd\] window kernel shape" $R/msd.log | tail -30
timeout 600 python tools/profile_stream.py 400 > $R/stream.log 2>&1; tail -4 $R/stream.log
