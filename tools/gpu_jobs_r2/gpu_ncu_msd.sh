#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 300 python tools/profile_msd.py 30000 5000 1 > gpurun_out/ncu/msd_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msd_window_wide -c 1 -o gpurun_out/ncu/prof_msd_wide -f python tools/profile_msd.py 30000 5000 1 > gpurun_out/ncu/ncu_msd_wide.log 2>&1
tail -3 gpurun_out/ncu/ncu_msd_wide.log
