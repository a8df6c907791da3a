#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msd_slab_commit_reg -s 3 -c 1 -o gpurun_out/ncu/prof_commit_reg -f python tools/profile_msd.py 100000 2048 1 > gpurun_out/ncu/ncu_commit_reg.log 2>&1
tail -2 gpurun_out/ncu/ncu_commit_reg.log
