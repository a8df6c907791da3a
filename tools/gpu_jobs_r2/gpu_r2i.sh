#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
AMOFB_LIB=experiments/build/libamofb_mkb8.so ncu --set full --clock-control none --import-source on -k regex:k_msd_window_soa -c 1 -o $O/prof_soa8 -f python tools/profile_msd.py 30000 5000 1 > $O/ncu.log 2>&1
ncu -i $O/prof_soa8.ncu-rep --page raw --csv > $O/prof_soa8_raw.csv; ncu -i $O/prof_soa8.ncu-rep --page source --csv > $O/prof_soa8_src.csv
python tools/ncu_summary.py $O/prof_soa8_raw.csv $O/prof_soa8_src.csv
