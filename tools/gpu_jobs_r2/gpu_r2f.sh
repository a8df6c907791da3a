#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
AMOFB_MSD_DEBUG=1 python tools/profile_msd.py 100000 5000 2 2>&1 | grep -v "^\[amofb msd\]" | tail -1; AMOFB_MSD_DEBUG=1 python tools/profile_msd.py 100000 5000 1 2>&1 | grep "amofb msd" | head -1
for shape in "13 2 8" "13 2 4" "11 3 5" "9 3 5" "7 4 4" "5 5 3" "13 1 16" "7 2 8"; do
  set -- $shape
  echo "== nwt $1 ng $2 wpg $3"; AMOFB_MSD_AP_NWT=$1 AMOFB_MSD_AP_NG=$2 AMOFB_MSD_AP_WPG=$3 python tools/profile_msd.py 100000 5000 2 2>&1 | tail -1
done
ncu --set full --clock-control none --import-source on -k regex:k_msd_window_soa -c 1 -o $O/prof_msd_soa -f python tools/profile_msd.py 30000 5000 1 > $O/ncu_soa.log 2>&1
ncu -i $O/prof_msd_soa.ncu-rep --page raw --csv > $O/prof_msd_soa_raw.csv; ncu -i $O/prof_msd_soa.ncu-rep --page source --csv > $O/prof_msd_soa_src.csv
python tools/ncu_summary.py $O/prof_msd_soa_raw.csv $O/prof_msd_soa_src.csv
ncu --set full --clock-control none --import-source on -k regex:k_msd_slab_commit -s 3 -c 1 -o $O/prof_msd_commit -f python tools/profile_msd.py 100000 2000 1 > $O/ncu_commit.log 2>&1
ncu -i $O/prof_msd_commit.ncu-rep --page raw --csv > $O/prof_msd_commit_raw.csv; ncu -i $O/prof_msd_commit.ncu-rep --page source --csv > $O/prof_msd_commit_src.csv
python tools/ncu_summary.py $O/prof_msd_commit_raw.csv $O/prof_msd_commit_src.csv | head -30
