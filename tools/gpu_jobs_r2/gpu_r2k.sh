#!/bin/bash
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bad.py -m gpu -x -q 2>&1 | tail -2
for w in 1.0 1.5; do echo "== widen $w"; AMOFB_BAD_CELL_WIDEN=$w python tools/profile_bad.py 1000 3 | tail -1; done | tee $O/sweep.log
AMOFB_BAD_CELL_WIDEN=1.0 ncu --set full --clock-control none --import-source on -k regex:k_bad_search -s 3 -c 1 -o $O/prof_search -f python tools/profile_bad.py 500 2 > $O/ncu.log 2>&1
ncu -i $O/prof_search.ncu-rep --page raw --csv > $O/prof_search_raw.csv; ncu -i $O/prof_search.ncu-rep --page source --csv > $O/prof_search_src.csv
python tools/ncu_summary.py $O/prof_search_raw.csv $O/prof_search_src.csv
