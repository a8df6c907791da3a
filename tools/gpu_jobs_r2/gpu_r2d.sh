#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pair.py tests/test_gpu_classes.py -m gpu -x -q > $O/pytest_pair.log 2>&1; tail -3 $O/pytest_pair.log
bash tools/sweep_variants.sh c2 214 k2 l2 l1 l2u2 l3 > $O/sweep_c2.log 2>&1; grep "===\|rep" $O/sweep_c2.log
bash tools/sweep_variants.sh c3 20 k2 l2 l1 l3 > $O/sweep_c3.log 2>&1; grep "===\|rep" $O/sweep_c3.log
AMOFB_LIB=experiments/build/libamofb_l2.so ncu --set full --clock-control none --import-source on -k regex:k_pair_tiled -s 1 -c 1 -o $O/prof_l2 -f python tools/profile_pair.py c2 107 2 > $O/ncu_l2.log 2>&1
ncu -i $O/prof_l2.ncu-rep --page raw --csv > $O/prof_l2_raw.csv; ncu -i $O/prof_l2.ncu-rep --page source --csv > $O/prof_l2_src.csv
python tools/ncu_summary.py $O/prof_l2_raw.csv $O/prof_l2_src.csv
