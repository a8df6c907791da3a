#!/bin/bash
# pair-kernel iteration: tests, variant sweep, one ncu capture (run under gpurun)
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pair.py tests/test_gpu_classes.py -m gpu -x -q > $O/pytest_pair.log 2>&1; tail -5 $O/pytest_pair.log
bash tools/sweep_variants.sh c2 214 k2 p3k2 p3k1 p3k2t1024 > $O/sweep_c2.log 2>&1; cat $O/sweep_c2.log
bash tools/sweep_variants.sh c3 20 k2 p3k2 p3k1 > $O/sweep_c3.log 2>&1; cat $O/sweep_c3.log
AMOFB_LIB=experiments/build/libamofb_p3k2.so ncu --set full --clock-control none --import-source on -k regex:k_pair_tiled -s 1 -c 1 -o $O/prof_p3k2 -f python tools/profile_pair.py c2 107 2 > $O/ncu_p3k2.log 2>&1
ncu -i $O/prof_p3k2.ncu-rep --page raw --csv > $O/prof_p3k2_raw.csv; ncu -i $O/prof_p3k2.ncu-rep --page source --csv > $O/prof_p3k2_src.csv
python tools/ncu_summary.py $O/prof_p3k2_raw.csv $O/prof_p3k2_src.csv
