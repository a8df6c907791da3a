#!/bin/bash
# Everything the round's records under profiles/ come from, in one GPU call (outputs under gpurun_out/rec/).
mkdir -p gpurun_out/rec
R=gpurun_out/rec
python -m pytest tests -m gpu -x -q > $R/pytest_gpu.log 2>&1
python bench.py > $R/bench_c2.json 2> $R/bench_c2.err
python bench.py --impl reference > $R/bench_ref_c2.json 2> $R/bench_ref_c2.err
python bench.py --workload c3 > $R/bench_c3.json 2> $R/bench_c3.err
python bench.py --workload c4 > $R/bench_c4.json 2> $R/bench_c4.err
python bench.py --workload c5 > $R/bench_c5_full.json 2> $R/bench_c5_full.err
python -c "import __graft_entry__ as g; g.smoke()" > $R/smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file $R/launches_c2.csv python bench.py --steps 2 --warmup 3 --frames 2000 > $R/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 150 --csv --log-file $R/launches_c4.csv python bench.py --workload c4 --frames 1000 --steps 2 > $R/ncu_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_msd -c 80 --csv --log-file $R/launches_c5.csv python bench.py --workload c5 --atoms 100000 --frames 5000 --steps 1 > $R/ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bad -s 3 -c 1 -o $R/prof_bad -f python tools/profile_bad.py 500 2 > $R/ncu_bad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msd_window -s 1 -c 1 -o $R/prof_msd_window_ap -f python bench.py --workload c5 --atoms 30000 --frames 5000 --steps 1 --warmup 1 > $R/ncu_msd.log 2>&1
python tools/profile_cn.py c2 1000 > $R/cn_c2.log 2>&1
python tools/profile_cn.py c3 100 > $R/cn_c3.log 2>&1
tail -2 $R/pytest_gpu.log; for f in c2 ref_c2 c3 c4 c5_full; do cut -c1-180 $R/bench_$f.json; done; tail -1 $R/smoke.log
