#!/bin/bash
# Round 2: everything the records under profiles/ come from, in one GPU call (outputs under gpurun_out/rec/).
mkdir -p gpurun_out/rec
R=gpurun_out/rec
python -m pytest tests -m gpu -x -q > $R/pytest_gpu.log 2>&1
( time python bench.py --steps 20 --warmup 5 > $R/bench_n1.json 2> $R/bench_n1.err ) 2> $R/time_n1.txt
( time python bench.py --impl reference --steps 20 --warmup 5 > $R/bench_ref_n1.json 2> $R/bench_ref_n1.err ) 2> $R/time_ref_n1.txt
python -c "import __graft_entry__ as g; g.smoke()" > $R/smoke.log 2>&1
# launch list of the bench command itself (shorter: 2 steps, 3 warm-up; per-launch times under ncu are serialised and cold)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $R/launches_bench.csv python bench.py --steps 2 --warmup 3 > $R/ncu_bench.log 2>&1
python tools/profile_bad.py 2000 3 > $R/bad.log 2>&1
python tools/profile_msd.py 100000 5000 3 > $R/msd.log 2>&1
# (python tools/profile_stream.py 400 > $R/stream.log: profiles/r02_stream.txt)
tail -2 $R/pytest_gpu.log; cut -c1-400 $R/bench_n1.json; cut -c1-300 $R/bench_ref_n1.json; tail -1 $R/smoke.log; tail -1 $R/bad.log; tail -1 $R/msd.log; grep real $R/time_n1.txt $R/time_ref_n1.txt
