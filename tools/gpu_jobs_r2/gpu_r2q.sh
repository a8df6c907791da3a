#!/bin/bash
mkdir -p gpurun_out/r2q
R=gpurun_out/r2q
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_msd_slab -c 40 --csv --log-file $R/col.csv python tools/profile_msd.py 100000 5000 1 > $R/col.log 2>&1
AMOFB_MSD_NO_COLUMN_COMMIT=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_msd_slab -c 40 --csv --log-file $R/blk.csv python tools/profile_msd.py 100000 5000 1 > $R/blk.log 2>&1
tail -2 $R/col.log
