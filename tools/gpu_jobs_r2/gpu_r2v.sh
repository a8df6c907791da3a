#!/bin/bash
mkdir -p gpurun_out/r2v
bash tools/sweep_variants.sh c2 214 base t256x4 t256x3 t384x2 t1024 2>&1 | tee gpurun_out/r2v/c2.log
bash tools/sweep_variants.sh c3 20 base t256x4 t384x2 2>&1 | tee gpurun_out/r2v/c3.log
