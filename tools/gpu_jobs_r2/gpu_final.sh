#!/bin/bash
mkdir -p gpurun_out/final
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest_gpu.log 2>&1; tail -2 gpurun_out/final/pytest_gpu.log
timeout 120 python tools/profile_stream.py 400 > gpurun_out/final/stream.log 2>&1; tail -3 gpurun_out/final/stream.log
