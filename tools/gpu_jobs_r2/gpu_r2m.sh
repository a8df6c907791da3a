#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 600 python tools/profile_stream.py 400 > $O/stream.log 2>&1; cat $O/stream.log
SEL="tests/test_gpu_pair.py::test_zif4_golden tests/test_gpu_pair.py::test_rdf_and_cn_random_boxes tests/test_gpu_pair.py::test_kernel_variants_agree tests/test_gpu_pair.py::test_cutoffs_beyond_rmax tests/test_gpu_bad.py tests/test_gpu_msd.py::test_streaming_path_matches_oracle tests/test_gpu_msd.py::test_window_msd"
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest $SEL -m gpu -x -q ) > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/sanitizer_memcheck.log; tail -6 $O/sanitizer_memcheck.log
( time timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_pair.py::test_zif4_golden "tests/test_gpu_pair.py::test_rdf_and_cn_random_boxes" tests/test_gpu_bad.py::test_zif4_golden tests/test_gpu_bad.py::test_random_boxes "tests/test_gpu_msd.py::test_streaming_path_matches_oracle" -m gpu -x -q ) > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/sanitizer_racecheck.log; tail -6 $O/sanitizer_racecheck.log
