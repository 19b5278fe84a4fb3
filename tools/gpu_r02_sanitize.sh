#!/bin/bash
# round 2: compute-sanitizer over what round 2 added to the kernels (tools/run_series_small.py) and
# memcheck over the GPU tests that reach them
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck; do
  timeout 900 $SAN --tool $tool --error-exitcode 9 python tools/run_series_small.py > gpurun_out/sanitize_r02_${tool}_series.log 2>&1; echo "$tool series rc=$?"
  tail -3 gpurun_out/sanitize_r02_${tool}_series.log
done
timeout 1200 $SAN --tool racecheck --racecheck-report analysis python tools/run_series_small.py > gpurun_out/sanitize_r02_racecheck_series.log 2>&1; echo "racecheck series rc=$?"
tail -5 gpurun_out/sanitize_r02_racecheck_series.log
timeout 1500 $SAN --tool memcheck --error-exitcode 9 python -m pytest -x -q tests/test_gpu_families.py tests/test_halo_bins.py tests/test_gpu_parity.py -m gpu -k "not full_size and not hundred" > gpurun_out/sanitize_r02_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"
tail -4 gpurun_out/sanitize_r02_memcheck_tests.log
