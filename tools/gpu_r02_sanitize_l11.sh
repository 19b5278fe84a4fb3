#!/bin/bash
# round 2: compute-sanitizer over the rewritten leauthaud11 occupation kernel (tools/run_l11_small.py)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 9 python tools/run_l11_small.py > gpurun_out/sanitize_r02_${tool}_l11.log 2>&1; echo "$tool l11 rc=$?"
  tail -2 gpurun_out/sanitize_r02_${tool}_l11.log
done
timeout 900 $SAN --tool racecheck --racecheck-report analysis --kernel-name kernel_substring=occupation_l11 python tools/run_l11_small.py > gpurun_out/sanitize_r02_racecheck_l11.log 2>&1; echo "racecheck l11 rc=$?"
tail -4 gpurun_out/sanitize_r02_racecheck_l11.log
