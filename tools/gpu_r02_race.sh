#!/bin/bash
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/race_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/race_smoke.log
for parts in theta occ massdep; do
  PARTS=$parts timeout 900 $SAN --tool racecheck --racecheck-report analysis python tools/run_series_small.py > gpurun_out/race_$parts.log 2>&1; echo "$parts rc=$?"; tail -2 gpurun_out/race_$parts.log
done
TC_TUNE_SERIES_FUSED=0 PARTS=theta timeout 900 $SAN --tool racecheck --racecheck-report analysis python tools/run_series_small.py > gpurun_out/race_theta_nodes.log 2>&1; echo "theta nodes rc=$?"; tail -2 gpurun_out/race_theta_nodes.log
