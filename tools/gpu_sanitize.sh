#!/bin/bash
# compute-sanitizer over the smoke run and the small-shape GPU parity tests (memcheck), and
# racecheck / synccheck over the smoke run.  Logs go to gpurun_out/sanitize_*.log.
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
tail -3 gpurun_out/sanitize_memcheck_smoke.log
timeout 1500 $SAN --tool memcheck --error-exitcode 9 python -m pytest -x -q tests/test_gpu_families.py tests/test_gpu_parity.py -m gpu -k "not full_size and not hundred" > gpurun_out/sanitize_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"
tail -4 gpurun_out/sanitize_memcheck_tests.log
timeout 900 $SAN --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?"
tail -3 gpurun_out/sanitize_synccheck_smoke.log
timeout 900 $SAN --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
tail -6 gpurun_out/sanitize_racecheck_smoke.log
