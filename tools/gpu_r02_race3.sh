#!/bin/bash
# racecheck of the smoke run of the code as it was before this round's kernel changes (copied to build/old_tree)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
cd build/old_tree
timeout 600 $SAN --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > ../../gpurun_out/race_smoke_old_tree.log 2>&1; echo "old tree rc=$?"; tail -2 ../../gpurun_out/race_smoke_old_tree.log
