#!/bin/bash
# racecheck of the smoke run over older builds (bisect of the hazards racecheck started to report)
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
cp tabcorr_b200/libtabcorr_b200.so /tmp/lib_keep.so
for v in 9834dc8 a8b1812 pre_massdep; do
  cp build/lib_$v.so tabcorr_b200/libtabcorr_b200.so
  timeout 600 $SAN --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/race_smoke_$v.log 2>&1; echo "$v rc=$?"; tail -1 gpurun_out/race_smoke_$v.log
done
cp /tmp/lib_keep.so tabcorr_b200/libtabcorr_b200.so
