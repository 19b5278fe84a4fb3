#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tcgen_contract -s 2 -c 1 -o gpurun_out/prof_tcgen -f python tools/run_tcgen_only.py > gpurun_out/ncu_tcgen.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/ncu_tcgen.log
