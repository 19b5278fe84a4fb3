#!/bin/bash
# ncu full capture of the predict kernel on the headline workload
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 3 -c 1 -o gpurun_out/prof_predict -f python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
