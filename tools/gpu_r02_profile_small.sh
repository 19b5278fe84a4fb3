#!/bin/bash
# round 2, first call: baseline of every shape (tools/bench_variants.py) + ncu full captures of the
# fused kernel at N=60 and N=120 (parameter draws and precomputed occupations)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/nvsmi.txt
timeout 600 python tools/bench_variants.py > gpurun_out/variants_base.jsonl 2> gpurun_out/variants_base.err; echo "variants rc=$?"
cat gpurun_out/variants_base.jsonl
for cfg in "theta 30 1 19" "occ 30 1 19" "theta 60 1 20"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 3 -c 1 \
    -o gpurun_out/prof_$1_$2x$3x$4 -f python tools/run_occ_input.py $1 $2 $3 $4 > gpurun_out/ncu_$1_$2x$3x$4.log 2>&1
  echo "ncu $cfg rc=$?"
done
