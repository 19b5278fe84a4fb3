#!/bin/bash
# ncu full capture of the fused kernel on precomputed occupations (contraction only) at N=60 and of the theta path
mkdir -p gpurun_out
for kind in occ theta; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 3 -c 1 \
    -o gpurun_out/prof2_${kind}_30x1x19 -f python tools/run_occ_input.py $kind 30 1 19 > gpurun_out/ncu2_${kind}_30x1x19.log 2>&1
echo "ncu $kind rc=$?"
done
