#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/tcgen_launches.csv python tools/tcgen_check.py 100000 > gpurun_out/tcgen_ncu.log 2>&1; echo "rc=$?"
grep -E "weights_image|tcgen_contract|finalize|predict_kernel" gpurun_out/tcgen_launches.csv | awk -F'","' '{print $5, $NF}' | head -60
