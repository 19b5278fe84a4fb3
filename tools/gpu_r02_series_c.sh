#!/bin/bash
# round 2: ncu capture of the fused kernel with series items at N=60 (30 x 1 x 19) and GPU tests
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 3 -c 1 \
    -o gpurun_out/prof_series_theta_30x1x19 -f python tools/run_occ_input.py theta 30 1 19 > gpurun_out/ncu_series_theta_30x1x19.log 2>&1
echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
