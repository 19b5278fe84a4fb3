#!/bin/bash
# GPU tests with the in-tree build, then A/B of two builds over the shapes of bench_variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh
