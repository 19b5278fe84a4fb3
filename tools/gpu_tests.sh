#!/bin/bash
# GPU parity tests only (optionally a -k expression in $K)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q ${K:+-k "$K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
