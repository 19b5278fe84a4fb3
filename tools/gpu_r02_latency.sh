#!/bin/bash
# round 2: latency of the one-draw / small-batch entries + GPU tests
mkdir -p gpurun_out
timeout 600 python tools/bench_small_batches.py > gpurun_out/small_batches_r02.jsonl 2> gpurun_out/small_batches_r02.err; echo "small rc=$?"
cut -c1-160 gpurun_out/small_batches_r02.jsonl
timeout 600 python tools/bench_interp_small.py > gpurun_out/interp_small_r02.jsonl 2> gpurun_out/interp_small_r02.err; echo "interp rc=$?"
cut -c1-200 gpurun_out/interp_small_r02.jsonl | head -12
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
