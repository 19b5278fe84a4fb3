#!/bin/bash
# GPU parity tests + headline bench + tuning-knob sweep over the 'wp' shapes (N=60/120/240/500)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python tools/bench_variants.py --reps 7 --only wp --tune "$1" > gpurun_out/variants.jsonl 2> gpurun_out/variants.err; echo "variants rc=$?"
python tools/show_variants.py gpurun_out/variants.jsonl 2>/dev/null || cat gpurun_out/variants.jsonl
tail -3 gpurun_out/variants.err
