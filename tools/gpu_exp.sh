#!/bin/bash
# experiment: GPU parity tests + kernel variants (tuning knobs in $1, shapes filter in $2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_variants.py --tune "$1" ${2:+--only "$2"} > gpurun_out/variants.jsonl 2> gpurun_out/variants.err; echo "variants rc=$?"
python tools/show_variants.py gpurun_out/variants.jsonl 2>/dev/null || cat gpurun_out/variants.jsonl
tail -3 gpurun_out/variants.err
