#!/bin/bash
# all BASELINE configurations on one GPU (+ N=500 tile-shape variants)
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py --sweep-draws 8388608 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
cat gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.err
timeout 300 python tools/bench_variants.py --only "cfg5" --tune "NBUF=2;NBUF=1" > gpurun_out/variants_n500.jsonl 2>&1
python tools/show_variants.py gpurun_out/variants_n500.jsonl
