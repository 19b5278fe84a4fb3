#!/bin/bash
# round 2: series evaluation of the occupations -- accuracy against the node path, then the parity
# tests, then the shapes of tools/bench_variants.py with and without the series
mkdir -p gpurun_out
timeout 900 python tools/series_check.py > gpurun_out/series_check.jsonl 2> gpurun_out/series_check.err; echo "series_check rc=$?"
cat gpurun_out/series_check.jsonl; tail -5 gpurun_out/series_check.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_variants.py --tune ";SERIES=0" > gpurun_out/variants_series.jsonl 2> gpurun_out/variants_series.err; echo "variants rc=$?"
python tools/show_variants.py gpurun_out/variants_series.jsonl
