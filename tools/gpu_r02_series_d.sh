#!/bin/bash
# round 2: series items with the compact node path / rolled term loops / halving reduction:
# shapes of tools/bench_variants.py, GPU tests, ncu capture at N=60
mkdir -p gpurun_out
timeout 600 python tools/bench_variants.py --tune "${TUNES:-}" > gpurun_out/variants_series_e.jsonl 2> gpurun_out/variants_series_e.err; echo "variants rc=$?"
python tools/show_variants.py gpurun_out/variants_series_e.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 3 -c 1 \
    -o gpurun_out/prof_series_theta_30x1x19 -f python tools/run_occ_input.py theta 30 1 19 > gpurun_out/ncu_series_theta_30x1x19.log 2>&1
echo "ncu rc=$?"
