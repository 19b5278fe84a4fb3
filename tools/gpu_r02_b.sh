#!/bin/bash
# round 2: bench line with the configs block + knob sweep on the small shapes
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-sample 2000 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"
cat gpurun_out/bench_r02a.json; tail -3 gpurun_out/bench_r02a.err
for only in "N=60 R=19 wp" "N=120" "cfg2 N=240"; do
timeout 600 python tools/bench_variants.py --reps 5 --only "$only" --tune ";CHUNKS=20;CHUNKS=40;CHUNKS=144;OCC_ITEMS=7;OCC_ITEMS=28;OCC_SPREAD=50;OCC_SPREAD=90" >> gpurun_out/variants_knobs.jsonl 2>> gpurun_out/variants_knobs.err; echo "variants rc=$?"
done
python tools/show_variants.py gpurun_out/variants_knobs.jsonl
