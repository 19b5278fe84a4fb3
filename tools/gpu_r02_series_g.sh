#!/bin/bash
# round 2: item-count / spread knobs with series items
mkdir -p gpurun_out
timeout 900 python tools/bench_variants.py --tune "${TUNES:-}" > gpurun_out/variants_knobs2.jsonl 2> gpurun_out/variants_knobs2.err; echo "variants rc=$?"
python tools/show_variants.py gpurun_out/variants_knobs2.jsonl
