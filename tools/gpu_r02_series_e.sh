#!/bin/bash
# round 2: build variants (build/lib_*.so) x item policy (SERIES_FUSED) over the shapes of bench_variants
mkdir -p gpurun_out
for v in default u1 u4 k4 d4u4; do
  cp build/lib_$v.so tabcorr_b200/libtabcorr_b200.so
  timeout 600 python tools/bench_variants.py --tune "SERIES_FUSED=1;SERIES_FUSED=0" > gpurun_out/variants_f_$v.jsonl 2> gpurun_out/variants_f_$v.err; echo "variants $v rc=$?"
  python tools/show_variants.py gpurun_out/variants_f_$v.jsonl
done
