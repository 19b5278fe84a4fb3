#!/bin/bash
# A/B of two builds on one box: build/lib_$A.so vs build/lib_$B.so over the shapes of bench_variants
mkdir -p gpurun_out
cp tabcorr_b200/libtabcorr_b200.so /tmp/lib_keep.so
for v in $A $B $A $B; do
  cp build/lib_$v.so tabcorr_b200/libtabcorr_b200.so
  timeout 600 python tools/bench_variants.py > gpurun_out/ab_$v.jsonl 2> gpurun_out/ab_$v.err; echo "$v rc=$?"
  python tools/show_variants.py gpurun_out/ab_$v.jsonl
done
cp /tmp/lib_keep.so tabcorr_b200/libtabcorr_b200.so
