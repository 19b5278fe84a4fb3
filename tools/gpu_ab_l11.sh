#!/bin/bash
# A/B of builds of the leauthaud11 kernel on one box: build/lib_$v.so for v in $VARIANTS (plus the
# in-tree build as "tree"): parity tests of the family, then tools/bench_families.py
mkdir -p gpurun_out
cp tabcorr_b200/libtabcorr_b200.so /tmp/lib_keep.so
for v in tree $VARIANTS tree $VARIANTS; do
  if [ "$v" = tree ]; then cp /tmp/lib_keep.so tabcorr_b200/libtabcorr_b200.so; else cp build/lib_$v.so tabcorr_b200/libtabcorr_b200.so; fi
  python -m pytest tests/test_gpu_families.py -x -q -m gpu -k "occupation_matches" 2>&1 | tail -1
  timeout 300 python tools/bench_families.py --only leauthaud11 2> gpurun_out/abl11_$v.err | tee -a gpurun_out/abl11_$v.jsonl | cut -c150-330
done
cp /tmp/lib_keep.so tabcorr_b200/libtabcorr_b200.so
