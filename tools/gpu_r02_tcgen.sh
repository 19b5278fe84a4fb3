#!/bin/bash
# round 2: first run of the tcgen05 contraction -- small batch first (bounded), then full size
mkdir -p gpurun_out
timeout 120 python tools/tcgen_check.py 4096 > gpurun_out/tcgen_small.jsonl 2> gpurun_out/tcgen_small.err; echo "small rc=$?"
cat gpurun_out/tcgen_small.jsonl; tail -5 gpurun_out/tcgen_small.err
timeout 300 python tools/tcgen_check.py 100000 > gpurun_out/tcgen_full.jsonl 2> gpurun_out/tcgen_full.err; echo "full rc=$?"
cat gpurun_out/tcgen_full.jsonl; tail -5 gpurun_out/tcgen_full.err
