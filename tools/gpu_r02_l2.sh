#!/bin/bash
# round 2: table stream with an L2 evict_last policy -- time over the shapes and DRAM bytes of the headline launch
mkdir -p gpurun_out
timeout 600 python tools/bench_variants.py > gpurun_out/variants_l2.jsonl 2> gpurun_out/variants_l2.err; echo "variants rc=$?"
python tools/show_variants.py gpurun_out/variants_l2.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:predict_kernel -s 3 -c 2 --csv --log-file gpurun_out/l2_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/l2_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/l2_metrics.csv | cut -d, -f5,13- | head -12
