#!/bin/bash
# Round 2, one gpurun call on 1 GPU: parity tests, smoke, bench (both arms), ncu launch list and
# full captures of the fused kernel (headline shape, series items) and of the occupation kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
cat gpurun_out/bench_r02.json; tail -3 gpurun_out/bench_r02.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_r02.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 3 -c 1 -o gpurun_out/prof_predict_r02 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:occupation_kernel -s 2 -c 1 -o gpurun_out/prof_occ_r02 -f python tools/run_occ_only.py > gpurun_out/ncu_occ.log 2>&1; echo "ncu occ rc=$?"
