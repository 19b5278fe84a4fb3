#!/bin/bash
# Round 2, end of round, one gpurun call on 1 GPU: parity tests, smoke, bench (both arms), ncu launch
# list of the bench command, final capture of the leauthaud11 kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02_final.json; tail -2 gpurun_out/bench_r02_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02_final.json 2> gpurun_out/bench_ref_r02_final.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_r02_final.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:occupation_l11 -s 2 -c 1 -o gpurun_out/prof_l11_r02_final -f python tools/run_l11_only.py > gpurun_out/ncu_l11_final.log 2>&1; echo "ncu l11 rc=$?"
