#!/bin/bash
# round 2, N GPUs (N = $1): bench.py under torchrun with the overlapped gather (default), the
# blocking NCCL gather and the peer slab, 20 steps each
N=${1:-2}
mkdir -p gpurun_out
for mode in ${MODES:-overlap nccl peer}; do
  extra="--no-configs --no-cpu"
  if [ "$mode" = overlap ]; then extra="--no-cpu"; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --gather $mode $extra > gpurun_out/bench_${N}gpu_$mode.json 2> gpurun_out/bench_${N}gpu_$mode.err; echo "$mode rc=$?"
  python - <<PY
import json
try:
    r = json.loads(open('gpurun_out/bench_${N}gpu_$mode.json').read().strip().splitlines()[-1])
    print('$mode', 'value', r['value'], 'ms/step', r['ms_per_step'], 'e2e', r['e2e']['value'], 'consistent', r.get('results_consistent'), 'kernel_ms', r['roofline']['kernel_ms'])
except Exception as e:
    print('$mode', 'no line', e)
PY
  tail -3 gpurun_out/bench_${N}gpu_$mode.err
done
