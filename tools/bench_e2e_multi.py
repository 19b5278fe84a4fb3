"""End-to-end (host in, host out on rank 0) predictions/s of predict_batch_sharded on the headline
workload for several chunk counts.  torchrun --nproc-per-node N tools/bench_e2e_multi.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import tabcorr_b200
from tabcorr_b200 import synthetic
from tabcorr_b200.distributed import predict_batch_sharded
world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', 1), ('RANK', 0), ('LOCAL_RANK', 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'], device=local)
draws = synthetic.make_draws(per_gpu * world, seed=1)
ref = None
for n_chunks in (3, 'host'):
    kw = dict(gather='host') if n_chunks == 'host' else dict(n_chunks=n_chunks)
    for _ in range(4):
        out = predict_batch_sharded(halotab, draws, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        out = predict_batch_sharded(halotab, draws, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    if rank == 0:
        if ref is None:
            ref = (out[0].copy(), out[1].copy())
        same = bool(np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]))
        print(json.dumps({'n_gpus': world, 'n_chunks': n_chunks, 'ms': dt * 1e3, 'preds_per_s': per_gpu * world / dt, 'same': same}), flush=True)
if world > 1:
    dist.destroy_process_group()
