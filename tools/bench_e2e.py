"""End-to-end (host numpy in, host numpy out) predictions/s of TabCorr.predict_batch on the headline
workload for several pipeline chunk sizes.  python tools/bench_e2e.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tabcorr_b200
from tabcorr_b200 import synthetic
n_draws = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
draws = synthetic.make_draws(n_draws, seed=1)
ref = None
for chunk in (0, 'auto', 100000, 50000, 25000, [10000, 80000, 10000], [5000, 90000, 5000], [10000, 40000, 40000, 10000], [20000, 60000, 20000], [6000, 44000, 44000, 6000]):
    for _ in range(3):
        out = halotab.predict_batch(draws, pipeline_chunk=chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        out = halotab.predict_batch(draws, pipeline_chunk=chunk)
    dt = (time.perf_counter() - t0) / reps
    if ref is None:
        ref = out
    same = bool(np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1]))
    print(json.dumps({'pipeline_chunk': chunk, 'ms': dt * 1e3, 'preds_per_s': n_draws / dt, 'same_as_unchunked': same}))
