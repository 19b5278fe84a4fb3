#!/usr/bin/env python
"""Host-to-host latency of small batches (the ensemble-sampler regime: tens to hundreds of walkers
per likelihood call) on the real bolplanck table (N=60) and the headline shape (N=240).

    python tools/bench_small_batches.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import tabcorr_b200
    from tabcorr_b200 import synthetic
    tables = {
        'bolplanck_wp N=60 R=19': tabcorr_b200.TabCorr.read(
            os.path.join(ROOT, 'tests', 'golden', 'bolplanck_wp.hdf5')),
    }
    tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
    tables['synthetic N=240 R=20'] = tabcorr_b200.TabCorr.from_arrays(
        tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
    model = tabcorr_b200.PrebuiltHodModelFactory('zheng07', threshold=-18)
    for name, halotab in tables.items():
        for _ in range(20):
            halotab.predict(model, check_consistency=False)
        t0 = time.perf_counter()
        for _ in range(500):
            halotab.predict(model, check_consistency=False)
        one = (time.perf_counter() - t0) / 500
        print(json.dumps({'table': name, 'call': 'predict(model)', 'us': one * 1e6}))
        for n in (1, 8, 32, 128, 512, 2048, 8192):
            draws = synthetic.make_draws(n, seed=2)
            for _ in range(10):
                halotab.predict_batch(draws)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 200
            for _ in range(reps):
                halotab.predict_batch(draws)
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({'table': name, 'call': 'predict_batch', 'n_draws': n,
                              'us': dt * 1e6, 'us_per_draw': dt * 1e6 / n}))


if __name__ == '__main__':
    main()
