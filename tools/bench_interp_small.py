#!/usr/bin/env python
"""Host-to-host latency of small Interpolator batches (database-style wp grid: T=16 tables,
N=120, R=14): predict(model) and predict_batch for 1..4096 draws.  python tools/bench_interp_small.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import tabcorr_b200 as tb
    from tabcorr_b200 import synthetic
    axes = {'alpha_s': np.linspace(0.8, 1.2, 4), 'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))}
    tables, param_table = synthetic.make_grid_tables(axes, n_mass=30, n_sec=2, n_r=14)
    interp = tb.Interpolator([tb.TabCorr.from_arrays(t['gal_type'], t['tpcf_matrix'],
                                                     t['tpcf_shape'], t['attrs'], upload=False)
                              for t in tables], param_table)
    model = tb.PrebuiltHodModelFactory('zheng07', threshold=-20)
    model.param_dict.update(alpha_s=1.03, log_eta=0.1)
    for _ in range(20):
        interp.predict(model)
    t0 = time.perf_counter()
    for _ in range(300):
        interp.predict(model)
    print(json.dumps({'call': 'Interpolator.predict(model)', 'grid_tables': 16, 'n_tracers': 120,
                      'us': (time.perf_counter() - t0) / 300 * 1e6}))
    extra = {k: (float(v.min()), float(v.max())) for k, v in axes.items()}
    for n in (1, 8, 64, 512, 4096):
        draws = synthetic.make_draws(n, seed=3, extra=extra)
        for _ in range(10):
            interp.predict_batch(draws)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(100):
            interp.predict_batch(draws)
        dt = (time.perf_counter() - t0) / 100
        print(json.dumps({'call': 'Interpolator.predict_batch', 'n_draws': n, 'us': dt * 1e6,
                          'us_per_draw': dt * 1e6 / n}))


if __name__ == '__main__':
    main()
