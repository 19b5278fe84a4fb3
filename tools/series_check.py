#!/usr/bin/env python
"""Series evaluation of the bin averages (csrc/occupation.cuh, occupation_item_series) against the
node-by-node path (TC_TUNE_SERIES=0) on the GPU: largest deviations of the occupations and of
(ngal, xi) over wide priors, and the time of the standalone occupation kernel for both.

    python tools/series_check.py [--draws 20000]
"""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [
    # name, n_mass, n_sec, n_r, mode, decorated, n_gauss
    ('N=240 G=10', 60, 2, 20, 'auto', False, 10),
    ('N=240 G=10 decorated', 60, 2, 20, 'auto', True, 10),
    ('N=240 G=100', 60, 2, 20, 'auto', False, 100),
    ('N=240 G=1', 60, 2, 20, 'auto', False, 1),
    ('N=240 G=3', 60, 2, 20, 'auto', True, 3),
    ('N=60 (30 wide bins) G=10', 30, 1, 19, 'auto', False, 10),
    ('N=60 (30 wide bins) G=10 decorated', 30, 1, 19, 'auto', True, 10),
    ('N=1104 cross G=10', 276, 2, 13, 'cross', True, 10),
    ('N=28 (7 bins) G=10', 7, 2, 5, 'auto', True, 10),
]


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--draws', type=int, default=20000)
    args = parser.parse_args()
    import torch
    import tabcorr_b200
    from tabcorr_b200 import synthetic
    from tabcorr_b200.models import ModelSpec, theta_from_params

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    rng = np.random.default_rng(5)
    for name, n_mass, n_sec, n_r, mode, decorated, n_gauss in SHAPES:
        tab = synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=n_r, mode=mode)
        n = args.draws
        draws = {
            'logMmin': rng.uniform(10.5, 15.0, n),
            'sigma_logM': np.where(rng.random(n) < 0.1, rng.uniform(1e-3, 0.1, n),
                                   rng.uniform(0.05, 1.5, n)),
            'logM0': rng.uniform(9.0, 15.5, n),
            'logM1': rng.uniform(12.0, 15.0, n),
            'alpha': np.where(rng.random(n) < 0.05, rng.uniform(-0.5, 6.0, n),
                              rng.uniform(0.0, 3.0, n)),
        }
        if decorated:
            for key in synthetic.ASSEMBIAS_KEYS:
                draws[key] = rng.uniform(-1.2, 1.2, n)
        spec = ModelSpec(decorated=decorated)
        theta = torch.from_numpy(theta_from_params(draws, None, spec)).cuda()
        results = {}
        for series in (1, 0):
            os.environ['TC_TUNE_SERIES'] = str(series)
            halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                                       tab['tpcf_shape'], tab['attrs'])
            group = halotab._ensure_device()
            occ = group.occupation(spec, n_gauss, theta)
            ngal = torch.empty((n, 1), dtype=torch.float64, device='cuda')
            xi = torch.empty((n, n_r, 1), dtype=torch.float64, device='cuda')
            group.predict_into(spec, n_gauss, theta, None, False, ngal, 0, xi, 0)
            t_occ = timed(lambda: group.occupation(spec, n_gauss, theta))
            t_pred = timed(lambda: group.predict_into(spec, n_gauss, theta, None, False, ngal, 0,
                                                      xi, 0))
            results[series] = (occ.cpu().numpy(), ngal.cpu().numpy(), xi.cpu().numpy(), t_occ,
                               t_pred)
        del os.environ['TC_TUNE_SERIES']
        (o1, g1, x1, t1, p1), (o0, g0, x0, t0, p0) = results[1], results[0]
        is_sat = np.asarray(tab['gal_type']['gal_type']) == b'satellites'
        scale = np.maximum(np.abs(o0), 1e-300)
        xi_scale = np.abs(x0).max(axis=1, keepdims=True)
        print(json.dumps({
            'shape': name, 'n_draws': n,
            'cen_max_abs_dev': float(np.abs(o1 - o0)[:, ~is_sat].max()),
            'sat_max_rel_dev': float((np.abs(o1 - o0) / scale)[:, is_sat].max()),
            'sat_max_abs_dev_over_rowmax': float(
                (np.abs(o1 - o0)[:, is_sat] /
                 np.maximum(np.abs(o0[:, is_sat]).max(axis=1, keepdims=True), 1e-300)).max()),
            'ngal_max_rel_dev': float(np.nanmax(np.abs(g1 / g0 - 1))),
            'xi_max_dev_over_max_xi': float(np.nanmax(np.abs(x1 - x0) / xi_scale)),
            'nonfinite_series': int((~np.isfinite(o1)).sum()),
            'nonfinite_nodes': int((~np.isfinite(o0)).sum()),
            'occupation_ms_series': t1, 'occupation_ms_nodes': t0,
            'predict_ms_series': p1, 'predict_ms_nodes': p0}))


if __name__ == '__main__':
    main()
