#!/usr/bin/env python
"""Device-timed throughput of the occupation families on the headline table shape (N=240, R=20,
G=10, 1e5 draws): zheng07 (fused kernel) against leauthaud11 / hearin15 (occupation kernel ->
contraction on the occupation input).  Exploration tool; numbers go to profiles/ with this command.

    python tools/bench_families.py [--draws 100000] [--reps 10]
"""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--draws', type=int, default=100000)
    parser.add_argument('--reps', type=int, default=10)
    parser.add_argument('--only', default=None, help='substring of the family name')
    args = parser.parse_args()
    import torch
    import tabcorr_b200
    from tabcorr_b200 import synthetic, models
    from tabcorr_b200.models import ModelSpec, theta_from_params
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = []
        for i in range(args.reps):
            flush.fill_(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
    halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                               tab['tpcf_shape'], tab['attrs'])
    group = halotab._ensure_device()
    ngal = torch.empty((args.draws, 1), dtype=torch.float64, device='cuda')
    xi = torch.empty((args.draws, 20, 1), dtype=torch.float64, device='cuda')
    cases = [
        ('zheng07', ModelSpec(), synthetic.make_draws(args.draws, seed=1)),
        ('zheng07 decorated', ModelSpec(decorated=True),
         synthetic.make_draws(args.draws, seed=1, decorated=True)),
        ('leauthaud11', ModelSpec(models.FAMILY_LEAUTHAUD11, False, True, 0.5, 10.5, 0.0),
         synthetic.make_draws_leauthaud11(args.draws, seed=1)),
        ('hearin15', ModelSpec(models.FAMILY_LEAUTHAUD11, True, True, 0.5, 10.5, 0.0),
         synthetic.make_draws_leauthaud11(args.draws, seed=1, decorated=True)),
    ]
    for name, spec, draws in cases:
        if args.only and args.only not in name:
            continue
        theta = torch.from_numpy(theta_from_params(draws, None, spec)).cuda()
        t_predict = timed(lambda: group.predict_into(spec, 10, theta, None, False, ngal, 0, xi, 0))
        t_occ = timed(lambda: group.occupation(spec, 10, theta))
        print(json.dumps({'family': name, 'n_tracers': 240, 'n_r': 20, 'n_gauss': 10,
                          'n_draws': args.draws, 'predict_ms': t_predict,
                          'preds_per_s': args.draws / t_predict * 1e3,
                          'occupation_kernel_ms': t_occ,
                          'occupation_evals_per_s': args.draws * 2400 / t_occ * 1e3,
                          'finite': bool(torch.isfinite(xi).all().item())}))


if __name__ == '__main__':
    main()
