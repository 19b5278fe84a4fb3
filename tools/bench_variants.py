#!/usr/bin/env python
"""Device-timed throughput of the fused kernel over table shapes and input kinds (not the headline
bench; an exploration tool whose numbers go to profiles/ with the command that made them).

    python tools/bench_variants.py [--draws 100000] [--reps 10]

For every shape: predictions/s with parameter draws (occupation + contraction), with precomputed
occupations (contraction only: what the DMMA loop alone achieves), executed-DMMA fraction of the
live peak, and the standalone occupation kernel.
"""

import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [
    # name, n_mass, n_sec, n_r, mode, kind, decorated
    ('cfg2 N=240 R=20 wp', 60, 2, 20, 'auto', 'wp', False),
    ('cfg2a N=120 R=20 wp', 60, 1, 20, 'auto', 'wp', False),
    ('cfg3 N=240 R=42 multipoles decorated', 60, 2, 42, 'auto', 'multipole', True),
    ('cfg5 N=500 R=20 wp', 125, 2, 20, 'auto', 'wp', False),
    ('bolplanck-like N=60 R=19 wp', 30, 1, 19, 'auto', 'wp', False),
    ('cross N=1104 R=13 ds', 276, 2, 13, 'cross', 'wp', False),
    ('cross N=60 R=19 ds', 30, 1, 19, 'cross', 'wp', False),
    ('cross N=240 R=19 ds', 60, 2, 19, 'cross', 'wp', False),
]


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--draws', type=int, default=100000)
    parser.add_argument('--reps', type=int, default=10)
    parser.add_argument('--only', default=None)
    parser.add_argument('--n-gauss', type=int, default=10)
    parser.add_argument('--tune', default='', help='semicolon-separated sets of TC_TUNE_* knobs, '
                        'e.g. "CHUNKS=72,OCC_ITEMS=28;CHUNKS=40"')
    args = parser.parse_args()
    import torch
    import tabcorr_b200
    from tabcorr_b200 import _lib, synthetic
    from tabcorr_b200.models import ModelSpec, theta_from_params
    lib = _lib.load()
    peak = ctypes.c_double()
    _lib.check(lib.tc_measure_dmma_peak(0, ctypes.byref(peak)))
    print(json.dumps({'dmma_peak_tflops': peak.value}))
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

    runs = [(tune, shape) for tune in args.tune.split(';') for shape in SHAPES]
    for tune, (name, n_mass, n_sec, n_r, mode, kind, decorated) in runs:
        if args.only and args.only not in name:
            continue
        for key in [k for k in os.environ if k.startswith('TC_TUNE_')]:
            del os.environ[key]
        for item in filter(None, tune.split(',')):
            key, value = item.split('=')
            os.environ['TC_TUNE_' + key] = value
        tab = synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=n_r, mode=mode, kind=kind)
        halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                                   tab['tpcf_shape'], tab['attrs'])
        group = halotab._ensure_device()
        n = len(tab['gal_type'])
        draws = synthetic.make_draws(args.draws, seed=1, decorated=decorated)
        spec = ModelSpec(decorated=decorated)
        theta = torch.from_numpy(theta_from_params(draws, None, spec)).cuda()
        ngal = torch.empty((args.draws, 1), dtype=torch.float64, device='cuda')
        xi = torch.empty((args.draws, n_r, 1), dtype=torch.float64, device='cuda')
        occ = group.occupation(spec, args.n_gauss, theta)
        t_theta = timed(lambda: group.predict_into(spec, args.n_gauss, theta, None, False, ngal, 0, xi, 0))
        t_occ = timed(lambda: group.predict_into(None, args.n_gauss, None, occ, False, ngal, 0, xi, 0))
        t_only = timed(lambda: group.occupation(spec, args.n_gauss, theta))
        t_tf32 = (timed(lambda: group.predict_into(spec, args.n_gauss, theta, None, False, ngal, 0,
                                                   xi, 0, precision=1)) if mode == 'auto' else None)
        n_pad = (n + 15) // 16 * 16
        if mode == 'auto':
            executed = 2.0 * n_r * 64.0 * (n_pad // 8) * (n_pad // 8 + 1) / 2
            algorithmic = 2.0 * n_r * n * n + 2.0 * n_r * n
        else:
            executed = 2.0 * ((n_r + 15) // 16 * 16) * n_pad
            algorithmic = 2.0 * n_r * n
        out = {
            'shape': name, 'tune': tune, 'n_gauss': args.n_gauss, 'n_draws': args.draws,
            'theta_ms': t_theta, 'theta_preds_per_s': args.draws / t_theta * 1e3,
            'occ_input_ms': t_occ, 'occ_input_preds_per_s': args.draws / t_occ * 1e3,
            'occupation_kernel_ms': t_only, 'tf32_theta_ms': t_tf32,
            'tf32_preds_per_s': args.draws / t_tf32 * 1e3 if t_tf32 else None,
            'occupation_evals_per_s': args.draws * n * args.n_gauss / t_only * 1e3,
            'executed_frac_theta': executed * args.draws / (t_theta * 1e-3) / (peak.value * 1e12),
            'executed_frac_occ_input': executed * args.draws / (t_occ * 1e-3) / (peak.value * 1e12),
            'algorithmic_tflops_theta': algorithmic * args.draws / (t_theta * 1e-3) * 1e-12,
        }
        print(json.dumps(out))


if __name__ == '__main__':
    main()
