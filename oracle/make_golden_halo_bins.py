#!/usr/bin/env python
"""Golden vectors of the halo-bin table from the reference's OWN
functions (``tabcorr.tabcorr.sort_into_bins`` / ``distribution_index`` imported unmodified through
``oracle/refstub.py``; the histogram is ``np.histogram2d`` as at tabcorr/tabcorr.py:194-199).
TEST INFRASTRUCTURE; runs only in the build container, writes ``tests/golden/halo_bins.npz``.

    python oracle/make_golden_halo_bins.py
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refstub  # noqa: E402


def make_halos(n_halos, seed):
    """Seeded halo catalogue with a steep mass function and uniform secondary percentiles."""
    rng = np.random.default_rng(seed)
    log_m = 10.7 + rng.exponential(0.45, n_halos)
    log_m = log_m[log_m < 15.0]
    prim = 10**log_m
    sec = rng.random(len(prim))
    return prim, sec


CASES = {
    # name: (n_halos, seed, primary bins, secondary bins) as the reference builds them (:160-184)
    'int30x1': (200000, 11, 30, None),
    'int60x2': (200000, 12, 60, 0.5),
    'int12x4': (5000, 13, 12, 4),
}


def bins_for(prim, prim_bins, sec_bins):
    log_bins = np.linspace(np.log10(np.amin(prim)) - 1e-3, np.log10(np.amax(prim)) + 1e-3,
                           prim_bins + 1)
    if sec_bins is None:
        pct_bins = np.array([-1e-3, 1 + 1e-3])
    elif isinstance(sec_bins, float):
        pct_bins = np.array([-1e-3, sec_bins, 1 + 1e-3])
    else:
        pct_bins = np.linspace(-1e-3, 1 + 1e-3, sec_bins + 1)
    return log_bins, pct_bins


def main():
    ref = refstub.load().tabcorr
    out = {}
    for name, (n_halos, seed, prim_bins, sec_bins) in CASES.items():
        prim, sec = make_halos(n_halos, seed)
        log_bins, pct_bins = bins_for(prim, prim_bins, sec_bins)
        n_h, log_bins, pct_bins = np.histogram2d(np.log10(prim), sec, bins=[log_bins, pct_bins])
        members = ref.sort_into_bins(np.log10(prim), log_bins, sec, pct_bins, prim)
        n_p, n_s = len(log_bins) - 1, len(pct_bins) - 1
        dist = np.zeros(n_p * n_s)
        mean = np.full(n_p * n_s, np.nan)
        for i in range(n_p * n_s):
            if len(members[i]) > 0:
                x_min, x_max = 10**log_bins[i % n_p], 10**log_bins[i % n_p + 1]
                mean[i] = np.mean(members[i])
                dist[i] = ref.distribution_index(x_min, x_max, mean[i])
        out[name + '/log_bins'] = log_bins
        out[name + '/pct_bins'] = pct_bins
        out[name + '/n_h'] = n_h.ravel(order='F')
        out[name + '/n_members'] = np.array([len(m) for m in members], dtype=np.float64)
        out[name + '/mean_prim'] = mean
        out[name + '/dist_index'] = dist
        out[name + '/case'] = np.array([n_halos, seed], dtype=np.int64)
    # distribution_index on its own, including the clipped ends
    x_max = np.array([1.2, 1.4058, 2.0, 10.0])
    x_mean = np.array([[1.0001, 1.05, 1.1, 1.15, 1.1999],
                       [1.0001, 1.1, 1.2, 1.3, 1.4057],
                       [1.001, 1.2, 1.5, 1.8, 1.999],
                       [1.001, 2.0, 5.0, 8.0, 9.999]])
    out['dist/x_max'] = x_max
    out['dist/x_mean'] = x_mean
    out['dist/n'] = np.array([[float(ref.distribution_index(1.0, xm, v)) for v in row]
                              for xm, row in zip(x_max, x_mean)])
    path = os.path.join(ROOT, 'tests', 'golden', 'halo_bins.npz')
    np.savez_compressed(path, **out)
    print('wrote {} arrays to {}'.format(len(out), path))


if __name__ == '__main__':
    main()
