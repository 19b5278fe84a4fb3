"""Shared case definitions: the synthetic tables/draws recorded in tests/golden by
oracle/make_golden.py (same seeds), as plain arrays."""

import hashlib

import numpy as np

from tabcorr_b200 import synthetic

THETA_M18 = dict(logMmin=11.35, sigma_logM=0.25, logM0=11.20, logM1=12.40, alpha=0.83)
THETA_M21 = dict(logMmin=12.79, sigma_logM=0.39, logM0=11.92, logM1=13.94, alpha=1.15)
THETA_AS = dict(logMmin=12.9, sigma_logM=0.25, logM0=11.20, logM1=14.1, alpha=1.2)

N_DRAWS = 12

SYNTHETIC = {
    'syn240': (dict(n_mass=60, n_sec=2, n_r=20), False),
    'syn240dec': (dict(n_mass=60, n_sec=2, n_r=20), True),
    'syn120': (dict(n_mass=60, n_sec=1, n_r=20), False),
    'syn36x3': (dict(n_mass=6, n_sec=3, n_r=5), True),
    'synmulti': (dict(n_mass=60, n_sec=2, n_r=42, kind='multipole', tpcf_shape=(3, 14)), True),
    'syncross': (dict(n_mass=60, n_sec=2, n_r=13, mode='cross'), True),
}

GRIDS = {
    'grid2d': ({'alpha_s': np.linspace(0.8, 1.2, 4),
                'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))},
               dict(n_mass=12, n_sec=2, n_r=14, mode='auto')),
    'grid3d': ({'alpha_c': np.linspace(0.0, 0.4, 4), 'alpha_s': np.linspace(0.8, 1.2, 5),
                'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))},
               dict(n_mass=8, n_sec=2, n_r=6, mode='auto', kind='multipole')),
    'grid1dx': ({'log_eta': np.linspace(-0.5, 0.5, 6)},
                dict(n_mass=10, n_sec=2, n_r=7, mode='cross')),
}


def synthetic_case(name, golden=None):
    kw, decorated = SYNTHETIC[name]
    tab = synthetic.make_table(**kw)
    if golden is not None:  # the generator must still produce what the golden run saw
        sha = np.frombuffer(hashlib.sha1(
            np.ascontiguousarray(tab['tpcf_matrix']).tobytes()).digest(), np.uint8)
        assert np.array_equal(sha, golden[name + '/matrix_sha1'])
        sha = np.frombuffer(hashlib.sha1(tab['gal_type'].tobytes()).digest(), np.uint8)
        assert np.array_equal(sha, golden[name + '/gal_type_sha1'])
    draws = synthetic.make_draws(N_DRAWS, seed=11, decorated=decorated)
    return tab, draws, decorated


def grid_case(name):
    axes, kw = GRIDS[name]
    tables, param_table = synthetic.make_grid_tables(axes, **kw)
    extra = {k: (float(np.min(v)), float(np.max(v))) for k, v in axes.items()}
    draws = synthetic.make_draws(N_DRAWS, seed=13, decorated=True, extra=extra)
    return tables, param_table, draws


def draws_row(draws, i):
    return {k: float(v[i]) for k, v in draws.items()}
