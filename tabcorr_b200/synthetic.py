"""Seeded synthetic tables and HOD draws of the shapes named in BASELINE.json / SURVEY.md §8(d).

There is no network and no halotools in the benchmark image, so tabulated correlation tables
cannot be produced by ``TabCorr.tabulate`` (reference ``tabcorr/tabcorr.py:23-372``).  The
generators below emit tables with the layout that function writes -- rows ordered all centrals
then all satellites, inside each type the secondary-percentile bin is the outer and the mass bin
the inner index (``tabcorr/tabcorr.py:199-234``); packed lower-triangular ``tpcf_matrix``
(``:770-806``) rounded to float32 as the default ``write`` does (``:418-419,448``).
"""

import numpy as np

GAL_TYPE_DTYPE = np.dtype([
    ('n_h', '<f8'), ('log_prim_haloprop_min', '<f8'), ('log_prim_haloprop_max', '<f8'),
    ('sec_haloprop_percentile_min', '<f8'), ('sec_haloprop_percentile_max', '<f8'),
    ('prim_haloprop', '<f8'), ('sec_haloprop_percentile', '<f8'),
    ('prim_haloprop_dist_index', '<f8'), ('gal_type', 'S10')])

ZHENG07_KEYS = ('logMmin', 'sigma_logM', 'logM0', 'logM1', 'alpha')
ASSEMBIAS_KEYS = ('mean_occupation_centrals_assembias_param1',
                  'mean_occupation_satellites_assembias_param1')


def make_gal_type(n_mass=60, n_sec=2, log_m_range=(10.7, 15.1), rng=None, n_h_scale=1.0):
    """Halo-bin table with ``2 * n_sec * n_mass`` rows (N=240 for the headline shape)."""
    if rng is None:
        rng = np.random.default_rng(20260117)
    edges = np.linspace(log_m_range[0], log_m_range[1], n_mass + 1)
    if n_sec == 1:
        sec_edges = np.array([-1e-3, 1 + 1e-3])
    else:
        sec_edges = np.concatenate([[-1e-3], np.linspace(0, 1, n_sec + 1)[1:-1], [1 + 1e-3]])
    log_m = 0.5 * (edges[1:] + edges[:-1])
    n_h_mass = 10**(-1.8 - 0.95 * (log_m - log_m_range[0])) * np.exp(-10**(log_m - 14.6))
    block = np.zeros(n_sec * n_mass, dtype=GAL_TYPE_DTYPE)
    for s in range(n_sec):
        sl = slice(s * n_mass, (s + 1) * n_mass)
        block['n_h'][sl] = n_h_scale * n_h_mass / n_sec
        block['log_prim_haloprop_min'][sl] = edges[:-1]
        block['log_prim_haloprop_max'][sl] = edges[1:]
        block['sec_haloprop_percentile_min'][sl] = sec_edges[s]
        block['sec_haloprop_percentile_max'][sl] = sec_edges[s + 1]
        block['sec_haloprop_percentile'][sl] = 0.5 * (sec_edges[s] + sec_edges[s + 1])
        dist = rng.uniform(-2.5, -1.5, n_mass)
        dist[-2:] = [-10.0, 10.0]  # the real tables clip at +-10 in the rarest bins
        block['prim_haloprop_dist_index'][sl] = dist
        block['prim_haloprop'][sl] = 10**log_m
    gal_type = np.concatenate([block, block])
    gal_type['gal_type'][:len(block)] = b'centrals'
    gal_type['gal_type'][len(block):] = b'satellites'
    return gal_type


def make_table(n_mass=60, n_sec=2, n_r=20, mode='auto', kind='wp', seed=20260117, n_h_scale=1.0,
               tpcf_shape=None, gal_type=None):
    """Synthetic table as a dict of plain arrays (``gal_type, tpcf_matrix, tpcf_shape, attrs``).

    ``kind='wp'``: positive lognormal entries with a one-halo boost on the same-mass-bin blocks and
    a few exact ``-2 * pi_max`` entries; ``kind='multipole'``: sign-mixed entries (xi_2, xi_4);
    ``mode='cross'``: an ``[R, N]`` matrix of Delta-Sigma-like magnitude.
    """
    rng = np.random.default_rng(seed)
    if gal_type is None:
        gal_type = make_gal_type(n_mass, n_sec, rng=rng, n_h_scale=n_h_scale)
    n = len(gal_type)
    if mode == 'auto':
        dense = np.empty((n_r, n, n))
        mass_bin = np.arange(n) % n_mass
        same = mass_bin[:, None] == mass_bin[None, :]
        for r in range(n_r):
            if kind == 'wp':
                m = rng.lognormal(2.0, 2.0, (n, n)) * (1.0 + 30.0 * same * np.exp(-0.4 * r))
                hit = rng.integers(0, n, (6, 2))
                m[hit[:, 0], hit[:, 1]] = -80.0
            else:
                m = rng.normal(0.0, 50.0 * np.exp(-0.15 * r), (n, n))
            m = np.tril(m) + np.tril(m, -1).T
            dense[r] = m
        dense = dense.astype(np.float32).astype(np.float64)
        rows, cols = np.tril_indices(n)
        matrix = dense[:, rows, cols]  # row-major lower triangle == symmetric_matrix_to_array
    elif mode == 'cross':
        matrix = (rng.lognormal(29.0, 1.5, (n_r, n)) *
                  np.exp(-0.25 * np.arange(n_r))[:, None]).astype(np.float32).astype(np.float64)
    else:
        raise ValueError("mode must be 'auto' or 'cross'")
    attrs = {'tpcf': kind if mode == 'auto' else 'mean_delta_sigma', 'mode': mode,
             'simname': 'synthetic', 'redshift': 0.0, 'Num_ptcl_requirement': 300,
             'prim_haloprop_key': 'halo_mvir', 'sec_haloprop_key': 'halo_nfw_conc'}
    if tpcf_shape is None:
        tpcf_shape = (n_r,)
    return {'gal_type': gal_type, 'tpcf_matrix': matrix, 'tpcf_shape': tuple(tpcf_shape),
            'attrs': attrs}


def make_draws(n_draws, seed=1, decorated=False, extra=None):
    """Independent uniform zheng07 draws, ``dict[str, ndarray[n_draws]]``."""
    rng = np.random.default_rng(seed)
    draws = {
        'logMmin': rng.uniform(11.0, 14.0, n_draws),
        'sigma_logM': rng.uniform(0.05, 1.0, n_draws),
        'logM0': rng.uniform(10.0, 13.5, n_draws),
        'logM1': rng.uniform(12.0, 15.0, n_draws),
        'alpha': rng.uniform(0.5, 1.5, n_draws),
    }
    if decorated:
        for key in ASSEMBIAS_KEYS:
            draws[key] = rng.uniform(-1.0, 1.0, n_draws)
    if extra:
        for key, (lo, hi) in extra.items():
            draws[key] = rng.uniform(lo, hi, n_draws)
    return draws


def make_draws_leauthaud11(n_draws, seed=1, decorated=False, extra=None):
    """Independent uniform leauthaud11 draws around the halotools defaults (Behroozi et al. 2010
    table 2, Leauthaud et al. 2011), ``dict[str, ndarray[n_draws]]``; ranges keep the
    stellar-to-halo-mass relation monotonic."""
    rng = np.random.default_rng(seed)
    ranges = {
        'smhm_m0_0': (10.4, 11.0), 'smhm_m0_a': (0.4, 0.8), 'smhm_m1_0': (12.0, 12.7),
        'smhm_m1_a': (0.1, 0.5), 'smhm_beta_0': (0.35, 0.5), 'smhm_beta_a': (0.1, 0.25),
        'smhm_delta_0': (0.4, 0.7), 'smhm_delta_a': (0.1, 0.25), 'smhm_gamma_0': (1.2, 1.9),
        'smhm_gamma_a': (2.0, 3.0), 'scatter_model_param1': (0.1, 0.4), 'alphasat': (0.8, 1.2),
        'bsat': (8.0, 13.0), 'bcut': (0.5, 3.0), 'betacut': (-0.3, 0.1), 'betasat': (0.6, 1.1),
    }
    draws = {key: rng.uniform(lo, hi, n_draws) for key, (lo, hi) in ranges.items()}
    if decorated:
        for key in ASSEMBIAS_KEYS:
            draws[key] = rng.uniform(-1.0, 1.0, n_draws)
    if extra:
        for key, (lo, hi) in extra.items():
            draws[key] = rng.uniform(lo, hi, n_draws)
    return draws


def make_grid_tables(axes, n_mass=30, n_sec=2, n_r=14, mode='auto', kind='wp', seed=7,
                     n_h_scale=1.0):
    """Tables on a full rectangular grid, as ``scripts/tabulate_snapshot.py:158-165,240-254`` writes
    them: one ``gal_type`` shared by all grid points, matrices varying smoothly with the grid
    coordinates.  ``axes``: dict ``{key: knots}`` in column order.  Returns
    ``(tables, param_table)`` with ``param_table`` a dict ``{key: values[T]}`` in file order
    (deliberately shuffled, since ``Interpolator.__init__`` must sort it)."""
    rng = np.random.default_rng(seed)
    gal_type = make_gal_type(n_mass, n_sec, rng=rng, n_h_scale=n_h_scale)
    base = make_table(n_mass, n_sec, n_r, mode, kind, seed=seed + 1, gal_type=gal_type)
    slopes = [make_table(n_mass, n_sec, n_r, mode, kind, seed=seed + 2 + d, gal_type=gal_type)
              for d in range(len(axes))]
    keys = list(axes.keys())
    grid = np.stack(np.meshgrid(*[np.asarray(axes[k], dtype=float) for k in keys],
                                indexing='ij'), axis=-1).reshape(-1, len(keys))
    grid = grid[rng.permutation(len(grid))]
    tables = []
    for point in grid:
        m = base['tpcf_matrix'].copy()
        for d, key in enumerate(keys):
            knots = np.asarray(axes[key], dtype=float)
            t = (point[d] - knots.mean()) / (knots.max() - knots.min())
            m = m + slopes[d]['tpcf_matrix'] * (0.3 * t + 0.5 * t**2 - 0.7 * t**3)
        m = m.astype(np.float32).astype(np.float64)
        tables.append({'gal_type': gal_type.copy(), 'tpcf_matrix': m,
                       'tpcf_shape': base['tpcf_shape'], 'attrs': dict(base['attrs'])})
    param_table = {key: grid[:, d].copy() for d, key in enumerate(keys)}
    return tables, param_table
