"""GPU parity tests: the CUDA path, called through the reference-shaped Python API (which goes
through the C ABI), against the golden vectors recorded from the reference's own source and
against the numpy oracle on the same seeded inputs.

Tolerance: north_star asks for rtol 1e-10 in FP64.  xi can cross zero for sign-mixed multipole
tables, so the comparison is |delta| <= RTOL * max(|xi_ref|, scale) with scale = max |xi_ref| over
the radial bins of that draw ("cancellation").
"""

import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-10


@pytest.fixture(scope='module')
def tb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import tabcorr_b200
    return tabcorr_b200


def close(actual, ref, rtol=RTOL):
    actual, ref = np.asarray(actual), np.asarray(ref)
    scale = np.abs(ref).max(axis=-1, keepdims=True) if ref.ndim else np.abs(ref)
    np.testing.assert_allclose(actual, ref, rtol=rtol, atol=float(rtol) * 1e-3 * np.max(scale))


def make_model(tb, theta, decorated=False, **kw):
    model = tb.models.Zheng07Model(decorated=decorated, **kw)
    model.param_dict.update(theta)
    return model


def table_from_dict(tb, tab):
    return tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                  tab['attrs'])


def check_golden(golden, name, result):
    ngal, xi = result
    if isinstance(ngal, dict):
        assert sorted(ngal) == sorted(k.split('/')[-1] for k in golden
                                      if k.startswith(name + '/ngal/'))
        assert sorted(xi) == sorted(k.split('/')[-1] for k in golden
                                    if k.startswith(name + '/xi/'))
        for k in ngal:
            assert isinstance(k, str)
            close(ngal[k], golden['{}/ngal/{}'.format(name, k)])
        total = np.sum([np.abs(golden['{}/xi/{}'.format(name, k)]) for k in xi], axis=0)
        for k in xi:
            ref = golden['{}/xi/{}'.format(name, k)]
            assert xi[k].shape == ref.shape
            np.testing.assert_allclose(xi[k], ref, rtol=RTOL, atol=RTOL * 1e-3 * total.max())
    else:
        close(ngal, golden[name + '/ngal'])
        assert np.shape(xi) == golden[name + '/xi'].shape
        close(xi, golden[name + '/xi'])


def test_bolplanck_wp(tb, golden, golden_dir):
    halotab = tb.TabCorr.read(os.path.join(golden_dir, 'bolplanck_wp.hdf5'))
    model = tb.PrebuiltHodModelFactory('zheng07', threshold=-18, redshift=0.0)
    assert model.param_dict == cases.THETA_M18
    occ = halotab.mean_occupation(model)
    np.testing.assert_allclose(occ, golden['bolplanck_wp/occ'], rtol=1e-11, atol=1e-15)
    for g in (1, 10, 100):
        ngal, xi = halotab.predict(model, n_gauss_prim=g)
        assert isinstance(ngal, np.floating) and xi.shape == (19,)
        check_golden(golden, 'bolplanck_wp/G{}'.format(g), (ngal, xi))
    check_golden(golden, 'bolplanck_wp/sep', halotab.predict(model, separate_gal_type=True))
    check_golden(golden, 'bolplanck_wp/m21', halotab.predict(
        tb.PrebuiltHodModelFactory('zheng07', threshold=-21)))
    # the ndarray branch of predict (tabcorr.py:616-621)
    check_golden(golden, 'bolplanck_wp/G10', halotab.predict(golden['bolplanck_wp/occ']))


def test_bolplanck_ds(tb, golden, golden_dir):
    halotab = tb.TabCorr.read(os.path.join(golden_dir, 'bolplanck_ds.hdf5'))
    model = tb.PrebuiltHodModelFactory('zheng07', threshold=-21)
    check_golden(golden, 'bolplanck_ds/G10', halotab.predict(model))
    check_golden(golden, 'bolplanck_ds/sep', halotab.predict(model, separate_gal_type=True))


def test_consistency_errors(tb, golden_dir):
    halotab = tb.TabCorr.read(os.path.join(golden_dir, 'bolplanck_wp.hdf5'))
    with pytest.raises(ValueError, match='redshift'):
        halotab.predict(tb.PrebuiltHodModelFactory('zheng07', threshold=-18, redshift=0.5))
    with pytest.raises(ValueError, match='primary halo'):
        halotab.predict(tb.PrebuiltHodModelFactory('zheng07', threshold=-18,
                                                   prim_haloprop_key='halo_m200b'))
    with pytest.raises(ValueError, match='secondary halo'):
        halotab.predict(tb.PrebuiltHodModelFactory('decorated-zheng07', threshold=-18,
                                                   sec_haloprop_key='halo_spin'))
    model = tb.PrebuiltHodModelFactory('zheng07', threshold=-18)
    model.gal_types = ['centrals']
    with pytest.raises(ValueError, match='galaxy types'):
        halotab.predict(model)
    halotab.predict(tb.PrebuiltHodModelFactory('zheng07', threshold=-18, redshift=0.5),
                    check_consistency=False)
    # **occ_kwargs (tabcorr.py:556-563): the keywords halotools' components read are either the
    # reference's own (a TypeError there too) or ignored next to them; others cannot be honoured
    model = tb.PrebuiltHodModelFactory('zheng07', threshold=-18)
    ngal, xi = halotab.predict(model)
    ngal2, xi2 = halotab.predict(model, table=None, sec_haloprop=np.ones(3))
    assert ngal == ngal2 and np.array_equal(xi, xi2)
    with pytest.raises(TypeError, match='multiple values'):
        halotab.predict(model, prim_haloprop=np.ones(3))
    with pytest.raises(NotImplementedError, match='custom_knob'):
        halotab.mean_occupation(model, custom_knob=1.0)


def test_ds_efficient_interpolator(tb, golden, golden_dir):
    interp = tb.Interpolator.read(os.path.join(golden_dir, 'ds_efficient.hdf5'))
    assert len(interp.tabcorr_list) == 4 and list(interp.unique_gal_type_index) == [0]
    knot = float(interp.xp[0][1])
    for tag, log_eta in (('a', 0.1), ('b', -0.3), ('knot', knot)):
        model = make_model(tb, dict(cases.THETA_AS, log_eta=log_eta), redshift=0.5,
                           prim_haloprop_key='halo_m258m')
        check_golden(golden, 'ds_efficient/' + tag, interp.predict(model))
        check_golden(golden, 'ds_efficient/{}_sep'.format(tag),
                     interp.predict(model, separate_gal_type=True))
    model = make_model(tb, dict(cases.THETA_AS, log_eta=0.6), redshift=0.5,
                       prim_haloprop_key='halo_m258m')
    with pytest.raises(ValueError, match='outside of the interpolation'):
        interp.predict(model)
    check_golden(golden, 'ds_efficient/extrap', interp.predict(model, extrapolate=True))
    model = make_model(tb, cases.THETA_AS, redshift=0.5, prim_haloprop_key='halo_m258m')
    with pytest.raises(ValueError, match='log_eta'):
        interp.predict(model)
    check_golden(golden, 'ds_efficient/table0', interp.tabcorr_list[0].predict(model))


@pytest.mark.parametrize('name', sorted(cases.SYNTHETIC))
def test_synthetic_tables(tb, golden, name):
    tab, draws, decorated = cases.synthetic_case(name, golden)
    halotab = table_from_dict(tb, tab)
    for g in ((1, 10, 100) if name == 'syn240dec' else (10,)):
        ngal, xi = halotab.predict_batch(draws, n_gauss_prim=g)
        assert xi.shape == (cases.N_DRAWS,) + tuple(tab['tpcf_shape'])
        close(ngal, golden['{}/G{}/ngal'.format(name, g)])
        ref = golden['{}/G{}/xi'.format(name, g)]
        close(xi.reshape(cases.N_DRAWS, -1), ref.reshape(cases.N_DRAWS, -1))
        occ = halotab.mean_occupation_batch(draws, n_gauss_prim=g).cpu().numpy()
        np.testing.assert_allclose(occ, golden['{}/G{}/occ'.format(name, g)], rtol=1e-11,
                                   atol=1e-15)
    model = make_model(tb, cases.draws_row(draws, 0), decorated=decorated)
    check_golden(golden, name + '/sep0', halotab.predict(model, separate_gal_type=True))
    # single-model API == row 0 of the batch, bitwise
    ngal0, xi0 = halotab.predict(model)
    ngal, xi = halotab.predict_batch(draws)
    assert ngal0 == ngal[0] and np.array_equal(xi0, xi[0])


@pytest.mark.parametrize('name', sorted(cases.GRIDS))
def test_synthetic_grids(tb, golden, name):
    tables, param_table, draws = cases.grid_case(name)
    interp = tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table)
    ngal, xi = interp.predict_batch(draws)
    close(ngal, golden[name + '/ngal'])
    close(xi.reshape(cases.N_DRAWS, -1), golden[name + '/xi'].reshape(cases.N_DRAWS, -1),
          rtol=1e-9)
    model = make_model(tb, cases.draws_row(draws, 0), decorated=True)
    check_golden(golden, name + '/sep0', interp.predict(model, separate_gal_type=True))


def test_interpolator_mixed_halo_tables(tb):
    """Grid tables with different gal_type tables form several device groups."""
    from oracle import tabcorr_oracle as orc
    axes = {'log_eta': np.linspace(-0.5, 0.5, 4)}
    tables, param_table = cases.synthetic.make_grid_tables(axes, n_mass=8, n_sec=2, n_r=5)
    tables[2]['gal_type'] = tables[2]['gal_type'].copy()
    tables[2]['gal_type']['n_h'] *= 1.25
    draws = cases.synthetic.make_draws(5, seed=3, decorated=True, extra={'log_eta': (-0.5, 0.5)})
    interp = tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table)
    assert len(interp.unique_gal_type_index) == 2
    ngal, xi = interp.predict_batch(draws)
    ref = orc.OracleInterpolator(
        [orc.OracleTable(t['gal_type'], t['tpcf_matrix'], t['tpcf_shape'], 'auto')
         for t in tables], param_table)
    for i in range(5):
        model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=True)
        ngal_ref, xi_ref = ref.predict(model)
        close(ngal[i], ngal_ref)
        close(xi[i], xi_ref, rtol=1e-9)


@pytest.mark.parametrize('n_draws', [1, 7, 200, 3000])
def test_batch_sizes_against_oracle(tb, n_draws):
    """Every draw-tile width (8, 16, 32, 64 draws per CTA) and ragged tails."""
    from oracle import tabcorr_oracle as orc
    tab = cases.synthetic.make_table(n_mass=20, n_sec=2, n_r=7, seed=5)
    draws = cases.synthetic.make_draws(n_draws, seed=n_draws, decorated=True)
    halotab = table_from_dict(tb, tab)
    ngal, xi = halotab.predict_batch(draws)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    for i in np.unique(np.linspace(0, n_draws - 1, 25).astype(int)):
        model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=True)
        ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model))
        close(ngal[i], ngal_ref)
        close(xi[i], xi_ref)


@pytest.mark.parametrize('n_mass,n_sec,mode', [
    (60, 2, 'auto'),     # n_pad 240: 56-draw tiles, two W buffers (the headline shape)
    (72, 2, 'auto'),     # n_pad 288: 48-draw tiles
    (80, 2, 'auto'),     # n_pad 320: 40-draw tiles
    (100, 2, 'auto'),    # n_pad 400: 32-draw tiles
    (125, 2, 'auto'),    # n_pad 512: 24-draw tiles (BASELINE configs[4] table shape)
    (150, 2, 'auto'),    # n_pad 608: 16-draw tiles
    (250, 2, 'auto'),    # n_pad 1008: 8-draw tiles
    (450, 2, 'auto'),    # n_pad 1808: 8-draw tiles, single W buffer (no overlap)
    (276, 2, 'cross'),   # n_pad 1104, the ds_efficient shape
    (500, 2, 'cross'),   # n_pad 2000: single W buffer
])
def test_every_tile_width_against_oracle(tb, n_mass, n_sec, mode):
    """Each draw-tile width / W-buffer count the launcher can pick, with several tiles per CTA
    (10^4 draws), against the oracle on a sample of draws; total and per-gal-type results."""
    from oracle import tabcorr_oracle as orc
    n_r = 3
    tab = cases.synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=n_r, seed=17, mode=mode)
    n_draws = 10000
    draws = cases.synthetic.make_draws(n_draws, seed=n_mass, decorated=True)
    halotab = table_from_dict(tb, tab)
    ngal, xi = halotab.predict_batch(draws)
    ngal_sep, xi_sep = halotab.predict_batch(draws, separate_gal_type=True)
    np.testing.assert_allclose(ngal_sep['centrals'] + ngal_sep['satellites'], ngal, rtol=1e-13)
    scale = sum(np.abs(v) for v in xi_sep.values()).max(axis=1, keepdims=True)
    assert np.all(np.abs(sum(xi_sep.values()) - xi) <= 1e-12 * scale)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], mode)
    for i in [0, 1, 55, 56, 4999, n_draws - 57, n_draws - 1]:
        model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=True)
        ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model))
        close(ngal[i], ngal_ref)
        close(xi[i], xi_ref)


def test_batch_invariance_bitwise(tb):
    """A draw's result does not depend on the batch it is in (tile width, position, schedule)."""
    tab = cases.synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
    halotab = table_from_dict(tb, tab)
    draws = cases.synthetic.make_draws(20000, seed=2)
    ngal, xi = halotab.predict_batch(draws)
    for sl in (slice(0, 1), slice(5, 14), slice(100, 500), slice(19000, 20000)):
        sub = {k: v[sl] for k, v in draws.items()}
        ngal_s, xi_s = halotab.predict_batch(sub)
        assert np.array_equal(ngal_s, ngal[sl]) and np.array_equal(xi_s, xi[sl])


def test_full_size_properties(tb):
    """BASELINE configs[1] at full size (N=240, R=20, B=1e5): size-independent properties plus a
    sample of draws against the oracle."""
    from oracle import tabcorr_oracle as orc
    tab = cases.synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
    halotab = table_from_dict(tb, tab)
    n_draws = 100000
    draws = cases.synthetic.make_draws(n_draws, seed=1)
    ngal, xi = halotab.predict_batch(draws)
    assert np.all(np.isfinite(ngal)) and np.all(np.isfinite(xi))
    # sum of the per-type parts equals the total (tests/test_general.py:8-28, rtol 1e-6 there)
    ngal_sep, xi_sep = halotab.predict_batch(draws, separate_gal_type=True)
    assert sorted(ngal_sep) == ['centrals', 'satellites']
    assert sorted(xi_sep) == ['centrals-centrals', 'centrals-satellites', 'satellites-satellites']
    np.testing.assert_allclose(ngal_sep['centrals'] + ngal_sep['satellites'], ngal, rtol=1e-13)
    total = sum(xi_sep.values())
    scale = sum(np.abs(v) for v in xi_sep.values()).max(axis=1, keepdims=True)
    assert np.all(np.abs(total - xi) <= 1e-12 * scale)
    # xi is invariant under a common rescaling of the tracer weights: scale the occupations
    occ = halotab.mean_occupation_batch(draws)
    ngal2, xi2 = halotab.predict_batch(None, occupation=occ * 3.0)
    np.testing.assert_allclose(ngal2, 3.0 * ngal, rtol=1e-13)
    np.testing.assert_allclose(xi2, xi, rtol=1e-12)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    for i in np.random.default_rng(0).integers(0, n_draws, 40):
        model = orc.Zheng07Oracle(cases.draws_row(draws, i))
        ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model))
        close(ngal[i], ngal_ref)
        close(xi[i], xi_ref)


def test_n_gauss_prim_convergence(tb, golden_dir):
    # tests/test_general.py:31-43
    halotab = tb.TabCorr.read(os.path.join(golden_dir, 'bolplanck_wp.hdf5'))
    model = tb.PrebuiltHodModelFactory('zheng07', threshold=-20)
    ngal_1, xi_1 = halotab.predict(model, n_gauss_prim=1)
    ngal_2, xi_2 = halotab.predict(model, n_gauss_prim=10)
    ngal_3, xi_3 = halotab.predict(model, n_gauss_prim=100)
    assert not np.isclose(ngal_1, ngal_2, atol=0, rtol=1e-6)
    assert not np.allclose(xi_1, xi_2, atol=0, rtol=1e-6)
    assert np.isclose(ngal_2, ngal_3, atol=0, rtol=1e-6)
    assert np.allclose(xi_2, xi_3, atol=0, rtol=1e-6)


def test_interpolator_matches_scipy_cubic(tb):
    # tests/test_general.py:46-69
    from scipy.interpolate import interp1d
    name = 'grid2d'
    tables, param_table, draws = cases.grid_case(name)
    axes = cases.GRIDS[name][0]
    interp = tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table)
    base = cases.draws_row(draws, 0)
    for key in axes:
        bins = axes[key]
        model = make_model(tb, dict(base, alpha_s=1.1, log_eta=0.1), decorated=True)
        xi_bins = []
        for x in bins:
            model.param_dict[key] = x
            xi_bins.append(interp.predict(model)[1])
        xi_bins = np.array(xi_bins)
        for x in np.linspace(np.amin(bins), np.amax(bins), 10):
            model.param_dict[key] = x
            xi_tabcorr = interp.predict(model)[1]
            xi_scipy = [interp1d(bins, xi_bins[:, i], kind='cubic')(x)
                        for i in range(len(xi_tabcorr))]
            assert np.allclose(xi_tabcorr, xi_scipy)


def test_modulate_with_cenocc_and_split(tb):
    from oracle import tabcorr_oracle as orc
    tab = cases.synthetic.make_table(n_mass=15, n_sec=2, n_r=4, seed=9)
    halotab = table_from_dict(tb, tab)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    draws = cases.synthetic.make_draws(6, seed=4, decorated=True)
    for kw in (dict(modulate_with_cenocc=True), dict(split=0.3),
               dict(split=0.7, modulate_with_cenocc=True)):
        spec = tb.models.ModelSpec(decorated=True, **kw)
        ngal, xi = halotab.predict_batch(draws, model=spec)
        for i in range(6):
            model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=True, **kw)
            ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model))
            close(ngal[i], ngal_ref)
            close(xi[i], xi_ref)


def test_legacy_table_without_dist_index(tb):
    from oracle import tabcorr_oracle as orc
    tab = cases.synthetic.make_table(n_mass=10, n_sec=1, n_r=3, seed=2)
    names = [n for n in tab['gal_type'].dtype.names if n != 'prim_haloprop_dist_index']
    legacy = np.zeros(len(tab['gal_type']), dtype=[(n, tab['gal_type'].dtype[n]) for n in names])
    for n in names:
        legacy[n] = tab['gal_type'][n]
    tab['gal_type'] = legacy
    halotab = table_from_dict(tb, tab)
    table = orc.OracleTable(legacy, tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    model = orc.Zheng07Oracle(cases.THETA_M21)
    ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model))
    ngal, xi = halotab.predict(make_model(tb, cases.THETA_M21))
    close(ngal, ngal_ref)
    close(xi, xi_ref)


def test_unsupported_model_fails_loudly(tb, golden_dir):
    halotab = tb.TabCorr.read(os.path.join(golden_dir, 'bolplanck_wp.hdf5'))

    class Other:
        param_dict = {'a': 1.0}

    with pytest.raises(NotImplementedError):
        halotab.predict(Other(), check_consistency=False)
    with pytest.raises(NotImplementedError):
        halotab.predict(tb.PrebuiltHodModelFactory('zheng07', threshold=-18), foo=1)


def test_device_math_accuracy(tb):
    """The table-driven erf / pow of the occupation kernel against mpmath-free references:
    scipy.special.erf (absolute error bar 3e-15) and numpy power (relative 1e-13)."""
    import torch
    from scipy.special import erf
    from tabcorr_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-7, 7, 200000), np.linspace(-6.3, 6.3, 100001),
                        [-100.0, 100.0, 0.0, -6.0, 6.0, 5.999999, -5.75, -1e-300, -15.99, 16.0,
                         1e20, -1e20, 2.0**60, -2.0**60, np.inf, -np.inf, 1e300, -1e300]])
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    _lib.check(lib.tc_debug_math(0, xd.data_ptr(), None, out.data_ptr(), len(x), None))
    got = out.cpu().numpy()
    assert np.all((got >= 0) & (got <= 1))
    assert np.max(np.abs(got - 0.5 * (1 + erf(x)))) < 3e-15      # incl. |x| beyond 2^49, +-inf
    t = 10**rng.uniform(-25, 8, 300000)
    alpha = rng.uniform(0.3, 2.5, len(t))
    td, ad = torch.from_numpy(t).cuda(), torch.from_numpy(alpha).cuda()
    out = torch.empty_like(td)
    _lib.check(lib.tc_debug_math(1, td.data_ptr(), ad.data_ptr(), out.data_ptr(), len(t), None))
    got = out.cpu().numpy()
    assert np.max(np.abs(got / t**alpha - 1)) < 1e-13
    # grouped erf of the leauthaud11 kernel: pairs close together share one coefficient column
    # (|x - y| <= 1 keeps most pairs on that path), pairs far apart take the per-argument path
    x = np.concatenate([rng.uniform(-8.5, 8.5, 200000), rng.uniform(-8.5, 8.5, 50000),
                        [-100.0, 100.0, 7.24, 7.26, -7.26, 16.0, 1e300, -1e300, np.inf]])
    y = np.concatenate([x[:200000] + rng.uniform(-1.0, 1.0, 200000), rng.uniform(-20, 20, 50000),
                        [-99.0, 7.0, 7.26, 6.3, -6.3, -16.0, 1.0, -1.0, -np.inf]])
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    for kind, arg in ((2, x), (3, y)):
        out = torch.empty_like(xd)
        _lib.check(lib.tc_debug_math(kind, xd.data_ptr(), yd.data_ptr(), out.data_ptr(), len(x),
                                     None))
        got = out.cpu().numpy()
        assert np.all(np.isfinite(got))
        assert np.max(np.abs(got - 0.5 * (1 + erf(arg)))) < 3e-15, (kind, np.max(np.abs(got - 0.5 * (1 + erf(arg)))))


@pytest.mark.parametrize('seed', range(12))
def test_random_ragged_shuffled_tables(tb, seed):
    """Randomised table shapes against the oracle: random numbers of mass / secondary bins, empty
    (mass, sec) cells dropped at random (tabcorr.py:346-351 does that), rows in RANDOM order (the
    kernels re-order centrals first internally), auto or cross, with or without the legacy
    dist-index column, random n_gauss_prim and batch size; total and per-gal-type results."""
    from oracle import tabcorr_oracle as orc
    rng = np.random.default_rng(1000 + seed)
    n_mass, n_sec = int(rng.integers(3, 40)), int(rng.integers(1, 4))
    mode = 'auto' if rng.random() < 0.65 else 'cross'
    n_r = int(rng.integers(1, 9))
    kind = 'multipole' if (mode == 'auto' and rng.random() < 0.3) else 'wp'
    tab = cases.synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=n_r, mode=mode, kind=kind,
                                     seed=seed)
    gal_type, n = tab['gal_type'], len(tab['gal_type'])
    if mode == 'auto':   # dense symmetric form to re-pack after dropping / permuting rows
        rows, cols = np.tril_indices(n)
        dense = np.zeros((n_r, n, n))
        dense[:, rows, cols] = tab['tpcf_matrix']
        dense[:, cols, rows] = tab['tpcf_matrix']
    keep = np.flatnonzero(rng.random(n) > 0.2)
    if len(keep) < 2:
        keep = np.arange(n)
    keep = rng.permutation(keep)
    gal_type = gal_type[keep]
    if mode == 'auto':
        sub = dense[:, keep][:, :, keep]
        rows, cols = np.tril_indices(len(keep))
        matrix = sub[:, rows, cols]
    else:
        matrix = tab['tpcf_matrix'][:, keep]
    if rng.random() < 0.3:   # legacy table without the mass-function slope column
        names = [k for k in gal_type.dtype.names if k != 'prim_haloprop_dist_index']
        legacy = np.zeros(len(gal_type), dtype=[(k, gal_type.dtype[k]) for k in names])
        for k in names:
            legacy[k] = gal_type[k]
        gal_type = legacy
    n_gauss = int(rng.choice([1, 2, 3, 5, 7, 10, 12]))
    n_draws = int(rng.integers(1, 300))
    decorated = bool(rng.random() < 0.6)
    draws = cases.synthetic.make_draws(n_draws, seed=seed, decorated=decorated)
    halotab = tb.TabCorr.from_arrays(gal_type, matrix, tab['tpcf_shape'], tab['attrs'])
    table = orc.OracleTable(gal_type, matrix, tab['tpcf_shape'], mode)
    ngal, xi = halotab.predict_batch(draws, n_gauss_prim=n_gauss)
    ngal_sep, xi_sep = halotab.predict_batch(draws, n_gauss_prim=n_gauss, separate_gal_type=True)
    occ = halotab.mean_occupation_batch(draws, n_gauss_prim=n_gauss).cpu().numpy()
    for i in sorted(set([0, n_draws // 2, n_draws - 1])):
        model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=decorated)
        occ_ref = orc.mean_occupation(table, model, n_gauss)
        np.testing.assert_allclose(occ[i], occ_ref, rtol=1e-11, atol=1e-13 * max(1.0, occ_ref.max()))
        ngal_ref, xi_ref = orc.predict(table, occ_ref)
        close(ngal[i], ngal_ref)
        close(xi[i], xi_ref)
        ngal_ref, xi_ref = orc.predict(table, occ_ref, separate_gal_type=True)
        assert sorted(xi_sep) == sorted(xi_ref) and sorted(ngal_sep) == sorted(ngal_ref)
        total = np.max(np.sum([np.abs(v) for v in xi_ref.values()], axis=0))
        for key in xi_ref:
            np.testing.assert_allclose(xi_sep[key][i], xi_ref[key], rtol=RTOL,
                                       atol=RTOL * 1e-3 * max(total, 1e-300))
        for key in ngal_ref:
            close(ngal_sep[key][i], ngal_ref[key])


def test_step_function_centrals_and_empty_batch(tb):
    """sigma_logM -> 0 makes <N_cen> the 0 / 1 step erf gives for infinite arguments (the
    reference divides by sigma_logM, tabcorr.py:556-559 -> halotools), and a batch of zero draws
    returns empty results without a launch (round-1 advice)."""
    from oracle import tabcorr_oracle as orc
    tab = cases.synthetic.make_table(n_mass=20, n_sec=2, n_r=6, seed=9)
    halotab = table_from_dict(tb, tab)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    draws = cases.synthetic.make_draws(40, seed=5)
    draws['sigma_logM'] = np.full(40, 1e-300)
    ngal, xi = halotab.predict_batch(draws)
    assert np.all(np.isfinite(ngal)) and np.all(np.isfinite(xi))
    for i in (0, 7, 39):
        model = orc.Zheng07Oracle(cases.draws_row(draws, i))
        ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model))
        close(ngal[i], ngal_ref)
        close(xi[i], xi_ref)
    empty = {k: v[:0] for k, v in draws.items()}
    ngal0, xi0 = halotab.predict_batch(empty)
    assert ngal0.shape == (0,) and xi0.shape == (0,) + tuple(tab['tpcf_shape'])
    ngal0, xi0 = halotab.predict_batch(empty, separate_gal_type=True)
    assert all(v.shape == (0,) for v in ngal0.values()) and len(xi0) == 3
