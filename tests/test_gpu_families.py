"""GPU parity tests of the leauthaud11 / hearin15 occupation family
against the numpy/scipy oracle on the same seeded inputs, through the reference-shaped Python API.

The occupation formulas restate halotools from memory (parity unpinned, see oracle/ header); what
these tests pin is that the CUDA kernel computes exactly what the oracle states, including
halotools' numerical inversion of the stellar-to-halo-mass relation (100-knot not-a-knot spline).
Tolerance: rtol 1e-10 on (ngal, xi) like the zheng07 path; occupations to 5e-11 absolute/relative.
"""

import ctypes

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


@pytest.fixture(scope='module')
def orc():
    from oracle import tabcorr_oracle
    return tabcorr_oracle


def close(actual, ref, rtol=RTOL):
    actual, ref = np.asarray(actual), np.asarray(ref)
    scale = np.abs(ref).max(axis=-1, keepdims=True) if ref.ndim else np.abs(ref)
    np.testing.assert_allclose(actual, ref, rtol=rtol, atol=float(rtol) * 1e-3 * np.max(scale))


def table_pair(tb, orc, name):
    kw, _ = cases.SYNTHETIC[name]
    tab = tb.synthetic.make_table(**kw)
    halotab = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                     tab['attrs'])
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                            tab['attrs']['mode'])
    return halotab, table


VARIANTS = [
    dict(threshold=10.5, redshift=0.0, decorated=False, modulate_with_cenocc=True),
    dict(threshold=11.0, redshift=0.5, decorated=False, modulate_with_cenocc=False),
    dict(threshold=10.5, redshift=0.0, decorated=True, modulate_with_cenocc=True),
    dict(threshold=10.0, redshift=1.0, decorated=True, modulate_with_cenocc=False, split=0.3),
]


@pytest.mark.parametrize('variant', range(len(VARIANTS)))
@pytest.mark.parametrize('n_gauss', [10, 3])
def test_occupation_matches_oracle(tb, orc, variant, n_gauss):
    kw = VARIANTS[variant]
    halotab, table = table_pair(tb, orc, 'syn240')
    draws = tb.synthetic.make_draws_leauthaud11(9, seed=21 + variant, decorated=kw['decorated'])
    spec = tb.models.ModelSpec(tb.models.FAMILY_LEAUTHAUD11, kw['decorated'],
                               kw['modulate_with_cenocc'], kw.get('split', 0.5),
                               kw['threshold'], kw['redshift'])
    occ = halotab.mean_occupation_batch(draws, n_gauss_prim=n_gauss, model=spec).cpu().numpy()
    assert occ.shape == (9, 240)
    for i in range(9):
        model = orc.Leauthaud11Oracle(cases.draws_row(draws, i), **kw)
        ref = orc.mean_occupation(table, model, n_gauss)
        np.testing.assert_allclose(occ[i], ref, rtol=5e-11, atol=5e-11 * max(1.0, ref.max()))
    assert occ.max() > 1.0 and 0.0 < occ[:, :120].max() <= 1.0 + 1e-12


@pytest.mark.parametrize('name', ['syn240', 'syncross', 'synmulti'])
def test_predict_batch_matches_oracle(tb, orc, name):
    halotab, table = table_pair(tb, orc, name)
    kw = dict(threshold=10.5, redshift=0.0, decorated=True, modulate_with_cenocc=True)
    draws = tb.synthetic.make_draws_leauthaud11(40, seed=5, decorated=True)
    stand_in = tb.PrebuiltHodModelFactory('hearin15', threshold=10.5, redshift=0.0)
    ngal, xi = halotab.predict_batch(draws, model=stand_in)
    ngal_sep, xi_sep = halotab.predict_batch(draws, model=stand_in, separate_gal_type=True)
    for i in range(0, 40, 3):
        model = orc.Leauthaud11Oracle(cases.draws_row(draws, i), **kw)
        occ = orc.mean_occupation(table, model)
        ngal_ref, xi_ref = orc.predict(table, occ)
        close(ngal[i], ngal_ref)
        close(xi[i].ravel(), np.ravel(xi_ref))
        ngal_ref, xi_ref = orc.predict(table, occ, separate_gal_type=True)
        total = np.max(np.sum([np.abs(v) for v in xi_ref.values()], axis=0))
        for key in xi_ref:
            np.testing.assert_allclose(xi_sep[key][i], xi_ref[key], rtol=RTOL,
                                       atol=RTOL * 1e-3 * total)
        for key in ngal_ref:
            close(ngal_sep[key][i], ngal_ref[key])


def test_model_api_and_array_input(tb, orc):
    halotab, table = table_pair(tb, orc, 'syn36x3')
    model = tb.PrebuiltHodModelFactory('leauthaud11', threshold=10.8, redshift=0.02)
    model.param_dict['alphasat'] = 1.1
    spec = tb.models.resolve_model(model)
    assert (spec.family, spec.decorated, spec.modulate_with_cenocc, spec.threshold) == \
        (1, False, True, 10.8)
    ref_model = orc.Leauthaud11Oracle(model.param_dict, threshold=10.8, redshift=0.02)
    occ_ref = orc.mean_occupation(table, ref_model)
    np.testing.assert_allclose(halotab.mean_occupation(model), occ_ref, rtol=5e-11, atol=5e-11)
    ngal_ref, xi_ref = orc.predict(table, occ_ref)
    ngal, xi = halotab.predict(model)
    assert isinstance(ngal, np.floating) and xi.shape == tuple(table.tpcf_shape)
    close(ngal, ngal_ref)
    close(xi, xi_ref)
    # [B, 16] array in kernel order == dict input; bare dicts of this family need model=
    theta = np.array([[model.param_dict[k] for k in tb.models.LEAUTHAUD11_KEYS]] * 3)
    ngal_a, xi_a = halotab.predict_batch(theta, model=model)
    assert np.array_equal(ngal_a, np.repeat(ngal_a[:1], 3)) and np.array_equal(xi_a[0], xi_a[2])
    close(ngal_a[0], ngal_ref)
    close(xi_a[0], xi_ref)
    with pytest.raises(ValueError, match='model='):
        halotab.predict_batch({k: np.array([v]) for k, v in model.param_dict.items()})
    with pytest.raises(ValueError, match='columns|ordered'):
        halotab.predict_batch(theta[:, :7], model=model)
    # the optional 3xTF32 contraction also accepts this family's occupations
    ngal_t, xi_t = halotab.predict_batch(theta, model=model, precision='3xtf32')
    close(xi_t[0], xi_ref, rtol=1e-6)


def test_interpolator_with_leauthaud11(tb, orc):
    tables, param_table, _ = cases.grid_case('grid2d')
    halotabs = [tb.TabCorr.from_arrays(t['gal_type'], t['tpcf_matrix'], t['tpcf_shape'],
                                       t['attrs'], upload=False) for t in tables]
    interp = tb.Interpolator(halotabs, param_table)
    oracle_tables = [orc.OracleTable(t['gal_type'], t['tpcf_matrix'], t['tpcf_shape'],
                                     t['attrs']['mode']) for t in tables]
    ointerp = orc.OracleInterpolator(oracle_tables, param_table)
    model = tb.PrebuiltHodModelFactory('hearin15', threshold=10.5)
    model.param_dict.update(alpha_s=1.03, log_eta=-0.11)
    ref_model = orc.Leauthaud11Oracle(
        {k: v for k, v in model.param_dict.items() if k not in ('alpha_s', 'log_eta')},
        threshold=10.5, decorated=True)
    ref_model.param_dict.update(alpha_s=1.03, log_eta=-0.11)
    ngal_ref, xi_ref = ointerp.predict(ref_model)
    ngal, xi = interp.predict(model)
    close(ngal, ngal_ref)
    close(xi, xi_ref)


def test_non_monotonic_relation_gives_nan(tb, orc):
    halotab, _ = table_pair(tb, orc, 'syn36x3')
    model = tb.PrebuiltHodModelFactory('leauthaud11')
    model.param_dict.update(smhm_beta_0=-3.0)   # halo mass falls with stellar mass
    with pytest.raises(ValueError):
        orc.Leauthaud11Oracle(model.param_dict).mean_log_stellar_mass(np.array([1e12]))
    occ = halotab.mean_occupation(model)
    assert np.all(np.isnan(occ))


def test_fused_entry_point_rejects_the_family(tb):
    from tabcorr_b200 import _lib
    halotab = tb.TabCorr.from_arrays(**{k: v for k, v in tb.synthetic.make_table(
        n_mass=4, n_sec=1, n_r=3).items()})
    group = halotab._ensure_device()
    group.plan(10)
    import torch
    theta = torch.zeros((2, 18), dtype=torch.float64, device='cuda')
    out = torch.zeros((2, 8), dtype=torch.float64, device='cuda')
    work = torch.zeros(1 << 20, dtype=torch.uint8, device='cuda')
    model = _lib.tc_model(1, 0, 1, 0, 0.5, 10.5, 0.0)
    status = group.lib.tc_predict_batch(group.handle, ctypes.byref(model), 10, theta.data_ptr(), 0,
                                        None, 2, 0, 0, out.data_ptr(), 1, out.data_ptr() + 64, 3,
                                        work.data_ptr(), work.numel(), None)
    assert status == -3 and b'tc_occupation_batch' in group.lib.tc_last_error()


def test_full_size_properties(tb):
    """BASELINE-size batch (1e5 hearin15 draws, N=240, R=20): the size-independent properties the
    reference's own tests use (sum of parts == total, rtol 1e-6 there; G convergence) plus
    invariance to the batch split."""
    kw, _ = cases.SYNTHETIC['syn240']
    tab = tb.synthetic.make_table(**kw)
    halotab = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                     tab['attrs'])
    model = tb.PrebuiltHodModelFactory('hearin15', threshold=10.5)
    n = 100000
    draws = tb.synthetic.make_draws_leauthaud11(n, seed=31, decorated=True)
    ngal, xi = halotab.predict_batch(draws, model=model)
    assert np.all(np.isfinite(ngal)) and np.all(np.isfinite(xi)) and np.all(ngal > 0)
    ngal_sep, xi_sep = halotab.predict_batch(draws, model=model, separate_gal_type=True)
    np.testing.assert_allclose(ngal_sep['centrals'] + ngal_sep['satellites'], ngal, rtol=1e-13)
    total = sum(xi_sep.values())
    scale = np.sum([np.abs(v) for v in xi_sep.values()], axis=0).max(axis=1, keepdims=True)
    assert np.max(np.abs(total - xi) / scale) < 1e-12
    # the same draws in two halves and in reverse order give the same numbers bit for bit
    half = {k: v[n // 2:] for k, v in draws.items()}
    ngal_h, xi_h = halotab.predict_batch(half, model=model)
    assert np.array_equal(ngal_h, ngal[n // 2:]) and np.array_equal(xi_h, xi[n // 2:])
    # Gauss-Legendre convergence at the default parameters, like the reference's own test (random
    # draws with a stellar-mass scatter of 0.1 dex have a central step sharper than a mass bin)
    ngal_10, xi_10 = halotab.predict(model, n_gauss_prim=10)
    ngal_100, xi_100 = halotab.predict(model, n_gauss_prim=100)
    np.testing.assert_allclose(ngal_10, ngal_100, rtol=1e-4)
    np.testing.assert_allclose(xi_10, xi_100, rtol=1e-3)


# ---------------------------------------------------------------------------------------------
# mass-dependent assembly bias (halotools HeavisideAssembias with *_abscissa keywords)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('mode', ['auto', 'cross'])
def test_mass_dependent_assembias_against_oracle(tb, orc, mode):
    """Strength ordinates per draw (mean_occupation_*_assembias_param1..n) and a splitting
    percentile that varies with halo mass, centrals and satellites with different control
    points: occupations, ngal and xi against the oracle's restatement of halotools'
    assembias_strength / percentile_splitting_function (spline of degree min(3, n - 1) over
    log10 M, clipped), through the batch, the one-model and the separate_gal_type entries."""
    from tabcorr_b200.models import ModelSpec, assembias_keys
    tab = cases.synthetic.make_table(n_mass=24, n_sec=2, n_r=5, seed=12, mode=mode)
    halotab = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                     tab['attrs'])
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], mode)
    knots = dict(strength_abscissa=((11.0, 12.5, 14.0), (11.5, 12.2, 13.0, 14.5)),
                 split_abscissa=((11.0, 14.5), (0.0,)), split_ordinates=((0.25, 0.7), (0.4,)))
    spec = ModelSpec(0, True, **knots)
    assert spec.n_theta == 5 + 3 + 4
    n_draws = 40
    rng = np.random.default_rng(8)
    draws = cases.synthetic.make_draws(n_draws, seed=6)
    for key in assembias_keys('centrals', 3) + assembias_keys('satellites', 4):
        draws[key] = rng.uniform(-1.6, 1.6, n_draws)      # beyond [-1, 1]: clipped per halo
    for n_gauss in (3, 10):
        ngal, xi = halotab.predict_batch(draws, model=spec, n_gauss_prim=n_gauss)
        ngal_sep, xi_sep = halotab.predict_batch(draws, model=spec, n_gauss_prim=n_gauss,
                                                 separate_gal_type=True)
        occ = halotab.mean_occupation_batch(draws, model=spec, n_gauss_prim=n_gauss).cpu().numpy()
        for i in (0, 7, n_draws - 1):
            model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=True, **knots)
            occ_ref = orc.mean_occupation(table, model, n_gauss)
            np.testing.assert_allclose(occ[i], occ_ref, rtol=1e-11,
                                       atol=1e-13 * max(1.0, occ_ref.max()))
            ngal_ref, xi_ref = orc.predict(table, occ_ref)
            close(ngal[i], ngal_ref)
            close(xi[i], xi_ref)
            ngal_ref_sep, _ = orc.predict(table, occ_ref, separate_gal_type=True)
            for key in ngal_ref_sep:
                close(ngal_sep[key][i], ngal_ref_sep[key])
    # equal ordinates and a flat split are the plain decorated model, to rounding
    flat = {k: v for k, v in draws.items() if 'assembias' not in k}
    for key in assembias_keys('centrals', 3) + assembias_keys('satellites', 4):
        flat[key] = np.full(n_draws, 0.6 if 'centrals' in key else -0.35)
    flat_spec = ModelSpec(0, True, split=0.4, strength_abscissa=knots['strength_abscissa'],
                          split_abscissa=((11.0, 14.0), ()), split_ordinates=((0.4, 0.4), ()))
    plain = dict({k: v for k, v in flat.items() if 'assembias' not in k},
                 mean_occupation_centrals_assembias_param1=np.full(n_draws, 0.6),
                 mean_occupation_satellites_assembias_param1=np.full(n_draws, -0.35))
    a = halotab.predict_batch(flat, model=flat_spec)
    b = halotab.predict_batch(plain, model=ModelSpec(0, True, split=0.4))
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-10, atol=1e-12 * np.abs(b[1]).max())
    # the reference's calling convention: predict(model) with a model object
    model = tb.models.PrebuiltHodModelFactory(
        'decorated-zheng07', threshold=-20, assembias_strength=[0.9, -0.5, 0.3],
        assembias_strength_abscissa=[11.0, 12.5, 14.0], split=[0.3, 0.6], split_abscissa=[11.0, 14.0])
    ngal1, xi1 = halotab.predict(model)
    resolved = tb.models.resolve_model(model)
    oracle_model = orc.Zheng07Oracle(dict(model.param_dict), decorated=True,
                                     strength_abscissa=resolved.strength_abscissa,
                                     split_abscissa=resolved.split_abscissa,
                                     split_ordinates=resolved.split_ordinates)
    ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, oracle_model))
    close(ngal1, ngal_ref)
    close(xi1, xi_ref)


def test_hearin15_mass_dependent_assembias_against_oracle(tb, orc):
    """The same keywords for the leauthaud11 family: strength ordinates per draw and a splitting
    percentile that varies with halo mass, different control points for centrals and satellites,
    against the oracle (Leauthaud11Oracle shares Zheng07Oracle's restatement of halotools'
    assembias_strength / percentile_splitting_function)."""
    from tabcorr_b200.models import ModelSpec, assembias_keys
    halotab, table = table_pair(tb, orc, 'syn36x3')
    knots = dict(strength_abscissa=((11.0, 12.5, 14.0), (11.5, 12.2, 13.0, 14.5)),
                 split_abscissa=((11.0, 14.5), (0.0,)), split_ordinates=((0.25, 0.7), (0.4,)))
    spec = ModelSpec(tb.models.FAMILY_LEAUTHAUD11, True, True, 0.5, 10.5, 0.0, **knots)
    assert spec.n_theta == 16 + 3 + 4
    n_draws = 37
    rng = np.random.default_rng(18)
    draws = tb.synthetic.make_draws_leauthaud11(n_draws, seed=16)
    for key in assembias_keys('centrals', 3) + assembias_keys('satellites', 4):
        draws[key] = rng.uniform(-1.6, 1.6, n_draws)      # beyond [-1, 1]: clipped per halo
    for n_gauss in (3, 10):
        occ = halotab.mean_occupation_batch(draws, model=spec, n_gauss_prim=n_gauss).cpu().numpy()
        ngal, xi = halotab.predict_batch(draws, model=spec, n_gauss_prim=n_gauss)
        for i in (0, 7, n_draws - 1):
            model = orc.Leauthaud11Oracle(cases.draws_row(draws, i), threshold=10.5, redshift=0.0,
                                          decorated=True, modulate_with_cenocc=True, **knots)
            occ_ref = orc.mean_occupation(table, model, n_gauss)
            np.testing.assert_allclose(occ[i], occ_ref, rtol=5e-11,
                                       atol=5e-11 * max(1.0, occ_ref.max()))
            ngal_ref, xi_ref = orc.predict(table, occ_ref)
            close(ngal[i], ngal_ref)
            close(xi[i].ravel(), np.ravel(xi_ref))
    # equal ordinates and a flat split are the plain hearin15 model, to rounding
    flat = {k: v for k, v in draws.items() if 'assembias' not in k}
    for key in assembias_keys('centrals', 3) + assembias_keys('satellites', 4):
        flat[key] = np.full(n_draws, 0.6 if 'centrals' in key else -0.35)
    flat_spec = ModelSpec(tb.models.FAMILY_LEAUTHAUD11, True, True, 0.4, 10.5, 0.0,
                          strength_abscissa=knots['strength_abscissa'],
                          split_abscissa=((11.0, 14.0), ()), split_ordinates=((0.4, 0.4), ()))
    plain = dict({k: v for k, v in flat.items() if 'assembias' not in k},
                 mean_occupation_centrals_assembias_param1=np.full(n_draws, 0.6),
                 mean_occupation_satellites_assembias_param1=np.full(n_draws, -0.35))
    a = halotab.mean_occupation_batch(flat, model=flat_spec).cpu().numpy()
    b = halotab.mean_occupation_batch(
        plain, model=ModelSpec(tb.models.FAMILY_LEAUTHAUD11, True, True, 0.4, 10.5, 0.0)).cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-13 * b.max())
    # the reference's calling convention: predict(model) with a model object
    model = tb.PrebuiltHodModelFactory(
        'hearin15', threshold=10.5, central_assembias_strength=[0.9, -0.5, 0.3],
        satellite_assembias_strength=[0.1, 0.2, 0.3], assembias_strength_abscissa=[11.0, 12.5, 14.0],
        split=[0.3, 0.6], split_abscissa=[11.0, 14.0])
    ngal1, xi1 = halotab.predict(model)
    resolved = tb.models.resolve_model(model)
    oracle_model = orc.Leauthaud11Oracle(dict(model.param_dict), threshold=10.5, decorated=True,
                                         strength_abscissa=resolved.strength_abscissa,
                                         split_abscissa=resolved.split_abscissa,
                                         split_ordinates=resolved.split_ordinates)
    ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, oracle_model))
    close(ngal1, ngal_ref)
    close(xi1, xi_ref)


@pytest.mark.parametrize('decorated', [False, True])
def test_mass_dependent_scatter_against_oracle(tb, orc, decorated):
    """halotools' LogNormalScatterModel with scatter_abscissa / scatter_ordinates: the stellar-mass
    scatter of a halo is the spline of scatter_model_param1..n over log10 M; alone and together
    with a mass-dependent decoration."""
    from tabcorr_b200.models import ModelSpec, assembias_keys
    halotab, table = table_pair(tb, orc, 'syn36x3')
    knots = dict(scatter_abscissa=(11.0, 13.0, 15.0))
    if decorated:
        knots.update(strength_abscissa=((11.0, 14.0), ()), split_abscissa=((), (11.0, 14.5)),
                     split_ordinates=((), (0.3, 0.6)))
    spec = ModelSpec(tb.models.FAMILY_LEAUTHAUD11, decorated, True, 0.5, 10.5, 0.0, **knots)
    assert spec.n_theta == 18 + (1 if decorated else 0) + 2
    n_draws = 35
    rng = np.random.default_rng(28)
    draws = tb.synthetic.make_draws_leauthaud11(n_draws, seed=26, decorated=decorated)
    draws['scatter_model_param2'] = rng.uniform(0.1, 0.35, n_draws)
    draws['scatter_model_param3'] = rng.uniform(0.1, 0.35, n_draws)
    if decorated:
        for key in assembias_keys('centrals', 2):
            draws[key] = rng.uniform(-1.2, 1.2, n_draws)
    for n_gauss in (3, 10):
        occ = halotab.mean_occupation_batch(draws, model=spec, n_gauss_prim=n_gauss).cpu().numpy()
        ngal, xi = halotab.predict_batch(draws, model=spec, n_gauss_prim=n_gauss)
        for i in (0, 11, n_draws - 1):
            model = orc.Leauthaud11Oracle(cases.draws_row(draws, i), threshold=10.5, redshift=0.0,
                                          decorated=decorated, modulate_with_cenocc=True, **knots)
            occ_ref = orc.mean_occupation(table, model, n_gauss)
            np.testing.assert_allclose(occ[i], occ_ref, rtol=5e-11,
                                       atol=5e-11 * max(1.0, occ_ref.max()))
            ngal_ref, xi_ref = orc.predict(table, occ_ref)
            close(ngal[i], ngal_ref)
            close(xi[i].ravel(), np.ravel(xi_ref))
    # equal ordinates are the constant scatter, to rounding
    if not decorated:
        flat = dict(draws, scatter_model_param2=draws['scatter_model_param1'],
                    scatter_model_param3=draws['scatter_model_param1'])
        a = halotab.mean_occupation_batch(flat, model=spec).cpu().numpy()
        b = halotab.mean_occupation_batch(
            draws, model=ModelSpec(tb.models.FAMILY_LEAUTHAUD11, False, True, 0.5, 10.5,
                                   0.0)).cpu().numpy()
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12 * b.max())
        # the reference's calling convention
        model = tb.PrebuiltHodModelFactory('leauthaud11', threshold=10.5,
                                           scatter_abscissa=[12.0, 15.0],
                                           scatter_ordinates=[0.3, 0.12])
        ngal1, xi1 = halotab.predict(model)
        oracle_model = orc.Leauthaud11Oracle(dict(model.param_dict), threshold=10.5,
                                             scatter_abscissa=(12.0, 15.0))
        ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, oracle_model))
        close(ngal1, ngal_ref)
        close(xi1, xi_ref)


@pytest.mark.parametrize('seed', range(8))
def test_ragged_shuffled_tables(tb, orc, seed):
    """The mass bins of the leauthaud11 kernel pair the centrals group and the satellites group
    over the same node masses: tables with (mass, secondary) cells dropped at random and rows in
    random order leave bins with one galaxy type only, groups with a single row and pairs whose
    rows carry different weights.  Random n_gauss_prim (incl. values the five-node unrolling does
    not divide), batch sizes around the 32-draw blocks, legacy tables, all model variants."""
    rng = np.random.default_rng(500 + seed)
    n_mass, n_sec = int(rng.integers(3, 40)), int(rng.integers(1, 4))
    tab = tb.synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=3, mode='cross', seed=seed)
    gal_type = tab['gal_type']
    keep = np.flatnonzero(rng.random(len(gal_type)) > 0.25)
    if len(keep) < 2:
        keep = np.arange(len(gal_type))
    keep = rng.permutation(keep)
    gal_type, matrix = gal_type[keep], tab['tpcf_matrix'][:, keep]
    if rng.random() < 0.3:   # legacy table without the mass-function slope column
        names = [k for k in gal_type.dtype.names if k != 'prim_haloprop_dist_index']
        legacy = np.zeros(len(gal_type), dtype=[(k, gal_type.dtype[k]) for k in names])
        for k in names:
            legacy[k] = gal_type[k]
        gal_type = legacy
    kw = VARIANTS[seed % len(VARIANTS)]
    n_gauss = int(rng.choice([1, 2, 3, 5, 7, 10, 12]))
    n_draws = int(rng.choice([1, 7, 31, 32, 33, 65, 100]))
    draws = tb.synthetic.make_draws_leauthaud11(n_draws, seed=seed, decorated=kw['decorated'])
    spec = tb.models.ModelSpec(tb.models.FAMILY_LEAUTHAUD11, kw['decorated'],
                               kw['modulate_with_cenocc'], kw.get('split', 0.5),
                               kw['threshold'], kw['redshift'])
    halotab = tb.TabCorr.from_arrays(gal_type, matrix, tab['tpcf_shape'], tab['attrs'])
    table = orc.OracleTable(gal_type, matrix, tab['tpcf_shape'], 'cross')
    occ = halotab.mean_occupation_batch(draws, n_gauss_prim=n_gauss, model=spec).cpu().numpy()
    assert occ.shape == (n_draws, len(gal_type))
    for i in sorted(set([0, n_draws // 2, n_draws - 1])):
        model = orc.Leauthaud11Oracle(cases.draws_row(draws, i), **kw)
        ref = orc.mean_occupation(table, model, n_gauss)
        np.testing.assert_allclose(occ[i], ref, rtol=5e-11, atol=5e-11 * max(1.0, ref.max()))


def test_not_finite_parameters_give_nan(tb, orc):
    """A parameter that is not finite, or a stellar-to-halo-mass table that is not increasing
    (halotools raises there), makes every occupation of that draw NaN and leaves the others."""
    halotab, _ = table_pair(tb, orc, 'syn36x3')
    draws = tb.synthetic.make_draws_leauthaud11(40, seed=3)
    spec = tb.models.ModelSpec(tb.models.FAMILY_LEAUTHAUD11, False, True, 0.5, 10.5, 0.0)
    clean = halotab.mean_occupation_batch(draws, model=spec).cpu().numpy()
    assert np.all(np.isfinite(clean))
    bad = {k: np.array(v, dtype=np.float64, copy=True) for k, v in draws.items()}
    bad['scatter_model_param1'][5] = np.nan
    bad['smhm_beta_0'][17] = -3.0      # decreasing relation: the spline table is not monotonic
    occ = halotab.mean_occupation_batch(bad, model=spec).cpu().numpy()
    assert np.all(np.isnan(occ[5])) and np.all(np.isnan(occ[17]))
    rest = np.setdiff1d(np.arange(40), [5, 17])
    assert np.array_equal(occ[rest], clean[rest])


def test_small_batch_and_one_draw_paths_equal_the_general_path(tb, orc):
    """Host batches of at most 4096 draws and ``predict(model)`` of this family go through
    persistent pinned buffers (occupation kernel -> contraction, no allocations or copy launches):
    bit for bit the results of the general path."""
    halotab, _ = table_pair(tb, orc, 'syn240')
    model = tb.PrebuiltHodModelFactory('hearin15', threshold=10.5)
    draws = tb.synthetic.make_draws_leauthaud11(100, seed=9, decorated=True)
    for separate in (False, True):
        small = halotab.predict_batch(draws, model=model, separate_gal_type=separate)
        general = halotab.predict_batch(draws, model=model, separate_gal_type=separate,
                                        as_numpy=False)
        if separate:
            for s, g in zip(small, general):
                assert set(s) == set(g)
                for key in s:
                    assert np.array_equal(s[key], g[key].cpu().numpy())
        else:
            assert np.array_equal(small[0], general[0].cpu().numpy())
            assert np.array_equal(small[1], general[1].cpu().numpy())
    # [B, 18] array input and a second call that reuses the buffers with fewer draws
    theta = tb.models.theta_from_params(draws, None, tb.models.resolve_model(model))
    again = halotab.predict_batch(theta[:37], model=model)
    assert np.array_equal(again[0], small_total(halotab, draws, model)[0][:37])
    # one draw: predict(model) against the same draw in a batch
    row = cases.draws_row(draws, 5)
    model.param_dict.update(row)
    ngal1, xi1 = halotab.predict(model)
    batch = halotab.predict_batch(draws, model=model)
    np.testing.assert_allclose(ngal1, batch[0][5], rtol=1e-13)
    np.testing.assert_allclose(xi1, batch[1][5], rtol=1e-12, atol=1e-14 * np.abs(batch[1][5]).max())


def small_total(halotab, draws, model):
    return halotab.predict_batch(draws, model=model)


def test_interpolator_latency_paths_equal_the_general_path(tb, orc):
    """``Interpolator.predict(model)`` and small host batches of this family (persistent buffers,
    occupation kernel -> contraction per table group) against the general batch path."""
    axes = {'alpha_s': np.linspace(0.8, 1.2, 4)}
    tables, param_table = tb.synthetic.make_grid_tables(axes, n_mass=12, n_sec=2, n_r=5, kind='wp',
                                                        seed=3)
    halotabs = [tb.TabCorr.from_arrays(t['gal_type'], t['tpcf_matrix'], t['tpcf_shape'],
                                       t['attrs']) for t in tables]
    interp = tb.Interpolator(halotabs, param_table)
    model = tb.PrebuiltHodModelFactory('hearin15', threshold=10.5)
    draws = tb.synthetic.make_draws_leauthaud11(50, seed=4, decorated=True)
    draws['alpha_s'] = np.random.default_rng(5).uniform(0.85, 1.15, 50)
    small = interp.predict_batch(draws, model=model)
    general = interp.predict_batch(draws, model=model, as_numpy=False)
    assert np.array_equal(small[0], general[0].cpu().numpy())
    assert np.array_equal(small[1], general[1].cpu().numpy())
    model.param_dict.update(cases.draws_row(draws, 7))
    model.param_dict['alpha_s'] = float(draws['alpha_s'][7])
    ngal1, xi1 = interp.predict(model)
    np.testing.assert_allclose(ngal1, small[0][7], rtol=1e-13)
    np.testing.assert_allclose(xi1, small[1][7], rtol=1e-12, atol=1e-14 * np.abs(small[1][7]).max())
    ngal_sep, xi_sep = interp.predict(model, separate_gal_type=True)
    np.testing.assert_allclose(sum(ngal_sep.values()), ngal1, rtol=1e-12)
