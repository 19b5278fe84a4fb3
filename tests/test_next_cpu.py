"""CPU tests of the host logic next to the hot path: multipole weights, chunk schedules,
TableSet grouping (with an oracle-backed stand-in table)."""

import numpy as np
import pytest

from tabcorr_b200 import multipole
from tabcorr_b200.tabcorr import chunk_schedule
from tabcorr_b200.tableset import TableSet


def test_multipole_weights_match_definition():
    mu_bins = np.linspace(0, 1, 41)
    centres = 0.5 * (mu_bins[1:] + mu_bins[:-1])
    legendre = np.polynomial.legendre.Legendre
    xi = np.stack([3.0 * legendre.basis(0)(centres) - 1.5 * legendre.basis(2)(centres) +
                   0.25 * legendre.basis(4)(centres), np.cos(centres)])
    for order, expected in ((0, 3.0), (2, -1.5), (4, 0.25)):
        value = multipole.tpcf_multipole(xi, mu_bins, order=order)
        # midpoint rule with 40 bins: percent-level agreement with the exact coefficient
        assert abs(value[0] - expected) < 2e-2 * max(1.0, abs(expected))
        # the literal loop of the published definition
        loop = (2 * order + 1) / 2 * sum(
            xi[:, m] * (mu_bins[m + 1] - mu_bins[m]) *
            (legendre.basis(order)(centres[m]) + legendre.basis(order)(-centres[m]))
            for m in range(40))
        np.testing.assert_allclose(value, loop, rtol=1e-13)
    # odd multipoles of an even function vanish identically
    assert np.all(multipole.tpcf_multipole(xi, mu_bins, order=1) == 0)


def test_chunk_schedule_covers_every_draw_once():
    for n in (0, 1, 9999, 30000, 30001, 100000, 2500000):
        for chunk in ('auto', 0, 7, 25000, [10, 20], [5000, 90000, 5000]):
            bounds = chunk_schedule(n, chunk)
            assert [b[0] for b in bounds] == [0] + [b[1] for b in bounds[:-1]] if bounds else n == 0
            assert (bounds[-1][1] if bounds else 0) == n
            assert all(hi > lo for lo, hi in bounds)
    assert chunk_schedule(100000) == [(0, 10000), (10000, 90000), (90000, 100000)]
    assert max(hi - lo for lo, hi in chunk_schedule(5000000)) <= 1 << 20
    with pytest.raises(ValueError):
        chunk_schedule(10, 'fast')
    with pytest.raises(ValueError):
        chunk_schedule(10, [0])


class _Fake:
    """predict_batch stand-in returning torch tensors: result = offset + params."""

    def __init__(self, offset):
        self.offset = offset

    def predict_batch(self, params, separate_gal_type=False, as_numpy=True, **kw):
        import torch
        a = torch.as_tensor(np.asarray(params['a'], dtype=np.float64)) + self.offset
        xi = torch.stack([a, 2 * a], dim=1)
        if separate_gal_type:
            return {'centrals': a, 'satellites': -a}, {'centrals-centrals': xi}
        return a, xi


def test_table_set_groups_and_scatters():
    ts = TableSet([_Fake(0.0), _Fake(100.0), _Fake(1000.0)])
    a = np.arange(12.0)
    index = np.array([2, 0, 1, 1, 0, 2, 2, 2, 0, 1, 0, 0])
    ngal, xi = ts.predict_batch({'a': a, 'scalar': 1.0}, index, as_numpy=False)
    expected = a + np.array([0.0, 100.0, 1000.0])[index]
    assert np.array_equal(ngal.numpy(), expected)
    assert np.array_equal(xi.numpy(), np.stack([expected, 2 * expected], axis=1))
    ngal, xi = ts.predict_batch({'a': a}, index.astype(float), separate_gal_type=True,
                                as_numpy=False)
    assert np.array_equal(ngal['satellites'].numpy(), -expected)
    assert list(xi) == ['centrals-centrals']
    with pytest.raises(ValueError):
        ts.predict_batch({'a': a}, index + 1)
    with pytest.raises(ValueError):
        ts.predict_batch({'a': a}, index + 0.5)
    with pytest.raises(ValueError):
        TableSet([])


def test_leauthaud11_oracle_inverts_the_smhm_relation():
    """The oracle's halotools-style inversion (100-knot interpolating spline) is the not-a-knot
    cubic spline and inverts mean_log_halo_mass to the spline's accuracy."""
    from scipy.interpolate import CubicSpline
    from oracle import tabcorr_oracle as orc
    model = orc.Leauthaud11Oracle(redshift=0.3)
    knots = np.linspace(8.5, 12.5, 100)
    table = model.mean_log_halo_mass(knots)
    assert np.all(np.diff(table) > 0)
    log_mh = np.linspace(table[0] - 0.2, 15.3, 57)   # includes the extrapolated low-mass end
    ours = model.mean_log_stellar_mass(10**log_mh)
    np.testing.assert_allclose(ours, CubicSpline(table, knots, bc_type='not-a-knot')(log_mh),
                               rtol=0, atol=2e-11)
    inside = (log_mh > table[0]) & (log_mh < table[-1])
    np.testing.assert_allclose(model.mean_log_halo_mass(ours[inside]), log_mh[inside], atol=1e-5)
    # occupations: centrals rise from 0 to 1 through 0.5 at M_h(threshold), satellites ~ power law
    m_thr = 10**model.mean_log_halo_mass(model.threshold)
    assert abs(model.mean_occupation_centrals(prim_haloprop=np.array([m_thr]))[0] - 0.5) < 1e-5
    sats = model.mean_occupation_satellites(prim_haloprop=np.array([1e14, 2e14]))
    assert 1.9 < sats[1] / sats[0] < 2.2


def test_resolve_model_families():
    from types import SimpleNamespace
    from tabcorr_b200 import models

    def component(name, **attrs):
        return type(name, (), {})().__class__ and _with(type(name, (), {})(), attrs)

    def _with(obj, attrs):
        for k, v in attrs.items():
            setattr(obj, k, v)
        return obj

    cens = component('Leauthaud11Cens', threshold=10.5, redshift=0.1, prim_haloprop_key='halo_mvir',
                     param_dict={'scatter_model_param1': 0.2})
    sats = component('Leauthaud11Sats', threshold=10.5, modulate_with_cenocc=True,
                     prim_haloprop_key='halo_mvir')
    model = SimpleNamespace(_input_model_dictionary={'centrals_occupation': cens,
                                                     'satellites_occupation': sats})
    spec = models.resolve_model(model)
    assert spec.key()[:6] == (1, False, True, 0.5, 10.5, 0.1) and spec.n_theta == 18
    assert spec.theta_keys[:16] == models.LEAUTHAUD11_KEYS
    cens.param_dict['scatter_model_param2'] = 0.3
    with pytest.raises(NotImplementedError, match='scatter'):
        models.resolve_model(model)
    model = models.PrebuiltHodModelFactory('hearin15', threshold=11.0, redshift=0.5)
    spec = models.resolve_model(model)
    assert spec.key()[:6] == (1, True, True, 0.5, 11.0, 0.5) and not spec.mass_dependent
    theta = models.theta_from_params(model.param_dict, 1, spec)
    assert theta.shape == (1, 18) and theta[0, 16] == 1.0 and theta[0, 17] == 0.2
    with pytest.raises(ValueError, match='missing occupation parameters'):
        models.theta_from_params({'alphasat': 1.0}, 1, spec)
    with pytest.raises(ValueError, match='model='):
        models.spec_from_params({k: 1.0 for k in models.LEAUTHAUD11_KEYS})
    with pytest.raises(NotImplementedError):
        models.PrebuiltHodModelFactory('tinker13')
    assert models.spec_from_params({k: 1.0 for k in models.THETA_KEYS}).decorated


def test_leauthaud11_oracle_known_answers():
    """Regression pins of the leauthaud11 restatement (computed by this oracle -- NOT reference
    values: halotools is absent, see the oracle header).  Satellites use Leauthaud11Sats' own
    littleh = 0.72, centrals Behroozi10SmHm's 0.7 (round-1 advice; the satellite pins changed by
    1-20 % when that was corrected)."""
    from oracle import tabcorr_oracle as orc
    mass = 10**np.array([11.5, 12.5, 13.5, 14.5])
    model = orc.Leauthaud11Oracle()
    np.testing.assert_allclose(
        model.mean_occupation_centrals(prim_haloprop=mass, sec_haloprop_percentile=np.full(4, .3)),
        [0.00041245588971217106, 0.59916378405937, 0.9768642259131576, 0.9995655939151888],
        rtol=1e-10)
    np.testing.assert_allclose(
        model.mean_occupation_satellites(prim_haloprop=mass, sec_haloprop_percentile=np.full(4, .3)),
        [1.3513600498370619e-08, 0.04373237926410932, 1.224272670440629, 13.223118443159517],
        rtol=1e-10)
    model = orc.Leauthaud11Oracle(threshold=11.0, redshift=0.5, decorated=True,
                                  modulate_with_cenocc=False)
    pct = np.array([.2, .8, .2, .8])
    np.testing.assert_allclose(
        model.mean_occupation_centrals(prim_haloprop=mass, sec_haloprop_percentile=pct),
        [5.797862190348724e-13, 0.006104036664566104, 0.15264881029766253, 0.9135468664557148],
        rtol=1e-9)
    np.testing.assert_allclose(
        model.mean_occupation_satellites(prim_haloprop=mass, sec_haloprop_percentile=pct),
        [7.82793499675266e-06, 0.007944292740829723, 0.03765846917308047, 1.1702449732622406],
        rtol=1e-10)


def test_mass_dependent_assembias_models():
    """halotools HeavisideAssembias with assembias_strength_abscissa / split_abscissa: the model
    objects map to kernel descriptors with per-type control points, the draw grows by the extra
    strength ordinates, and the oracle's restatement reduces to the constant model when all
    ordinates are equal (tabcorr/tabcorr.py:556-563 passes whatever model it is given)."""
    from types import SimpleNamespace
    from oracle import tabcorr_oracle as orc
    from tabcorr_b200 import models

    model = models.PrebuiltHodModelFactory(
        'decorated-zheng07', threshold=-20, assembias_strength=[0.8, -0.3, 0.1],
        assembias_strength_abscissa=[11.0, 12.5, 14.0], split=[0.3, 0.6], split_abscissa=[11.0, 14.0])
    spec = models.resolve_model(model)
    assert spec.mass_dependent and not spec.latency_paths and spec.n_strength == (3, 3)
    assert spec.theta_keys[5:8] == models.assembias_keys('centrals', 3)
    assert spec.theta_keys[8:] == models.assembias_keys('satellites', 3) and spec.n_theta == 11
    theta = models.theta_from_params(model.param_dict, 1, spec)
    assert list(theta[0, 5:]) == [0.8, -0.3, 0.1, 0.8, -0.3, 0.1]
    assert spec.split_abscissa == ((11.0, 14.0),) * 2 and spec.split_ordinates == ((0.3, 0.6),) * 2

    # a halotools-shaped model: class names + the attributes HeavisideAssembias keeps
    def component(name, **attrs):
        obj = type(name, (), {})()
        for k, v in attrs.items():
            setattr(obj, k, v)
        return obj
    cens = component('AssembiasZheng07Cens', _assembias_strength_abscissa=[12.0, 13.0],
                     _split_abscissa=[2], _split_ordinates=[0.5])
    sats = component('AssembiasZheng07Sats', _assembias_strength_abscissa=[2],
                     _split_abscissa=[11.0, 12.0, 13.0], _split_ordinates=[0.2, 0.5, 0.7],
                     modulate_with_cenocc=False)
    spec = models.resolve_model(SimpleNamespace(
        _input_model_dictionary={'centrals_occupation': cens, 'satellites_occupation': sats}))
    assert spec.strength_abscissa == ((12.0, 13.0), ()) and spec.n_strength == (2, 1)
    assert spec.split_abscissa == ((0.0,), (11.0, 12.0, 13.0))
    assert spec.split_ordinates == ((0.5,), (0.2, 0.5, 0.7))
    with pytest.raises(NotImplementedError, match='control points'):
        models.ModelSpec(0, True, strength_abscissa=((11, 12, 13, 14, 15), ()))
    with pytest.raises(NotImplementedError, match='decorated'):
        models.ModelSpec(1, False, strength_abscissa=((11, 12), ()))
    # the leauthaud11 family (hearin15) takes the same keywords
    spec = models.ModelSpec(1, True, threshold=10.5, strength_abscissa=((11, 12), ()),
                            split_abscissa=((), (11.0, 13.0)), split_ordinates=((), (0.3, 0.6)))
    assert spec.mass_dependent and spec.n_strength == (2, 1) and spec.n_theta == 16 + 2 + 1
    assert spec.theta_keys[16:18] == models.assembias_keys('centrals', 2)
    model = models.PrebuiltHodModelFactory(
        'hearin15', threshold=10.8, central_assembias_strength=[0.9, 0.2, -0.4],
        satellite_assembias_strength=0.3, assembias_strength_abscissa=[11.5, 12.5, 14.0],
        split=[0.3, 0.6], split_abscissa=[11.0, 14.0])
    spec = models.resolve_model(model)
    assert spec.family == 1 and spec.mass_dependent and spec.n_strength == (3, 3)
    theta = models.theta_from_params(model.param_dict, 1, spec)
    assert theta.shape == (1, 22) and list(theta[0, 16:]) == [0.9, 0.2, -0.4, 0.3, 0.3, 0.3]
    # mass-dependent stellar-mass scatter (halotools LogNormalScatterModel keywords)
    spec = models.ModelSpec(1, False, True, 0.5, 10.5, 0.0, scatter_abscissa=(11.0, 13.0, 15.0))
    assert spec.n_theta == 18 + 2 and spec.theta_keys[18:] == ('scatter_model_param2',
                                                               'scatter_model_param3')
    assert not spec.mass_dependent and spec.scatter_keys == spec.theta_keys[18:]
    model = models.PrebuiltHodModelFactory('leauthaud11', scatter_abscissa=[12, 15],
                                           scatter_ordinates=[0.3, 0.1])
    spec = models.resolve_model(model)
    theta = models.theta_from_params(model.param_dict, 1, spec)
    assert spec.scatter_abscissa == (12.0, 15.0) and theta.shape == (1, 19)
    assert theta[0, 10] == 0.3 and theta[0, 18] == 0.1
    with pytest.raises(ValueError, match='scatter'):
        models.theta_from_params({k: v for k, v in model.param_dict.items()
                                  if k != 'scatter_model_param2'}, 1, spec)
    with pytest.raises(NotImplementedError, match='leauthaud11'):
        models.ModelSpec(0, False, scatter_abscissa=(11.0, 13.0))
    smhm = SimpleNamespace(scatter_model=SimpleNamespace(abscissa=[12.0, 15.0]))
    cens = component('Leauthaud11Cens', threshold=10.5, redshift=0.0, smhm_model=smhm,
                     param_dict={'scatter_model_param1': 0.3, 'scatter_model_param2': 0.1})
    sats = component('Leauthaud11Sats', threshold=10.5, modulate_with_cenocc=True)
    spec = models.resolve_model(SimpleNamespace(
        _input_model_dictionary={'centrals_occupation': cens, 'satellites_occupation': sats}))
    assert spec.scatter_abscissa == (12.0, 15.0) and spec.n_theta == 19
    cens = component('AssembiasLeauthaud11Cens', _assembias_strength_abscissa=[12.0, 13.0],
                     _split_abscissa=[2], _split_ordinates=[0.5], threshold=10.5, redshift=0.0)
    sats = component('AssembiasLeauthaud11Sats', _assembias_strength_abscissa=[2],
                     _split_abscissa=[11.0, 13.0], _split_ordinates=[0.2, 0.7], threshold=10.5,
                     modulate_with_cenocc=True)
    spec = models.resolve_model(SimpleNamespace(
        _input_model_dictionary={'centrals_occupation': cens, 'satellites_occupation': sats}))
    assert spec.family == 1 and spec.n_strength == (2, 1)
    assert spec.split_abscissa == ((0.0,), (11.0, 13.0))

    # oracle: equal ordinates == the constant model; a real mass dependence changes the result
    mass = 10**np.linspace(11.0, 14.5, 9)
    pct = np.tile([0.25, 0.75], 5)[:9]
    params = dict(orc.Zheng07Oracle(decorated=True).param_dict)
    constant = orc.Zheng07Oracle(params, decorated=True, split=0.4)
    flat = dict(params)
    for t in ('centrals', 'satellites'):
        for k in range(3):
            flat['mean_occupation_{}_assembias_param{}'.format(t, k + 1)] = 0.5
    same = orc.Zheng07Oracle(flat, decorated=True, split=0.4,
                             strength_abscissa=((11.0, 12.5, 14.0),) * 2,
                             split_abscissa=((11.0, 14.0),) * 2, split_ordinates=((0.4, 0.4),) * 2)
    for fn in ('mean_occupation_centrals', 'mean_occupation_satellites'):
        a = getattr(constant, fn)(prim_haloprop=mass, sec_haloprop_percentile=pct)
        b = getattr(same, fn)(prim_haloprop=mass, sec_haloprop_percentile=pct)
        np.testing.assert_allclose(b, a, rtol=1e-13)
    flat['mean_occupation_centrals_assembias_param2'] = -0.9
    varied = orc.Zheng07Oracle(flat, decorated=True, split=0.4,
                               strength_abscissa=((11.0, 12.5, 14.0),) * 2)
    assert not np.allclose(varied.mean_occupation_centrals(prim_haloprop=mass,
                                                           sec_haloprop_percentile=pct),
                           constant.mean_occupation_centrals(prim_haloprop=mass,
                                                             sec_haloprop_percentile=pct))
