"""Parity of the occupation arithmetic with a REAL halotools install -- runs automatically wherever
halotools imports (marker ``halotools``) and is skipped otherwise.

halotools is the un-vendored, unpinned dependency behind ``tabcorr/tabcorr.py:556-563``
(``model.mean_occupation_centrals/satellites``); it is absent from this image, so the oracle's
``Zheng07Oracle`` / ``Leauthaud11Oracle`` are restatements ("parity unpinned", oracle header).
These tests close that gap the day the package is present: the model construction follows the
reference's own fixtures (``/root/reference/tests/conftest.py:27-36``, ``README.md:47``).

* CPU part: the oracle classes against ``PrebuiltHodModelFactory('zheng07' | 'leauthaud11' |
  'hearin15')`` and a decorated zheng07 built from ``AssembiasZheng07Cens/Sats``.
* GPU part (also marked ``gpu``): ``TabCorr.mean_occupation`` / ``predict`` driven by the real
  model objects (``models.resolve_model`` maps them by component class) against the oracle's
  ``mean_occupation`` fed with the same real model, i.e. kernel == halotools through the
  reference's own quadrature.
"""

import numpy as np
import pytest

halotools = pytest.importorskip('halotools')
from halotools.empirical_models import PrebuiltHodModelFactory  # noqa: E402

import cases  # noqa: E402
from oracle import tabcorr_oracle as orc  # noqa: E402

pytestmark = pytest.mark.halotools

MASS = 10**np.linspace(10.8, 15.2, 300)


def perturb(param_dict, rng, scale=0.05):
    for key in list(param_dict):
        if 'assembias' in key:
            param_dict[key] = float(rng.uniform(-1, 1))
        else:
            param_dict[key] = float(param_dict[key] * (1 + scale * rng.uniform(-1, 1)))


def real_zheng07(threshold=-21, **kw):
    return PrebuiltHodModelFactory('zheng07', threshold=threshold, redshift=0.5,
                                   prim_haloprop_key='halo_m258m', mdef='258m', **kw)


def real_decorated_zheng07(threshold=-20):
    from halotools.empirical_models import (AssembiasZheng07Cens, AssembiasZheng07Sats,
                                            HodModelFactory, NFWPhaseSpace, TrivialPhaseSpace)
    return HodModelFactory(
        centrals_occupation=AssembiasZheng07Cens(threshold=threshold),
        satellites_occupation=AssembiasZheng07Sats(threshold=threshold),
        centrals_profile=TrivialPhaseSpace(), satellites_profile=NFWPhaseSpace())


@pytest.mark.parametrize('threshold', [-18, -20, -21])
@pytest.mark.parametrize('modulate', [False, True])
def test_zheng07_oracle_equals_halotools(threshold, modulate):
    model = real_zheng07(threshold, modulate_with_cenocc=modulate)
    rng = np.random.default_rng(threshold + 100)
    for _ in range(4):
        perturb(model.param_dict, rng)
        oracle = orc.Zheng07Oracle(dict(model.param_dict), modulate_with_cenocc=modulate)
        np.testing.assert_allclose(oracle.mean_occupation_centrals(prim_haloprop=MASS),
                                   model.mean_occupation_centrals(prim_haloprop=MASS),
                                   rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(oracle.mean_occupation_satellites(prim_haloprop=MASS),
                                   model.mean_occupation_satellites(prim_haloprop=MASS),
                                   rtol=1e-13, atol=1e-300)


def test_decorated_zheng07_oracle_equals_halotools():
    model = real_decorated_zheng07()
    rng = np.random.default_rng(3)
    pct = rng.uniform(0, 1, len(MASS))
    for _ in range(6):
        perturb(model.param_dict, rng)
        oracle = orc.Zheng07Oracle(dict(model.param_dict), decorated=True)
        for name in ('centrals', 'satellites'):
            ours = getattr(oracle, 'mean_occupation_' + name)(
                prim_haloprop=MASS, sec_haloprop_percentile=pct)
            theirs = getattr(model, 'mean_occupation_' + name)(
                prim_haloprop=MASS, sec_haloprop_percentile=pct)
            np.testing.assert_allclose(ours, theirs, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize('name,decorated', [('leauthaud11', False), ('hearin15', True)])
def test_leauthaud11_oracle_equals_halotools(name, decorated):
    model = PrebuiltHodModelFactory(name, threshold=10.5, redshift=0.3)
    rng = np.random.default_rng(11)
    pct = rng.uniform(0, 1, len(MASS))
    for _ in range(4):
        perturb(model.param_dict, rng, scale=0.02)
        oracle = orc.Leauthaud11Oracle(dict(model.param_dict), threshold=10.5, redshift=0.3,
                                       decorated=decorated)
        for gal in ('centrals', 'satellites'):
            ours = getattr(oracle, 'mean_occupation_' + gal)(
                prim_haloprop=MASS, sec_haloprop_percentile=pct)
            theirs = getattr(model, 'mean_occupation_' + gal)(
                prim_haloprop=MASS, sec_haloprop_percentile=pct)
            np.testing.assert_allclose(ours, theirs, rtol=1e-10, atol=1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize('kind', ['zheng07', 'decorated_zheng07', 'leauthaud11', 'hearin15'])
def test_kernel_equals_halotools_through_the_reference_quadrature(kind):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import tabcorr_b200
    tab = cases.synthetic.make_table(n_mass=30, n_sec=2, n_r=9, seed=5)
    tab['attrs'] = dict(tab['attrs'], prim_haloprop_key='halo_mvir', redshift=0.0)
    halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                               tab['tpcf_shape'], tab['attrs'])
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    if kind == 'zheng07':
        model = PrebuiltHodModelFactory('zheng07', threshold=-20)
    elif kind == 'decorated_zheng07':
        model = real_decorated_zheng07()
    else:
        model = PrebuiltHodModelFactory(kind, threshold=10.5, redshift=0.0)
    rng = np.random.default_rng(17)
    for _ in range(3):
        perturb(model.param_dict, rng, scale=0.02)
        occ_ref = orc.mean_occupation(table, model)         # the real halotools model object
        occ = halotab.mean_occupation(model, check_consistency=False)
        np.testing.assert_allclose(occ, occ_ref, rtol=1e-10, atol=1e-14)
        ngal_ref, xi_ref = orc.predict(table, occ_ref)
        ngal, xi = halotab.predict(model, check_consistency=False)
        np.testing.assert_allclose(ngal, ngal_ref, rtol=1e-10)
        np.testing.assert_allclose(xi, xi_ref, rtol=1e-10, atol=1e-13 * np.abs(xi_ref).max())


def test_mass_dependent_assembias_oracle_equals_halotools():
    """HeavisideAssembias with assembias_strength_abscissa / split_abscissa: the oracle's
    restatement (spline of degree min(3, n - 1) over log10 M, clipped) and the attribute mapping of
    ``models.resolve_model`` against the real components."""
    from halotools.empirical_models import (AssembiasZheng07Cens, AssembiasZheng07Sats,
                                            HodModelFactory, NFWPhaseSpace, TrivialPhaseSpace)
    from tabcorr_b200 import models
    cens = AssembiasZheng07Cens(threshold=-20, assembias_strength=[0.8, -0.3, 0.1],
                                assembias_strength_abscissa=[11.0, 12.5, 14.0])
    sats = AssembiasZheng07Sats(threshold=-20, assembias_strength=[0.5, -0.5],
                                assembias_strength_abscissa=[12.0, 14.0],
                                split=[0.3, 0.6], split_abscissa=[11.0, 14.0])
    model = HodModelFactory(centrals_occupation=cens, satellites_occupation=sats,
                            centrals_profile=TrivialPhaseSpace(),
                            satellites_profile=NFWPhaseSpace())
    spec = models.resolve_model(model)
    assert spec.mass_dependent and spec.n_strength == (3, 2)
    oracle = orc.Zheng07Oracle(dict(model.param_dict), decorated=True,
                               strength_abscissa=spec.strength_abscissa,
                               split_abscissa=spec.split_abscissa,
                               split_ordinates=spec.split_ordinates)
    oracle.split = spec.split
    pct = np.random.default_rng(5).uniform(0, 1, len(MASS))
    for name in ('centrals', 'satellites'):
        ours = getattr(oracle, 'mean_occupation_' + name)(prim_haloprop=MASS,
                                                          sec_haloprop_percentile=pct)
        theirs = getattr(model, 'mean_occupation_' + name)(prim_haloprop=MASS,
                                                           sec_haloprop_percentile=pct)
        np.testing.assert_allclose(ours, theirs, rtol=1e-12, atol=1e-300)


def test_mass_dependent_hearin15_oracle_equals_halotools():
    """The same keywords on the leauthaud11 family: AssembiasLeauthaud11Cens / Sats with
    assembias_strength_abscissa / split_abscissa against Leauthaud11Oracle and the attribute
    mapping of ``models.resolve_model``."""
    from halotools.empirical_models import (AssembiasLeauthaud11Cens, AssembiasLeauthaud11Sats,
                                            HodModelFactory, NFWPhaseSpace, TrivialPhaseSpace)
    from tabcorr_b200 import models
    cens = AssembiasLeauthaud11Cens(threshold=10.5, assembias_strength=[0.8, -0.3, 0.1],
                                    assembias_strength_abscissa=[11.0, 12.5, 14.0])
    sats = AssembiasLeauthaud11Sats(threshold=10.5, assembias_strength=[0.5, -0.5],
                                    assembias_strength_abscissa=[12.0, 14.0],
                                    split=[0.3, 0.6], split_abscissa=[11.0, 14.0])
    model = HodModelFactory(centrals_occupation=cens, satellites_occupation=sats,
                            centrals_profile=TrivialPhaseSpace(),
                            satellites_profile=NFWPhaseSpace())
    spec = models.resolve_model(model)
    assert spec.family == 1 and spec.mass_dependent and spec.n_strength == (3, 2)
    oracle = orc.Leauthaud11Oracle(dict(model.param_dict), threshold=10.5, decorated=True,
                                   modulate_with_cenocc=spec.modulate_with_cenocc,
                                   strength_abscissa=spec.strength_abscissa,
                                   split_abscissa=spec.split_abscissa,
                                   split_ordinates=spec.split_ordinates)
    oracle.split = spec.split
    pct = np.random.default_rng(6).uniform(0, 1, len(MASS))
    for name in ('centrals', 'satellites'):
        ours = getattr(oracle, 'mean_occupation_' + name)(prim_haloprop=MASS,
                                                          sec_haloprop_percentile=pct)
        theirs = getattr(model, 'mean_occupation_' + name)(prim_haloprop=MASS,
                                                           sec_haloprop_percentile=pct)
        np.testing.assert_allclose(ours, theirs, rtol=1e-10, atol=1e-300)


def test_mass_dependent_scatter_oracle_equals_halotools():
    """Leauthaud11Cens with scatter_abscissa / scatter_ordinates (LogNormalScatterModel): the
    oracle's per-halo scatter and the attribute path ``models.resolve_model`` reads
    (``smhm_model.scatter_model.abscissa``) against the real components."""
    from halotools.empirical_models import (HodModelFactory, Leauthaud11Cens, Leauthaud11Sats,
                                            NFWPhaseSpace, TrivialPhaseSpace)
    from tabcorr_b200 import models
    kw = dict(threshold=10.5, scatter_abscissa=[12.0, 15.0], scatter_ordinates=[0.3, 0.12])
    cens = Leauthaud11Cens(**kw)
    sats = Leauthaud11Sats(**kw)
    model = HodModelFactory(centrals_occupation=cens, satellites_occupation=sats,
                            centrals_profile=TrivialPhaseSpace(),
                            satellites_profile=NFWPhaseSpace())
    spec = models.resolve_model(model)
    assert spec.family == 1 and spec.scatter_abscissa == (12.0, 15.0) and spec.n_theta == 19
    oracle = orc.Leauthaud11Oracle(dict(model.param_dict), threshold=10.5,
                                   modulate_with_cenocc=spec.modulate_with_cenocc,
                                   scatter_abscissa=spec.scatter_abscissa)
    for name in ('centrals', 'satellites'):
        ours = getattr(oracle, 'mean_occupation_' + name)(prim_haloprop=MASS)
        theirs = getattr(model, 'mean_occupation_' + name)(prim_haloprop=MASS)
        np.testing.assert_allclose(ours, theirs, rtol=1e-10, atol=1e-300)
