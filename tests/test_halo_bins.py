"""Halo-bin table of the tabulation side (tabcorr/tabcorr.py:192-234,
676-767): oracle and host code against golden vectors recorded from the reference's own functions
(oracle/make_golden_halo_bins.py), and the CUDA reduction against both."""

import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from make_golden_halo_bins import CASES, make_halos  # noqa: E402  (seeded catalogue generator)


@pytest.fixture(scope='module')
def halo_golden():
    with np.load(os.path.join(GOLDEN_DIR, 'halo_bins.npz')) as f:
        return {k: f[k] for k in f.files}


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_restatement_matches_reference_functions(halo_golden, name):
    from oracle import tabcorr_oracle as orc
    n_halos, seed = CASES[name][:2]
    prim, sec = make_halos(n_halos, seed)
    out = orc.halo_bin_table(prim, sec, halo_golden[name + '/log_bins'],
                             halo_golden[name + '/pct_bins'])
    assert np.array_equal(out['n_h'], halo_golden[name + '/n_h'])
    np.testing.assert_allclose(out['mean_prim'], halo_golden[name + '/mean_prim'], rtol=1e-14)
    np.testing.assert_allclose(out['prim_haloprop_dist_index'], halo_golden[name + '/dist_index'],
                               rtol=1e-12, atol=1e-12)
    n_p = len(halo_golden[name + '/log_bins']) - 1
    assert np.array_equal(out['log_prim_haloprop_min'][:n_p], halo_golden[name + '/log_bins'][:-1])
    assert out['n_h'].sum() == len(prim)


def test_host_distribution_index_matches_reference(halo_golden):
    """The product's own not-a-knot cubic (no scipy) against the reference's interp1d values,
    including the clipped ends (tabcorr/tabcorr.py:764-767)."""
    from tabcorr_b200.halo_bins import distribution_index, notaknot_cubic
    x_max, x_mean, n_ref = (halo_golden['dist/' + k] for k in ('x_max', 'x_mean', 'n'))
    for xm, row, ref in zip(x_max, x_mean, n_ref):
        got = distribution_index(np.ones(len(row)), np.full(len(row), xm), row)
        np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)
    assert distribution_index(1.0, 1.4, 1.0)[0] == -10.0 and distribution_index(1.0, 1.4, 1.4)[0] == 10.0
    # the spline itself against scipy on irregular knots
    from scipy.interpolate import make_interp_spline
    rng = np.random.default_rng(1)
    x = np.sort(rng.uniform(0, 5, 12))
    y = rng.normal(size=12)
    xq = np.linspace(x[0], x[-1], 57)
    np.testing.assert_allclose(notaknot_cubic(x, y, xq), make_interp_spline(x, y, k=3)(xq),
                               rtol=1e-11, atol=1e-12)


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    import ctypes
    from tabcorr_b200 import _lib
    lib = _lib.load()
    edges = (ctypes.c_double * 3)(0.0, 1.0, 0.5)      # not increasing
    out = (ctypes.c_double * 2)()
    rc = lib.tc_halo_bins(0, None, None, None, 0, edges, 2, edges, 2, out, out, out, None)
    assert rc == -1 and b"increase" in lib.tc_last_error()   # TC_EINVAL


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(CASES))
def test_gpu_halo_bins_match_golden(halo_golden, name):
    import torch
    from tabcorr_b200 import halo_bins
    n_halos, seed = CASES[name][:2]
    prim, sec = make_halos(n_halos, seed)
    log_bins, pct_bins = halo_golden[name + '/log_bins'], halo_golden[name + '/pct_bins']
    n_h, members, mean = halo_bins.halo_bin_counts(prim, sec, log_bins, pct_bins)
    assert np.array_equal(n_h, halo_golden[name + '/n_h'])            # integer work: bit-exact
    assert np.array_equal(members, halo_golden[name + '/n_members'])
    np.testing.assert_allclose(mean, halo_golden[name + '/mean_prim'], rtol=1e-13)
    table = halo_bins.halo_bin_table(prim, sec, log_bins, pct_bins)
    cells = len(n_h)
    assert len(table) == 2 * cells
    assert list(table['gal_type'][:cells]) == ['centrals'] * cells
    assert list(table['gal_type'][cells:]) == ['satellites'] * cells
    for half in (slice(0, cells), slice(cells, None)):
        assert np.array_equal(table['n_h'].data[half], halo_golden[name + '/n_h'])
        np.testing.assert_allclose(table['prim_haloprop_dist_index'].data[half],
                                   halo_golden[name + '/dist_index'], rtol=1e-9, atol=1e-9)
    # bit-reproducible whatever the grid: device-resident inputs, shuffled order
    perm = np.random.default_rng(0).permutation(len(prim))
    again = halo_bins.halo_bin_counts(torch.from_numpy(prim[perm]).cuda(),
                                      torch.from_numpy(sec[perm]).cuda(), log_bins, pct_bins)
    assert np.array_equal(again[0], n_h) and np.array_equal(again[1], members)
    np.testing.assert_allclose(again[2], mean, rtol=1e-13)   # device log10 may move edge haloes


@pytest.mark.gpu
def test_gpu_halo_bins_edges_and_outliers():
    """np.histogram2d keeps values on the last edge, np.digitize drops them; values outside the
    bins and NaN are ignored; empty cells have dist_index 0 (tabcorr/tabcorr.py:216-218)."""
    from oracle import tabcorr_oracle as orc
    from tabcorr_b200 import halo_bins
    log_bins = np.array([11.0, 11.5, 12.0, 13.0])
    pct_bins = np.array([0.0, 0.5, 1.0])
    log_m = np.array([11.0, 11.2, 11.5, 12.0, 12.999, 13.0, 13.5, 10.9, 11.7, np.nan, 11.3])
    sec = np.array([0.0, 0.2, 0.5, 1.0, 0.7, 0.3, 0.1, 0.4, 0.75, 0.2, np.nan])
    prim = 10**log_m
    n_h, members, mean = halo_bins.halo_bin_counts(prim, sec, log_bins, pct_bins)
    ref_nh = np.histogram2d(np.log10(prim[:9]), sec[:9], bins=[log_bins, pct_bins])[0].ravel(order='F')
    assert np.array_equal(n_h, ref_nh)
    ref_members = orc.sort_into_bins(np.log10(prim[:9]), log_bins, sec[:9], pct_bins, prim[:9])
    assert np.array_equal(members, [len(m) for m in ref_members])
    for got, m in zip(mean, ref_members):
        if len(m):
            np.testing.assert_allclose(got, np.mean(m), rtol=1e-13)
        else:
            assert np.isnan(got)
    table = halo_bins.halo_bin_table(prim, sec, log_bins, pct_bins)
    assert np.all(table['prim_haloprop_dist_index'].data[np.concatenate([members, members]) == 0] == 0)
