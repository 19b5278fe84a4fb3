"""GPU tests of the rows next to the hot path (BASELINE configs[2..4]):
write -> read round trip on the device, (s, mu) -> multipole tables, per-draw table selection,
sweeps with device-side draws, the chunked host-to-host pipeline."""

import numpy as np
import pytest

import cases
from test_gpu_parity import RTOL, close, table_from_dict, tb  # noqa: F401

pytestmark = pytest.mark.gpu


def test_write_read_round_trip_predicts_identically(tb, tmp_path):
    tab = cases.synthetic.make_table(n_mass=20, n_sec=2, n_r=9, seed=4)
    halotab = table_from_dict(tb, tab)
    draws = cases.synthetic.make_draws(300, seed=3, decorated=True)
    ngal, xi = halotab.predict_batch(draws)
    halotab.write(tmp_path / 'table.hdf5')                 # float32 matrix on disk (default)
    again = tb.TabCorr.read(tmp_path / 'table.hdf5')
    ngal2, xi2 = again.predict_batch(draws)
    # the synthetic matrix is float32-representable, so nothing is lost
    assert np.array_equal(ngal, ngal2) and np.array_equal(xi, xi2)


def test_interpolator_write_read_round_trip(tb, tmp_path):
    tables, param_table, draws = cases.grid_case('grid2d')
    tabs = [table_from_dict(tb, t) for t in tables]
    interp = tb.Interpolator(tabs, param_table)
    ngal, xi = interp.predict_batch(draws)
    interp.write(tmp_path / 'grid.hdf5')
    again = tb.Interpolator.read(tmp_path / 'grid.hdf5')
    ngal2, xi2 = again.predict_batch(draws)
    assert np.array_equal(ngal, ngal2) and np.array_equal(xi, xi2)


def test_s_mu_table_to_multipoles(tb):
    """Linearity: multipoles of the (s, mu) prediction == prediction of the multipole table, and
    the transform equals the column-by-column loop of scripts/tabulate_snapshot.py:102-113."""
    n_s, n_mu = 6, 10
    mu_bins = np.linspace(0, 1, n_mu + 1)
    tab = cases.synthetic.make_table(n_mass=15, n_sec=2, n_r=n_s * n_mu, kind='multipole', seed=9,
                                     tpcf_shape=(n_s, n_mu))
    halotab = table_from_dict(tb, tab)
    draws = cases.synthetic.make_draws(200, seed=6, decorated=True)
    ngal, xi_s_mu = halotab.predict_batch(draws)
    assert xi_s_mu.shape == (200, n_s, n_mu)
    scale = np.abs(xi_s_mu).max()
    for order in (0, 2, 4):
        mult = tb.tabcorr_s_mu_to_multipole(halotab, mu_bins, order)
        assert mult.tpcf_shape == (n_s,) and mult.tpcf_matrix.shape == (n_s, tab['tpcf_matrix'].shape[1])
        loop = np.zeros_like(mult.tpcf_matrix)
        for i in range(tab['tpcf_matrix'].shape[1]):
            loop[:, i] = tb.tpcf_multipole(tab['tpcf_matrix'][:, i].reshape(n_s, n_mu), mu_bins,
                                           order=order)
        np.testing.assert_allclose(mult.tpcf_matrix, loop, rtol=1e-13, atol=1e-13 * np.abs(loop).max())
        ngal_l, xi_l = mult.predict_batch(draws)
        assert np.array_equal(ngal_l, ngal)
        expected = tb.tpcf_multipole(xi_s_mu, mu_bins, order=order)
        np.testing.assert_allclose(xi_l, expected, rtol=0, atol=1e-11 * scale * (2 * order + 1))
    # the (s, mu) table is unchanged
    assert halotab.tpcf_shape == (n_s, n_mu)


def test_table_set_per_draw_cosmology(tb):
    """BASELINE configs[3]: every draw names one of C table sets (cosmologies)."""
    axes = {'alpha_s': np.linspace(0.8, 1.2, 4), 'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))}
    interps = []
    for c in range(3):
        tables, param_table = cases.synthetic.make_grid_tables(
            axes, n_mass=10, n_sec=2, n_r=7, seed=40 + c, n_h_scale=1.0 + 0.2 * c)
        interps.append(tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table))
    table_set = tb.TableSet(interps)
    n_draws = 500
    extra = {k: (float(v.min()), float(v.max())) for k, v in axes.items()}
    draws = cases.synthetic.make_draws(n_draws, seed=77, extra=extra)
    index = np.random.default_rng(5).integers(0, 3, n_draws)
    ngal, xi = table_set.predict_batch(draws, index)
    ngal_sep, xi_sep = table_set.predict_batch(draws, index, separate_gal_type=True)
    assert ngal.shape == (n_draws,) and xi.shape == (n_draws, 7)
    for c in range(3):
        rows = np.flatnonzero(index == c)
        sub = {k: v[rows] for k, v in draws.items()}
        ngal_c, xi_c = interps[c].predict_batch(sub)
        assert np.array_equal(ngal[rows], ngal_c) and np.array_equal(xi[rows], xi_c)
        ngal_c, xi_c = interps[c].predict_batch(sub, separate_gal_type=True)
        for key in ngal_c:
            assert np.array_equal(ngal_sep[key][rows], ngal_c[key])
        for key in xi_c:
            assert np.array_equal(xi_sep[key][rows], xi_c[key])
    with pytest.raises(ValueError):
        table_set.predict_batch(draws, index + 1)
    # one draw outside the knot hull of the LAST group: the deferred range check still raises
    outside = {k: v.copy() for k, v in draws.items()}
    outside['alpha_s'][np.flatnonzero(index == 2)[-1]] = 1.5
    with pytest.raises(ValueError, match='interpolation range'):
        table_set.predict_batch(outside, index)
    ngal_x, xi_x = table_set.predict_batch(outside, index, extrapolate=True)
    assert np.all(np.isfinite(xi_x)) and np.array_equal(ngal_x[index == 0], ngal[index == 0])
    ngal_f, xi_f, flag = interps[0].predict_batch({k: v[:3] for k, v in draws.items()},
                                                  as_numpy=False, defer_range_check=True)
    assert int(flag.item()) == 0 and ngal_f.is_cuda
    # plain TabCorr members work too
    plain = tb.TableSet([interp.tabcorr_list[0] for interp in interps])
    ngal_p, xi_p = plain.predict_batch(draws, index)
    ref = interps[1].tabcorr_list[0].predict_batch({k: v[index == 1] for k, v in draws.items()})
    assert np.array_equal(ngal_p[index == 1], ref[0]) and np.array_equal(xi_p[index == 1], ref[1])


def test_sweep_device_side_draws(tb):
    """BASELINE configs[4] in miniature: chunked sweep == one batch over the same draws."""
    import torch
    from tabcorr_b200 import sweep
    tab = cases.synthetic.make_table(n_mass=25, n_sec=2, n_r=8, seed=12)
    halotab = table_from_dict(tb, tab)
    prior = sweep.UniformPrior(sweep.ZHENG07_PRIOR, seed=3)
    n_draws, chunk = 10000, 4096
    ngal, xi = sweep.predict_sweep(halotab, prior, n_draws, chunk=chunk)
    assert ngal.shape == (n_draws,) and xi.shape == (n_draws, 8)
    theta = torch.cat([prior.sample(c, hi - lo, 'cuda')
                       for c, (lo, hi) in enumerate(sweep.chunk_bounds(n_draws, chunk))])
    ngal_b, xi_b = halotab.predict_batch(theta.cpu().numpy())
    assert np.array_equal(ngal, ngal_b) and np.array_equal(xi, xi_b)
    assert theta[:, 0].min() >= 11.0 and theta[:, 0].max() <= 14.0
    seen = []
    done = sweep.predict_sweep(halotab, prior, n_draws, chunk=chunk,
                               consume=lambda lo, hi, slab: seen.append((lo, hi, float(slab[:, 0].sum()))))
    assert done == n_draws and [s[:2] for s in seen] == sweep.chunk_bounds(n_draws, chunk)
    np.testing.assert_allclose(sum(s[2] for s in seen), ngal.sum(), rtol=1e-12)


@pytest.mark.parametrize('chunk', [0, 'auto', 7, 4000, [100, 50000, 100], [35000]])
def test_pipeline_chunk_schedules_agree_bitwise(tb, chunk):
    tab = cases.synthetic.make_table(n_mass=20, n_sec=2, n_r=6, seed=21)
    halotab = table_from_dict(tb, tab)
    draws = cases.synthetic.make_draws(70001 if chunk != 7 else 50, seed=8, decorated=True)
    ref = halotab.predict_batch(draws, pipeline_chunk=0)
    out = halotab.predict_batch(draws, pipeline_chunk=chunk)
    assert np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1])
    sep = halotab.predict_batch(draws, separate_gal_type=True, pipeline_chunk=chunk)
    np.testing.assert_allclose(sep[0]['centrals'] + sep[0]['satellites'], ref[0], rtol=1e-13)


@pytest.mark.parametrize('shape', [(60, 2, 6, 'auto'), (20, 1, 6, 'auto'), (40, 2, 19, 'cross')])
def test_schedule_perturbation_leaves_results_bitwise_equal(tb, shape, monkeypatch):
    """The W tiles of the fused kernel are handed from occupation items to contraction chunks
    (and back, when a buffer is reused) through the `full` / `empty` counters, not through block
    barriers -- something compute-sanitizer's racecheck cannot follow (it reports every such
    hand-over as a hazard, profiles/r02_sanitizer.md).  This test perturbs the schedule instead:
    TC_TUNE_STRESS makes every occupation item and every chunk start after a pseudo-random delay
    of up to 50 us, so writers routinely finish long after readers are ready and the other way
    round; a missing or misplaced wait would read stale weights.  Several tiles per CTA, series
    items (N = 240, cross) and node items (N = 40), batch and precomputed-occupation input."""
    n_mass, n_sec, n_r, mode = shape
    tab = cases.synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=n_r, seed=33, mode=mode)
    halotab = table_from_dict(tb, tab)
    n_draws = 30000
    draws = cases.synthetic.make_draws(n_draws, seed=17, decorated=True)
    ref = halotab.predict_batch(draws, pipeline_chunk=0)
    occ = halotab.mean_occupation_batch(draws).cpu().numpy()
    ref_occ = halotab.predict_batch(None, occupation=occ, pipeline_chunk=0)
    for stress in ('50000', '3000'):
        monkeypatch.setenv('TC_TUNE_STRESS', stress)
        out = halotab.predict_batch(draws, pipeline_chunk=0)
        out_occ = halotab.predict_batch(None, occupation=occ, pipeline_chunk=0)
        monkeypatch.delenv('TC_TUNE_STRESS')
        assert np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1])
        assert np.array_equal(ref_occ[0], out_occ[0]) and np.array_equal(ref_occ[1], out_occ[1])


def test_cfg3_decorated_multipoles_against_oracle(tb):
    """BASELINE configs[2]: xi_0,2,4 (R = 3 x 14), decorated zheng07, n_gauss_prim = 10, at batch
    size 2e4; a sample of draws against the oracle plus the G-convergence property."""
    from oracle import tabcorr_oracle as orc
    tab = cases.synthetic.make_table(n_mass=60, n_sec=2, n_r=42, kind='multipole',
                                     tpcf_shape=(3, 14))
    halotab = table_from_dict(tb, tab)
    n_draws = 20000
    draws = cases.synthetic.make_draws(n_draws, seed=31, decorated=True)
    ngal, xi = halotab.predict_batch(draws)
    assert xi.shape == (n_draws, 3, 14)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    for i in np.random.default_rng(1).integers(0, n_draws, 25):
        model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=True)
        ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model))
        close(ngal[i], ngal_ref)
        # sign-mixed entries: tolerance relative to the magnitude of the terms
        np.testing.assert_allclose(xi[i], xi_ref, rtol=RTOL, atol=RTOL * np.abs(xi_ref).max())
    # n_gauss_prim is honoured (tests/test_general.py:31-43): G = 1 differs, G = 100 matches the
    # oracle's G = 100 (random draws with sigma_logM ~ 0.05 are not converged at G = 10, so the
    # reference's "G = 10 equals G = 100" property is tested on its fiducial model only)
    ngal_1, _ = halotab.predict_batch(draws, n_gauss_prim=1)
    assert not np.allclose(ngal_1, ngal, rtol=1e-6, atol=0)
    ngal_100, xi_100 = halotab.predict_batch(draws, n_gauss_prim=100)
    for i in (0, 1234, n_draws - 1):
        model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=True)
        ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model, 100))
        close(ngal_100[i], ngal_ref)
        np.testing.assert_allclose(xi_100[i], xi_ref, rtol=RTOL, atol=RTOL * np.abs(xi_ref).max())


def _oracle_check(tb, tab, draws, decorated, mode='auto', rows=(0, 1, -1), n_gauss=10):
    from oracle import tabcorr_oracle as orc
    halotab = table_from_dict(tb, tab)
    ngal, xi = halotab.predict_batch(draws, n_gauss_prim=n_gauss)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], mode)
    n_draws = len(ngal)
    for i in rows:
        i = i % n_draws
        model = orc.Zheng07Oracle(cases.draws_row(draws, i), decorated=decorated)
        ngal_ref, xi_ref = orc.predict(table, orc.mean_occupation(table, model, n_gauss))
        close(ngal[i], ngal_ref)
        close(xi[i].ravel(), np.ravel(xi_ref))
    return halotab, ngal, xi


@pytest.mark.parametrize('n_mass,n_sec,n_r,mode', [
    (2, 1, 1, 'auto'), (2, 2, 3, 'auto'), (2, 1, 1, 'cross'), (3, 3, 2, 'auto'), (7, 2, 17, 'cross'),
    (8, 1, 16, 'auto'), (9, 2, 33, 'auto'), (31, 2, 5, 'cross'), (64, 2, 2, 'auto'),
])
def test_edge_shapes_against_oracle(tb, n_mass, n_sec, n_r, mode):
    """Smallest tables, R = 1, row counts around the 16-row tile and 4-row k-step boundaries."""
    tab = cases.synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=n_r, seed=n_mass + n_r,
                                     mode=mode)
    draws = cases.synthetic.make_draws(70, seed=n_r, decorated=True)
    _oracle_check(tb, tab, draws, True, mode, rows=(0, 7, 8, 63, 64, -1))


def test_single_galaxy_type_tables(tb):
    """Tables holding only centrals or only satellites (the reference accepts any gal_type set)."""
    full = cases.synthetic.make_table(n_mass=12, n_sec=2, n_r=4, seed=2)
    draws = cases.synthetic.make_draws(40, seed=5, decorated=True)
    n = len(full['gal_type'])
    rows_i, cols_i = np.tril_indices(n)
    for name in ('centrals', 'satellites'):
        keep = np.flatnonzero(full['gal_type']['gal_type'] == name.encode())
        dense = np.zeros((4, n, n))
        dense[:, rows_i, cols_i] = full['tpcf_matrix']
        dense = dense + np.transpose(np.tril(dense, -1), (0, 2, 1))
        sub = dense[:, keep][:, :, keep]
        r2, c2 = np.tril_indices(len(keep))
        tab = dict(full, gal_type=full['gal_type'][keep], tpcf_matrix=sub[:, r2, c2])
        halotab, ngal, xi = _oracle_check(tb, tab, draws, True)
        ngal_sep, xi_sep = halotab.predict_batch(draws, separate_gal_type=True)
        assert list(ngal_sep) == [name] and list(xi_sep) == ['{0}-{0}'.format(name)]
        np.testing.assert_allclose(ngal_sep[name], ngal, rtol=1e-13)
        np.testing.assert_allclose(xi_sep['{0}-{0}'.format(name)], xi, rtol=1e-12)


def test_table_too_large_for_the_tile_fails_loudly(tb):
    """A table whose padded rows do not fit one 8-draw W tile in shared memory is refused with an
    error naming the limit -- not computed some other way."""
    n_mass = 950   # N = 3800 rows -> 3808 padded rows x 8 draws x 8 B > 227 KB
    gal_type = cases.synthetic.make_gal_type(n_mass, 2)
    matrix = np.ones((1, len(gal_type)))
    attrs = dict(cases.synthetic.make_table(n_mass=2, n_sec=1, n_r=1, mode='cross')['attrs'])
    halotab = tb.TabCorr.from_arrays(gal_type, matrix, (1,), attrs)
    with pytest.raises(Exception, match='too large|workspace|unsupported'):
        halotab.predict_batch(cases.synthetic.make_draws(10, seed=1))


def test_non_finite_parameters_do_not_poison_neighbours(tb):
    """A NaN / inf parameter set yields a non-finite result for that draw only."""
    tab = cases.synthetic.make_table(n_mass=20, n_sec=2, n_r=6, seed=3)
    halotab = table_from_dict(tb, tab)
    draws = cases.synthetic.make_draws(200, seed=9)
    ref = halotab.predict_batch(draws)
    bad = {k: v.copy() for k, v in draws.items()}
    bad['logMmin'][17] = np.nan
    bad['alpha'][101] = np.inf
    out = halotab.predict_batch(bad)
    good = np.ones(200, dtype=bool)
    good[[17, 101]] = False
    assert np.array_equal(out[0][good], ref[0][good]) and np.array_equal(out[1][good], ref[1][good])
    assert not np.isfinite(out[0][17])


# ---------------------------------------------------------------------------------------------
# optional 3xTF32 mode (north_star: "an optional 3xTF32 mode at rtol 1e-6")
# ---------------------------------------------------------------------------------------------
TF32_RTOL = 1e-6


def _abs_table(tb, tab):
    """Same table with |M|: its prediction is the sum of the term magnitudes of xi, the scale a
    relative error bound of a sign-mixed sum has to refer to."""
    return table_from_dict(tb, dict(tab, tpcf_matrix=np.abs(tab['tpcf_matrix'])))


@pytest.mark.parametrize('kw,decorated', [
    (dict(n_mass=60, n_sec=2, n_r=20), False),
    (dict(n_mass=60, n_sec=2, n_r=42, kind='multipole', tpcf_shape=(3, 14)), True),
    (dict(n_mass=13, n_sec=1, n_r=3), True),
    (dict(n_mass=125, n_sec=2, n_r=4), False),
])
def test_3xtf32_mode_within_1e6_of_fp64(tb, kw, decorated):
    tab = cases.synthetic.make_table(**kw)
    halotab = table_from_dict(tb, tab)
    draws = cases.synthetic.make_draws(3000, seed=41, decorated=decorated)
    ngal, xi = halotab.predict_batch(draws)
    ngal_t, xi_t = halotab.predict_batch(draws, precision='3xtf32')
    assert xi_t.shape == xi.shape
    np.testing.assert_allclose(ngal_t, ngal, rtol=TF32_RTOL)
    _, scale = _abs_table(tb, tab).predict_batch(draws)
    err = np.abs(xi_t - xi) / scale
    assert err.max() < TF32_RTOL, err.max()
    assert not np.array_equal(xi_t, xi)          # it really is the other arithmetic
    # per-gal-type split, device-resident and chunked host paths agree with each other bitwise
    ngal_s, xi_s = halotab.predict_batch(draws, separate_gal_type=True, precision='3xtf32')
    total = sum(xi_s.values())
    assert (np.abs(total - xi) / scale).max() < 3 * TF32_RTOL
    again = halotab.predict_batch(draws, precision='3xtf32', pipeline_chunk=[700, 1500, 800])
    assert np.array_equal(again[1], xi_t) and np.array_equal(again[0], ngal_t)


def test_3xtf32_mode_scope(tb, golden_dir):
    import os
    # cross tables: the request is accepted and served in FP64 (occupation-bound path)
    ds = tb.TabCorr.read(os.path.join(golden_dir, 'bolplanck_ds.hdf5'))
    draws = cases.synthetic.make_draws(100, seed=2)
    a, b = ds.predict_batch(draws), ds.predict_batch(draws, precision='3xtf32')
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # the real auto table (float32 on disk: the TF32 split of the table is exact)
    wp = tb.TabCorr.read(os.path.join(golden_dir, 'bolplanck_wp.hdf5'))
    a, b = wp.predict_batch(draws), wp.predict_batch(draws, precision='3xtf32')
    np.testing.assert_allclose(b[0], a[0], rtol=TF32_RTOL)
    scale = table_from_dict(tb, dict(gal_type=wp.gal_type.as_array(), attrs=wp.attrs,
                                     tpcf_matrix=np.abs(wp.tpcf_matrix),
                                     tpcf_shape=wp.tpcf_shape)).predict_batch(draws)[1]
    assert (np.abs(b[1] - a[1]) / scale).max() < TF32_RTOL
    with pytest.raises(ValueError):
        wp.predict_batch(draws, precision='fp16')
    # Interpolator
    tables, param_table, grid_draws = cases.grid_case('grid2d')
    interp = tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table)
    a = interp.predict_batch(grid_draws)
    b = interp.predict_batch(grid_draws, precision='3xtf32')
    np.testing.assert_allclose(b[0], a[0], rtol=TF32_RTOL)
    np.testing.assert_allclose(b[1], a[1], rtol=1e-5, atol=1e-6 * np.abs(a[1]).max())


def test_predict_batch_out_argument(tb):
    """out=(ngal, xi): results land in caller-provided host memory (pipelined and direct path)."""
    import torch
    tab = tb.synthetic.make_table(n_mass=12, n_sec=2, n_r=7)
    halotab = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                     tab['attrs'])
    draws = tb.synthetic.make_draws(40000, seed=3)
    ngal_ref, xi_ref = halotab.predict_batch(draws)
    ngal = torch.empty((40000, 1), dtype=torch.float64, pin_memory=True)
    xi = np.empty((40000, 7, 1))
    got = halotab.predict_batch(draws, out=(ngal, xi))
    assert np.array_equal(got[0], ngal_ref) and np.array_equal(got[1], xi_ref)
    assert np.array_equal(ngal.numpy()[:, 0], ngal_ref) and np.array_equal(xi[:, :, 0], xi_ref)
    assert got[1].base is not None   # a view of the caller's buffer
    xi[:] = 0
    halotab.predict_batch(draws, out=(ngal, xi), pipeline_chunk=0)
    assert np.array_equal(xi[:, :, 0], xi_ref)
    sep = (np.empty((40000, 2)), np.empty((40000, 7, 3)))
    ngal_d, xi_d = halotab.predict_batch(draws, separate_gal_type=True, out=sep)
    ref_d = halotab.predict_batch(draws, separate_gal_type=True)
    for key in xi_d:
        assert np.array_equal(xi_d[key], ref_d[1][key])
    with pytest.raises(ValueError, match='out must hold'):
        halotab.predict_batch(draws, out=(np.empty((40000, 2)), xi))
    with pytest.raises(ValueError, match='as_numpy'):
        halotab.predict_batch(draws, out=(ngal, xi), as_numpy=False)


def test_small_batches_take_the_zero_copy_path_and_agree_bitwise(tb):
    """Batches up to SMALL_BATCH draws run through persistent pinned buffers; same kernels, so
    the results equal those of the general path bit for bit."""
    from tabcorr_b200 import tabcorr as tc_mod
    tab = tb.synthetic.make_table(n_mass=12, n_sec=2, n_r=7)
    halotab = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                     tab['attrs'])
    big = tb.synthetic.make_draws(tc_mod.SMALL_BATCH + 1000, seed=9, decorated=True)
    ngal_ref, xi_ref = halotab.predict_batch(big)            # general (pipelined) path
    sep_ref = halotab.predict_batch(big, separate_gal_type=True)
    for n in (1, 7, 64, tc_mod.SMALL_BATCH):
        small = {k: v[:n] for k, v in big.items()}
        ngal, xi = halotab.predict_batch(small)
        assert halotab._ensure_device()._small is not None
        assert ngal.shape == (n,) and xi.shape == (n, 7)
        assert np.array_equal(ngal, ngal_ref[:n]) and np.array_equal(xi, xi_ref[:n])
        ngal_d, xi_d = halotab.predict_batch(small, separate_gal_type=True)
        for key in xi_d:
            assert np.array_equal(xi_d[key], sep_ref[1][key][:n])
        for key in ngal_d:
            assert np.array_equal(ngal_d[key], sep_ref[0][key][:n])
    # array input, scalars broadcast, results are copies (a second call must not overwrite them)
    theta = np.stack([big[k][:5] for k in tb.models.THETA_KEYS], axis=1)
    decorated = tb.models.ModelSpec(decorated=True)   # arrays carry no keys: the model says so
    first = halotab.predict_batch(theta, model=decorated)
    second = halotab.predict_batch(theta[::-1].copy(), model=decorated)
    assert np.array_equal(first[1], xi_ref[:5]) and np.array_equal(second[1], xi_ref[:5][::-1])
    mixed = dict({k: v[:4] for k, v in big.items()}, alpha=1.0)
    ref = halotab.predict_batch(dict({k: v[:4] for k, v in big.items()}, alpha=np.ones(4)),
                                pipeline_chunk=0, as_numpy=False)
    assert np.array_equal(halotab.predict_batch(mixed)[1], ref[1].cpu().numpy())


def test_sweep_over_interpolator_and_other_family(tb):
    """Sweeps with interpolation coordinates in the prior (an Interpolator) and with the
    leauthaud11 family: the chunked device-side sweep equals one batch over the same draws."""
    import torch
    from tabcorr_b200 import sweep
    tables, param_table, _ = cases.grid_case('grid2d')
    interp = tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table)
    bounds = dict(sweep.ZHENG07_PRIOR)
    bounds.update({k: (0.0, 0.0) for k in tb.models.ASSEMBIAS_KEYS})
    bounds[tb.models.ASSEMBIAS_KEYS[0]] = (-1.0, 1.0)
    bounds[tb.models.ASSEMBIAS_KEYS[1]] = (-1.0, 1.0)
    bounds['log_eta'] = (float(np.min(param_table['log_eta'])), float(np.max(param_table['log_eta'])))
    bounds['alpha_s'] = (0.8, 1.2)
    prior = sweep.UniformPrior(bounds, seed=11)
    n_draws, chunk = 5000, 2048
    ngal, xi = sweep.predict_sweep(interp, prior, n_draws, chunk=chunk)
    sample = torch.cat([prior.sample(c, hi - lo, 'cuda')
                        for c, (lo, hi) in enumerate(sweep.chunk_bounds(n_draws, chunk))]).cpu().numpy()
    params = {k: sample[:, j] for j, k in enumerate(prior.keys)}
    ngal_ref, xi_ref = interp.predict_batch(params)
    assert np.array_equal(ngal, ngal_ref) and np.array_equal(xi, xi_ref)
    with pytest.raises(ValueError, match='interpolation coordinates'):
        sweep.predict_sweep(interp, sweep.UniformPrior(sweep.ZHENG07_PRIOR), 10)
    # leauthaud11 on a single table
    tab = cases.synthetic.make_table(n_mass=25, n_sec=2, n_r=8, seed=12)
    halotab = table_from_dict(tb, tab)
    model = tb.PrebuiltHodModelFactory('leauthaud11', threshold=10.5)
    l11 = sweep.UniformPrior({'smhm_m1_0': (12.1, 12.6), 'alphasat': (0.9, 1.1),
                              'scatter_model_param1': (0.15, 0.3)}, seed=2)
    with pytest.raises(ValueError):   # parameters the prior does not name would be 0
        pass_through = {k: (v, v) for k, v in model.param_dict.items()}
        pass_through['not_a_parameter'] = (0.0, 1.0)
        sweep.predict_sweep(halotab, sweep.UniformPrior(pass_through), 10, model=model)
    full = {k: (v, v) for k, v in model.param_dict.items()}
    full.update({'smhm_m1_0': (12.1, 12.6), 'alphasat': (0.9, 1.1)})
    l11 = sweep.UniformPrior(full, seed=2)
    ngal, xi = sweep.predict_sweep(halotab, l11, 3000, chunk=1024, model=model)
    sample = torch.cat([l11.sample(c, hi - lo, 'cuda')
                        for c, (lo, hi) in enumerate(sweep.chunk_bounds(3000, 1024))]).cpu().numpy()
    ref = halotab.predict_batch({k: sample[:, j] for j, k in enumerate(l11.keys)}, model=model,
                                pipeline_chunk=0)
    assert np.array_equal(ngal, ref[0]) and np.array_equal(xi, ref[1])


def test_interpolator_small_batches_agree_bitwise(tb):
    """Interpolator.predict_batch with few host draws takes the zero-copy path (persistent pinned
    buffers); the results equal the general path's bit for bit, errors are reported alike."""
    tables, param_table, _ = cases.grid_case('grid2d')
    interp = tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table)
    cap = interp._small_capacity()
    assert 1 <= cap <= 4096
    extra = {'alpha_s': (0.8, 1.2), 'log_eta': (float(np.min(param_table['log_eta'])),
                                               float(np.max(param_table['log_eta'])))}
    big = cases.synthetic.make_draws(cap + 500, seed=14, decorated=True, extra=extra)
    ngal_ref, xi_ref = interp.predict_batch(big)             # general path
    sep_ref = interp.predict_batch(big, separate_gal_type=True)
    for n in (1, 5, 64, cap):
        small = {k: v[:n] for k, v in big.items()}
        ngal, xi = interp.predict_batch(small)
        assert interp._one is not None and ngal.shape == (n,)
        assert np.array_equal(ngal, ngal_ref[:n]) and np.array_equal(xi, xi_ref[:n])
        ngal_d, xi_d = interp.predict_batch(small, separate_gal_type=True)
        for key in xi_d:
            assert np.array_equal(xi_d[key], sep_ref[1][key][:n])
        for key in ngal_d:
            assert np.array_equal(ngal_d[key], sep_ref[0][key][:n])
    outside = {k: v[:9].copy() for k, v in big.items()}
    outside['alpha_s'][4] = 1.7
    with pytest.raises(ValueError, match='interpolation range'):
        interp.predict_batch(outside)
    ngal_x, xi_x = interp.predict_batch(outside, extrapolate=True)
    assert np.all(np.isfinite(xi_x)) and np.array_equal(ngal_x[:4], ngal_ref[:4])
    # scalars broadcast; the single-model API still agrees with the batch
    model = tb.PrebuiltHodModelFactory('decorated-zheng07', threshold=-20)
    model.param_dict.update({k: float(v[3]) for k, v in big.items()})
    ngal_1, xi_1 = interp.predict(model)
    assert ngal_1 == ngal_ref[3] and np.array_equal(xi_1, xi_ref[3])


def test_concurrent_host_threads(tb):
    """SURVEY 8(b) ownership/threading: the reference mutates `self` lazily on the first call;
    here the derived tables are immutable and the persistent latency buffers are guarded, so host
    threads may share one table (and one Interpolator) -- every call returns what a serial call
    returns."""
    import threading
    tab = tb.synthetic.make_table(n_mass=20, n_sec=2, n_r=9)
    halotab = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                     tab['attrs'])
    tables, param_table, _ = cases.grid_case('grid1dx')
    interp = tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table)
    lo, hi = float(np.min(param_table['log_eta'])), float(np.max(param_table['log_eta']))
    n_threads, n_calls = 4, 40
    draws = [tb.synthetic.make_draws(7 + 13 * k, seed=50 + k, decorated=True,
                                     extra={'log_eta': (lo, hi)}) for k in range(n_threads)]
    serial = [(halotab.predict_batch(d), halotab.predict_batch(d, n_gauss_prim=3),
               interp.predict_batch(d)) for d in draws]
    errors = []

    def worker(k):
        try:
            for call in range(n_calls):
                a = halotab.predict_batch(draws[k])
                b = halotab.predict_batch(draws[k], n_gauss_prim=3)
                c = interp.predict_batch(draws[k])
                for got, ref in zip((a, b, c), serial[k]):
                    if not (np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])):
                        errors.append((k, call))
                        return
        except Exception as exc:   # pragma: no cover - reported below
            errors.append((k, repr(exc)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert errors == []


def test_two_streams_share_a_table(tb):
    """Launches on different CUDA streams may overlap: each stream has its own scratch buffer."""
    import threading
    import torch
    tab = tb.synthetic.make_table(n_mass=30, n_sec=2, n_r=6)
    halotab = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                     tab['attrs'])
    draws = [tb.synthetic.make_draws(20000, seed=60 + k) for k in range(2)]
    theta = [torch.from_numpy(tb.models.theta_from_params(d, None, tb.models.ModelSpec())).cuda()
             for d in draws]
    serial = [halotab.predict_batch(t, as_numpy=False) for t in theta]
    torch.cuda.synchronize()
    results = [None, None]

    def worker(k):
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for _ in range(10):
                results[k] = halotab.predict_batch(theta[k], as_numpy=False)
        stream.synchronize()

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in range(2):
        assert torch.equal(results[k][0], serial[k][0]) and torch.equal(results[k][1], serial[k][1])
    assert len(halotab._ensure_device()._workspace) >= 2


# ---------------------------------------------------------------------------------------------
# the 3xTF32 mode on tcgen05 / TMEM / TMA (csrc/tcgen05_contract.cuh) against the other paths
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('kw,n_draws', [
    (dict(n_mass=60, n_sec=2, n_r=20), 5000),        # headline shape: 4 column blocks, 4 K segments
    (dict(n_mass=60, n_sec=2, n_r=7), 300),          # odd number of radial bins (half-used MMA)
    (dict(n_mass=30, n_sec=1, n_r=19), 129),         # one column block, one segment, 2 tiles
    (dict(n_mass=13, n_sec=1, n_r=3), 1),            # 26 rows: padding everywhere, one draw
    (dict(n_mass=47, n_sec=1, n_r=5), 777),          # 94 rows: partial second column block
    (dict(n_mass=64, n_sec=2, n_r=2), 2500),         # 256 rows: the largest eligible table
])
def test_tcgen05_contraction_against_fp64_and_warp_level_tf32(tb, kw, n_draws, monkeypatch):
    """Batches on auto tables of at most 256 padded rows run the 3xTF32 mode on the
    5th-generation tensor cores.  Same accuracy bar as the warp-level path (1e-6 of the sum of the
    term magnitudes); the two TF32 paths agree with each other to 2e-6; results do not depend on
    how a batch is cut; ngal is the FP64 value."""
    import torch
    from tabcorr_b200.models import ModelSpec, theta_from_params
    tab = cases.synthetic.make_table(**kw)
    halotab = table_from_dict(tb, tab)
    draws = cases.synthetic.make_draws(n_draws, seed=77)
    theta = torch.from_numpy(theta_from_params(draws, None, ModelSpec())).cuda()
    ngal, xi = halotab.predict_batch(theta, as_numpy=False)
    _, scale = _abs_table(tb, tab).predict_batch(theta, as_numpy=False)
    ngal_t, xi_t = halotab.predict_batch(theta, as_numpy=False, precision='3xtf32')
    assert torch.isfinite(xi_t).all()
    err = ((xi_t - xi).abs() / scale).max().item()
    assert err < TF32_RTOL, err
    assert not torch.equal(xi_t, xi)
    np.testing.assert_allclose(ngal_t.cpu().numpy(), ngal.cpu().numpy(), rtol=1e-13)
    # the warp-level TF32 MMA path (TC_TUNE_TCGEN=0) is the other arithmetic of the same mode
    monkeypatch.setenv('TC_TUNE_TCGEN', '0')
    _, xi_w = halotab.predict_batch(theta, as_numpy=False, precision='3xtf32')
    monkeypatch.delenv('TC_TUNE_TCGEN')
    assert not torch.equal(xi_w, xi_t)
    assert ((xi_w - xi_t).abs() / scale).max().item() < 2 * TF32_RTOL
    # bitwise independent of the batch: a sub-batch, and the host path cut into chunks
    if n_draws > 3:
        lo, hi = n_draws // 3, n_draws // 3 + max(1, n_draws // 2)
        _, xi_sub = halotab.predict_batch(theta[lo:hi].contiguous(), as_numpy=False,
                                          precision='3xtf32')
        assert torch.equal(xi_sub, xi_t[lo:hi])
        host = halotab.predict_batch(draws, precision='3xtf32')
        assert np.array_equal(host[1], xi_t.cpu().numpy())


def test_tcgen05_interpolator_group(tb):
    """An Interpolator group stacks its tables as extra radial bins of one launch; the spline of
    the 3xTF32 results stays within the mode's tolerance of the FP64 prediction."""
    tables, param_table, grid_draws = cases.grid_case('grid2d')
    interp = tb.Interpolator([table_from_dict(tb, t) for t in tables], param_table)
    big = {k: np.tile(v, 40) for k, v in grid_draws.items()}
    a = interp.predict_batch(big)
    b = interp.predict_batch(big, precision='3xtf32')
    np.testing.assert_allclose(b[0], a[0], rtol=TF32_RTOL)
    np.testing.assert_allclose(b[1], a[1], rtol=1e-5, atol=1e-6 * np.abs(a[1]).max())
    assert not np.array_equal(a[1], b[1])
