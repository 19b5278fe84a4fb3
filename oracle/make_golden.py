"""Record golden vectors by running the UNMODIFIED reference source (container only).

TEST INFRASTRUCTURE.  Run as ``python oracle/make_golden.py`` where /root/reference exists.  It

1. copies the three data fixtures the reference ships for this path into ``tests/golden/``
   (binary table data, not source: ``docs/examples/bolplanck_{wp,ds}.hdf5`` and
   ``tests/AbacusSummit/base_c000_ph000/0p50/ds_efficient.hdf5``),
2. runs the reference's own ``TabCorr.predict`` / ``Interpolator.predict`` (``oracle/refstub.py``)
   on them and on seeded synthetic tables (``tabcorr_b200/synthetic.py``), with the halotools
   occupation restatement ``Zheng07Oracle`` standing in for the (absent) halotools model,
3. writes every output to ``tests/golden/reference_outputs.npz``.

``tests/test_oracle.py`` pins ``oracle/tabcorr_oracle.py`` against that file; the GPU tests pin
the CUDA path against both.
"""

import hashlib
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refstub  # noqa: E402
from oracle.tabcorr_oracle import Zheng07Oracle  # noqa: E402
from tabcorr_b200 import h5mini, synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
FIXTURES = {
    'bolplanck_wp.hdf5': 'docs/examples/bolplanck_wp.hdf5',
    'bolplanck_ds.hdf5': 'docs/examples/bolplanck_ds.hdf5',
    'ds_efficient.hdf5': 'tests/AbacusSummit/base_c000_ph000/0p50/ds_efficient.hdf5',
}

THETA_M18 = dict(logMmin=11.35, sigma_logM=0.25, logM0=11.20, logM1=12.40, alpha=0.83)
THETA_M21 = dict(logMmin=12.79, sigma_logM=0.39, logM0=11.92, logM1=13.94, alpha=1.15)
THETA_AS = dict(logMmin=12.9, sigma_logM=0.25, logM0=11.20, logM1=14.1, alpha=1.2)


def ref_table_from_group(group):
    attrs = {k: group.attrs[k] for k in group.attrs}
    return refstub.make_tabcorr(group['gal_type'][()], group['tpcf_matrix'][()],
                                group['tpcf_shape'][()], attrs)


def ref_table_from_dict(tab):
    return refstub.make_tabcorr(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'],
                                tab['attrs'])


def put_result(out, name, result):
    ngal, xi = result
    if isinstance(ngal, dict):
        for key in ngal:
            out['{}/ngal/{}'.format(name, key)] = np.float64(ngal[key])
        for key in xi:
            out['{}/xi/{}'.format(name, key)] = np.asarray(xi[key], dtype=np.float64)
    else:
        out['{}/ngal'.format(name)] = np.float64(ngal)
        out['{}/xi'.format(name)] = np.asarray(xi, dtype=np.float64)


def draws_row(draws, i):
    return {k: float(v[i]) for k, v in draws.items()}


def main():
    if not refstub.available():
        raise SystemExit('reference checkout not found; golden vectors can only be made in the '
                         'build container')
    os.makedirs(GOLDEN, exist_ok=True)
    for name, rel in FIXTURES.items():
        shutil.copyfile(os.path.join(refstub.REFERENCE_ROOT, rel), os.path.join(GOLDEN, name))
    out = {}

    # ---- real fixtures ------------------------------------------------------------------
    wp = ref_table_from_group(h5mini.File(os.path.join(GOLDEN, 'bolplanck_wp.hdf5')))
    model = Zheng07Oracle(THETA_M18)
    out['bolplanck_wp/occ'] = wp.mean_occupation(model, check_consistency=False)
    for g in (1, 10, 100):
        put_result(out, 'bolplanck_wp/G{}'.format(g),
                   wp.predict(model, n_gauss_prim=g, check_consistency=False))
    put_result(out, 'bolplanck_wp/sep',
               wp.predict(model, separate_gal_type=True, check_consistency=False))
    put_result(out, 'bolplanck_wp/m21', wp.predict(Zheng07Oracle(THETA_M21),
                                                    check_consistency=False))

    ds = ref_table_from_group(h5mini.File(os.path.join(GOLDEN, 'bolplanck_ds.hdf5')))
    model = Zheng07Oracle(THETA_M21)
    put_result(out, 'bolplanck_ds/G10', ds.predict(model, check_consistency=False))
    put_result(out, 'bolplanck_ds/sep',
               ds.predict(model, separate_gal_type=True, check_consistency=False))

    f = h5mini.File(os.path.join(GOLDEN, 'ds_efficient.hdf5'))
    param = f['param_dict_table'][()]
    order = np.argsort(param['tabcorr_index'])
    tables = [ref_table_from_group(f['tabcorr_{}'.format(i)]) for i in range(len(param))]
    interp = refstub.make_interpolator(tables, {'log_eta': param['log_eta'][order]})
    for tag, log_eta in (('a', 0.1), ('b', -0.3), ('knot', float(param['log_eta'][order][1]))):
        model = Zheng07Oracle(dict(THETA_AS, log_eta=log_eta))
        put_result(out, 'ds_efficient/{}'.format(tag),
                   interp.predict(model, check_consistency=False))
        put_result(out, 'ds_efficient/{}_sep'.format(tag),
                   interp.predict(model, separate_gal_type=True, check_consistency=False))
    model = Zheng07Oracle(dict(THETA_AS, log_eta=0.6))
    put_result(out, 'ds_efficient/extrap',
               interp.predict(model, extrapolate=True, check_consistency=False))
    put_result(out, 'ds_efficient/table0',
               tables[0].predict(Zheng07Oracle(THETA_AS), check_consistency=False))

    # ---- synthetic tables (regenerated from seeds by the tests; checksums guard the generator)
    n_draws = 12
    for case, kw, decorated in (
            ('syn240', dict(n_mass=60, n_sec=2, n_r=20), False),
            ('syn240dec', dict(n_mass=60, n_sec=2, n_r=20), True),
            ('syn120', dict(n_mass=60, n_sec=1, n_r=20), False),
            ('syn36x3', dict(n_mass=6, n_sec=3, n_r=5), True),
            ('synmulti', dict(n_mass=60, n_sec=2, n_r=42, kind='multipole',
                              tpcf_shape=(3, 14)), True),
            ('syncross', dict(n_mass=60, n_sec=2, n_r=13, mode='cross'), True)):
        tab = synthetic.make_table(**kw)
        out[case + '/matrix_sha1'] = np.frombuffer(
            hashlib.sha1(np.ascontiguousarray(tab['tpcf_matrix']).tobytes()).digest(), np.uint8)
        out[case + '/gal_type_sha1'] = np.frombuffer(
            hashlib.sha1(tab['gal_type'].tobytes()).digest(), np.uint8)
        ref = ref_table_from_dict(tab)
        draws = synthetic.make_draws(n_draws, seed=11, decorated=decorated)
        for g in ((1, 10, 100) if case == 'syn240dec' else (10,)):
            ngal, xi, occ = [], [], []
            for i in range(n_draws):
                model = Zheng07Oracle(draws_row(draws, i), decorated=decorated)
                occ.append(ref.mean_occupation(model, n_gauss_prim=g, check_consistency=False))
                res = ref.predict(model, n_gauss_prim=g, check_consistency=False)
                ngal.append(res[0])
                xi.append(res[1])
            out['{}/G{}/ngal'.format(case, g)] = np.array(ngal)
            out['{}/G{}/xi'.format(case, g)] = np.array(xi)
            out['{}/G{}/occ'.format(case, g)] = np.array(occ)
        model = Zheng07Oracle(draws_row(draws, 0), decorated=decorated)
        put_result(out, case + '/sep0',
                   ref.predict(model, separate_gal_type=True, check_consistency=False))

    # ---- synthetic interpolator grids -------------------------------------------------------
    for case, axes, kw in (
            ('grid2d', {'alpha_s': np.linspace(0.8, 1.2, 4),
                        'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))},
             dict(n_mass=12, n_sec=2, n_r=14, mode='auto')),
            ('grid3d', {'alpha_c': np.linspace(0.0, 0.4, 4), 'alpha_s': np.linspace(0.8, 1.2, 5),
                        'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))},
             dict(n_mass=8, n_sec=2, n_r=6, mode='auto', kind='multipole')),
            ('grid1dx', {'log_eta': np.linspace(-0.5, 0.5, 6)},
             dict(n_mass=10, n_sec=2, n_r=7, mode='cross'))):
        tables, param_table = synthetic.make_grid_tables(axes, **kw)
        interp = refstub.make_interpolator([ref_table_from_dict(t) for t in tables], param_table)
        extra = {k: (float(np.min(v)), float(np.max(v))) for k, v in axes.items()}
        draws = synthetic.make_draws(n_draws, seed=13, decorated=True, extra=extra)
        ngal, xi = [], []
        for i in range(n_draws):
            model = Zheng07Oracle(draws_row(draws, i), decorated=True)
            res = interp.predict(model, check_consistency=False)
            ngal.append(res[0])
            xi.append(res[1])
        out[case + '/ngal'] = np.array(ngal)
        out[case + '/xi'] = np.array(xi)
        model = Zheng07Oracle(draws_row(draws, 0), decorated=True)
        put_result(out, case + '/sep0',
                   interp.predict(model, separate_gal_type=True, check_consistency=False))

    # spline matrix and evaluation on their own
    ref = refstub.load()
    xp = np.array([-1.0, -0.2, 0.1, 0.9, 1.7, 2.0])
    out['spline/xp'] = xp
    out['spline/a'] = ref.interpolator.spline_interpolation_matrix(xp)
    yp = np.random.default_rng(3).normal(size=(6, 5))
    out['spline/yp'] = yp
    out['spline/x'] = np.array([-1.0, -0.5, 0.1, 1.0, 2.0])
    out['spline/y'] = np.array([ref.interpolator.spline_interpolate(
        x, xp, out['spline/a'], yp) for x in out['spline/x']])

    np.savez_compressed(os.path.join(GOLDEN, 'reference_outputs.npz'), **out)
    print('wrote {} arrays to tests/golden/reference_outputs.npz'.format(len(out)))
    print('KA1 ngal', out['bolplanck_wp/G10/ngal'], 'wp[0]', out['bolplanck_wp/G10/xi'][0])
    print('KA2 ngal', out['bolplanck_ds/G10/ngal'], 'ds[0]', out['bolplanck_ds/G10/xi'][0])
    print('KA3 ngal', out['ds_efficient/a/ngal'], 'ds[0]', out['ds_efficient/a/xi'][0])


if __name__ == '__main__':
    main()
