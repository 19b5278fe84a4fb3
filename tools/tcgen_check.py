"""Error and time of the tcgen05 3xTF32 contraction against FP64 (and against the warp-level TF32
MMA path, TC_TUNE_TCGEN=0) for the K-segment counts TC_TUNE_TCGEN_SEG.  python tools/tcgen_check.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tabcorr_b200
from tabcorr_b200 import synthetic
from tabcorr_b200.models import ModelSpec, theta_from_params

shapes = [dict(n_mass=60, n_sec=2, n_r=20), dict(n_mass=30, n_sec=1, n_r=19),
          dict(n_mass=60, n_sec=2, n_r=42, kind='multipole'), dict(n_mass=60, n_sec=1, n_r=20)]
if os.environ.get('TCGEN_SHAPES'):
    shapes = [shapes[int(i)] for i in os.environ['TCGEN_SHAPES'].split(',')]
n_draws = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for shape in shapes:
    tab = synthetic.make_table(**shape)
    mk = lambda t: tabcorr_b200.TabCorr.from_arrays(t['gal_type'], t['tpcf_matrix'], t['tpcf_shape'], t['attrs'])
    halotab, habs = mk(tab), mk(dict(tab, tpcf_matrix=np.abs(tab['tpcf_matrix'])))
    draws = synthetic.make_draws(n_draws, seed=1)
    theta = torch.from_numpy(theta_from_params(draws, None, ModelSpec())).cuda()
    ngal, xi = halotab.predict_batch(theta, as_numpy=False)
    _, scale = habs.predict_batch(theta, as_numpy=False)
    torch.cuda.synchronize()
    for knobs in [dict(kv.split('=') for kv in item.split(',')) for item in os.environ.get('TCGEN_KNOBS', 'TCGEN=0;TCGEN_SEG=1;TCGEN_SEG=2;TCGEN_SEG=4').split(';')]:
        for key in [k for k in os.environ if k.startswith('TC_TUNE_')]:
            del os.environ[key]
        for k, v in knobs.items():
            os.environ['TC_TUNE_' + k] = v
        ms = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ngal_t, xi_t = halotab.predict_batch(theta, as_numpy=False, precision='3xtf32')
            b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        err = ((xi_t - xi).abs() / scale)
        bad = int(torch.isnan(xi_t).sum())
        print(json.dumps({'shape': shape, 'knobs': knobs, 'ms': float(np.median(ms[2:])),
                          'preds_per_s': n_draws / float(np.median(ms[2:])) * 1e3, 'nan': bad,
                          'max_err': float(err.nan_to_num(9.0).max()), 'mean_err': float(err.nan_to_num(9.0).mean()),
                          'mean_signed': float(((xi_t - xi) / scale).nan_to_num(9.0).mean()),
                          'ngal_err': float(((ngal_t - ngal).abs() / ngal).max()),
                          'first': [float(v) for v in xi_t[0, :3].flatten()[:3]],
                          'first_ref': [float(v) for v in xi[0, :3].flatten()[:3]]}), flush=True)
