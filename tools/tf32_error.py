"""Error and time of the 3xTF32 mode against FP64 for several accumulation-chain lengths
(TC_TUNE_TF32_SEG).  python tools/tf32_error.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tabcorr_b200
from tabcorr_b200 import synthetic
from tabcorr_b200.models import ModelSpec, theta_from_params
for shape in (dict(n_mass=60, n_sec=2, n_r=20), dict(n_mass=60, n_sec=2, n_r=42, kind='multipole'),
              dict(n_mass=125, n_sec=2, n_r=20)):
    tab = synthetic.make_table(**shape)
    mk = lambda t: tabcorr_b200.TabCorr.from_arrays(t['gal_type'], t['tpcf_matrix'], t['tpcf_shape'], t['attrs'])
    halotab, habs = mk(tab), mk(dict(tab, tpcf_matrix=np.abs(tab['tpcf_matrix'])))
    draws = synthetic.make_draws(100000, seed=1)
    theta = torch.from_numpy(theta_from_params(draws, None, ModelSpec())).cuda()
    ngal, xi = halotab.predict_batch(theta, as_numpy=False)
    _, scale = habs.predict_batch(theta, as_numpy=False)
    for seg in (8, 1000000):
        os.environ['TC_TUNE_TF32_SEG'] = str(seg)
        ms = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ngal_t, xi_t = halotab.predict_batch(theta, as_numpy=False, precision='3xtf32')
            b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        err = ((xi_t - xi).abs() / scale)
        print(json.dumps({'shape': shape, 'segment': seg, 'ms': float(np.median(ms[2:])),
                          'max_err': float(err.max()), 'mean_err': float(err.mean()),
                          'mean_signed': float(((xi_t - xi) / scale).mean()),
                          'ngal_err': float(((ngal_t - ngal).abs() / ngal).max())}), flush=True)
