"""Launch the tcgen05 3xTF32 path a few times: a target for ncu.  python tools/run_tcgen_only.py [n_mass n_sec n_r]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tabcorr_b200
from tabcorr_b200 import synthetic
from tabcorr_b200.models import ModelSpec, theta_from_params
n_mass, n_sec, n_r = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (60, 2, 20)
tab = synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=n_r)
halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
draws = synthetic.make_draws(100000, seed=1)
theta = torch.from_numpy(theta_from_params(draws, None, ModelSpec())).cuda()
for _ in range(5):
    ngal, xi = halotab.predict_batch(theta, as_numpy=False, precision='3xtf32')
torch.cuda.synchronize()
print('done', float(xi.sum()))
