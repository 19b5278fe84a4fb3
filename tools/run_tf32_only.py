"""Launch the fused kernel in 3xTF32 mode a few times on the headline workload (target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tabcorr_b200
from tabcorr_b200 import synthetic
from tabcorr_b200.models import ModelSpec, theta_from_params
tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
draws = synthetic.make_draws(100000, seed=1)
theta = torch.from_numpy(theta_from_params(draws, None, ModelSpec())).cuda()
for _ in range(5):
    out = halotab.predict_batch(theta, as_numpy=False, precision='3xtf32')
torch.cuda.synchronize()
print(float(out[0].sum()))
