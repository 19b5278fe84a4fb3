"""A few launches of the leauthaud11 occupation kernel for ncu (N=240, G=10, 1e5 draws)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tabcorr_b200
from tabcorr_b200 import synthetic, models
from tabcorr_b200.models import ModelSpec, theta_from_params
tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
group = halotab._ensure_device()
spec = ModelSpec(models.FAMILY_LEAUTHAUD11, False, True, 0.5, 10.5, 0.0)
theta = torch.from_numpy(theta_from_params(synthetic.make_draws_leauthaud11(100000, seed=1), None, spec)).cuda()
for _ in range(4):
    occ = group.occupation(spec, 10, theta)
torch.cuda.synchronize()
print(float(occ.sum()))
