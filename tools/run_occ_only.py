"""Launch the standalone occupation kernel a few times (target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tabcorr_b200
from tabcorr_b200 import synthetic
from tabcorr_b200.models import ModelSpec, theta_from_params
n_mass = int(sys.argv[1]) if len(sys.argv) > 1 else 60
tab = synthetic.make_table(n_mass=n_mass, n_sec=2, n_r=4)
halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
draws = synthetic.make_draws(100000, seed=1)
theta = torch.from_numpy(theta_from_params(draws, None, ModelSpec())).cuda()
group = halotab._ensure_device()
for _ in range(4):
    occ = group.occupation(ModelSpec(), 10, theta)
torch.cuda.synchronize()
print(float(occ.sum()))
