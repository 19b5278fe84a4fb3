"""Launch the fused kernel a few times on precomputed occupations (contraction only) or on draws:
a target for ncu.  python tools/run_occ_input.py [theta|occ] [n_mass n_sec n_r]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tabcorr_b200
from tabcorr_b200 import synthetic
from tabcorr_b200.models import ModelSpec, theta_from_params
kind = sys.argv[1] if len(sys.argv) > 1 else 'occ'
n_mass, n_sec, n_r = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (60, 2, 20)
tab = synthetic.make_table(n_mass=n_mass, n_sec=n_sec, n_r=n_r)
halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
group = halotab._ensure_device()
draws = synthetic.make_draws(100000, seed=1)
spec = ModelSpec()
theta = torch.from_numpy(theta_from_params(draws, None, spec)).cuda()
occ = group.occupation(spec, 10, theta)
ngal = torch.empty((100000, 1), dtype=torch.float64, device='cuda')
xi = torch.empty((100000, n_r, 1), dtype=torch.float64, device='cuda')
for _ in range(5):
    if kind == 'occ':
        group.predict_into(None, 10, None, occ, False, ngal, 0, xi, 0)
    else:
        group.predict_into(spec, 10, theta, None, False, ngal, 0, xi, 0)
torch.cuda.synchronize()
print('done', float(xi.sum()))
