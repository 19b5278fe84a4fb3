"""Small launches of everything round 2 added to the kernels, as a target for compute-sanitizer:
series occupation items with their shared-memory queue (fused auto kernel at N=240, cross kernel,
standalone occupation kernel, tiny-batch item shape), the mass-dependent occupation kernel, the
coalesced occupation-input items and the halo-bin histogram."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tabcorr_b200
from tabcorr_b200 import halo_bins, synthetic
from tabcorr_b200.models import ModelSpec, assembias_keys

total = 0.0
PARTS = os.environ.get('PARTS', 'theta,occ,massdep,halo').split(',')
for mode, n_mass in (('auto', 60), ('cross', 60)):
    tab = synthetic.make_table(n_mass=n_mass, n_sec=2, n_r=5, mode=mode)
    halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                               tab['tpcf_shape'], tab['attrs'])
    for n_draws in (3, 40, 700):
        draws = synthetic.make_draws(n_draws, seed=n_draws, decorated=True)
        draws['sigma_logM'][::7] = 0.01          # small sigma: every pair of the draw is queued
        ngal, xi = halotab.predict_batch(draws) if 'theta' in PARTS else (0.0, 0.0)
        occ = halotab.mean_occupation_batch(draws)
        ngal2, xi2 = halotab.predict_batch(None, occupation=occ.cpu().numpy()) \
            if (mode == 'auto' and 'occ' in PARTS) else (ngal, xi)
        total += float(np.sum(ngal)) + float(np.sum(xi)) + float(occ.sum()) + float(np.sum(xi2))
    spec = ModelSpec(0, True, strength_abscissa=((11.0, 12.5, 14.0), (11.5, 13.0)),
                     split_abscissa=((11.0, 14.5), ()), split_ordinates=((0.25, 0.7), ()))
    draws = synthetic.make_draws(50, seed=2)
    rng = np.random.default_rng(0)
    for key in assembias_keys('centrals', 3) + assembias_keys('satellites', 2):
        draws[key] = rng.uniform(-1, 1, 50)
    ngal, xi = halotab.predict_batch(draws, model=spec) if 'massdep' in PARTS else (0.0, 0.0)
    total += float(np.sum(ngal)) + float(np.sum(xi))
rng = np.random.default_rng(1)
prim = 10**(10.7 + rng.exponential(0.45, 20000))
n_h, members, mean = halo_bins.halo_bin_counts(prim, rng.random(20000),
                                               np.linspace(10.6, 16.0, 31), np.array([-1e-3, 0.5, 1.001]))
total += float(n_h.sum())
torch.cuda.synchronize()
print('done', total)
