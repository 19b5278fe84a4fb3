"""A small hearin15 batch (leauthaud11 occupation kernel + contraction) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tabcorr_b200 as tb
from tabcorr_b200 import synthetic
tab = synthetic.make_table(n_mass=12, n_sec=2, n_r=5)
h = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
m = tb.PrebuiltHodModelFactory('hearin15')
d = synthetic.make_draws_leauthaud11(150, seed=1, decorated=True)
print(h.predict_batch(d, model=m)[0][:3])
