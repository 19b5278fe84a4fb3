import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch, tabcorr_b200
from tabcorr_b200 import synthetic
tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
h = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
h60 = tabcorr_b200.TabCorr.read('/root/repo/tests/golden/bolplanck_wp.hdf5')
for t in (h60, h):
    for n in (1, 64, 1024):
        d = synthetic.make_draws(n, seed=2)
        for _ in range(3):
            t.predict_batch(d)
