"""A few small batches (1, 64, 1024 draws; N=60 and N=240) for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/run_latency_shapes.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, tabcorr_b200
from tabcorr_b200 import synthetic
tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
h = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
h60 = tabcorr_b200.TabCorr.read(os.path.join(ROOT, 'tests', 'golden', 'bolplanck_wp.hdf5'))
for t in (h60, h):
    for n in (1, 64, 1024):
        d = synthetic.make_draws(n, seed=2)
        for _ in range(3):
            t.predict_batch(d)
