import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import tabcorr_b200 as tb
from tabcorr_b200 import synthetic
tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
h = tb.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
for name in ('zheng07', 'leauthaud11', 'hearin15'):
    m = tb.PrebuiltHodModelFactory(name) if name != 'zheng07' else tb.PrebuiltHodModelFactory('zheng07', threshold=-20)
    for _ in range(20): h.predict(m)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 300
    for i in range(n):
        m.param_dict['alphasat' if name != 'zheng07' else 'alpha'] = 1.0 + 1e-4 * i
        h.predict(m)
    dt = (time.perf_counter() - t0) / n
    print(name, 'predict(model) latency %.1f us' % (dt * 1e6))
