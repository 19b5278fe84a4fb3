"""CUDA-event durations of predict_kernel / finalize_kernel for 1, 64 and 1024 draws with the
parameters and the results in device memory or in mapped pinned host memory (what the zero-copy
paths cost inside the kernels).  python tools/latency_kernels.py"""
import os, sys, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, tabcorr_b200
from tabcorr_b200 import _lib, synthetic
from tabcorr_b200.models import ModelSpec, theta_from_params
lib = _lib.load()
for name in ('N60', 'N240'):
    if name == 'N60':
        h = tabcorr_b200.TabCorr.read(os.path.join(ROOT, 'tests', 'golden', 'bolplanck_wp.hdf5'))
    else:
        tab = synthetic.make_table(n_mass=60, n_sec=2, n_r=20)
        h = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], tab['attrs'])
    g = h._ensure_device()
    spec = ModelSpec()
    for n in (1, 64, 1024):
        th_host = torch.from_numpy(theta_from_params(synthetic.make_draws(n, seed=2), None, spec)).pin_memory()
        th_dev = th_host.cuda()
        for where, th in (('device', th_dev), ('pinned', th_host)):
            for outw in ('device', 'pinned'):
                if outw == 'device':
                    ngal = torch.empty((n, 1), dtype=torch.float64, device='cuda'); xi = torch.empty((n, g.n_r, 1), dtype=torch.float64, device='cuda')
                else:
                    ngal = torch.empty((n, 1), dtype=torch.float64).pin_memory(); xi = torch.empty((n, g.n_r, 1), dtype=torch.float64).pin_memory()
                _lib.check(lib.tc_profile_enable(1))
                ks, fs = [], []
                for i in range(30):
                    g.predict_into(spec, 10, th, None, False, ngal, 0, xi, 0)
                    torch.cuda.synchronize()
                    a, b = ctypes.c_float(), ctypes.c_float()
                    _lib.check(lib.tc_profile_read(ctypes.byref(a), ctypes.byref(b)))
                    if i >= 5:
                        ks.append(a.value); fs.append(b.value)
                _lib.check(lib.tc_profile_enable(0))
                print(json.dumps({'table': name, 'n': n, 'theta': where, 'out': outw, 'predict_us': 1e3 * float(np.median(ks)), 'finalize_us': 1e3 * float(np.median(fs))}))
