#!/usr/bin/env python
"""Where one small host-to-host prediction spends its time (N=60 real table, B=1 and 64): Python
before the library call, the (asynchronous) tc_predict_batch call, the stream synchronisation, and
Python after it; next to the floor of this box (an empty torch kernel + synchronise)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import tabcorr_b200
    from tabcorr_b200 import _lib, synthetic, tabcorr as tc_mod
    from tabcorr_b200.models import ModelSpec, theta_columns
    halotab = tabcorr_b200.TabCorr.read(os.path.join(ROOT, 'tests', 'golden', 'bolplanck_wp.hdf5'))
    group = halotab._ensure_device()
    x = torch.zeros(1, device='cuda')
    stream = torch.cuda.current_stream()
    for _ in range(100):
        x.add_(1.0)
        stream.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        x.add_(1.0)
        stream.synchronize()
    print(json.dumps({'floor_us_torch_kernel_plus_sync': (time.perf_counter() - t0) / 2000 * 1e6}))
    for n in (1, 64):
        draws = synthetic.make_draws(n, seed=2)
        halotab.predict_batch(draws)
        buf = group._small
        spec = ModelSpec()
        model = group._model_struct(spec)
        reps = 2000
        acc = np.zeros(4)
        for _ in range(reps):
            t0 = time.perf_counter()
            columns = theta_columns(draws, spec)
            for j, column in enumerate(columns):
                buf.theta_np[j, :n] = column
            t1 = time.perf_counter()
            _lib.check(group.lib.tc_predict_batch(
                group.handle, ctypes.byref(model), 10, buf.theta.data_ptr(), buf.capacity, None, n,
                0, 0, buf.ngal.data_ptr(), 1, buf.xi.data_ptr(), group.n_r,
                buf.workspace.data_ptr(), buf.workspace.numel(), stream.cuda_stream))
            t2 = time.perf_counter()
            stream.synchronize()
            t3 = time.perf_counter()
            ngal = buf.ngal_np[:n].copy()
            xi = buf.xi_np[:n * group.n_r].reshape(n, group.n_r).copy()
            t4 = time.perf_counter()
            acc += (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
        acc *= 1e6 / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            halotab.predict_batch(draws)
        total = (time.perf_counter() - t0) / reps * 1e6
        print(json.dumps({'n_draws': n, 'python_before_us': acc[0], 'library_call_us': acc[1],
                          'synchronize_us': acc[2], 'python_after_us': acc[3],
                          'predict_batch_total_us': total}))


if __name__ == '__main__':
    main()
