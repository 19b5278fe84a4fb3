#!/usr/bin/env python
"""Halo-bin reduction (tc_halo_bins, csrc/halo_bins.cuh) on the GPU against its HBM roofline, with
the reference's CPU path (np.histogram2d + sort_into_bins + per-cell mean, tabcorr/tabcorr.py:
194-227 as restated in oracle/) timed beside it on a sample.

    python tools/bench_halo_bins.py [--halos 50000000]
"""

import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--halos', type=int, default=50_000_000)
    parser.add_argument('--cpu-sample', type=int, default=5_000_000)
    args = parser.parse_args()
    import torch
    from tabcorr_b200 import _lib, halo_bins
    from oracle import tabcorr_oracle as orc
    lib = _lib.load()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    n = args.halos
    gen = torch.Generator(device='cuda').manual_seed(1)
    log_m = 10.7 + torch.empty(n, dtype=torch.float64, device='cuda').exponential_(1 / 0.45, generator=gen)
    log_m.clamp_(max=15.0)
    prim = torch.pow(10.0, log_m)
    sec = torch.rand(n, dtype=torch.float64, device='cuda', generator=gen)
    for n_prim, n_sec in ((30, 1), (60, 2), (100, 4)):
        log_bins = np.linspace(10.7 - 1e-3, 15.0 + 1e-3, n_prim + 1)
        pct_bins = np.linspace(-1e-3, 1 + 1e-3, n_sec + 1)
        pe = log_bins.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        se = pct_bins.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        out = [np.empty(n_prim * n_sec) for _ in range(3)]
        ptrs = [o.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) for o in out]
        stream = torch.cuda.current_stream().cuda_stream

        def run():
            _lib.check(lib.tc_halo_bins(0, log_m.data_ptr(), sec.data_ptr(), prim.data_ptr(), n,
                                        pe, n_prim, se, n_sec, ptrs[0], ptrs[1], ptrs[2], stream))
        for _ in range(2):
            run()
        ms = []
        for _ in range(5):   # the call synchronises: wall clock of the whole call (kernel + tiny copies)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            ms.append((time.perf_counter() - t0) * 1e3)
        ms = float(np.median(ms))
        gbs = 24.0 * n / (ms * 1e-3) / 1e9
        line = {'n_halos': n, 'cells': [n_prim, n_sec], 'ms': ms, 'halos_per_s': n / (ms * 1e-3),
                'algorithmic_GB_per_s': gbs, 'bytes_per_halo': 24,
                'n_h_total': float(out[0].sum())}
        hbm = peaks.get('hbm_gbs') or peaks.get('hbm_gbps')
        if isinstance(hbm, dict):
            hbm = hbm.get('burst') or hbm.get('sustained')
        if hbm:
            line['hbm_peak_GB_per_s'] = hbm
            line['frac_of_hbm_peak'] = gbs / hbm
        if (n_prim, n_sec) == (60, 2):
            m = min(args.cpu_sample, n)
            prim_h, sec_h = prim[:m].cpu().numpy(), sec[:m].cpu().numpy()
            t0 = time.perf_counter()
            ref = orc.halo_bin_table(prim_h, sec_h, log_bins, pct_bins)
            cpu_s = time.perf_counter() - t0
            line['cpu_reference_halos_per_s'] = m / cpu_s
            line['cpu_sample'] = m
            got = halo_bins.halo_bin_counts(prim[:m], sec[:m], log_bins, pct_bins)
            line['counts_equal_on_sample'] = bool(np.array_equal(got[0], ref['n_h']))
        print(json.dumps(line))


if __name__ == '__main__':
    main()
