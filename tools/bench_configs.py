#!/usr/bin/env python
"""Device-timed throughput of every BASELINE.json configuration that runs without halotools
(SURVEY.md section 8(d)); one JSON line per configuration.  Not the headline bench (bench.py is);
the lines go to profiles/ with the command that made them.

    python tools/bench_configs.py [--draws 100000] [--reps 5] [--only cfg3]
    torchrun --nproc-per-node N tools/bench_configs.py --only cfg5 --sweep-draws 100000000
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
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--draws', type=int, default=100000)
    parser.add_argument('--reps', type=int, default=5)
    parser.add_argument('--only', default=None)
    parser.add_argument('--sweep-draws', type=int, default=1 << 23)
    parser.add_argument('--sweep-chunk', type=int, default=1 << 20)
    args = parser.parse_args()
    import torch
    import torch.distributed as dist
    import tabcorr_b200
    from tabcorr_b200 import _lib, synthetic, sweep
    from tabcorr_b200.models import Zheng07Model

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = _lib.load()
    peak = ctypes.c_double()
    _lib.check(lib.tc_measure_dmma_peak(local_rank, ctypes.byref(peak)))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')

    def emit(**line):
        if rank == 0:
            line['dmma_peak_tflops'] = peak.value
            print(json.dumps(line), flush=True)

    def timed(fn, reps=None):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = []
        for i in range(reps or args.reps):
            flush.fill_(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    def executed_flops(n, n_r, mode):
        n_pad = (n + 15) // 16 * 16
        if mode == 'auto':
            return 2.0 * n_r * 64.0 * (n_pad // 8) * (n_pad // 8 + 1) / 2
        return 2.0 * ((n_r + 15) // 16 * 16) * n_pad

    def algorithmic_flops(n, n_r, mode):
        return 2.0 * n_r * n * n + 2.0 * n_r * n if mode == 'auto' else 2.0 * n_r * n

    def want(name):
        return args.only is None or args.only in name or name in args.only

    def table_of(tab):
        return tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                                tab['tpcf_shape'], tab['attrs'])

    def device_run(halotab, draws, decorated=False, **kw):
        """ms of one device-resident predict_batch over the draws (parameters already in HBM)."""
        from tabcorr_b200.models import ModelSpec, theta_from_params
        theta = torch.from_numpy(theta_from_params(draws, None, ModelSpec(decorated=decorated))).cuda()
        spec = ModelSpec(decorated=decorated)
        return timed(lambda: halotab.predict_batch(theta, model=spec, as_numpy=False, **kw))

    n_draws = args.draws
    golden = os.path.join(ROOT, 'tests', 'golden')

    # ---- cfg1: the README workflow's real tables (what TabCorr.tabulate wrote), zheng07 ------------
    if want('cfg1') and world == 1:
        for fname, threshold in (('bolplanck_wp.hdf5', -18), ('bolplanck_ds.hdf5', -21)):
            halotab = tabcorr_b200.TabCorr.read(os.path.join(golden, fname))
            model = Zheng07Model(threshold=threshold, redshift=0.0)
            halotab.predict(model)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_calls = 300
            for _ in range(n_calls):
                halotab.predict(model)
            latency_us = (time.perf_counter() - t0) / n_calls * 1e6
            draws = synthetic.make_draws(n_draws, seed=1)
            ms = device_run(halotab, draws)
            n, n_r = len(halotab.gal_type), int(np.prod(halotab.tpcf_shape))
            mode = halotab.attrs['mode']
            emit(config='cfg1 ' + fname, n_tracers=n, n_r=n_r, mode=mode, n_draws=n_draws,
                 predict_model_latency_us=latency_us, batch_ms=ms,
                 preds_per_s=n_draws / ms * 1e3,
                 executed_frac=executed_flops(n, n_r, mode) * n_draws / (ms * 1e-3) / (peak.value * 1e12))

    # ---- cfg2 / cfg3: synthetic N = 240 tables -------------------------------------------------------
    shapes = [
        ('cfg2 N=240 R=20 wp zheng07', dict(n_mass=60, n_sec=2, n_r=20), False, 10),
        ('cfg2a N=120 R=20 wp zheng07', dict(n_mass=60, n_sec=1, n_r=20), False, 10),
        ('cfg3 N=240 R=3x14 multipoles decorated G=10',
         dict(n_mass=60, n_sec=2, n_r=42, kind='multipole', tpcf_shape=(3, 14)), True, 10),
        ('cfg3 same, G=1', dict(n_mass=60, n_sec=2, n_r=42, kind='multipole', tpcf_shape=(3, 14)),
         True, 1),
        ('cfg3 same, G=100', dict(n_mass=60, n_sec=2, n_r=42, kind='multipole', tpcf_shape=(3, 14)),
         True, 100),
        ('cfg5-shape N=500 R=20 wp zheng07', dict(n_mass=125, n_sec=2, n_r=20), False, 10),
    ]
    for name, kw, decorated, n_gauss in shapes:
        if not want(name) or world > 1:
            continue
        tab = synthetic.make_table(**kw)
        halotab = table_of(tab)
        draws = synthetic.make_draws(n_draws, seed=1, decorated=decorated)
        ms = device_run(halotab, draws, decorated, n_gauss_prim=n_gauss)
        ms_tf32 = device_run(halotab, draws, decorated, n_gauss_prim=n_gauss, precision='3xtf32')
        n, n_r = len(tab['gal_type']), kw['n_r']
        emit(config=name, n_tracers=n, n_r=n_r, mode='auto', n_gauss_prim=n_gauss, n_draws=n_draws,
             batch_ms=ms, preds_per_s=n_draws / ms * 1e3,
             batch_ms_3xtf32=ms_tf32, preds_per_s_3xtf32=n_draws / ms_tf32 * 1e3,
             algorithmic_tflops=algorithmic_flops(n, n_r, 'auto') * n_draws / (ms * 1e-3) * 1e-12,
             executed_frac=executed_flops(n, n_r, 'auto') * n_draws / (ms * 1e-3) / (peak.value * 1e12))

    # ---- cfg4: database-style Interpolators, per-draw cosmology -------------------------------------
    if want('cfg4') and world == 1 and 'real' not in (args.only or ''):
        axes_wp = {'alpha_s': np.linspace(0.8, 1.2, 4), 'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))}
        axes_xi = {'alpha_c': np.linspace(0.0, 0.4, 4), 'alpha_s': np.linspace(0.8, 1.2, 4),
                   'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))}
        for label, axes, kind, n_cosmo in (('wp T=16', axes_wp, 'wp', 8), ('xi_0 T=64', axes_xi, 'multipole', 4)):
            interps = []
            for c in range(n_cosmo):
                tables, param_table = synthetic.make_grid_tables(
                    axes, n_mass=30, n_sec=2, n_r=14, kind=kind, seed=100 + c, n_h_scale=1 + 0.1 * c)
                interps.append(tabcorr_b200.Interpolator([table_of(t) for t in tables], param_table))
            table_set = tabcorr_b200.TableSet(interps)
            extra = {k: (float(v.min()), float(v.max())) for k, v in axes.items()}
            draws = synthetic.make_draws(n_draws, seed=2, extra=extra)
            index = np.random.default_rng(3).integers(0, n_cosmo, n_draws)
            t_tables = int(np.prod([len(v) for v in axes.values()]))
            table_set.predict_batch(draws, index)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                table_set.predict_batch(draws, index)
            ms = (time.perf_counter() - t0) / args.reps * 1e3
            one = {k: v[index == 0] for k, v in draws.items()}
            ms_one = timed(lambda: interps[0].predict_batch(one, as_numpy=False))
            n_one = int((index == 0).sum())
            emit(config='cfg4 database-style {} x {} cosmologies, N=120 R=14, per-draw cosmology'.format(
                     label, n_cosmo), n_tracers=120, n_r=14, grid_tables=t_tables, n_draws=n_draws,
                 host_to_host_ms=ms, preds_per_s_host_to_host=n_draws / ms * 1e3,
                 single_interpolator_device_ms=ms_one, single_interpolator_draws=n_one,
                 single_interpolator_preds_per_s=n_one / ms_one * 1e3,
                 executed_frac_single=executed_flops(120, 14 * t_tables, 'auto') * n_one /
                 (ms_one * 1e-3) / (peak.value * 1e12))
    if want('cfg4') and world == 1:
        interp = tabcorr_b200.Interpolator.read(os.path.join(golden, 'ds_efficient.hdf5'))
        model = Zheng07Model(threshold=-21, redshift=0.5, prim_haloprop_key='halo_m258m')
        model.param_dict['log_eta'] = 0.1
        interp.predict(model)
        t0 = time.perf_counter()
        for _ in range(200):
            interp.predict(model)
        emit(config='cfg4 real ds_efficient.hdf5 Interpolator.predict(model) latency',
             predict_model_latency_us=(time.perf_counter() - t0) / 200 * 1e6)
        draws = synthetic.make_draws(n_draws, seed=4, extra={'log_eta': (-0.47, 0.47)})
        ms = timed(lambda: interp.predict_batch(draws, as_numpy=False))
        emit(config='cfg4 real ds_efficient.hdf5 Interpolator (T=4, cross, N=1104, R=13)',
             n_tracers=1104, n_r=13, grid_tables=4, n_draws=n_draws, batch_ms_incl_h2d=ms,
             preds_per_s=n_draws / ms * 1e3)

    # ---- cfg5: MCMC-scale sweep, N = 500, device-side draws, gather to rank 0 -----------------------
    if want('cfg5'):
        tab = synthetic.make_table(n_mass=125, n_sec=2, n_r=20)
        halotab = table_of(tab)
        prior = sweep.UniformPrior(sweep.ZHENG07_PRIOR, seed=5)
        totals = {'n': 0, 'ngal': None}

        def consume(lo, hi, slab):   # on rank 0: running column sums stay on the device
            s = slab.sum(dim=0)
            totals['ngal'] = s if totals['ngal'] is None else totals['ngal'] + s
            totals['n'] += hi - lo

        def run(n):
            totals['n'], totals['ngal'] = 0, None
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sweep.predict_sweep(halotab, prior, n, chunk=args.sweep_chunk, consume=consume)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            return time.perf_counter() - t0

        run(min(args.sweep_draws, world * args.sweep_chunk))   # warm-up round
        seconds = run(args.sweep_draws)
        if world > 1:
            t = torch.tensor([seconds], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            seconds = float(t.item())
        n = 500
        if rank == 0:
            assert totals['n'] == args.sweep_draws
            emit(config='cfg5 sweep N=500 R=20 wp zheng07, device-side draws, gather to rank 0',
                 n_gpus=world, n_tracers=n, n_r=20, n_draws=args.sweep_draws, chunk=args.sweep_chunk,
                 seconds=seconds, preds_per_s=args.sweep_draws / seconds,
                 mean_ngal=float(totals['ngal'][0].item() / totals['n']),
                 executed_frac_per_gpu=executed_flops(n, 20, 'auto') * args.sweep_draws / seconds /
                 (peak.value * 1e12) / world)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
