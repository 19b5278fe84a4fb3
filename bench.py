#!/usr/bin/env python
"""Benchmark of the TabCorr prediction hot path on B200 (contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): a synthetic
bolplanck-shaped w_p table with N = 240 tracer rows (60 mass bins x 2 secondary-percentile bins x
centrals/satellites), R = 20 radial bins, n_gauss_prim = 10, and a batch of 1e5 zheng07 draws per
GPU.  One step = one prediction (ngal + w_p) of every draw of the batch.  Under torchrun every rank
holds a replica of the table and its own 1e5 draws (weak scaling); the results are collected on
rank 0 with one NCCL gather per step, inside the timed region.

Prints ONE JSON line.  `value` is device-timed throughput with the parameters already in HBM;
`e2e` is the same metric through the public API ``TabCorr.predict_batch`` with host numpy inputs
and outputs (pinned H2D + D2H inside the timed region, wall clock); on N > 1 GPUs through
``predict_batch_sharded(..., gather='host')``: every rank copies its rows into one shared,
CUDA-registered host segment over its own PCIe link and rank 0 reads the whole result there.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_MASS, N_SEC, N_R, N_GAUSS = 60, 2, 20, 10
DRAWS_PER_GPU = 100000
# dram__bytes_read.sum + dram__bytes_write.sum of predict_kernel for this workload (one launch =
# 1e5 draws).  NOT measured by this run: the constant is copied from the committed ncu --set full
# capture named in NCU_TRAFFIC_SOURCE; only reported for the default batch size
NCU_DRAM_BYTES_PER_LAUNCH = 10725632 + 4377344
NCU_TRAFFIC_SOURCE = 'profiles/r02_predict_kernel_N240_R20.md'
# algorithmic HBM bytes per launch: 7 parameters in, 1 + R results out per draw, the table once
ALGORITHMIC_BYTES_PER_DRAW = 8 * 7 + 8 * (1 + N_R)
METRIC = 'HOD predictions/sec (ngal+wp)'
UNIT = 'predictions/s'


def workload_config(n_draws):
    n = 2 * N_SEC * N_MASS
    return {
        'workload': 'BASELINE configs[1]: synthetic bolplanck-shaped wp table, N={} tracer rows '
                    '({} mass x {} sec x cen/sat), R={} r_p bins, n_gauss_prim={}, {} zheng07 '
                    'draws per GPU'.format(n, N_MASS, N_SEC, N_R, N_GAUSS, n_draws),
        'n_tracers': n, 'n_r': N_R, 'n_gauss_prim': N_GAUSS, 'draws_per_gpu': n_draws,
        'l2': 'flushed between timed steps (256 MiB device write)',
    }


def algorithmic_flops(n, r):
    """dense W.M_r plus row-dot, per prediction."""
    return 2.0 * r * n * n + 2.0 * r * n


def executed_flops(n, r, mode='auto'):
    """Tensor flops the kernel issues per prediction: the lower triangle of the padded table in
    8 x 8 DMMA blocks (auto: the reference's packed prefactor-2 sum, tabcorr/tabcorr.py:626-647);
    cross tables: a 16-bin tile of radial bins times the padded rows."""
    n_pad = (n + 15) // 16 * 16
    if mode == 'auto':
        return 2.0 * r * 64.0 * (n_pad // 8) * (n_pad // 8 + 1) / 2
    return 2.0 * ((r + 15) // 16 * 16) * n_pad


# ---------------------------------------------------------------------------------------------
# CPU arms: the oracle port of the reference loop (README.md:72-74 idiom)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """Time the reference algorithm (numpy port) on draws[lo:hi]; returns (n, seconds)."""
    lo, hi, repeats = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from oracle import tabcorr_oracle as orc
    from tabcorr_b200 import synthetic
    tab = synthetic.make_table(n_mass=N_MASS, n_sec=N_SEC, n_r=N_R)
    draws = synthetic.make_draws(DRAWS_PER_GPU, seed=1)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    model = orc.Zheng07Oracle()
    orc.predict(table, orc.mean_occupation(table, model, N_GAUSS))  # warm caches
    t0 = time.perf_counter()
    for _ in range(repeats):
        for i in range(lo, hi):
            for key, values in draws.items():
                model.param_dict[key] = values[i]
            orc.predict(table, orc.mean_occupation(table, model, N_GAUSS))
    return (hi - lo) * repeats, time.perf_counter() - t0


def cpu_baseline_single(n_sample):
    n, seconds = _cpu_worker((0, n_sample, 1))
    return {'value': n / seconds, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': 'first {} of the {} draws, oracle/tabcorr_oracle.py (numpy port of '
                      'TabCorr.predict incl. mean_occupation), 1 process, 1 thread'.format(
                          n_sample, DRAWS_PER_GPU)}


def parity_vs_cpu(ngal_gpu, xi_gpu, n_check=256):
    """Max relative deviation of the GPU results from the CPU port on the first draws of the
    workload (BASELINE.md section 3 asks for it beside the timings)."""
    from oracle import tabcorr_oracle as orc
    from tabcorr_b200 import synthetic
    tab = synthetic.make_table(n_mass=N_MASS, n_sec=N_SEC, n_r=N_R)
    draws = synthetic.make_draws(DRAWS_PER_GPU, seed=1)
    table = orc.OracleTable(tab['gal_type'], tab['tpcf_matrix'], tab['tpcf_shape'], 'auto')
    model = orc.Zheng07Oracle()
    n_check = min(n_check, len(ngal_gpu))
    dev_ngal = dev_xi = 0.0
    for i in range(n_check):
        for key, values in draws.items():
            model.param_dict[key] = values[i]
        ngal, xi = orc.predict(table, orc.mean_occupation(table, model, N_GAUSS))
        dev_ngal = max(dev_ngal, abs(ngal_gpu[i] / ngal - 1.0))
        dev_xi = max(dev_xi, float(np.max(np.abs(xi_gpu[i] - xi)) / np.max(np.abs(xi))))
    return {'draws_checked': n_check, 'max_rel_dev_ngal': dev_ngal, 'max_rel_dev_xi': dev_xi,
            'tolerance': 1e-10, 'ok': bool(dev_ngal < 1e-10 and dev_xi < 1e-10)}


def run_reference_arm(args):
    """`--impl reference`: the reference's algorithm on all host cores (multiprocessing fan-out,
    one table copy per worker), same workload/metric; each step is a bounded sample."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_worker = 150
    ctx = mp.get_context('fork')
    jobs = [(w * per_worker, (w + 1) * per_worker, 1) for w in range(cores)]
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_worker, [(0, 20, 1)] * cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_worker, jobs)
        seconds = time.perf_counter() - t0
    n_total = per_worker * cores * args.steps
    value = n_total / seconds
    sample = ('{} draws per step ({} per worker x {} workers) of the same table and draw set, '
              'numpy port of the reference loop (the reference itself needs h5py/astropy/'
              'halotools, absent here)'.format(per_worker * cores, per_worker, cores))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * seconds / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(DRAWS_PER_GPU),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, short device-timed runs (the `configs` block)
# ---------------------------------------------------------------------------------------------
# FP64 instructions the occupation kernel issues per quadrature-node evaluation (DFMA + DADD +
# DMUL of the ncu capture in profiles/r01_occupation_kernel_N240.md: 94.0e6 warp instructions for
# 1.2e8 evaluations x 32 lanes); each counted as one FMA = 2 flops against the DFMA peak
FP64_OPS_PER_EVALUATION = 25.1


def run_configs(args, world, rank, local_rank, lib, peak_dmma):
    """BASELINE.json configs[0..4] beside the headline: real bolplanck tables (cfg1), decorated
    multipoles (cfg3), database-style Interpolator + per-draw cosmology (cfg4) and, on any
    number of GPUs, the fixed-size sweep of configs[4].  One dict per configuration."""
    import ctypes
    import torch
    import torch.distributed as dist
    import tabcorr_b200
    from tabcorr_b200 import _lib, synthetic, sweep
    from tabcorr_b200.models import ModelSpec, theta_from_params

    out = []
    n_draws = args.config_draws
    reps = 3
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    golden = os.path.join(ROOT, 'tests', 'golden')

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = []
        for i in range(reps):
            flush.fill_(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    def table_of(tab):
        return tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                                tab['tpcf_shape'], tab['attrs'],
                                                device=local_rank)

    def device_ms(halotab, draws, decorated=False, **kw):
        spec = ModelSpec(decorated=decorated)
        theta = torch.from_numpy(theta_from_params(draws, None, spec)).cuda()
        return timed(lambda: halotab.predict_batch(theta, model=spec, as_numpy=False, **kw))

    def frac(n, r, mode, ms, n_d):
        return executed_flops(n, r, mode) * n_d / (ms * 1e-3) / (peak_dmma * 1e12)

    if rank == 0:
        peak_dfma = ctypes.c_double()
        _lib.check(lib.tc_measure_dfma_peak(local_rank, ctypes.byref(peak_dfma)))
        # ---- cfg1: the README workflow's real tables ------------------------------------------
        draws = synthetic.make_draws(n_draws, seed=1)
        halotab = tabcorr_b200.TabCorr.read(os.path.join(golden, 'bolplanck_wp.hdf5'),
                                            device=local_rank)
        ms = device_ms(halotab, draws)
        out.append({'config': 'cfg1 real bolplanck_wp.hdf5 (auto, N=60, R=19), zheng07',
                    'n_draws': n_draws, 'ms': ms, 'preds_per_s': n_draws / ms * 1e3,
                    'executed_frac': frac(60, 19, 'auto', ms, n_draws)})
        halotab = tabcorr_b200.TabCorr.read(os.path.join(golden, 'bolplanck_ds.hdf5'),
                                            device=local_rank)
        ms = device_ms(halotab, draws)
        evals = n_draws * 60.0 * N_GAUSS / (ms * 1e-3)
        out.append({'config': 'cfg1 real bolplanck_ds.hdf5 (cross, N=60, R=19), zheng07',
                    'n_draws': n_draws, 'ms': ms, 'preds_per_s': n_draws / ms * 1e3,
                    'bound': 'fp64 alu (occupation arithmetic; 2RN contraction flops per draw)',
                    'occupation_evaluations_per_s': evals,
                    'fp64_ops_per_evaluation': FP64_OPS_PER_EVALUATION,
                    'dfma_peak_tflops': peak_dfma.value,
                    'fp64_alu_frac': evals * FP64_OPS_PER_EVALUATION * 2.0 /
                    (peak_dfma.value * 1e12)})
        # ---- cfg3: multipoles, decorated zheng07 ----------------------------------------------
        tab = synthetic.make_table(n_mass=N_MASS, n_sec=N_SEC, n_r=42, kind='multipole',
                                   tpcf_shape=(3, 14))
        draws_dec = synthetic.make_draws(n_draws, seed=1, decorated=True)
        ms = device_ms(table_of(tab), draws_dec, decorated=True, n_gauss_prim=N_GAUSS)
        out.append({'config': 'cfg3 N=240 R=3x14 multipoles, decorated zheng07, G=10',
                    'n_draws': n_draws, 'ms': ms, 'preds_per_s': n_draws / ms * 1e3,
                    'executed_frac': frac(240, 42, 'auto', ms, n_draws)})
        # ---- the other occupation family on the headline table -----------------
        from tabcorr_b200 import models
        tab2 = synthetic.make_table(n_mass=N_MASS, n_sec=N_SEC, n_r=N_R)
        halotab2 = table_of(tab2)
        for name, dec in (('leauthaud11', False), ('hearin15', True)):
            spec = ModelSpec(models.FAMILY_LEAUTHAUD11, dec, True, 0.5, 10.5, 0.0)
            theta = torch.from_numpy(theta_from_params(
                synthetic.make_draws_leauthaud11(n_draws, seed=1, decorated=dec), None,
                spec)).cuda()
            ms = timed(lambda: halotab2.predict_batch(theta, model=spec, as_numpy=False,
                                                      n_gauss_prim=N_GAUSS))
            ms_occ = timed(lambda: halotab2.mean_occupation_batch(theta, n_gauss_prim=N_GAUSS,
                                                                  model=spec))
            out.append({'config': 'cfg2 table (N=240, R=20, G=10), {}: occupation kernel -> '
                                  'contraction'.format(name),
                        'n_draws': n_draws, 'ms': ms, 'preds_per_s': n_draws / ms * 1e3,
                        'occupation_kernel_ms': ms_occ,
                        'executed_frac': frac(2 * N_MASS * N_SEC, N_R, 'auto', ms, n_draws)})
        # ---- cfg4: database-style Interpolator, per-draw cosmology ----------------------------
        axes = {'alpha_s': np.linspace(0.8, 1.2, 4), 'log_eta': np.log10(np.geomspace(1 / 3, 3, 4))}
        interps = []
        for c in range(8):
            tables, param_table = synthetic.make_grid_tables(
                axes, n_mass=30, n_sec=2, n_r=14, kind='wp', seed=100 + c, n_h_scale=1 + 0.1 * c)
            interps.append(tabcorr_b200.Interpolator([table_of(t) for t in tables], param_table))
        table_set = tabcorr_b200.TableSet(interps)
        extra = {k: (float(v.min()), float(v.max())) for k, v in axes.items()}
        draws4 = synthetic.make_draws(n_draws, seed=2, extra=extra)
        index = np.random.default_rng(3).integers(0, 8, n_draws)
        table_set.predict_batch(draws4, index)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            table_set.predict_batch(draws4, index)
        ms_set = (time.perf_counter() - t0) / reps * 1e3
        one = {k: v[index == 0] for k, v in draws4.items()}
        n_one = int((index == 0).sum())
        ms_one = timed(lambda: interps[0].predict_batch(one, as_numpy=False))
        out.append({'config': 'cfg4 database-style wp Interpolator T=16 (N=120, R=14), 8 cosmologies '
                              'by TableSet, per-draw cosmology',
                    'n_draws': n_draws, 'tableset_host_to_host_ms': ms_set,
                    'tableset_preds_per_s_host_to_host': n_draws / ms_set * 1e3,
                    'single_interpolator_draws': n_one, 'single_interpolator_ms': ms_one,
                    'single_interpolator_preds_per_s': n_one / ms_one * 1e3,
                    'executed_frac_single_interpolator': frac(120, 14 * 16, 'auto', ms_one, n_one)})
        del table_set, interps

    # ---- cfg5: fixed-size sweep sharded over the GPUs (strong scaling, gather to rank 0) -----
    tab = synthetic.make_table(n_mass=125, n_sec=2, n_r=20)
    halotab = table_of(tab)
    prior = sweep.UniformPrior(sweep.ZHENG07_PRIOR, seed=5)
    totals = {'n': 0}

    def consume(lo, hi, slab):
        totals['n'] += hi - lo

    def run(n):
        totals['n'] = 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sweep.predict_sweep(halotab, prior, n, chunk=args.sweep_chunk, consume=consume)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return time.perf_counter() - t0

    run(min(args.sweep_draws, world * args.sweep_chunk))
    seconds = run(args.sweep_draws)
    if world > 1:
        t = torch.tensor([seconds], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
    if rank == 0:
        out.append({'config': 'cfg5 sweep N=500 R=20 wp zheng07: {} draws in total (fixed, strong '
                              'scaling) generated on the device, chunks of {}, gather to rank 0'
                              .format(args.sweep_draws, args.sweep_chunk),
                    'n_gpus': world, 'n_draws': args.sweep_draws, 'seconds': seconds,
                    'preds_per_s': args.sweep_draws / seconds,
                    'executed_frac_per_gpu': executed_flops(500, 20) * args.sweep_draws / seconds /
                    (peak_dmma * 1e12) / world})
    return out


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks and throttle reasons while the timed region runs: NVML in a background
    thread (a sample every 5 ms; the main thread releases the GIL while it waits on CUDA events),
    `nvidia-smi -lms` as a fallback when pynvml is missing."""

    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None

    def _nvml_loop(self, nvml, handle):
        names = {'hw_slowdown': nvml.nvmlClocksEventReasonHwSlowdown,
                 'hw_thermal_slowdown': nvml.nvmlClocksEventReasonHwThermalSlowdown,
                 'sw_thermal_slowdown': nvml.nvmlClocksEventReasonSwThermalSlowdown,
                 'sw_power_cap': nvml.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nvml.nvmlDeviceGetClockInfo(handle, nvml.NVML_CLOCK_SM)))
                mask = nvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.005)

    def start(self):
        try:
            import pynvml as nvml
            nvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            index = self.index
            if visible:
                entries = [v.strip() for v in visible.split(',') if v.strip()]
                if self.index < len(entries) and entries[self.index].isdigit():
                    index = int(entries[self.index])
            handle = nvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(nvml.nvmlDeviceGetMaxClockInfo(handle, nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join()
            return {'sm_mhz': float(np.median(self.sm)) if self.sm else None,
                    'sm_max_mhz': self.sm_max, 'samples': len(self.sm),
                    'reasons': sorted(self.reasons), 'source': 'nvml'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.06)
        self.proc.terminate()
        self.proc.wait()
        sm, sm_max, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        with open(self.path) as f:
            for row in f:
                cells = [c.strip() for c in row.split(',')]
                if len(cells) < 7:
                    continue
                try:
                    sm.append(float(cells[0]))
                    sm_max.append(float(cells[1]))
                except ValueError:
                    continue
                for name, cell in zip(names, cells[3:7]):
                    if cell.lower().startswith('active'):
                        reasons.add(name)
        os.unlink(self.path)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(np.max(sm_max)) if sm_max else None,
                'samples': len(sm), 'reasons': sorted(reasons), 'source': 'nvidia-smi'}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import tabcorr_b200
    from tabcorr_b200 import _lib, synthetic
    from tabcorr_b200.models import ModelSpec, theta_from_params
    from tabcorr_b200.distributed import (OverlappedGather, PeerCopyGather, PeerSlab,
                                          gather_slab_chunks, predict_batch_sharded)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device; the hot path has no CPU fallback '
                         '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    n_draws = args.draws
    tab = synthetic.make_table(n_mass=N_MASS, n_sec=N_SEC, n_r=N_R)
    n_rows = len(tab['gal_type'])
    table_bytes = int(np.asarray(tab['tpcf_matrix']).size * 8)
    halotab = tabcorr_b200.TabCorr.from_arrays(tab['gal_type'], tab['tpcf_matrix'],
                                               tab['tpcf_shape'], tab['attrs'], device=local_rank)
    all_draws = synthetic.make_draws(n_draws * world, seed=1)
    draws = {k: np.ascontiguousarray(v[rank * n_draws:(rank + 1) * n_draws])
             for k, v in all_draws.items()}
    group = halotab._ensure_device()
    spec = ModelSpec()
    theta = torch.from_numpy(theta_from_params(draws, None, spec)).to(device)
    ngal = torch.empty((n_draws, 1), dtype=torch.float64, device=device)
    xi = torch.empty((n_draws, N_R, 1), dtype=torch.float64, device=device)
    result = torch.empty((n_draws, 1 + N_R), dtype=torch.float64, device=device)
    full, peer, pipe = None, None, None
    if world > 1 and args.gather == 'peer':
        # rank 0's [world * B, 1 + R] result slab, mapped into every rank (CUDA IPC over NVLink)
        peer = PeerSlab(n_draws * world, 1 + N_R, dst=0, device=local_rank)
        full = peer.tensor
        my_rows = peer.rows(rank * n_draws, (rank + 1) * n_draws)
    elif world > 1 and args.gather in ('overlap', 'overlap-nccl'):
        # double-buffered slabs: the collection of step k runs behind the kernels of step k + 1,
        # moved by the copy engines into rank 0's peer-mapped slab ('overlap') or by NCCL
        try:
            pipe = (PeerCopyGather if args.gather == 'overlap' else OverlappedGather)(
                n_draws, 1 + N_R, device, dst=0)
        except RuntimeError as err:   # no CUDA IPC / peer access on this node: all ranks agree
            if rank == 0:
                print('bench.py: {}; falling back to --gather overlap-nccl'.format(err),
                      file=sys.stderr)
            args.gather = 'overlap-nccl'
            pipe = OverlappedGather(n_draws, 1 + N_R, device, dst=0)
    elif world > 1 and rank == 0:
        full = torch.empty((n_draws * world, 1 + N_R), dtype=torch.float64, device=device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    lib = _lib.load()
    n_chunks = args.gather_chunks if (world > 1 and peer is None and pipe is None) else 1
    draws_per_launch = n_draws - n_draws * (n_chunks - 1) // n_chunks   # the last chunk's

    landed = [None]   # rank 0: the buffer the latest overlapped gather lands in

    def compute_chunk(c0, c1):
        halotab.predict_into_slab(theta[c0:c1], result[c0:c1], N_GAUSS)

    def step():
        if world == 1:
            group.predict_into(spec, N_GAUSS, theta, None, False, ngal, 0, xi, 0)
        elif pipe is not None:
            # the one collective of the path (results to rank 0) is queued asynchronously and runs
            # while the next step's kernels do; every gather completes inside the timed region
            halotab.predict_into_slab(theta, pipe.begin(), N_GAUSS)
            landed[0] = pipe.submit()
        elif peer is not None:
            # the collection of the results on rank 0 is fused into the prediction: the epilogue
            # kernel stores this rank's rows into rank 0's memory over NVLink; a barrier remains
            halotab.predict_into_slab(theta, my_rows, N_GAUSS)
            dist.barrier(device_ids=[local_rank])
        else:
            # the kernels write their rows of the [B, 1 + R] slab directly; the one collective of
            # the path (results to rank 0) is issued per chunk and overlaps the next chunk's kernels
            gather_slab_chunks(compute_chunk, result, full, n_chunks=n_chunks, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed throughput -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    _lib.check(lib.tc_profile_enable(1))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kernel_ms, finalize_ms = [], []

    def read_profile():
        a, b = ctypes.c_float(), ctypes.c_float()
        _lib.check(lib.tc_profile_read(ctypes.byref(a), ctypes.byref(b)))
        kernel_ms.append(a.value)
        finalize_ms.append(b.value)

    barrier()
    if pipe is None:
        for i in range(args.steps):
            flush.fill_(i & 0xFF)  # evict the table and the draws from L2 between timed steps
            starts[i].record()
            step()
            stops[i].record()
            stops[i].synchronize()
            read_profile()
        total_ms = float(sum(s.elapsed_time(e) for s, e in zip(starts, stops)))
    else:
        # overlapped gather: ONE timed region from the first step's start to the completion of the
        # last gather, L2 flushes and host launch gaps included -- a gather may only hide behind
        # work that is itself inside the region
        starts[0].record()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            step()
        pipe.finish()            # the last step's gather is not hidden behind anything
        stops[-1].record()
        stops[-1].synchronize()
        total_ms = float(starts[0].elapsed_time(stops[-1]))
        full = landed[0]
        for i in range(3):       # kernel times for the roofline block, outside the timed region
            flush.fill_(i & 0xFF)
            step()
            torch.cuda.synchronize()
            read_profile()
        pipe.finish()
        torch.cuda.synchronize()
        full = landed[0]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    _lib.check(lib.tc_profile_enable(0))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = n_draws * world * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API (host numpy in, host numpy out) --------------------
    def e2e_step():
        if world == 1:
            return halotab.predict_batch(draws, n_gauss_prim=N_GAUSS)
        return predict_batch_sharded(halotab, all_draws, n_gauss_prim=N_GAUSS, dst=0,
                                     gather='host')

    for _ in range(4):  # also warms torch's pinned-memory cache (results of two calls stay alive)
        host_result = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_result = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n_draws * world * args.steps / e2e_s
    h2d = n_draws * 7 * 8
    d2h = n_draws * (1 + N_R) * 8

    # numerics sanity inside the bench: the e2e results equal the device-timed ones
    same = True
    if rank == 0:
        ngal_host, xi_host = host_result
        if world == 1:
            same = (np.array_equal(ngal_host[:n_draws], ngal[:, 0].cpu().numpy()) and
                    np.array_equal(xi_host[:n_draws], xi[:, :, 0].cpu().numpy()))
        else:   # the gathered device result of all ranks against the host-to-host result
            gathered = full.cpu().numpy()
            same = (np.array_equal(ngal_host, gathered[:, 0]) and
                    np.array_equal(xi_host, gathered[:, 1:]))

    peak = ctypes.c_double()
    _lib.check(lib.tc_measure_dmma_peak(local_rank, ctypes.byref(peak)))
    # free the headline workload's buffers before the other configurations allocate theirs
    del flush
    configs = None
    if not args.no_configs:
        configs = run_configs(args, world, rank, local_rank, lib, peak.value)

    if rank == 0:
        k_ms = float(np.mean(kernel_ms))
        executed = executed_flops(n_rows, N_R) * draws_per_launch
        dense = algorithmic_flops(n_rows, N_R) * draws_per_launch
        achieved = executed / (k_ms * 1e-3) * 1e-12
        roofline = {
            'bound': 'tensor', 'kernel': 'predict_kernel<7, auto> (fused occupation + DMMA quadratic '
                                         'form; 2 W tiles of 56 draws at N=240)',
            'achieved': achieved, 'peak': peak.value, 'unit': 'TFLOP/s',
            'frac': achieved / peak.value,
            'flops_per_prediction': executed_flops(n_rows, N_R),
            'note': 'achieved = tensor flops the kernel executes (lower triangle of the symmetric '
                    'table, 2*R*64*T8(T8+1)/2 with T8 = n_pad/8: what ncu counts as DMMA) / kernel '
                    'time; the symmetric-minimal count R*N*(N+1) is {:.0f} per prediction'.format(
                        N_R * n_rows * (n_rows + 1.0)),
            'peak_source': 'FP64 DMMA (mma.sync m8n8k4 f64) peak measured live by '
                           'tc_measure_dmma_peak; MEASURED_PEAKS.json has no FP64 figure',
            'dense_equivalent_tflops': dense / (k_ms * 1e-3) * 1e-12,
            'dense_equivalent_note': 'SURVEY 8(d) dense count 2RN^2+2RN per prediction over the '
                                     'same time; exceeds the peak because only half of the '
                                     'symmetric product is executed -- not a roofline fraction',
            'traffic': NCU_DRAM_BYTES_PER_LAUNCH,
            'traffic_unit': 'bytes of DRAM read + write per predict_kernel launch (its reads are '
                            'the draws + the table; the results, 16.8 MB of the algorithmic bytes, '
                            'are written by finalize_kernel)',
            'traffic_source': 'constant copied from the ncu --set full capture summarised in ' +
                              NCU_TRAFFIC_SOURCE + ' (not measured by this run)',
            'algorithmic_bytes_per_launch': ALGORITHMIC_BYTES_PER_DRAW * draws_per_launch + table_bytes,
            'draws_per_launch': draws_per_launch,
            'kernel_ms': k_ms, 'finalize_ms': float(np.mean(finalize_ms)),
            'kernel_share_of_step': float(np.sum(kernel_ms) / total_ms) if world == 1 else None,
        }
        if draws_per_launch != DRAWS_PER_GPU:
            roofline['traffic'] = None
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': dict(workload_config(n_draws), **(
                {'collection': ("results stored by the epilogue kernel straight into rank 0's "
                                'memory (CUDA IPC peer slab over NVLink) + one barrier per step'
                                if peer is not None else
                                ('copy-engine transfer of every rank\'s rows into rank 0\'s peer-'
                                 'mapped slab (CUDA IPC, NVLink)' if args.gather == 'overlap' else
                                 'asynchronous NCCL gather to rank 0') +
                                " per step, overlapping the next step's kernels (double-buffered "
                                'slabs); one timed region over all steps incl. the L2 flushes, '
                                'closed after the last transfer and a barrier'
                                if pipe is not None else
                                'NCCL gather to rank 0 in {} chunk(s) per step'.format(n_chunks))}
                if world > 1 else {})),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'timing': 'wall clock around predict_batch calls'},
            # predict_kernel + finalize_kernel per chunk of a timed step
            'gpu_launches': 2 * n_chunks * args.steps,
            'roofline': roofline, 'clocks': clocks,
            'results_consistent': bool(same),
        }
        if configs is not None:
            line['configs'] = configs
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline_single(args.cpu_sample)
            if n_draws == DRAWS_PER_GPU:
                line['parity_vs_cpu'] = parity_vs_cpu(ngal[:, 0].cpu().numpy(),
                                                      xi[:, :, 0].cpu().numpy())
        emit(line)
    if peer is not None:
        del full
        peer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def emit(line):
    """Print the ONE JSON line on the real stdout (see main: fd 1 is pointed at stderr while the
    benchmark runs, so that nothing a library prints -- e.g. NCCL's version banner -- lands beside
    the JSON line)."""
    text = json.dumps(line) + '\n'
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, text.encode())
    else:
        sys.stdout.write(text)
        sys.stdout.flush()


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=20)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    parser.add_argument('--draws', type=int, default=DRAWS_PER_GPU, help='draws per GPU')
    parser.add_argument('--cpu-sample', type=int, default=8000,
                        help='draws of the workload timed for cpu_baseline')
    parser.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    parser.add_argument('--gather', default='overlap',
                        choices=['overlap', 'overlap-nccl', 'peer', 'nccl'],
                        help="N > 1: 'overlap' = copy-engine transfer of step k's rows into rank "
                             "0's peer-mapped slab behind the kernels of step k + 1, "
                             "'overlap-nccl' = the same with an asynchronous NCCL gather, "
                             "'peer' = results stored straight into rank "
                             "0's memory by the epilogue kernel (CUDA IPC / NVLink), 'nccl' = "
                             "blocking (optionally chunked) NCCL gather inside every step")
    parser.add_argument('--gather-chunks', type=int, default=1,
                        help='N > 1: chunks per step whose gather overlaps the next chunk\'s kernels')
    parser.add_argument('--no-configs', action='store_true',
                        help='skip the `configs` block (the other BASELINE.json configurations)')
    parser.add_argument('--config-draws', type=int, default=100000,
                        help='draws of the cfg1/cfg3/cfg4 runs of the `configs` block')
    parser.add_argument('--sweep-draws', type=int, default=1 << 24,
                        help='total draws of the cfg5 sweep (fixed for any number of GPUs)')
    parser.add_argument('--sweep-chunk', type=int, default=1 << 20)
    args = parser.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
