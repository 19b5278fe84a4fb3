"""Halo-bin table of the tabulation side: ``n_h`` histogram and ``prim_haloprop_dist_index``.

The cheap half of SURVEY.md section 8(f) #4 -- what ``TabCorr.tabulate`` computes before any pair
counting (``tabcorr/tabcorr.py:194-234``): the number of haloes per (primary-property bin,
secondary-percentile bin) cell and, per cell, the power-law index ``n`` whose distribution
``p(x) = x^n`` over the bin reproduces the mean primary property of the cell's haloes
(``distribution_index``, ``tabcorr/tabcorr.py:740-767``).  These are exactly the ``n_h`` and the
quadrature weights ``M^(n+1)`` the occupation kernel consumes (``tabcorr/tabcorr.py:568-574``).

The reduction over the halo catalogue (24 bytes per halo) runs on the GPU (``tc_halo_bins``,
``csrc/halo_bins.cuh``); the ~100 cells are finished on the host.  Pair counting
(``tabcorr/tabcorr.py:236-372``) stays with the reference.
"""

import ctypes

import numpy as np

from . import _lib
from .table import Table


def notaknot_cubic(x, y, xq):
    """Interpolating cubic spline with not-a-knot end conditions through ``(x, y)`` evaluated at
    ``xq`` -- what ``scipy.interpolate.interp1d(kind='cubic')`` builds (``make_interp_spline`` with
    ``k=3``, default boundary conditions).  ``x`` strictly increasing, at least 4 knots."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xq = np.asarray(xq, dtype=np.float64)
    n = len(x)
    if n < 4:
        raise ValueError('a not-a-knot cubic needs at least 4 knots')
    h = np.diff(x)
    slope = np.diff(y) / h
    system = np.zeros((n, n))
    rhs = np.zeros(n)
    for i in range(1, n - 1):   # continuity of the first derivative, in second derivatives m
        system[i, i - 1:i + 2] = h[i - 1], 2.0 * (h[i - 1] + h[i]), h[i]
        rhs[i] = 6.0 * (slope[i] - slope[i - 1])
    # not-a-knot: the third derivative is continuous across the second and second-to-last knot
    system[0, :3] = h[1], -(h[0] + h[1]), h[0]
    system[-1, -3:] = h[-1], -(h[-2] + h[-1]), h[-2]
    m = np.linalg.solve(system, rhs)
    seg = np.clip(np.searchsorted(x, xq, side='right') - 1, 0, n - 2)
    t = xq - x[seg]
    hs = h[seg]
    a = (m[seg + 1] - m[seg]) / (6.0 * hs)
    b = 0.5 * m[seg]
    c = slope[seg] - hs * (2.0 * m[seg] + m[seg + 1]) / 6.0
    return ((a * t + b) * t + c) * t + y[seg]


def distribution_index(x_min, x_max, x_mean):
    """Effective power-law index ``n`` with ``<x> = x_mean`` for ``p(x) = x^n`` on
    ``[x_min, x_max]``, clipped to ``[-10, 10]`` (``tabcorr/tabcorr.py:740-767``: a cubic spline
    through 100 tabulated indices).  Vectorised over cells."""
    x_min = np.asarray(x_min, dtype=np.float64)
    x_max = np.atleast_1d(np.asarray(x_max, dtype=np.float64) / x_min)
    x_mean = np.atleast_1d(np.asarray(x_mean, dtype=np.float64) / x_min)
    n_interp = np.linspace(-10, +10, 100)
    out = np.empty(x_max.shape)
    for i in range(x_max.size):
        x_interp = ((n_interp + 1) / (n_interp + 2) * (x_max.flat[i]**(n_interp + 2) - 1) /
                    (x_max.flat[i]**(n_interp + 1) - 1))
        if x_mean.flat[i] < x_interp[0]:
            out.flat[i] = -10.0
        elif x_mean.flat[i] > x_interp[-1]:
            out.flat[i] = +10.0
        else:
            out.flat[i] = notaknot_cubic(x_interp, n_interp, x_mean.flat[i:i + 1])[0]
    return out


def halo_bin_counts(prim_haloprop, sec_haloprop_percentile, log_prim_haloprop_bins,
                    sec_haloprop_percentile_bins, device=None):
    """GPU reduction of a halo catalogue into cells: ``(n_h, n_members, mean_prim)``, each of
    length ``n_sec * n_prim`` in the reference's row order (secondary bin outer, primary bin inner:
    ``n_h.ravel(order='F')``, ``tabcorr/tabcorr.py:199``).  ``n_h`` follows ``np.histogram2d``;
    ``n_members`` / ``mean_prim`` follow ``sort_into_bins`` (``np.digitize``), NaN in empty cells.
    Inputs may be numpy arrays or CUDA tensors."""
    import torch
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError('tabcorr_b200 needs a CUDA device (there is no CPU fallback)')
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)

    def to_dev(a):
        if isinstance(a, torch.Tensor):
            return a.to(device=device, dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)

    if isinstance(prim_haloprop, torch.Tensor):
        prim = to_dev(prim_haloprop)
        log_prim = torch.log10(prim)
    else:
        # np.log10 on the host, like the reference: the bin a halo on an edge falls into must
        # not depend on a last-bit difference between two log10 implementations
        prim_host = np.ascontiguousarray(prim_haloprop, dtype=np.float64)
        prim = to_dev(prim_host)
        log_prim = to_dev(np.log10(prim_host))
    sec = to_dev(sec_haloprop_percentile)
    if not (prim.ndim == 1 and sec.shape == prim.shape):
        raise ValueError('prim_haloprop and sec_haloprop_percentile must be 1-d of equal length')
    pe = np.ascontiguousarray(log_prim_haloprop_bins, dtype=np.float64)
    se = np.ascontiguousarray(sec_haloprop_percentile_bins, dtype=np.float64)
    n_prim, n_sec = len(pe) - 1, len(se) - 1
    n_h = np.empty(n_prim * n_sec)
    members = np.empty(n_prim * n_sec)
    mean = np.empty(n_prim * n_sec)
    as_p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    stream = torch.cuda.current_stream(device).cuda_stream
    _lib.check(lib.tc_halo_bins(device.index or 0, log_prim.data_ptr(), sec.data_ptr(),
                                prim.data_ptr(), prim.shape[0], as_p(pe), n_prim, as_p(se), n_sec,
                                as_p(n_h), as_p(members), as_p(mean), stream))
    return n_h, members, mean


def halo_bin_table(prim_haloprop, sec_haloprop_percentile, log_prim_haloprop_bins,
                   sec_haloprop_percentile_bins, device=None):
    """The ``gal_type`` table ``TabCorr.tabulate`` builds from the halo catalogue
    (``tabcorr/tabcorr.py:192-234``) before it counts pairs: per cell ``n_h`` (raw counts; the
    reference divides by the box volume at the very end, ``:354``), the bin edges, the bin centre
    ``prim_haloprop``, ``sec_haloprop_percentile`` and ``prim_haloprop_dist_index``; the block is
    stacked twice, for centrals and satellites."""
    pe = np.asarray(log_prim_haloprop_bins, dtype=np.float64)
    se = np.asarray(sec_haloprop_percentile_bins, dtype=np.float64)
    n_h, members, mean = halo_bin_counts(prim_haloprop, sec_haloprop_percentile, pe, se, device)
    n_prim, n_sec = len(pe) - 1, len(se) - 1
    # np.meshgrid(log_prim_bins, sec_bins): row index = secondary bin, column = primary bin
    log_min = np.tile(pe[:-1], n_sec)
    log_max = np.tile(pe[1:], n_sec)
    pct_min = np.repeat(se[:-1], n_prim)
    pct_max = np.repeat(se[1:], n_prim)
    dist = np.zeros(n_prim * n_sec)
    filled = members > 0
    if filled.any():
        dist[filled] = distribution_index(10**log_min[filled], 10**log_max[filled], mean[filled])
    columns = {
        'n_h': n_h, 'log_prim_haloprop_min': log_min, 'log_prim_haloprop_max': log_max,
        'sec_haloprop_percentile_min': pct_min, 'sec_haloprop_percentile_max': pct_max,
        'prim_haloprop': 10**(0.5 * (log_min + log_max)),
        'sec_haloprop_percentile': 0.5 * (pct_min + pct_max),
        'prim_haloprop_dist_index': dist}
    table = Table({k: np.concatenate([v, v]) for k, v in columns.items()})
    table['gal_type'] = np.concatenate([np.repeat('centrals', n_prim * n_sec),
                                        np.repeat('satellites', n_prim * n_sec)])
    return table
