"""Drop-in ``TabCorr`` for the prediction path, backed by the sm_100a kernels.

Mirrors the reference class for everything on the hot path (``tabcorr/tabcorr.py:374-416`` read,
``:465-578`` mean_occupation, ``:580-683`` predict): same signatures, same return types (numpy
scalars / arrays / dicts keyed by ``str``), same ``ValueError`` conditions.  The table itself is
device resident from ``read`` on, and ``predict_batch`` evaluates B parameter sets in one fused
launch.  Tabulation (``TabCorr.tabulate``) stays on the reference.
"""

import ctypes
import os
import threading

import numpy as np

from . import _lib
from . import h5mini
from .models import (ModelSpec, THETA_KEYS, resolve_model, spec_from_params, theta_columns,
                     theta_from_params)
from .table import Table

try:  # h5py is optional: used when present, otherwise the built-in reader
    import h5py as _h5py
except ImportError:  # pragma: no cover - h5py is absent in the benchmark image
    _h5py = None

_GROUP_TYPES = (h5mini.Group,) + ((_h5py.Group,) if _h5py is not None else ())


def _torch():
    import torch
    return torch


def _require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError(
            'tabcorr_b200 needs a CUDA device (B200, sm_100a): predictions are computed by the '
            'CUDA extension only and there is no CPU fallback.')
    if device is None:
        return torch.cuda.current_device()
    return torch.device(device).index if not isinstance(device, int) else device


def leggauss01(n_gauss_prim):
    """Gauss-Legendre nodes mapped to [0, 1] and weights (``tabcorr/tabcorr.py:543-546``)."""
    x, w = np.polynomial.legendre.leggauss(int(n_gauss_prim))
    return np.ascontiguousarray((x + 1) / 2), np.ascontiguousarray(w)


def chunk_schedule(n_draws, chunk='auto', edge=None, largest=1 << 20):
    """Chunk boundaries ``[(lo, hi), ...]`` of the pipelined host-to-host batch path.

    ``chunk``: int (fixed size), sequence of sizes (the last repeats) or 'auto': one chunk up to
    30 000 draws; otherwise a head and a tail of ``edge`` draws (default: a tenth of the batch,
    at most 65 536) around equal chunks of at most ``largest`` draws."""
    n_draws = int(n_draws)
    if edge is None:
        edge = min(max(n_draws // 10, 1), 65536)
    if isinstance(chunk, str):
        if chunk != 'auto':
            raise ValueError("pipeline_chunk must be 'auto', an int or a sequence of ints")
        if n_draws <= 30000:
            sizes = [max(n_draws, 1)]
        else:
            middle = n_draws - 2 * edge
            pieces = -(-middle // largest)
            sizes = [edge] + [middle // pieces + (1 if i < middle % pieces else 0)
                              for i in range(pieces)] + [edge]
    elif isinstance(chunk, (list, tuple)):
        sizes = [int(c) for c in chunk]
        if not sizes or min(sizes) <= 0:
            raise ValueError('chunk sizes must be positive')
    else:
        sizes = [int(chunk)] if int(chunk) > 0 else [max(n_draws, 1)]
    bounds, lo, i = [], 0, 0
    while lo < n_draws:
        hi = min(n_draws, lo + sizes[min(i, len(sizes) - 1)])
        bounds.append((lo, hi))
        lo, i = hi, i + 1
    return bounds


# serialises the first upload of a table / interpolator shared by several host threads
_UPLOAD_LOCK = threading.RLock()


def batch_size(columns):
    """Number of draws of a list of parameter columns (arrays ``[B]`` and broadcast scalars): the
    common array length -- 0 for empty arrays -- or 1 when every column is a scalar."""
    lengths = [np.shape(c)[0] for c in columns if np.ndim(c) > 0]
    return max(lengths) if lengths else 1


MAX_STREAM_WORKSPACES = 8


class DeviceTableGroup:
    """One gal_type table with one or more correlation matrices on the device (``tc_table``)."""

    def __init__(self, gal_type, matrices, mode, n_r, device=None):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.mode = mode
        self.n_rows = len(gal_type)
        self.n_r = int(n_r)
        self.n_tables = len(matrices)
        names = np.asarray(gal_type['gal_type'].data)
        is_sat = np.ascontiguousarray((names != 'centrals').astype(np.int32))  # tabcorr.py:555
        self.is_sat = is_sat

        def column(name):
            return np.ascontiguousarray(gal_type[name].data, dtype=np.float64)

        n_h, log_min, log_max = (column('n_h'), column('log_prim_haloprop_min'),
                                 column('log_prim_haloprop_max'))
        pct = column('sec_haloprop_percentile')
        dist = (column('prim_haloprop_dist_index')
                if 'prim_haloprop_dist_index' in gal_type.colnames else None)
        mats = [np.ascontiguousarray(m, dtype=np.float64) for m in matrices]
        n_cols = self.n_rows * (self.n_rows + 1) // 2 if mode == 'auto' else self.n_rows
        for m in mats:
            if m.shape != (self.n_r, n_cols):
                raise ValueError('tpcf_matrix has shape {}, expected {}'.format(
                    m.shape, (self.n_r, n_cols)))
        c_double_p = ctypes.POINTER(ctypes.c_double)
        mat_ptrs = (c_double_p * len(mats))(*[_lib.as_double_p(m) for m in mats])
        handle = ctypes.c_void_p()
        _lib.check(self.lib.tc_table_create(
            ctypes.byref(handle), _lib.TC_MODE_AUTO if mode == 'auto' else _lib.TC_MODE_CROSS,
            self.n_rows, self.n_r, self.n_tables, _lib.as_double_p(n_h),
            _lib.as_double_p(log_min), _lib.as_double_p(log_max), _lib.as_double_p(pct),
            _lib.as_double_p(dist) if dist is not None else None, _lib.as_int32_p(is_sat),
            mat_ptrs, self.device))
        self.handle = handle
        self._planned = set()
        self._workspace = None
        self._copy_streams = None
        self._single = None
        self._small = None
        self._small_occ = None
        self._lock = threading.Lock()

    def __del__(self):
        handle = getattr(self, 'handle', None)
        if handle:
            try:
                self.lib.tc_table_destroy(handle)
            except Exception:  # interpreter shutdown
                pass
            self.handle = None

    def copy_streams(self):
        """Two side streams (host-to-device, device-to-host) for the pipelined batch path."""
        if self._copy_streams is None:
            torch = _torch()
            self._copy_streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        return self._copy_streams

    def plan(self, n_gauss):
        if n_gauss not in self._planned:
            x01, w = leggauss01(n_gauss)
            _lib.check(self.lib.tc_table_plan(self.handle, int(n_gauss), _lib.as_double_p(x01),
                                              _lib.as_double_p(w)))
            self._planned.add(n_gauss)

    def n_comp(self, separate):
        if not separate:
            return 1
        return 3 if self.mode == 'auto' else 2

    def _workspace_for(self, n_draws, separate, precision=_lib.TC_PRECISION_FP64):
        """Scratch buffer of the current stream (kernels of different streams may overlap, so
        every stream that launches on this table gets its own)."""
        torch = _torch()
        need = int(self.lib.tc_predict_workspace_bytes_for(self.handle, int(n_draws),
                                                           int(separate), int(precision)))
        if need == 0 and n_draws > 0:
            raise _lib.TabCorrB200Error(
                'table too large for the CUDA kernel: the weights of 8 draws x {} halo bins do not '
                'fit the 227 KB shared-memory tile (at most ~3500 bins are supported)'.format(
                    self.n_rows))
        if self._workspace is None:
            self._workspace = {}
        stream = torch.cuda.current_stream(self.device).cuda_stream
        current = self._workspace.pop(stream, None)
        if current is None or current.numel() < need:
            current = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._workspace[stream] = current          # most recently used last
        while len(self._workspace) > MAX_STREAM_WORKSPACES:
            # short-lived streams must not pin a buffer each for the table's lifetime; a buffer of
            # an evicted stream stays alive until its kernels finish (caching allocator semantics)
            self._workspace.pop(next(iter(self._workspace)))
        return current

    @staticmethod
    def _model_struct(spec):
        model = _lib.tc_model(spec.family, int(spec.decorated), int(spec.modulate_with_cenocc), 0,
                              spec.split, spec.threshold, spec.redshift)
        if len(spec.scatter_abscissa) > 1:   # leauthaud11: mass-dependent stellar-mass scatter
            model.n_scatter = len(spec.scatter_abscissa)
            for k, value in enumerate(spec.scatter_abscissa):
                model.scatter_abscissa[k] = value
        if not spec.mass_dependent and not any(len(a) for a in spec.split_abscissa):
            return model
        for t in range(2):   # mass-dependent decoration (centrals, satellites)
            absc = spec.strength_abscissa[t]
            model.n_strength[t] = len(absc) if len(absc) > 1 else 0
            for k, value in enumerate(absc):
                model.strength_abscissa[t][k] = value
            model.n_split[t] = len(spec.split_abscissa[t])
            for k, (a, o) in enumerate(zip(spec.split_abscissa[t], spec.split_ordinates[t])):
                model.split_abscissa[t][k] = a
                model.split_ordinates[t][k] = o
        return model

    def occupation(self, spec, n_gauss, theta, theta_columns=False):
        """theta: CUDA tensor ``[B, n_theta]`` (or ``[n_theta, B]`` with ``theta_columns``) ->
        CUDA tensor ``[B, n_rows]`` (mean_occupation)."""
        torch = _torch()
        self.plan(n_gauss)
        n_draws = theta.shape[1] if theta_columns else theta.shape[0]
        n_theta = theta.shape[0] if theta_columns else theta.shape[1]
        if n_theta != spec.n_theta:
            raise ValueError('the parameter array has {} columns, the model family needs {} '
                             '({})'.format(n_theta, spec.n_theta, ', '.join(spec.theta_keys)))
        occ = torch.zeros((n_draws, self.n_rows), dtype=torch.float64, device=self.device)
        model = self._model_struct(spec)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.tc_occupation_batch(
            self.handle, ctypes.byref(model), int(n_gauss), theta.data_ptr(),
            theta.stride(0) if theta_columns else 0, n_draws, occ.data_ptr(), stream))
        return occ

    def predict_one(self, spec, n_gauss, values, separate):
        """One parameter set (``values``: the 7 kernel parameters) -> host arrays
        ``ngal [T, 1|2]``, ``xi [T, R, C]`` (copies).  Latency path, see ``_SingleDrawBuffers``."""
        torch = _torch()
        self.plan(n_gauss)
        if not spec.latency_paths:
            # families outside the fused kernel: occupation kernel -> contraction on its output,
            # with the same persistent buffers (parameters and results in pinned host memory the
            # kernels access directly, the occupations in a device buffer): two ctypes calls and
            # one stream synchronisation, no allocations, no copy launches
            values = np.asarray(values, dtype=np.float64)
            if values.shape != (spec.n_theta,):
                raise ValueError('the parameter array has {} entries, the model family needs {} '
                                 '({})'.format(values.size, spec.n_theta,
                                               ', '.join(spec.theta_keys)))
            with self._lock:
                if self._single is None:
                    self._single = _SingleDrawBuffers(self)
                buf = self._single
                buf.theta_any_np[:spec.n_theta] = values
                n_ng, n_comp = (2 if separate else 1), self.n_comp(separate)
                model = self._model_struct(spec)
                stream = torch.cuda.current_stream(self.device)
                _lib.check(self.lib.tc_occupation_batch(
                    self.handle, ctypes.byref(model), int(n_gauss), buf.theta_any.data_ptr(), 0, 1,
                    buf.occ.data_ptr(), stream.cuda_stream))
                _lib.check(self.lib.tc_predict_batch(
                    self.handle, ctypes.byref(model), int(n_gauss), None, 0, buf.occ.data_ptr(), 1,
                    int(separate), _lib.TC_PRECISION_FP64, buf.ngal.data_ptr(),
                    self.n_tables * n_ng, buf.xi.data_ptr(), self.n_tables * self.n_r * n_comp,
                    buf.workspace.data_ptr(), buf.workspace.numel(), stream.cuda_stream))
                stream.synchronize()
                ngal = buf.ngal_np[:self.n_tables * n_ng].reshape(self.n_tables, n_ng).copy()
                xi = buf.xi_np[:self.n_tables * self.n_r * n_comp].reshape(
                    self.n_tables, self.n_r, n_comp).copy()
            return ngal, xi
        with self._lock:
            if self._single is None:
                self._single = _SingleDrawBuffers(self)
            buf = self._single
            buf.theta_np[:] = values
            n_ng, n_comp = (2 if separate else 1), self.n_comp(separate)
            model = self._model_struct(spec)
            stream = torch.cuda.current_stream(self.device)
            _lib.check(self.lib.tc_predict_one(
                self.handle, ctypes.byref(model), int(n_gauss), buf.theta.data_ptr(),
                int(separate), _lib.TC_PRECISION_FP64, buf.ngal.data_ptr(), self.n_tables * n_ng,
                buf.xi.data_ptr(), self.n_tables * self.n_r * n_comp, buf.workspace.data_ptr(),
                buf.workspace.numel(), stream.cuda_stream))
            stream.synchronize()
            ngal = buf.ngal_np[:self.n_tables * n_ng].reshape(self.n_tables, n_ng).copy()
            xi = buf.xi_np[:self.n_tables * self.n_r * n_comp].reshape(
                self.n_tables, self.n_r, n_comp).copy()
        return ngal, xi

    def predict_small(self, spec, n_gauss, columns, n_draws, separate, precision):
        """Up to ``SMALL_BATCH`` parameter sets (``columns``: one array or scalar per kernel
        parameter) -> host arrays ``ngal [B, T * (1|2)]``, ``xi [B, T * R * C]`` (copies).  The
        ensemble-sampler regime: like :meth:`predict_one` the parameters and the results live in
        persistent pinned host memory that the kernels read and write directly, so a call is one
        ctypes call and one stream synchronisation -- no allocations, no copy launches."""
        torch = _torch()
        self.plan(n_gauss)
        with self._lock:
            if self._small is None:
                self._small = _SmallBatchBuffers(self, SMALL_BATCH)
            buf = self._small
            for j, column in enumerate(columns):
                buf.theta_np[j, :n_draws] = column
            n_ng, n_comp = (2 if separate else 1), self.n_comp(separate)
            ngal_cols, xi_cols = self.n_tables * n_ng, self.n_tables * self.n_r * n_comp
            model = self._model_struct(spec)
            stream = torch.cuda.current_stream(self.device)
            if n_draws == 1:   # parameters in the launch arguments
                buf.one_np[:] = buf.theta_np[:, 0]
                _lib.check(self.lib.tc_predict_one(
                    self.handle, ctypes.byref(model), int(n_gauss), buf.one.data_ptr(),
                    int(separate), int(precision), buf.ngal.data_ptr(), ngal_cols,
                    buf.xi.data_ptr(), xi_cols, buf.workspace.data_ptr(), buf.workspace.numel(),
                    stream.cuda_stream))
            else:
                _lib.check(self.lib.tc_predict_batch(
                    self.handle, ctypes.byref(model), int(n_gauss), buf.theta.data_ptr(),
                    buf.capacity, None, int(n_draws), int(separate), int(precision),
                    buf.ngal.data_ptr(), ngal_cols, buf.xi.data_ptr(), xi_cols,
                    buf.workspace.data_ptr(), buf.workspace.numel(), stream.cuda_stream))
            stream.synchronize()
            ngal = buf.ngal_np[:n_draws * ngal_cols].reshape(n_draws, ngal_cols).copy()
            xi = buf.xi_np[:n_draws * xi_cols].reshape(n_draws, xi_cols).copy()
        return ngal, xi

    def predict_small_occupation(self, spec, n_gauss, columns, n_draws, separate):
        """:meth:`predict_small` for the families that run as occupation kernel -> contraction
        (leauthaud11 / hearin15, mass-dependent decoration): the same persistent pinned buffers,
        the occupations of the batch in a persistent device buffer between the two library calls."""
        torch = _torch()
        self.plan(n_gauss)
        with self._lock:
            if self._small_occ is None:
                self._small_occ = _SmallBatchBuffers(self, SMALL_BATCH, occupation=True)
            buf = self._small_occ
            for j, column in enumerate(columns):
                buf.theta_np[j, :n_draws] = column
            n_ng, n_comp = (2 if separate else 1), self.n_comp(separate)
            ngal_cols, xi_cols = self.n_tables * n_ng, self.n_tables * self.n_r * n_comp
            model = self._model_struct(spec)
            stream = torch.cuda.current_stream(self.device)
            _lib.check(self.lib.tc_occupation_batch(
                self.handle, ctypes.byref(model), int(n_gauss), buf.theta.data_ptr(), buf.capacity,
                int(n_draws), buf.occ.data_ptr(), stream.cuda_stream))
            _lib.check(self.lib.tc_predict_batch(
                self.handle, ctypes.byref(model), int(n_gauss), None, 0, buf.occ.data_ptr(),
                int(n_draws), int(separate), _lib.TC_PRECISION_FP64, buf.ngal.data_ptr(), ngal_cols,
                buf.xi.data_ptr(), xi_cols, buf.workspace.data_ptr(), buf.workspace.numel(),
                stream.cuda_stream))
            stream.synchronize()
            ngal = buf.ngal_np[:n_draws * ngal_cols].reshape(n_draws, ngal_cols).copy()
            xi = buf.xi_np[:n_draws * xi_cols].reshape(n_draws, xi_cols).copy()
        return ngal, xi

    def predict_into_raw(self, spec, n_gauss, theta, ngal_ptr, ngal_stride, xi_ptr, xi_stride,
                         precision=_lib.TC_PRECISION_FP64):
        """Total prediction of the draws ``theta`` (CUDA ``[B, n_theta]``) with the outputs given
        as raw device pointers and strides in doubles (``tc_predict_batch`` semantics) -- rows of a
        result slab, possibly in the memory of another GPU of the node."""
        torch = _torch()
        if spec.family != 0 or spec.mass_dependent:
            occ = self.occupation(spec, n_gauss, theta)
            theta = None
        else:
            occ = None
            self.plan(n_gauss)
        n_draws = theta.shape[0] if theta is not None else occ.shape[0]
        with self._lock:
            workspace = self._workspace_for(n_draws, False, precision)
            model = self._model_struct(spec)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self.lib.tc_predict_batch(
                self.handle, ctypes.byref(model), int(n_gauss),
                theta.data_ptr() if theta is not None else None, 0,
                occ.data_ptr() if occ is not None else None, n_draws, 0, int(precision),
                int(ngal_ptr), int(ngal_stride), int(xi_ptr), int(xi_stride),
                workspace.data_ptr(), workspace.numel(), stream))

    def predict_into(self, spec, n_gauss, theta, occ, separate, ngal, ngal_offset, xi, xi_offset,
                     theta_columns=False, precision=_lib.TC_PRECISION_FP64):
        """Fused launch writing this group's tables into the ``[B, T_total, ...]`` buffers ``ngal``
        and ``xi`` starting at table offset ``*_offset`` (in doubles within a draw).  ``theta`` is
        ``[B, n_theta]``, or ``[n_theta, B]`` (one contiguous column per parameter) with
        ``theta_columns``.  Families the fused kernel does not implement (leauthaud11) run the
        occupation kernel first and contract its output."""
        torch = _torch()
        if theta is not None and spec is not None and (spec.family != 0 or spec.mass_dependent):
            occ = self.occupation(spec, n_gauss, theta, theta_columns=theta_columns)
            theta, theta_columns = None, False
        n_draws = (theta.shape[1] if theta_columns else theta.shape[0]) if theta is not None \
            else occ.shape[0]
        theta_ld = theta.stride(0) if (theta is not None and theta_columns) else 0
        if theta is not None:
            self.plan(n_gauss)
        with self._lock:
            workspace = self._workspace_for(n_draws, separate, precision)
            model = self._model_struct(spec if spec is not None else ModelSpec())
            stream = torch.cuda.current_stream(self.device).cuda_stream
            ngal_flat = ngal.view(n_draws, -1)
            xi_flat = xi.view(n_draws, -1)
            _lib.check(self.lib.tc_predict_batch(
                self.handle, ctypes.byref(model), int(n_gauss),
                theta.data_ptr() if theta is not None else None, theta_ld,
                occ.data_ptr() if occ is not None else None, n_draws, int(separate),
                int(precision), ngal_flat.data_ptr() + 8 * ngal_offset, ngal_flat.stride(0),
                xi_flat.data_ptr() + 8 * xi_offset, xi_flat.stride(0),
                workspace.data_ptr(), workspace.numel(), stream))


class _SingleDrawBuffers:
    """Persistent buffers of the one-draw fast path (``TabCorr.predict(model)``): the parameters
    and the results live in pinned host memory that the kernels access directly (unified
    addressing), so a call is one ctypes call (two launches) and one stream synchronisation -- no
    copies, no allocations."""

    def __init__(self, group):
        torch = _torch()
        f64 = torch.float64
        self.theta = torch.zeros(len(THETA_KEYS), dtype=f64, pin_memory=True)
        # families that run as occupation kernel -> contraction: any parameter vector, occupations
        self.theta_any = torch.zeros(32, dtype=f64, pin_memory=True)   # >= TC_N_THETA_MAX
        self.theta_any_np = self.theta_any.numpy()
        self.occ = torch.zeros((1, group.n_rows), dtype=f64, device=group.device)
        self.ngal = torch.zeros(2 * group.n_tables, dtype=f64, pin_memory=True)
        self.xi = torch.zeros(group.n_tables * group.n_r * 3, dtype=f64, pin_memory=True)
        self.theta_np, self.ngal_np, self.xi_np = (self.theta.numpy(), self.ngal.numpy(),
                                                   self.xi.numpy())
        need = max(int(group.lib.tc_predict_workspace_bytes(group.handle, 1, sep))
                   for sep in (0, 1))
        self.workspace = torch.empty(max(need, 8), dtype=torch.uint8, device=group.device)


SMALL_BATCH = 4096   # batches up to this size take the zero-copy path of predict_batch (measured
#                      against the copy pipeline: 2048 draws 198 -> 109 us, 4096 223 -> 195 us, 8192 286 -> 446 us)


class _SmallBatchBuffers:
    """Persistent pinned buffers of the small-batch fast path (``DeviceTableGroup.predict_small``):
    parameters as one contiguous column per parameter (``theta_ld = capacity``), results for
    ``capacity`` draws, workspace on the device."""

    def __init__(self, group, capacity, occupation=False):
        torch = _torch()
        f64 = torch.float64
        self.capacity = int(capacity)
        # occupation=True: the buffers of predict_small_occupation (any family's parameter vector,
        # occupations of the batch on the device)
        n_theta = 32 if occupation else len(THETA_KEYS)   # 32 >= TC_N_THETA_MAX
        self.theta = torch.zeros((n_theta, self.capacity), dtype=f64, pin_memory=True)
        if occupation:
            self.occ = torch.zeros((self.capacity, group.n_rows), dtype=f64, device=group.device)
        self.one = torch.zeros(len(THETA_KEYS), dtype=f64)   # plain host memory: read at call time
        self.one_np = self.one.numpy()
        self.ngal = torch.zeros(self.capacity * 2 * group.n_tables, dtype=f64, pin_memory=True)
        self.xi = torch.zeros(self.capacity * group.n_tables * group.n_r * 3, dtype=f64,
                              pin_memory=True)
        self.theta_np, self.ngal_np, self.xi_np = (self.theta.numpy(), self.ngal.numpy(),
                                                   self.xi.numpy())
        # large enough for either precision: the 3xTF32 mode then takes the same (tcgen05) path for
        # small and large batches, so its results do not depend on how a batch is cut
        need = max(int(group.lib.tc_predict_workspace_bytes_for(group.handle, self.capacity, sep,
                                                                prec))
                   for sep in (0, 1) for prec in (_lib.TC_PRECISION_FP64, _lib.TC_PRECISION_3XTF32))
        self.workspace = torch.empty(max(need, 8), dtype=torch.uint8, device=group.device)


class _PinnedStage:
    """Pinned host staging buffer whose numpy view is filled in place, then copied H2D."""

    def __init__(self):
        self.tensor = None

    def __call__(self, shape):
        torch = _torch()
        self.tensor = torch.empty(tuple(shape), dtype=torch.float64, pin_memory=True)
        return self.tensor.numpy()

    def to(self, device):
        return self.tensor.to(device=device, non_blocking=True)


def _to_device_f64(array, device):
    """Host array -> CUDA float64 tensor through pinned memory (asynchronous H2D copy)."""
    torch = _torch()
    if isinstance(array, torch.Tensor):
        return array.to(device=device, dtype=torch.float64, non_blocking=True).contiguous()
    array = np.asarray(array, dtype=np.float64)
    stage = _PinnedStage()
    stage(array.shape)[...] = array
    return stage.to(device)


def theta_to_device(params, spec, device):
    """Parameter dict -> CUDA ``[B, 7]`` tensor.  The columns are staged contiguously in pinned
    memory (filling ``[B, 7]`` rows column by column costs 4x more host time), copied
    asynchronously and transposed on the device."""
    torch = _torch()
    columns = theta_columns(params, spec)
    n_draws = batch_size(columns)
    stage = torch.empty((len(columns), n_draws), dtype=torch.float64, pin_memory=n_draws > 0)
    stage_np = stage.numpy()
    for j, column in enumerate(columns):
        stage_np[j] = column
    return stage.to(device=device, non_blocking=True).t().contiguous()


def _to_host(tensor):
    """CUDA tensor -> numpy array through pinned memory (one asynchronous D2H copy + sync)."""
    torch = _torch()
    if not tensor.is_cuda:  # already on the host (gloo tests of the gather plumbing)
        return tensor.numpy()
    host = torch.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
    host.copy_(tensor, non_blocking=True)
    torch.cuda.current_stream(tensor.device).synchronize()
    return host.numpy()


class TabCorr:
    """Tabulated halo correlation functions with device-resident tables."""

    def __init__(self):
        self.init = False
        self._device_group = None
        self._device = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_arrays(cls, gal_type, tpcf_matrix, tpcf_shape, attrs, tpcf_args=(), tpcf_kwargs=None,
                    device=None, upload=True):
        """Build a table from plain arrays (what ``read`` does after parsing the file).

        With ``upload`` (default) the table is copied to the device right away when one is
        present; otherwise on its first prediction (an ``Interpolator`` uploads whole groups)."""
        halotab = cls()
        halotab.attrs = dict(attrs)
        halotab.tpcf_matrix = np.asarray(tpcf_matrix).astype(np.float64)  # tabcorr.py:399
        halotab.tpcf_args = tuple(tpcf_args)
        halotab.tpcf_kwargs = dict(tpcf_kwargs or {})
        halotab.tpcf_shape = tuple(int(s) for s in tpcf_shape)
        halotab.gal_type = gal_type if isinstance(gal_type, Table) else Table(gal_type)
        halotab._device = device
        halotab.init = True
        if upload and _torch().cuda.is_available():
            halotab._ensure_device()
        return halotab

    @classmethod
    def read(cls, fname, device=None, upload=True):
        """Read tabulated correlation functions from the disk (``tabcorr/tabcorr.py:374-416``).

        Parameters
        ----------
        fname : string, h5py.Group or tabcorr_b200.h5mini.Group
            Name of the file or group containing the TabCorr object.
        device : int or torch.device, optional
            CUDA device that holds the table.  Default is the current device.
        """
        if isinstance(fname, _GROUP_TYPES):
            fstream, close = fname, False
        else:
            fname = os.fspath(fname)
            fstream = _h5py.File(fname, 'r') if _h5py is not None else h5mini.File(fname)
            close = True
        try:
            attrs = {}
            for key in fstream.attrs.keys():
                attrs[key] = fstream.attrs[key]
            tpcf_matrix = fstream['tpcf_matrix'][()]
            # (a table whose arguments were all dropped by max_args_size has no such group)
            tpcf_args = (tuple(fstream['tpcf_args'][key][()] for key in fstream['tpcf_args'].keys())
                         if 'tpcf_args' in fstream else ())
            tpcf_kwargs = {}
            if 'tpcf_kwargs' in fstream:
                for key in fstream['tpcf_kwargs'].keys():
                    tpcf_kwargs[key] = fstream['tpcf_kwargs'][key][()]
            tpcf_shape = tuple(fstream['tpcf_shape'][()])
            gal_type = Table(fstream['gal_type'][()])
        finally:
            if close:
                fstream.close()
        return cls.from_arrays(gal_type, tpcf_matrix, tpcf_shape, attrs, tpcf_args, tpcf_kwargs,
                               device=device, upload=upload)

    @classmethod
    def tabulate(cls, *args, **kwargs):
        raise NotImplementedError(
            'tabulation (halotools/Corrfunc pair counting, tabcorr/tabcorr.py:23-372) is outside '
            'the accelerated path; tabulate with the reference package and TabCorr.read the file.')

    def _ensure_device(self):
        if self._device_group is None:
            with _UPLOAD_LOCK:   # host threads sharing a fresh table upload it once
                if self._device_group is None:
                    self._device_group = DeviceTableGroup(
                        self.gal_type, [self.tpcf_matrix], self.attrs['mode'],
                        int(np.prod(self.tpcf_shape)), device=self._device)
        return self._device_group

    # ------------------------------------------------------------------ model handling
    def _check_consistency(self, model):
        # same conditions and messages as tabcorr/tabcorr.py:496-535
        if sorted(model.gal_types) != sorted(['centrals', 'satellites']):
            raise ValueError(
                'The model instance must only have centrals and satellites as galaxy types. '
                'Check the `gal_types` attribute of the model instance.')
        components = model._input_model_dictionary
        for name in ('centrals_occupation', 'satellites_occupation'):
            if components[name].prim_haloprop_key != self.attrs['prim_haloprop_key']:
                raise ValueError('Mismatch in the primary halo properties of the model and the '
                                 'TabCorr instance.')
        for name in ('centrals_occupation', 'satellites_occupation'):
            if (hasattr(components[name], 'sec_haloprop_key') and
                    components[name].sec_haloprop_key != self.attrs['sec_haloprop_key']):
                raise ValueError('Mismatch in the secondary halo properties of the model and the '
                                 'TabCorr instance.')
        if not np.abs(model.redshift - self.attrs['redshift']) < 0.05:
            raise ValueError('Mismatch in the redshift of the model and the TabCorr instance.')

    @staticmethod
    def _no_occ_kwargs(occ_kwargs):
        """``**occ_kwargs`` of ``mean_occupation`` / ``predict`` (``tabcorr/tabcorr.py:556-563``:
        passed on to ``model.mean_occupation_<gal_type>`` together with ``prim_haloprop`` and
        ``sec_haloprop_percentile``).  The occupation components of the implemented families
        (halotools Zheng07Cens/Sats, Leauthaud11Cens/Sats, their HeavisideAssembias decorations)
        read exactly four keywords: the two the reference supplies itself -- repeating them is the
        reference's ``TypeError`` -- and ``table`` / ``sec_haloprop``, which halotools ignores once
        those two are present.  Anything else can only belong to a user-defined component the
        kernel does not implement."""
        for key in occ_kwargs:
            if key in ('prim_haloprop', 'sec_haloprop_percentile'):
                raise TypeError("mean_occupation() got multiple values for keyword argument "
                                "'{}'".format(key))
        unknown = [k for k in occ_kwargs if k not in ('table', 'sec_haloprop')]
        if unknown:
            raise NotImplementedError(
                'keyword arguments for the occupation functions ({}) are not supported by the '
                'CUDA occupation kernel'.format(', '.join(unknown)))

    # ------------------------------------------------------------------ reference API
    def mean_occupation(self, model, n_gauss_prim=10, check_consistency=True, **occ_kwargs):
        """Mean occupation of each halo/galaxy bin (``tabcorr/tabcorr.py:465-578``).

        Returns a numpy array with the length of ``self.gal_type``.
        """
        self._no_occ_kwargs(occ_kwargs)
        if check_consistency:
            self._check_consistency(model)
        spec = resolve_model(model)
        group = self._ensure_device()
        theta = _to_device_f64(theta_from_params(model.param_dict, 1, spec), group.device)
        return group.occupation(spec, int(n_gauss_prim), theta)[0].cpu().numpy()

    def mean_occupation_batch(self, params, n_gauss_prim=10, model=None):
        """``mean_occupation`` for B parameter sets: CUDA tensor ``[B, N]``."""
        spec, theta = self._spec_and_theta(params, model)
        return self._ensure_device().occupation(spec, int(n_gauss_prim), theta)

    def _spec_and_theta(self, params, model):
        torch = _torch()
        group = self._ensure_device()
        if isinstance(params, dict):
            spec = resolve_model(model) if model is not None else spec_from_params(params)
            theta = theta_to_device(params, spec, group.device)
        else:
            spec = resolve_model(model) if model is not None else ModelSpec()
            theta = _to_device_f64(params, group.device)
            if theta.ndim != 2 or theta.shape[1] not in (spec.n_base, spec.n_theta):
                raise ValueError('params must be a dict of arrays or a [B, {}|{}] array ordered as '
                                 '{}'.format(spec.n_base, spec.n_theta,
                                             ', '.join(spec.theta_keys)))
            if theta.shape[1] == spec.n_base:
                theta = torch.cat([theta, torch.zeros((theta.shape[0], spec.n_theta - spec.n_base),
                                                      dtype=torch.float64, device=theta.device)],
                                  dim=1)
        return spec, theta

    def predict_batch(self, params, separate_gal_type=False, n_gauss_prim=10, model=None,
                      occupation=None, as_numpy=True, pipeline_chunk='auto', precision='fp64',
                      out=None):
        """Predict number density and correlation function for B parameter sets at once.

        Parameters
        ----------
        params : dict of arrays ``[B]`` or array/tensor ``[B, 5|7]``
            Occupation parameters keyed by their halotools names (``logMmin, sigma_logM, logM0,
            logM1, alpha`` and, for decorated models, the two ``*_assembias_param1``), or the
            same as columns in that order.  With a leauthaud11 / hearin15 ``model`` the keys are
            ``models.LEAUTHAUD11_KEYS`` (16 | 18 columns).  Ignored when ``occupation`` is given.
        separate_gal_type : bool, optional
            Split the result by galaxy type like ``predict`` does.
        n_gauss_prim : int, optional
            Gauss-Legendre points per mass bin.
        model : model instance or ModelSpec, optional
            Supplies the occupation family (decoration, split, ``modulate_with_cenocc``).  By
            default the family is inferred from the keys of ``params``.
        occupation : array ``[B, N]``, optional
            Precomputed mean occupations (the ndarray branch of ``predict``).
        as_numpy : bool, optional
            Return host numpy arrays (default) or leave the results on the device.
        precision : 'fp64' or '3xtf32', optional
            'fp64' (default) matches the reference to rtol 1e-10.  '3xtf32' contracts
            auto-correlation tables on the TF32 tensor cores with split operands (relative error
            ~1e-7, about twice the throughput); occupations stay FP64.  Cross-correlation tables
            always use FP64 (they are bound by the occupation arithmetic).
        out : tuple of two host tensors/arrays, optional
            Pre-allocated destination of the host results: ``ngal [B, 1|2]`` and ``xi [B, R, C]``
            float64 (C = 1, or 3 / 2 with ``separate_gal_type``), ideally pinned or
            ``cudaHostRegister``-ed memory (e.g. the shared segment ``predict_batch_sharded`` lets
            every rank write its slice into).  The returned arrays are views of them.
        pipeline_chunk : 'auto', int or sequence of int, optional
            With host inputs and host outputs the draws are cut into chunks whose host-to-device
            copy, kernels and device-to-host copy overlap on three CUDA streams.  An int is a
            fixed chunk size, a sequence gives the chunk sizes explicitly (the last one repeats),
            0 disables chunking.  'auto' (default) uses a short first and last chunk -- only the
            first chunk's upload and the last chunk's download are not hidden behind a kernel --
            and few large chunks in between, because every launch pays an un-overlapped
            occupation phase for its first tile.  Results do not depend on it.

        Returns
        -------
        ngal : ``[B]`` (or dict of ``[B]``), xi : ``[B, *tpcf_shape]`` (or dict thereof)
        """
        torch = _torch()
        group = self._ensure_device()
        separate = bool(separate_gal_type)
        precision = _lib.precision_code(precision)
        if group.mode != 'auto':
            precision = _lib.TC_PRECISION_FP64
        if out is not None and not as_numpy:
            raise ValueError('out= holds host results; it cannot be combined with as_numpy=False')
        if (occupation is None and as_numpy and out is None and
                not isinstance(params, torch.Tensor)):
            small = self._predict_batch_small(params, model, separate, int(n_gauss_prim), precision)
            if small is not None:
                return small
        if (occupation is None and as_numpy and not isinstance(params, torch.Tensor) and
                (isinstance(pipeline_chunk, (str, list, tuple)) or
                 (pipeline_chunk and pipeline_chunk > 0))):
            return self._predict_batch_pipelined(params, model, separate, int(n_gauss_prim),
                                                 pipeline_chunk, precision, out)
        if occupation is not None:
            occ = _to_device_f64(occupation, group.device)
            if occ.ndim != 2 or occ.shape[1] != group.n_rows:
                raise ValueError('occupation must have shape [B, {}]'.format(group.n_rows))
            spec, theta, n_draws = None, None, occ.shape[0]
        else:
            spec, theta = self._spec_and_theta(params, model)
            occ, n_draws = None, theta.shape[0]
        n_comp = group.n_comp(separate)
        ngal = torch.empty((n_draws, 2 if separate else 1), dtype=torch.float64,
                           device=group.device)
        xi = torch.empty((n_draws, group.n_r, n_comp), dtype=torch.float64, device=group.device)
        group.predict_into(spec, int(n_gauss_prim), theta, occ, separate, ngal, 0, xi, 0,
                           precision=precision)
        if out is not None:
            ngal_out, xi_out = self._host_out(out, n_draws, ngal.shape[1], n_comp)
            ngal_out.copy_(ngal, non_blocking=True)
            xi_out.copy_(xi, non_blocking=True)
            torch.cuda.current_stream(group.device).synchronize()
            return self._format_batch(ngal_out.numpy(), xi_out.numpy(), separate, False)
        return self._format_batch(ngal, xi, separate, as_numpy)

    def predict_into_slab(self, theta, slab, n_gauss_prim=10, model=None, precision='fp64'):
        """Device-resident batch whose results go straight into rows of a ``[B, 1 + R]`` CUDA
        slab (column 0: ngal, columns 1..R: xi) -- the layout the multi-GPU gather moves
        (``distributed.gather_slab_chunks``).  The kernels write through the output strides of
        ``tc_predict_batch``; nothing is packed or concatenated afterwards.  ``theta``: CUDA
        ``[B, n_theta]`` tensor in kernel order; ``slab`` may be a row range of a larger tensor."""
        group = self._ensure_device()
        spec = resolve_model(model) if model is not None else ModelSpec()
        n_draws = theta.shape[0]
        code = _lib.precision_code(precision) if group.mode == 'auto' else _lib.TC_PRECISION_FP64
        if hasattr(slab, 'ptr'):   # distributed.SlabRows: rows of a peer slab (maybe another GPU's)
            if (slab.n_rows, slab.width) != (n_draws, 1 + group.n_r):
                raise ValueError('slab rows must be [{}, {}]'.format(n_draws, 1 + group.n_r))
            group.predict_into_raw(spec, int(n_gauss_prim), theta, slab.ptr, slab.width,
                                   slab.ptr + 8, slab.width, precision=code)
            return
        if tuple(slab.shape) != (n_draws, 1 + group.n_r) or slab.stride(1) != 1:
            raise ValueError('slab must be a [{}, {}] tensor with unit column stride'.format(
                n_draws, 1 + group.n_r))
        group.predict_into_raw(spec, int(n_gauss_prim), theta, slab.data_ptr(), slab.stride(0),
                               slab.data_ptr() + 8, slab.stride(0), precision=code)

    def _predict_batch_small(self, params, model, separate, n_gauss, precision):
        """The zero-copy path for host batches of at most ``SMALL_BATCH`` draws of a family the
        fused kernel implements; None when the batch does not qualify."""
        if isinstance(params, dict):
            spec = resolve_model(model) if model is not None else spec_from_params(params)
            if not spec.latency_paths and precision != _lib.TC_PRECISION_FP64:
                return None
            columns = theta_columns(params, spec)
        else:
            spec = resolve_model(model) if model is not None else ModelSpec()
            array = np.asarray(params, dtype=np.float64)
            if not spec.latency_paths:
                if (precision != _lib.TC_PRECISION_FP64 or array.ndim != 2 or
                        array.shape[1] not in (spec.n_base, spec.n_theta)):
                    return None   # the general path reports shape errors
            elif array.ndim != 2 or array.shape[1] not in (5, 7):
                return None
            columns = [array[:, j] for j in range(array.shape[1])]
            columns += [np.float64(0.0)] * (spec.n_theta - len(columns))
        n_draws = batch_size(columns)
        if (n_draws == 0 or n_draws > SMALL_BATCH or
                any(np.ndim(c) > 0 and c.shape[0] != n_draws for c in columns)):
            return None
        group = self._ensure_device()
        if spec.latency_paths:
            ngal, xi = group.predict_small(spec, n_gauss, columns, n_draws, separate, precision)
        else:
            ngal, xi = group.predict_small_occupation(spec, n_gauss, columns, n_draws, separate)
        return self._format_batch(ngal, xi.reshape(n_draws, group.n_r, group.n_comp(separate)),
                                  separate, False)

    def _host_out(self, out, n_draws, n_ng, n_comp):
        """Validate ``out=(ngal, xi)`` and return it as two float64 host tensors."""
        torch = _torch()
        group = self._ensure_device()
        tensors = []
        for member, shape in zip(out, ((n_draws, n_ng), (n_draws, group.n_r, n_comp))):
            tensor = member if isinstance(member, torch.Tensor) else torch.from_numpy(member)
            if (tensor.dtype != torch.float64 or tensor.is_cuda or not tensor.is_contiguous() or
                    tuple(tensor.shape) != shape):
                raise ValueError('out must hold contiguous float64 host arrays of shapes {} and '
                                 '{}'.format((n_draws, n_ng), (n_draws, group.n_r, n_comp)))
            tensors.append(tensor)
        return tensors

    def _predict_batch_pipelined(self, params, model, separate, n_gauss, chunk,
                                 precision=_lib.TC_PRECISION_FP64, out=None):
        """Host parameters in, host results out: the draws are cut into chunks; chunk i + 1 is
        staged in pinned memory and copied to the device while chunk i is evaluated and chunk
        i - 1 is copied back (copy streams + events; the kernels stay on the current stream)."""
        torch = _torch()
        group = self._ensure_device()
        device = group.device
        if isinstance(params, dict):
            spec = resolve_model(model) if model is not None else spec_from_params(params)
            columns = theta_columns(params, spec)
        else:
            spec = resolve_model(model) if model is not None else ModelSpec()
            array = np.asarray(params, dtype=np.float64)
            if array.ndim != 2 or array.shape[1] not in (spec.n_base, spec.n_theta):
                raise ValueError('params must be a dict of arrays or a [B, {}|{}] array ordered as '
                                 '{}'.format(spec.n_base, spec.n_theta,
                                             ', '.join(spec.theta_keys)))
            columns = [array[:, j] for j in range(array.shape[1])]
            columns += [np.float64(0.0)] * (spec.n_theta - len(columns))
        n_draws = batch_size(columns)
        n_ng, n_comp = (2 if separate else 1), group.n_comp(separate)
        if n_draws == 0:   # an empty batch gives empty results (no launch)
            if out is not None:
                self._host_out(out, 0, n_ng, n_comp)
            return self._format_batch(np.empty((0, n_ng)), np.empty((0, group.n_r, n_comp)),
                                      separate, False)
        f64 = torch.float64
        # parameters are staged one contiguous column per parameter (a strided fill of [B, 7] rows
        # costs 4x more host time); chunk [lo, hi) owns the block [7 lo, 7 hi) viewed as
        # [7, hi - lo], so that every chunk is one contiguous host-to-device copy
        n_theta = spec.n_theta
        theta_pin = torch.empty(n_draws * n_theta, dtype=f64, pin_memory=True)
        if out is not None:
            ngal_pin, xi_pin = self._host_out(out, n_draws, n_ng, n_comp)
        else:
            ngal_pin = torch.empty((n_draws, n_ng), dtype=f64, pin_memory=True)
            xi_pin = torch.empty((n_draws, group.n_r, n_comp), dtype=f64, pin_memory=True)
        theta_np = theta_pin.numpy()
        theta = torch.empty(n_draws * n_theta, dtype=f64, device=device)
        ngal = torch.empty((n_draws, n_ng), dtype=f64, device=device)
        xi = torch.empty((n_draws, group.n_r, n_comp), dtype=f64, device=device)
        compute = torch.cuda.current_stream(device)
        h2d, d2h = group.copy_streams()
        # the fresh device buffers may be recycled memory with work pending on the compute stream
        h2d.wait_stream(compute)
        d2h.wait_stream(compute)
        for lo, hi in chunk_schedule(n_draws, chunk):
            block = theta_np[n_theta * lo:n_theta * hi].reshape(n_theta, hi - lo)
            for j, column in enumerate(columns):
                block[j] = column[lo:hi] if np.ndim(column) > 0 else column
            theta_chunk = theta[n_theta * lo:n_theta * hi]
            with torch.cuda.stream(h2d):
                theta_chunk.copy_(theta_pin[n_theta * lo:n_theta * hi], non_blocking=True)
                staged = torch.cuda.Event()
                staged.record(h2d)
            compute.wait_event(staged)
            group.predict_into(spec, n_gauss, theta_chunk.view(n_theta, hi - lo), None, separate,
                               ngal[lo:hi], 0, xi[lo:hi], 0, theta_columns=True,
                               precision=precision)
            done = torch.cuda.Event()
            done.record(compute)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                ngal_pin[lo:hi].copy_(ngal[lo:hi], non_blocking=True)
                xi_pin[lo:hi].copy_(xi[lo:hi], non_blocking=True)
        d2h.synchronize()
        return self._format_batch(ngal_pin.numpy(), xi_pin.numpy(), separate, False)

    def _format_batch(self, ngal, xi, separate, as_numpy):
        """``ngal [B, 1|2]``, ``xi [B, R, C]`` device tensors -> reference-shaped outputs."""
        shape = tuple(self.tpcf_shape)
        if as_numpy:
            ngal, xi = _to_host(ngal), _to_host(xi)
        if not separate:
            return ngal[:, 0], xi[:, :, 0].reshape((xi.shape[0],) + shape)
        ngal_keys, xi_keys = self._separate_keys()
        ngal_dict = {key: ngal[:, j] for j, key in ngal_keys}
        xi_dict = {key: xi[:, :, j].reshape((xi.shape[0],) + shape) for j, key in xi_keys}
        return ngal_dict, xi_dict

    def _separate_keys(self):
        """Dictionary keys of ``separate_gal_type`` results in the reference's order
        (``np.unique`` of the gal_type column, ``tabcorr/tabcorr.py:660-681``)."""
        names = [str(n) for n in np.unique(self.gal_type['gal_type'].data)]
        for name in names:
            if name not in ('centrals', 'satellites'):
                raise NotImplementedError(
                    "galaxy type '{}': only 'centrals' and 'satellites' are supported".format(name))
        index = {'centrals': 0, 'satellites': 1}
        ngal_keys = [(index[n], n) for n in names]
        if self.attrs['mode'] == 'auto':
            xi_keys = []
            for i, n1 in enumerate(names):
                for n2 in names[i:]:
                    xi_keys.append((index[n1] + index[n2], '%s-%s' % (n1, n2)))
        else:
            xi_keys = ngal_keys
        return ngal_keys, xi_keys

    def predict(self, model, separate_gal_type=False, n_gauss_prim=10, check_consistency=True,
                **occ_kwargs):
        """Predict the number density and correlation function for a model
        (``tabcorr/tabcorr.py:580-683``).

        ``model`` is a halotools-style model instance or a numpy array with the mean occupation
        of each halo bin.  Returns ``(ngal, xi)`` as a numpy scalar and an array of shape
        ``tpcf_shape``, or two dictionaries if ``separate_gal_type`` is True.
        """
        if isinstance(model, np.ndarray):
            result = self.predict_batch(None, separate_gal_type, n_gauss_prim,
                                        occupation=model[np.newaxis, :])
        else:
            self._no_occ_kwargs(occ_kwargs)
            if check_consistency:
                self._check_consistency(model)
            spec = resolve_model(model)
            values = theta_from_params(model.param_dict, 1, spec)[0]
            ngal, xi = self._ensure_device().predict_one(spec, int(n_gauss_prim), values,
                                                         bool(separate_gal_type))
            result = self._format_batch(ngal, xi, bool(separate_gal_type), False)
        ngal, xi = result
        if separate_gal_type:
            return ({k: v[0] for k, v in ngal.items()}, {k: v[0] for k, v in xi.items()})
        return ngal[0], xi[0]

    def write(self, fname, overwrite=False, max_args_size=1000000, matrix_dtype=np.float32):
        """Write the table in the reference's HDF5 layout (``tabcorr/tabcorr.py:418-463``)."""
        from . import h5write
        h5write.write_tabcorr(self, fname, overwrite=overwrite, max_args_size=max_args_size,
                              matrix_dtype=matrix_dtype)
