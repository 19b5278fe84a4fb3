"""Drop-in ``Interpolator``: tensor-product cubic spline over a grid of device-resident tables.

Mirrors ``tabcorr/interpolator.py``: ``__init__`` (:14-70) validates that ``param_dict_table``
describes a full rectangular grid, sorts it and de-duplicates identical halo tables; ``predict``
(:124-216) evaluates every grid table and applies ``spline_interpolate`` (:275-331).  Here all
tables that share a halo table form one device table group, so a prediction is one fused
occupation + contraction launch over ``T * R`` stacked radial bins followed by the spline kernel.
"""

import ctypes
import threading

import numpy as np

from . import _lib
from . import h5mini
from .models import ModelSpec, resolve_model, spec_from_params
from .tabcorr import (TabCorr, DeviceTableGroup, _to_device_f64, _torch, _h5py,
                      theta_to_device)
from .table import Table


def spline_interpolation_matrix(xp):
    """Matrix form of the not-a-knot cubic spline through knots ``xp``
    (``tabcorr/interpolator.py:219-272``).

    Returns ``a`` with shape ``[len(xp) - 1, 4, len(xp)]`` such that the spline on segment ``i``
    is ``sum_p sum_k a[i, p, k] y[k] x**p``.  The 4n polynomial coefficients (n segments) solve a
    linear system whose rows are, in the reference's order: value at the left knot of each
    segment, value at the right knot, continuity of the first and of the second derivative at the
    interior knots, and continuity of the third derivative at the second and second-to-last knot.
    """
    xp = np.asarray(xp, dtype=np.float64)
    if len(xp) < 4:
        raise ValueError('Cannot perform spline interpolation with less than 4 values.')
    n = len(xp) - 1
    powers = np.arange(4)

    def basis(x, derivative):
        # d^derivative/dx^derivative of (1, x, x^2, x^3)
        coeff = np.array([np.prod(np.arange(p, p - derivative, -1)) if p >= derivative else 0.0
                          for p in powers], dtype=np.float64)
        expo = np.clip(powers - derivative, 0, None)
        return coeff * x**expo

    system = np.zeros((4 * n, 4 * n))
    for seg in range(n):
        cols = slice(4 * seg, 4 * seg + 4)
        system[seg, cols] = basis(xp[seg], 0)
        system[n + seg, cols] = basis(xp[seg + 1], 0)
    for seg in range(n - 1):
        left, right = slice(4 * seg, 4 * seg + 4), slice(4 * seg + 4, 4 * seg + 8)
        for derivative, row in ((1, 2 * n + seg), (2, 3 * n - 1 + seg)):
            system[row, left] = basis(xp[seg + 1], derivative)
            system[row, right] = -basis(xp[seg + 1], derivative)
    # The reference scales these two rows by the knot position (interpolator.py:262-265), which
    # leaves the solution unchanged; keep the scaling for bit-level agreement of the inverse, except
    # where the knot is 0 and the reference's system would be singular.
    scale_hi = xp[n - 1] if xp[n - 1] != 0 else 1.0
    scale_lo = xp[1] if xp[1] != 0 else 1.0
    system[4 * n - 2, 4 * (n - 2):4 * (n - 1)] = scale_hi * basis(xp[n - 1], 3)
    system[4 * n - 2, 4 * (n - 1):4 * n] = -scale_hi * basis(xp[n - 1], 3)
    system[4 * n - 1, 0:4] = scale_lo * basis(xp[1], 3)
    system[4 * n - 1, 4:8] = -scale_lo * basis(xp[1], 3)

    inverse = np.linalg.inv(system)
    # right-hand side: rows [0, n) equal y[0..n-1], rows [n, 2n) equal y[1..n], the rest zero
    a = np.zeros((4 * n, len(xp)))
    a[:, :n] += inverse[:, :n]
    a[:, 1:] += inverse[:, n:2 * n]
    return a.reshape(n, 4, len(xp))


class Interpolator:
    """Interpolation of multiple TabCorr instances on a rectangular parameter grid."""

    def __init__(self, tabcorr_list, param_dict_table, device=None):
        if not isinstance(param_dict_table, Table):
            param_dict_table = Table(param_dict_table)
        if len(tabcorr_list) != len(param_dict_table):
            raise ValueError("The number of TabCorr instances does not match the number of "
                             "entries in 'param_dict_table'.")
        self.tabcorr_list = tabcorr_list
        self.param_dict_table = param_dict_table.copy()
        self._keys = list(self.param_dict_table.colnames)

        self.xp = []
        self.a = []
        for key in self._keys:
            self.xp.append(np.sort(np.unique(param_dict_table[key].data)))
            self.a.append(spline_interpolation_matrix(self.xp[-1]))

        records = self.param_dict_table.as_array()
        if (int(np.prod([len(xp) for xp in self.xp])) != len(self.param_dict_table) or
                len(np.unique(records)) != len(records)):
            raise ValueError("The 'param_dict_table' does not describe a grid.")

        self.param_dict_table['tabcorr_index'] = np.arange(len(self.param_dict_table))
        self.param_dict_table.sort(self.param_dict_table.colnames)

        # identical halo tables share their occupations (interpolator.py:63-70)
        all_gal_type = [np.array(tabcorr.gal_type.as_array().tolist()).ravel()
                        for tabcorr in tabcorr_list]
        unique = np.unique(all_gal_type, axis=0, return_index=True, return_inverse=True)
        self.unique_gal_type_index = unique[1]
        self.unique_gal_type_inverse = np.asarray(unique[2]).ravel()

        first = tabcorr_list[0]
        for tabcorr in tabcorr_list:
            if (tabcorr.attrs['mode'] != first.attrs['mode'] or
                    tuple(tabcorr.tpcf_shape) != tuple(first.tpcf_shape)):
                raise ValueError('All TabCorr instances must share mode and tpcf_shape.')
        self._device = device
        self._groups = None
        self._interp = None
        self._one = None
        self._one_lock = threading.Lock()

    @property
    def tpcf_shape(self):
        """Shape of one prediction (shared by all grid tables)."""
        return tuple(self.tabcorr_list[0].tpcf_shape)

    # ------------------------------------------------------------------ I/O
    @classmethod
    def read(cls, fname, device=None):
        """Read a TabCorr interpolator from the disk (``tabcorr/interpolator.py:72-96``)."""
        import os
        fname = os.fspath(fname)
        fstream = _h5py.File(fname, 'r') if _h5py is not None else h5mini.File(fname)
        try:
            param_dict_table = Table(fstream['param_dict_table'][()])
            param_dict_table.sort('tabcorr_index')
            param_dict_table.remove_column('tabcorr_index')
            # device groups are built by the Interpolator: keep the single tables host-side
            tabcorr_list = [TabCorr.read(fstream['tabcorr_{}'.format(i)], device=device,
                                         upload=False)
                            for i in range(len(param_dict_table))]
        finally:
            fstream.close()
        return cls(tabcorr_list, param_dict_table, device=device)

    def write(self, fname, overwrite=False, max_args_size=1000000, matrix_dtype=np.float32):
        """Write the interpolator in the reference's layout (``tabcorr/interpolator.py:98-122``)."""
        from . import h5write
        h5write.write_interpolator(self, fname, overwrite=overwrite, max_args_size=max_args_size,
                                   matrix_dtype=matrix_dtype)

    # ------------------------------------------------------------------ device state
    def _ensure_device(self):
        if self._groups is not None:
            return
        from .tabcorr import _UPLOAD_LOCK
        with _UPLOAD_LOCK:   # host threads sharing a fresh interpolator upload it once
            if self._groups is None:
                self._upload()

    def _upload(self):
        lib = _lib.load()
        first = self.tabcorr_list[0]
        n_r = int(np.prod(first.tpcf_shape))
        groups = []
        for u in range(len(self.unique_gal_type_index)):
            members = [k for k in range(len(self.tabcorr_list))
                       if self.unique_gal_type_inverse[k] == u]
            group = DeviceTableGroup(
                self.tabcorr_list[members[0]].gal_type,
                [self.tabcorr_list[k].tpcf_matrix for k in members], first.attrs['mode'], n_r,
                device=self._device)
            groups.append((group, members))
        # position of every table in the stacked [B, T, ...] buffers: groups back to back
        slot_of_table = np.zeros(len(self.tabcorr_list), dtype=np.int64)
        slot = 0
        for group, members in groups:
            for k in members:
                slot_of_table[k] = slot
                slot += 1
        grid_to_slot = np.ascontiguousarray(
            slot_of_table[np.asarray(self.param_dict_table['tabcorr_index'].data)], dtype=np.int32)
        n_knots = np.ascontiguousarray([len(xp) for xp in self.xp], dtype=np.int32)
        knots = np.ascontiguousarray(np.concatenate(self.xp), dtype=np.float64)
        a = np.ascontiguousarray(np.concatenate([m.ravel() for m in self.a]), dtype=np.float64)
        handle = ctypes.c_void_p()
        _lib.check(lib.tc_interp_create(
            ctypes.byref(handle), len(self.xp), _lib.as_int32_p(n_knots), _lib.as_double_p(knots),
            _lib.as_double_p(a), _lib.as_int32_p(grid_to_slot), groups[0][0].device))
        self._lib = lib
        self._interp = handle
        self._groups = groups

    def __del__(self):
        handle = getattr(self, '_interp', None)
        if handle:
            try:
                self._lib.tc_interp_destroy(handle)
            except Exception:
                pass
            self._interp = None

    # ------------------------------------------------------------------ prediction
    def _small_capacity(self):
        """Draws the persistent buffers of :meth:`_predict_small` hold: ``SMALL_BATCH``, less for
        large grids (the per-table results of all draws stay below 64 MB of device memory)."""
        from .tabcorr import SMALL_BATCH
        self._ensure_device()
        per_draw = len(self.tabcorr_list) * self._groups[0][0].n_r * 3 * 8
        return int(max(1, min(SMALL_BATCH, (64 << 20) // per_draw)))

    def _predict_small(self, spec, n_gauss, columns, x_columns, n_draws, separate, extrapolate,
                       precision=_lib.TC_PRECISION_FP64):
        """Latency path of :meth:`predict` and of small :meth:`predict_batch` calls (at most
        ``SMALL_BATCH`` host draws): persistent buffers; the parameters (one pinned column per
        parameter; in the launch arguments for a single draw), the coordinates, the results and
        the out-of-range flag live in pinned host memory that the kernels access directly; one
        stream synchronisation, no copies, no allocations.  ``columns`` / ``x_columns``: one array
        ``[B]`` or scalar per kernel parameter / interpolation axis."""
        torch = _torch()
        self._ensure_device()
        first_group = self._groups[0][0]
        device = first_group.device
        n_tables = len(self.tabcorr_list)
        n_r = first_group.n_r
        n_axes = len(self._keys)
        with self._one_lock:
            if self._one is None:
                f64 = torch.float64
                cap = self._small_capacity()
                self._one = {
                    'cap': cap,
                    'theta': torch.zeros((32, cap), dtype=f64, pin_memory=True),   # >= TC_N_THETA_MAX
                    'occ': [None] * len(self._groups),   # occupation-kernel families, on demand
                    'one': torch.zeros(7, dtype=f64),
                    'x': torch.zeros((cap, n_axes), dtype=f64, pin_memory=True),
                    'flag': torch.zeros(1, dtype=torch.int32, pin_memory=True),
                    'ngal': torch.zeros(cap * 2, dtype=f64, pin_memory=True),
                    'xi': torch.zeros(cap * n_r * 3, dtype=f64, pin_memory=True),
                    'ngal_t': torch.zeros(cap * n_tables * 2, dtype=f64, device=device),
                    'xi_t': torch.zeros(cap * n_tables * n_r * 3, dtype=f64, device=device),
                    'workspace': [torch.empty(max(8, max(
                        int(self._lib.tc_predict_workspace_bytes(group.handle, cap, sep))
                        for sep in (0, 1))), dtype=torch.uint8, device=device)
                        for group, _ in self._groups],
                }
            buf = self._one
            cap = buf['cap']
            theta_np = buf['theta'].numpy()
            for j, column in enumerate(columns):
                theta_np[j, :n_draws] = column
            x_np = buf['x'].numpy()
            for d, column in enumerate(x_columns):
                x_np[:n_draws, d] = column
            buf['flag'].numpy()[0] = 0
            n_ng, n_comp = (2 if separate else 1), first_group.n_comp(separate)
            n_cols = n_r * n_comp
            stream = torch.cuda.current_stream(device)
            model = DeviceTableGroup._model_struct(spec)
            slot = 0
            for index, ((group, members), workspace) in enumerate(zip(self._groups,
                                                                      buf['workspace'])):
                group.plan(n_gauss)
                ngal_ptr = buf['ngal_t'].data_ptr() + 8 * slot * n_ng
                xi_ptr = buf['xi_t'].data_ptr() + 8 * slot * n_cols
                if not spec.latency_paths:
                    # occupation kernel -> contraction on its output (leauthaud11 / hearin15,
                    # mass-dependent decoration), the occupations in a persistent device buffer
                    if buf['occ'][index] is None:
                        buf['occ'][index] = torch.zeros((cap, group.n_rows), dtype=torch.float64,
                                                        device=device)
                    occ = buf['occ'][index]
                    _lib.check(self._lib.tc_occupation_batch(
                        group.handle, ctypes.byref(model), int(n_gauss), buf['theta'].data_ptr(),
                        cap, int(n_draws), occ.data_ptr(), stream.cuda_stream))
                    _lib.check(self._lib.tc_predict_batch(
                        group.handle, ctypes.byref(model), int(n_gauss), None, 0, occ.data_ptr(),
                        int(n_draws), int(separate), _lib.TC_PRECISION_FP64, ngal_ptr,
                        n_tables * n_ng, xi_ptr, n_tables * n_cols, workspace.data_ptr(),
                        workspace.numel(), stream.cuda_stream))
                elif n_draws == 1:
                    buf['one'].numpy()[:] = theta_np[:7, 0]
                    _lib.check(self._lib.tc_predict_one(
                        group.handle, ctypes.byref(model), int(n_gauss), buf['one'].data_ptr(),
                        int(separate), int(precision), ngal_ptr, n_tables * n_ng, xi_ptr,
                        n_tables * n_cols, workspace.data_ptr(), workspace.numel(),
                        stream.cuda_stream))
                else:
                    _lib.check(self._lib.tc_predict_batch(
                        group.handle, ctypes.byref(model), int(n_gauss), buf['theta'].data_ptr(),
                        cap, None, int(n_draws), int(separate), int(precision), ngal_ptr,
                        n_tables * n_ng, xi_ptr, n_tables * n_cols, workspace.data_ptr(),
                        workspace.numel(), stream.cuda_stream))
                slot += len(members)
            for data, cols, out in ((buf['ngal_t'], n_ng, buf['ngal']), (buf['xi_t'], n_cols, buf['xi'])):
                _lib.check(self._lib.tc_interp_apply_batch(
                    self._interp, buf['x'].data_ptr(), int(n_draws), data.data_ptr(), cols,
                    out.data_ptr(), int(bool(extrapolate)), buf['flag'].data_ptr(),
                    stream.cuda_stream))
            stream.synchronize()
            if int(buf['flag'].numpy()[0]) != 0:
                raise ValueError('The x-coordinates are outside of the interpolation range and '
                                 'extrapolation is turned off.')
            ngal = buf['ngal'].numpy()[:n_draws * n_ng].reshape(n_draws, n_ng).copy()
            xi = buf['xi'].numpy()[:n_draws * n_cols].reshape(n_draws, n_r, n_comp).copy()
        return self.tabcorr_list[0]._format_batch(ngal, xi, separate, False)

    def _predict_one(self, spec, n_gauss, values, x_values, separate, extrapolate):
        """One parameter set through :meth:`_predict_small`."""
        return self._predict_small(spec, n_gauss, list(values), list(x_values), 1, separate,
                                   extrapolate)

    def predict_batch(self, params, separate_gal_type=False, n_gauss_prim=10, extrapolate=False,
                      model=None, as_numpy=True, precision='fp64', defer_range_check=False):
        """Interpolated predictions for B parameter sets.

        ``params`` is a dict of arrays ``[B]`` holding the occupation parameters and one entry per
        interpolation axis (the column names of ``param_dict_table``).  Returns ``(ngal [B],
        xi [B, *tpcf_shape])`` or per-gal-type dicts, like :meth:`predict`.

        ``defer_range_check`` (device results only): do not synchronise to test the
        out-of-range flag; return ``(ngal, xi, flag)`` with ``flag`` a one-element int32 CUDA
        tensor that is non-zero when a draw lay outside the knot hull without ``extrapolate``
        (its outputs are NaN) -- callers that issue several batches check once at the end.
        """
        torch = _torch()
        self._ensure_device()
        for key in self._keys:
            if key not in params:
                raise ValueError('The key {} is not present in the parameter dictionary of the '
                                 'model.'.format(key))
        spec = resolve_model(model) if model is not None else spec_from_params(params)
        if as_numpy and not defer_range_check and (
                spec.latency_paths or _lib.precision_code(precision) == _lib.TC_PRECISION_FP64
                or self._groups[0][0].mode != 'auto'):
            from .models import theta_columns
            columns = theta_columns(params, spec)
            x_columns = [np.asarray(params[key], dtype=np.float64) for key in self._keys]
            sizes = [c.shape[0] for c in columns + x_columns if np.ndim(c) > 0]
            n_small = max(sizes + [1])
            first_mode = self._groups[0][0].mode
            if n_small <= self._small_capacity() and all(size == n_small for size in sizes):
                code = _lib.precision_code(precision) if first_mode == 'auto' \
                    else _lib.TC_PRECISION_FP64
                return self._predict_small(spec, int(n_gauss_prim), columns, x_columns, n_small,
                                           bool(separate_gal_type), extrapolate, code)
        device = self._groups[0][0].device
        theta = theta_to_device(params, spec, device)
        n_draws = theta.shape[0]
        coordinates = [np.asarray(params[key], dtype=np.float64) for key in self._keys]
        if any(c.ndim > 0 and c.shape[0] != n_draws for c in coordinates):
            raise ValueError('interpolation coordinates and occupation parameters differ in length')
        x = _to_device_f64(np.stack([np.broadcast_to(c, (n_draws,)) for c in coordinates]),
                           device).t().contiguous()
        return self.predict_batch_tensors(theta, x, spec, separate_gal_type, n_gauss_prim,
                                          extrapolate, as_numpy, precision, defer_range_check)

    def predict_batch_tensors(self, theta, x, model=None, separate_gal_type=False,
                              n_gauss_prim=10, extrapolate=False, as_numpy=False,
                              precision='fp64', defer_range_check=False):
        """:meth:`predict_batch` for parameters that already live on the device: ``theta``
        ``[B, n_theta]`` in the kernel order of the model's family (``ModelSpec.theta_keys``) and
        the interpolation coordinates ``x [B, D]`` in the column order of ``param_dict_table``
        (device-side sweeps, ``tabcorr_b200.sweep``)."""
        torch = _torch()
        self._ensure_device()
        spec = resolve_model(model) if model is not None else ModelSpec()
        device = self._groups[0][0].device
        theta = theta.to(device=device, dtype=torch.float64).contiguous()
        x = x.to(device=device, dtype=torch.float64).contiguous()
        n_draws = theta.shape[0]
        if theta.ndim != 2 or theta.shape[1] != spec.n_theta:
            raise ValueError('theta must have shape [B, {}] ({})'.format(
                spec.n_theta, ', '.join(spec.theta_keys)))
        if x.shape != (n_draws, len(self._keys)):
            raise ValueError('x must have shape [B, {}] ({})'.format(
                len(self._keys), ', '.join(self._keys)))

        separate = bool(separate_gal_type)
        first_group = self._groups[0][0]
        precision = _lib.precision_code(precision)
        if first_group.mode != 'auto':
            precision = _lib.TC_PRECISION_FP64
        n_tables = len(self.tabcorr_list)
        n_comp = first_group.n_comp(separate)
        n_ng = 2 if separate else 1
        n_r = first_group.n_r
        ngal_t = torch.empty((n_draws, n_tables, n_ng), dtype=torch.float64, device=device)
        xi_t = torch.empty((n_draws, n_tables, n_r * n_comp), dtype=torch.float64, device=device)
        slot = 0
        for group, members in self._groups:
            group.predict_into(spec, int(n_gauss_prim), theta, None, separate, ngal_t,
                               slot * n_ng, xi_t, slot * n_r * n_comp, precision=precision)
            slot += len(members)

        flag = torch.zeros(1, dtype=torch.int32, device=device)
        ngal = torch.empty((n_draws, n_ng), dtype=torch.float64, device=device)
        xi = torch.empty((n_draws, n_r * n_comp), dtype=torch.float64, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        for data, n_cols, out in ((ngal_t, n_ng, ngal), (xi_t, n_r * n_comp, xi)):
            _lib.check(self._lib.tc_interp_apply_batch(
                self._interp, x.data_ptr(), n_draws, data.data_ptr(), n_cols, out.data_ptr(),
                int(bool(extrapolate)), flag.data_ptr(), stream))
        if defer_range_check:
            if as_numpy:
                raise ValueError('defer_range_check needs as_numpy=False')
            return self.tabcorr_list[0]._format_batch(
                ngal, xi.view(n_draws, n_r, n_comp), separate, False) + (flag,)
        if int(flag.item()) != 0:
            raise ValueError('The x-coordinates are outside of the interpolation range and '
                             'extrapolation is turned off.')
        return self.tabcorr_list[0]._format_batch(ngal, xi.view(n_draws, n_r, n_comp), separate,
                                                  as_numpy)

    def predict(self, model, separate_gal_type=False, n_gauss_prim=10, extrapolate=False,
                check_consistency=True, **occ_kwargs):
        """Interpolate the predictions from multiple TabCorr instances
        (``tabcorr/interpolator.py:124-216``).  The values of the parameters to interpolate must
        be in ``model.param_dict``."""
        TabCorr._no_occ_kwargs(occ_kwargs)
        for key in self._keys:
            if key not in model.param_dict:
                raise ValueError('The key {} is not present in the parameter dictionary of the '
                                 'model.'.format(key))
        if check_consistency:
            for i in self.unique_gal_type_index:
                self.tabcorr_list[i]._check_consistency(model)
        spec = resolve_model(model)
        from .models import theta_from_params
        values = theta_from_params(model.param_dict, 1, spec)[0]
        x_values = [np.float64(model.param_dict[key]) for key in self._keys]
        ngal, xi = self._predict_one(spec, int(n_gauss_prim), values, x_values,
                                     bool(separate_gal_type), extrapolate)
        if separate_gal_type:
            return ({k: v[0] for k, v in ngal.items()}, {k: v[0] for k, v in xi.items()})
        return ngal[0], xi[0]
