"""Per-draw selection among several tables ("per-draw cosmology", BASELINE.json configs[3]).

In the reference, cosmology and simulation phase only select WHICH file ``database.read`` opens
(``tabcorr/database.py:283-286``); nothing interpolates across cosmologies.
A sampler that also varies the cosmology therefore holds one ``Interpolator`` per cosmology and
calls the one its current draw names.  ``TableSet`` is the batched form of that: draw ``b`` carries
an integer ``table_index[b]``; the draws are sorted by index on the device, every table evaluates
its contiguous segment (one fused launch + spline kernel per table), and the results are scattered
back into draw order on the device.
"""

import numpy as np


class TableSet:
    """A list of ``TabCorr`` / ``Interpolator`` instances with identical output shapes."""

    def __init__(self, tables):
        self.tables = list(tables)
        if not self.tables:
            raise ValueError('a TableSet needs at least one table')

    def __len__(self):
        return len(self.tables)

    def __getitem__(self, i):
        return self.tables[i]

    def predict_batch(self, params, table_index, separate_gal_type=False, as_numpy=True,
                      **predict_kwargs):
        """``predict_batch`` of table ``table_index[b]`` for every draw ``b``.

        ``params``: dict of ``[B]`` arrays (scalars are broadcast); ``table_index``: ``[B]``
        integers in ``range(len(self))``.  Returns what the tables' ``predict_batch`` returns,
        in draw order.

        The draws are segmented on the DEVICE: parameters, interpolation coordinates and indices
        are uploaded once (pinned staging), one stable ``torch.sort`` of the indices orders the
        draws by table, every table evaluates its contiguous segment from device-resident
        tensors (no per-table host gathers, uploads or synchronisations; the only host round trip
        is the ``len(self)`` segment sizes), and one ``index_copy_`` per output restores the draw
        order.
        """
        import torch
        from .models import resolve_model, spec_from_params
        from .tabcorr import _to_device_f64, _to_host, theta_to_device
        table_index = np.asarray(table_index)
        if table_index.ndim != 1:
            raise ValueError('table_index must be one-dimensional')
        if not np.issubdtype(table_index.dtype, np.integer):
            if not np.all(table_index == np.round(table_index)):
                raise ValueError('table_index must hold integers')
            table_index = table_index.astype(np.int64)
        n_draws = len(table_index)
        if n_draws == 0:
            raise ValueError('empty batch')
        if table_index.min() < 0 or table_index.max() >= len(self.tables):
            raise ValueError('table_index outside range(0, {})'.format(len(self.tables)))
        first = self.tables[0]
        is_interp = hasattr(first, 'tabcorr_list')
        if not torch.cuda.is_available() or not hasattr(first, '_ensure_device'):
            return self._predict_batch_host_grouped(params, table_index, separate_gal_type,
                                                    as_numpy, predict_kwargs)
        model = predict_kwargs.pop('model', None)
        spec = resolve_model(model) if model is not None else spec_from_params(params)
        first._ensure_device()
        device = (first._groups[0][0].device if is_interp else first._device_group.device)
        theta = theta_to_device(params, spec, device)
        if theta.shape[0] != n_draws:
            theta = theta.expand(n_draws, -1)
        keys = list(first._keys) if is_interp else []
        x = None
        if is_interp:
            for table in self.tables:
                if list(table._keys) != keys:
                    raise ValueError('the Interpolators of a TableSet must share their axes')
            for key in keys:
                if key not in params:
                    raise ValueError('The key {} is not present in the parameter dictionary of '
                                     'the model.'.format(key))
            x = _to_device_f64(np.stack([np.broadcast_to(np.asarray(params[k], dtype=np.float64),
                                                         (n_draws,)) for k in keys]),
                               device).t().contiguous()
        pinned = torch.empty(n_draws, dtype=torch.int64, pin_memory=True)
        pinned.numpy()[:] = table_index
        index_dev = pinned.to(device, non_blocking=True)
        _, order = torch.sort(index_dev, stable=True)
        counts = torch.bincount(index_dev, minlength=len(self.tables)).cpu().numpy()
        theta_sorted = theta.index_select(0, order)
        x_sorted = x.index_select(0, order) if x is not None else None
        out, out_spec, flags = None, None, []
        start = 0
        for t, count in enumerate(counts):
            count = int(count)
            if count == 0:
                continue
            table = self.tables[t]
            segment = slice(start, start + count)
            if is_interp:
                result = table.predict_batch_tensors(
                    theta_sorted[segment], x_sorted[segment], spec,
                    separate_gal_type=separate_gal_type, as_numpy=False, defer_range_check=True,
                    **predict_kwargs)
                flags.append(result[2])
                result = result[:2]
            else:
                result = table.predict_batch(theta_sorted[segment], model=spec,
                                             separate_gal_type=separate_gal_type, as_numpy=False,
                                             **predict_kwargs)
            flat, spec_t = _flatten(result)
            if out is None:
                out = [torch.empty((n_draws,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                       for v in flat]
                out_spec = spec_t
            elif spec_t != out_spec:
                raise ValueError('tables of a TableSet must produce identically shaped results')
            for dst, src in zip(out, flat):
                dst.index_copy_(0, order[segment], src)
            start += count
        if flags and int(torch.stack([f.reshape(()) for f in flags]).max().item()) != 0:
            raise ValueError('The x-coordinates are outside of the interpolation range and '
                             'extrapolation is turned off.')
        if as_numpy:
            out = [_to_host(v) for v in out]
        return _unflatten(out, out_spec)

    def _predict_batch_host_grouped(self, params, table_index, separate_gal_type, as_numpy,
                                    predict_kwargs):
        """Grouping on the host, one ``predict_batch`` per table (stand-in tables of the CPU tests;
        the device path above is what runs on a GPU)."""
        import torch
        n_draws = len(table_index)
        order = np.argsort(table_index, kind='stable')
        counts = np.bincount(table_index, minlength=len(self.tables))
        out, out_spec, start = None, None, 0
        for t, count in enumerate(counts):
            if count == 0:
                continue
            rows = order[start:start + count]
            start += count
            sub = {k: (np.asarray(v)[rows] if np.ndim(v) > 0 else v) for k, v in params.items()}
            result = self.tables[t].predict_batch(sub, separate_gal_type=separate_gal_type,
                                                  as_numpy=False, **predict_kwargs)
            flat, spec = _flatten(result)
            flat = [v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
                    for v in flat]
            if out is None:
                out = [torch.empty((n_draws,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                       for v in flat]
                out_spec = spec
            elif spec != out_spec:
                raise ValueError('tables of a TableSet must produce identically shaped results')
            index = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int64))
            for dst, src in zip(out, flat):
                dst.index_copy_(0, index.to(dst.device), src)
        if as_numpy:
            out = [v.cpu().numpy() for v in out]
        return _unflatten(out, out_spec)


def _flatten(result):
    """(ngal, xi) with array or dict members -> (list of tensors, structure spec)."""
    flat, spec = [], []
    for member in result:
        if isinstance(member, dict):
            keys = list(member.keys())
            spec.append(tuple(keys))
            flat.extend(member[k] for k in keys)
        else:
            spec.append(None)
            flat.append(member)
    return flat, tuple(spec)


def _unflatten(flat, spec):
    out, i = [], 0
    for keys in spec:
        if keys is None:
            out.append(flat[i])
            i += 1
        else:
            out.append({k: flat[i + j] for j, k in enumerate(keys)})
            i += len(keys)
    return tuple(out)
