"""Per-draw selection among several tables ("per-draw cosmology", BASELINE.json configs[3]).

In the reference, cosmology and simulation phase only select WHICH file ``database.read`` opens
(``tabcorr/database.py:283-286``); nothing interpolates across cosmologies.
A sampler that also varies the cosmology therefore holds one ``Interpolator`` per cosmology and
calls the one its current draw names.  ``TableSet`` is the batched form of that: draw ``b`` carries
an integer ``table_index[b]``; the draws are grouped by index on the host, every group goes through
its table's ``predict_batch`` (one fused launch + spline kernel per table set), and the results are
scattered back into draw order on the device.
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
        """
        import torch
        from .tabcorr import _to_host
        table_index = np.asarray(table_index)
        if table_index.ndim != 1:
            raise ValueError('table_index must be one-dimensional')
        if not np.issubdtype(table_index.dtype, np.integer):
            if not np.all(table_index == np.round(table_index)):
                raise ValueError('table_index must hold integers')
            table_index = table_index.astype(np.int64)
        n_draws = len(table_index)
        if n_draws and (table_index.min() < 0 or table_index.max() >= len(self.tables)):
            raise ValueError('table_index outside range(0, {})'.format(len(self.tables)))
        order = np.argsort(table_index, kind='stable')
        counts = np.bincount(table_index, minlength=len(self.tables))
        out = None
        start = 0
        flags = []
        order_dev = None   # uploaded once through pinned memory: a pageable copy per group would
        #                    block the host until the previous group's kernels have finished
        for t, count in enumerate(counts):
            if count == 0:
                continue
            rows = order[start:start + count]
            start += count
            sub = {k: (np.asarray(v)[rows] if np.ndim(v) > 0 else v) for k, v in params.items()}
            table = self.tables[t]
            if hasattr(table, 'tabcorr_list'):
                # an Interpolator: no synchronisation per table, the out-of-range flags of all
                # groups are tested once at the end, so the host prepares group t + 1 while the
                # device evaluates group t
                result = table.predict_batch(sub, separate_gal_type=separate_gal_type,
                                             as_numpy=False, defer_range_check=True,
                                             **predict_kwargs)
                flags.append(result[2])
                result = result[:2]
            else:
                result = table.predict_batch(sub, separate_gal_type=separate_gal_type,
                                             as_numpy=False, **predict_kwargs)
            flat, spec = _flatten(result)
            if out is None:
                out = [torch.empty((n_draws,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
                       for x in flat]
                out_spec = spec
            elif spec != out_spec:
                raise ValueError('tables of a TableSet must produce identically shaped results')
            if order_dev is None:
                if flat[0].is_cuda:
                    pinned = torch.empty(n_draws, dtype=torch.int64, pin_memory=True)
                    pinned.numpy()[:] = order
                    order_dev = pinned.to(flat[0].device, non_blocking=True)
                else:
                    order_dev = torch.from_numpy(np.ascontiguousarray(order, dtype=np.int64))
            index = order_dev[start - count:start]
            for dst, src in zip(out, flat):
                dst.index_copy_(0, index, src)
        if out is None:
            raise ValueError('empty batch')
        if flags and int(torch.stack([f.reshape(()) for f in flags]).max().item()) != 0:
            raise ValueError('The x-coordinates are outside of the interpolation range and '
                             'extrapolation is turned off.')
        if as_numpy:
            out = [_to_host(x) for x in out]
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
