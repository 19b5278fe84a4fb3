"""Minimal column table used where the reference uses ``astropy.table.Table``.

The reference keeps ``TabCorr.gal_type`` and ``Interpolator.param_dict_table`` as astropy tables
(``tabcorr/tabcorr.py:414``, ``tabcorr/interpolator.py:37``).  astropy is not available in the
benchmark image, and the hot path only needs named columns, so this class offers the part of the
astropy interface that the reference and its users touch: ``colnames``, ``len``, column access
(``table['n_h'].data``), column assignment/removal, ``copy``, ``sort``, ``as_array``, row
iteration.  Byte-string columns are decoded to ``str`` the way astropy does when it reads HDF5,
which is what makes ``gal_type['gal_type'] == 'centrals'`` (``tabcorr/tabcorr.py:555``) work.
"""

import numpy as np


class Column(np.ndarray):
    """ndarray with the ``.data`` attribute astropy columns have."""

    def __new__(cls, values, name=None):
        obj = np.asarray(values).view(cls)
        obj.name = name
        return obj

    def __array_finalize__(self, obj):
        self.name = getattr(obj, 'name', None)

    @property
    def data(self):
        return self.view(np.ndarray)


class Table:
    def __init__(self, data=None, names=None):
        self._columns = {}
        if data is None:
            return
        if isinstance(data, Table):
            for name in data.colnames:
                self[name] = np.array(data[name].data)
        elif isinstance(data, dict):
            for name, values in data.items():
                self[name] = values
        elif isinstance(data, np.ndarray) and data.dtype.names is not None:
            for name in data.dtype.names:
                self[name] = data[name]
        else:
            if names is None:
                raise ValueError('names are required for a list of columns')
            for name, values in zip(names, data):
                self[name] = values

    # -- columns --------------------------------------------------------------------------
    @property
    def colnames(self):
        return list(self._columns.keys())

    def __contains__(self, name):
        return name in self._columns

    def __getitem__(self, key):
        if isinstance(key, str):
            return Column(self._columns[key], name=key)
        if isinstance(key, (int, np.integer)):
            return self.as_array()[key]
        out = Table()
        for name, values in self._columns.items():
            out._columns[name] = values[key].copy()
        return out

    def __setitem__(self, name, values):
        values = np.array(values)
        if values.dtype.kind == 'S':
            values = np.char.decode(values, 'utf-8')
        elif values.dtype.kind == 'O':
            values = values.astype(str)
        if values.ndim == 0:
            values = np.full(len(self), values)
        if self._columns and len(values) != len(self):
            raise ValueError('column {} has the wrong length'.format(name))
        self._columns[name] = values

    def remove_column(self, name):
        del self._columns[name]

    # -- rows -----------------------------------------------------------------------------
    def __len__(self):
        for values in self._columns.values():
            return len(values)
        return 0

    def __iter__(self):
        records = self.as_array()
        for i in range(len(self)):
            yield records[i]

    def as_array(self):
        out = np.zeros(len(self), dtype=[(k, v.dtype) for k, v in self._columns.items()])
        for name, values in self._columns.items():
            out[name] = values
        return out

    def copy(self):
        return Table(self)

    def sort(self, keys):
        if isinstance(keys, str):
            keys = [keys]
        order = np.lexsort([self._columns[k] for k in reversed(keys)])
        for name in self._columns:
            self._columns[name] = self._columns[name][order]

    def __repr__(self):
        return 'Table(rows={}, columns={})'.format(len(self), self.colnames)
