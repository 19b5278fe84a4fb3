"""Import the UNMODIFIED reference source from /root/reference with stub third-party modules.

TEST INFRASTRUCTURE (see ``oracle/tabcorr_oracle.py``).  Works only where /root/reference exists
(the build container); nothing that runs on the GPU box may import this module.  It is used by
``oracle/make_golden.py`` to record golden input/output vectors and by the container-only tests
that compare the numpy restatement with the reference itself.

``import tabcorr`` fails in this image because h5py, astropy and halotools are not installed
(``tabcorr/tabcorr.py:3-17``).  ``predict``, ``mean_occupation``, ``Interpolator.__init__`` /
``predict`` and the spline helpers only touch numpy/scipy at run time, so empty stand-ins for the
missing modules are enough to execute the reference's own code on tables supplied as numpy
structured arrays.
"""

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('TABCORR_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'tabcorr', 'tabcorr.py'))


class _Column:
    """What ``gal_type['col']`` must offer to the reference: ``.data``, ``==``, ``len``, indexing."""

    def __init__(self, data):
        self.data = data

    def __eq__(self, other):
        return self.data == other

    def __len__(self):
        return len(self.data)

    def __getitem__(self, key):
        return self.data[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)


class StubTable:
    """Minimal stand-in for ``astropy.table.Table`` over a numpy structured array.

    Offers exactly what the reference touches: ``colnames``, ``len``, column access returning an
    object with ``.data``, column assignment, row iteration, ``copy``, ``sort(keys)``,
    ``as_array`` (``tabcorr/interpolator.py:37-65``, ``tabcorr/tabcorr.py:537-578,623,660-681``).
    String columns are decoded to ``str`` like astropy does for HDF5 byte strings.
    """

    def __init__(self, array):
        array = np.asarray(array)
        self._cols = {}
        for name in array.dtype.names:
            col = array[name]
            if col.dtype.kind == 'S':
                col = np.char.decode(col, 'utf-8')
            self._cols[name] = np.array(col)

    @property
    def colnames(self):
        return list(self._cols.keys())

    def __len__(self):
        return len(next(iter(self._cols.values())))

    def __getitem__(self, key):
        return _Column(self._cols[key])

    def __setitem__(self, key, values):
        self._cols[key] = np.asarray(values)

    def __iter__(self):
        # astropy rows convert to 0-d structured arrays, so that ``np.stack(table)`` is a 1-d
        # structured array and ``np.unique`` on it counts whole rows (interpolator.py:52-54).
        records = self.as_array()
        for i in range(len(self)):
            yield records[i]

    def copy(self):
        out = StubTable.__new__(StubTable)
        out._cols = {k: v.copy() for k, v in self._cols.items()}
        return out

    def sort(self, keys):
        if isinstance(keys, str):
            keys = [keys]
        order = np.lexsort([self._cols[k] for k in keys[::-1]])
        for k in self._cols:
            self._cols[k] = self._cols[k][order]

    def as_array(self):
        dtype = [(k, v.dtype) for k, v in self._cols.items()]
        out = np.zeros(len(self), dtype=dtype)
        for k, v in self._cols.items():
            out[k] = v
        return out


_reference = None


def load():
    """Return the reference's ``tabcorr`` package (modules ``tabcorr.tabcorr``, ``.interpolator``)."""
    global _reference
    if _reference is not None:
        return _reference
    if not available():
        raise RuntimeError('the reference checkout is not present at ' + REFERENCE_ROOT)

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        for key, value in attrs.items():
            setattr(mod, key, value)
        sys.modules[name] = mod
        return mod

    placeholder = type('placeholder', (), {})
    if 'h5py' not in sys.modules:
        stub('h5py', Group=placeholder, File=placeholder)
    stub('astropy')
    stub('astropy.table', Table=StubTable, vstack=None)
    stub('astropy.units')
    sys.modules['astropy'].units = sys.modules['astropy.units']
    stub('astropy.cosmology', Flatw0waCDM=placeholder, FlatwCDM=placeholder, Planck15=None,
         Parameter=lambda *a, **k: None)
    stub('halotools')
    stub('halotools.sim_manager', sim_defaults=types.SimpleNamespace(Num_ptcl_requirement=300))
    stub('halotools.empirical_models', HodModelFactory=placeholder,
         model_defaults=types.SimpleNamespace(prim_haloprop_key='halo_mvir',
                                              sec_haloprop_key='halo_nfw_conc'),
         TrivialPhaseSpace=placeholder, Zheng07Cens=placeholder, NFWPhaseSpace=placeholder,
         Zheng07Sats=placeholder)
    stub('halotools.mock_observables', return_xyz_formatted_array=None)
    stub('halotools.utils', crossmatch=None)
    stub('halotools.utils.table_utils', compute_conditional_percentiles=None)

    saved = sys.modules.pop('tabcorr', None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import tabcorr as reference  # noqa: the reference package itself
    finally:
        sys.path.remove(REFERENCE_ROOT)
    del saved
    _reference = reference
    return reference


def make_tabcorr(gal_type, tpcf_matrix, tpcf_shape, attrs):
    """Build a reference ``TabCorr`` instance by hand from plain arrays (bypasses h5py)."""
    reference = load()
    halotab = reference.TabCorr()
    halotab.attrs = dict(attrs)
    halotab.tpcf_matrix = np.asarray(tpcf_matrix).astype(np.float64)  # tabcorr/tabcorr.py:399
    halotab.tpcf_shape = tuple(int(s) for s in tpcf_shape)
    halotab.tpcf_args = ()
    halotab.tpcf_kwargs = {}
    halotab.gal_type = StubTable(gal_type)
    return halotab


def make_interpolator(tabcorr_list, param_table):
    """``param_table``: dict ``{key: values[T]}`` in column order."""
    reference = load()
    dtype = [(k, np.float64) for k in param_table]
    arr = np.zeros(len(tabcorr_list), dtype=dtype)
    for k, v in param_table.items():
        arr[k] = v
    return reference.Interpolator(tabcorr_list, StubTable(arr))
