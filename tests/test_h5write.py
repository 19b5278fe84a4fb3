"""TabCorr.write / Interpolator.write (reference tabcorr/tabcorr.py:418-463,
tabcorr/interpolator.py:98-122): files written by the dependency-free writer are read back with
the dependency-free reader and must reproduce the shipped fixtures' content; structural details
are compared with what h5py wrote into those fixtures."""

import os
import struct

import numpy as np
import pytest

import tabcorr_b200
from tabcorr_b200 import h5mini, h5write


def same_table(a, b, matrix_exact=True):
    assert a.attrs.keys() == b.attrs.keys()
    for key in a.attrs:
        assert a.attrs[key] == b.attrs[key], key
        assert type(a.attrs[key]) is type(b.attrs[key]) or isinstance(a.attrs[key], str), key
    assert a.tpcf_shape == b.tpcf_shape
    assert a.tpcf_matrix.dtype == b.tpcf_matrix.dtype == np.float64
    if matrix_exact:
        assert np.array_equal(a.tpcf_matrix, b.tpcf_matrix)
    assert len(a.tpcf_args) == len(b.tpcf_args)
    for x, y in zip(a.tpcf_args, b.tpcf_args):
        assert np.array_equal(x, y)
    assert a.tpcf_kwargs.keys() == b.tpcf_kwargs.keys()
    for key in a.tpcf_kwargs:
        assert np.array_equal(a.tpcf_kwargs[key], b.tpcf_kwargs[key])
    assert a.gal_type.colnames == b.gal_type.colnames
    for name in a.gal_type.colnames:
        assert np.array_equal(a.gal_type[name].data, b.gal_type[name].data), name


@pytest.mark.parametrize('name', ['bolplanck_wp.hdf5', 'bolplanck_ds.hdf5'])
def test_tabcorr_round_trip(tmp_path, golden_dir, name):
    original = tabcorr_b200.TabCorr.read(os.path.join(golden_dir, name), upload=False)
    out = tmp_path / name
    original.write(out)
    again = tabcorr_b200.TabCorr.read(out, upload=False)
    # the fixture matrices are float32 on disk, so the default float32 write is lossless
    same_table(original, again)
    with pytest.raises(FileExistsError):
        original.write(out)
    original.write(out, overwrite=True)
    # on-disk dtypes follow the reference: float32 matrix, int64 shape, S-type gal_type names
    with h5mini.File(out) as f:
        assert f['tpcf_matrix'][()].dtype == np.float32
        assert f['tpcf_shape'][()].dtype == np.int64
        assert f['gal_type'][()].dtype['gal_type'].kind == 'S'
        assert list(f.keys()) == sorted(f.keys())
    original.write(out, overwrite=True, matrix_dtype=np.float64)
    with h5mini.File(out) as f:
        assert f['tpcf_matrix'][()].dtype == np.float64


def test_max_args_size(tmp_path, golden_dir):
    original = tabcorr_b200.TabCorr.read(os.path.join(golden_dir, 'bolplanck_wp.hdf5'),
                                         upload=False)
    out = tmp_path / 'small.hdf5'
    original.write(out, max_args_size=10)   # arg_0 has 20 entries: dropped (tabcorr.py:450-453)
    with h5mini.File(out) as f:
        assert 'tpcf_args' not in f
        assert 'tpcf_kwargs/pi_max' in f
    again = tabcorr_b200.TabCorr.read(out, upload=False)
    assert again.tpcf_args == ()


def test_interpolator_round_trip(tmp_path, golden_dir):
    original = tabcorr_b200.Interpolator.read(os.path.join(golden_dir, 'ds_efficient.hdf5'))
    out = tmp_path / 'ds.hdf5'
    original.write(out)
    again = tabcorr_b200.Interpolator.read(out)
    assert again.param_dict_table.colnames == original.param_dict_table.colnames
    for name in original.param_dict_table.colnames:
        assert np.array_equal(again.param_dict_table[name].data,
                              original.param_dict_table[name].data)
    assert len(again.tabcorr_list) == len(original.tabcorr_list)
    for a, b in zip(original.tabcorr_list, again.tabcorr_list):
        same_table(a, b)
    for xa, xb in zip(original.xp, again.xp):
        assert np.array_equal(xa, xb)


def test_many_links_use_a_two_level_btree(tmp_path):
    """More than 2 * 16 * 8 = 256 links in one group need an internal B-tree level."""
    out = tmp_path / 'many.hdf5'
    with h5write.File(out, 'w') as f:
        for i in range(700):
            f['grid/table_{}'.format(i)] = np.arange(3) + i
        f.create_group('empty')
        f.attrs['note'] = 'two levels'
        f.attrs['count'] = 700
        f.attrs['ratio'] = 0.25
        f.attrs['raw'] = b'bytes'
    with h5mini.File(out) as f:
        assert len(f['grid']) == 700
        for i in (0, 1, 255, 256, 699):
            assert np.array_equal(f['grid/table_{}'.format(i)][()], np.arange(3) + i)
        assert len(f['empty']) == 0
        assert f.attrs['note'] == 'two levels' and f.attrs['count'] == 700
        assert f.attrs['ratio'] == 0.25 and f.attrs['raw'] == 'bytes'
        # keys of the B-tree are in strcmp order
        assert list(f['grid'].keys()) == sorted('table_{}'.format(i) for i in range(700))


def test_structures_mirror_h5py_fixture(tmp_path, golden_dir):
    """Byte-level comparison with what h5py wrote: superblock fields, datatype and dataspace
    messages, attribute encoding."""
    original = tabcorr_b200.TabCorr.read(os.path.join(golden_dir, 'bolplanck_wp.hdf5'),
                                         upload=False)
    out = tmp_path / 'mirror.hdf5'
    original.write(out)
    ref = open(os.path.join(golden_dir, 'bolplanck_wp.hdf5'), 'rb').read()
    new = open(out, 'rb').read()
    assert new[:24] == ref[:24]                      # signature, versions, sizes, K values
    assert struct.unpack_from('<Q', new, 40)[0] == len(new)   # end-of-file address
    f_ref, f_new = h5mini.File(os.path.join(golden_dir, 'bolplanck_wp.hdf5')), h5mini.File(out)

    def messages(f, path, mtype):
        obj = f[path] if path else f
        return [bytes(obj._r.buf[body:body + size]) for t, _, body, size in obj._messages
                if t == mtype]

    for path in ('tpcf_matrix', 'tpcf_shape', 'gal_type', 'tpcf_args/arg_0',
                 'tpcf_kwargs/pi_max'):
        for mtype in (0x0001, 0x0003, 0x0005):       # dataspace, datatype, fill value
            assert messages(f_new, path, mtype) == messages(f_ref, path, mtype), (path, mtype)
    # attributes: same encoded messages except for the global-heap address of vlen strings and
    # 'simname', which h5py stored as a fixed-length string and we store as a vlen string
    ref_attrs = {m[8:8 + m[8:].index(b'\x00')]: m for m in messages(f_ref, '', 0x000C)}
    new_attrs = {m[8:8 + m[8:].index(b'\x00')]: m for m in messages(f_new, '', 0x000C)}
    assert ref_attrs.keys() == new_attrs.keys()
    for key in (b'redshift', b'Num_ptcl_requirement'):
        assert ref_attrs[key] == new_attrs[key]
    for key in (b'tpcf', b'mode', b'prim_haloprop_key', b'sec_haloprop_key'):
        assert len(ref_attrs[key]) == len(new_attrs[key])
        assert ref_attrs[key][:-12] == new_attrs[key][:-12]   # up to the heap address + index


def test_reader_rejects_damaged_files_cleanly(tmp_path, golden_dir):
    """Truncated or corrupted files raise (H5FormatError / a plain exception), they never hang or
    return silently wrong tables."""
    raw = open(os.path.join(golden_dir, 'bolplanck_ds.hdf5'), 'rb').read()
    bad = tmp_path / 'bad.hdf5'
    for cut in (0, 7, 95, 500, 2000):
        bad.write_bytes(raw[:cut])
        with pytest.raises(Exception):
            tabcorr_b200.TabCorr.read(bad, upload=False)
    bad.write_bytes(b'\x89HDF\r\n\x1a\n' + b'\x02' + raw[9:])      # superblock version 2
    with pytest.raises(h5mini.H5FormatError, match='superblock version'):
        tabcorr_b200.TabCorr.read(bad, upload=False)
    bad.write_bytes(b'not an hdf5 file at all' * 10)
    with pytest.raises(h5mini.H5FormatError, match='signature'):
        tabcorr_b200.TabCorr.read(bad, upload=False)
    with pytest.raises(OSError):
        tabcorr_b200.TabCorr.read(tmp_path / 'missing.hdf5', upload=False)
