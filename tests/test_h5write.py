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


# ---------------------------------------------------------------------------------------------
# region-by-region comparison of write(read(fixture)) with the h5py-written fixture
# ---------------------------------------------------------------------------------------------
_MESSAGE_NAMES = {0x0000: 'nil', 0x0001: 'dataspace', 0x0003: 'datatype', 0x0005: 'fill_value',
                  0x0008: 'layout', 0x000C: 'attribute', 0x0010: 'continuation',
                  0x0011: 'symbol_table', 0x0012: 'modification_time'}


def _regions(fname):
    """Decompose a classic-format HDF5 file into address-free regions:
    {(object path, region name): bytes}.  Addresses inside messages (data address of the layout
    message, B-tree / heap addresses of the symbol-table message, global-heap references of
    variable-length string attributes) are replaced by what they point to, so that two files
    with the same content but a different placement of their blocks compare equal."""
    f = h5mini.File(fname)
    r = f._r
    out = {('', 'superblock[0:24]'): bytes(r.buf[0:24]),
           ('', 'superblock.addresses'): bytes(r.buf[24:40]) + bytes(r.buf[48:56]),
           ('', 'superblock.eof == file size'): struct.pack('<?', r.u64(40) == len(r.raw))}

    def visit(obj, path):
        counts = {}
        for mtype, flags, body, size in obj._messages:
            name = _MESSAGE_NAMES.get(mtype, hex(mtype))
            data = bytes(r.buf[body:body + size])
            if mtype in (0x0000, 0x0010):      # padding / continuation: placement only
                counts[name] = counts.get(name, 0) + 1
                continue
            if mtype == 0x0012:
                counts[name] = counts.get(name, 0) + 1
                continue
            if mtype == 0x0008:                # layout v3 contiguous: version, class, address, size
                assert data[0] == 3 and data[1] == 1
                address, nbytes = struct.unpack_from('<QQ', data, 2)
                out[(path, 'layout')] = data[:2] + struct.pack('<Q', nbytes)
                out[(path, 'raw data')] = bytes(r.buf[address + r.base:address + r.base + nbytes])
                continue
            if mtype == 0x0011:
                btree, heap = struct.unpack_from('<QQ', data, 0)
                links = r.group_links(btree + r.base, heap + r.base)
                out[(path, 'links in B-tree order')] = '\x00'.join(links).encode()
                continue
            if mtype == 0x000C:
                version = data[0]
                name_size, dt_size, ds_size = struct.unpack_from('<HHH', data, 2)
                attr = data[8:8 + name_size].rstrip(b'\x00').decode()
                p = 8 + (name_size + 7) // 8 * 8
                dtype = data[p:p + dt_size]
                p += (dt_size + 7) // 8 * 8
                space = data[p:p + ds_size]
                p += (ds_size + 7) // 8 * 8
                value = data[p:]
                if dtype[0] & 0x0f == 9:       # vlen: length, global heap address, index
                    length, gaddr, gidx = struct.unpack_from('<IQI', value, 0)
                    value = struct.pack('<I', length) + r.global_heap_object(gaddr + r.base, gidx)
                out[(path, 'attribute {!r}: header'.format(attr))] = bytes([version]) + data[2:8]
                out[(path, 'attribute {!r}: datatype'.format(attr))] = dtype
                out[(path, 'attribute {!r}: dataspace'.format(attr))] = space
                out[(path, 'attribute {!r}: value'.format(attr))] = value
                continue
            out[(path, name)] = bytes([flags]) + data
        for name, count in counts.items():
            out[(path, 'count of {} messages'.format(name))] = struct.pack('<I', count)
        out[(path, 'object header version')] = bytes(r.buf[obj._addr:obj._addr + 2])
        if isinstance(obj, h5mini.Group):
            for key in obj.keys():
                visit(obj[key], (path + '/' + key) if path else key)

    visit(f, '')
    return out


@pytest.mark.parametrize('name', ['bolplanck_wp.hdf5', 'bolplanck_ds.hdf5', 'ds_efficient.hdf5'])
def test_rewritten_fixture_equals_the_h5py_file_region_by_region(tmp_path, golden_dir, name):
    """``write(read(fixture))`` against the file h5py + astropy wrote (the reference's
    ``TabCorr.write``, tabcorr/tabcorr.py:418-463, ``Interpolator.write``,
    tabcorr/interpolator.py:98-122): every dataspace, datatype, fill-value and layout message,
    every attribute (header, datatype, dataspace, value), every group's links and every
    dataset's raw bytes are IDENTICAL; the complete list of differences is enumerated below and
    is placement/bookkeeping only (libhdf5 is absent here, so this is the strongest evidence
    available that libhdf5 will read these files as it reads its own)."""
    source = os.path.join(golden_dir, name)
    if name == 'ds_efficient.hdf5':
        obj = tabcorr_b200.Interpolator.read(source)
    else:
        obj = tabcorr_b200.TabCorr.read(source, upload=False)
    out = tmp_path / name
    obj.write(out)
    ref, new = _regions(source), _regions(out)
    if name == 'bolplanck_ds.hdf5':
        # the fixture holds arg_1, arg_2 (arg_0, the particle positions, fell under the
        # max_args_size rule, tabcorr.py:450-453); read() collects the existing keys in order
        # (:401-404) and write() enumerates them from 0 again (:450) -- in the reference too
        renamed = {}
        for (path, region), value in ref.items():
            path = {'tpcf_args/arg_1': 'tpcf_args/arg_0', 'tpcf_args/arg_2': 'tpcf_args/arg_1'}.get(
                path, path)
            if (path, region) == ('tpcf_args', 'links in B-tree order'):
                value = b'arg_0\x00arg_1'
            renamed[(path, region)] = value
        ref = renamed

    differences = set()
    for key in sorted(set(ref) | set(new)):
        if ref.get(key) != new.get(key):
            differences.add(key)
    # what may differ, and why:
    allowed = set()
    for path, region in differences:
        # libhdf5 stamps every dataset with a modification time and pads object headers with
        # NIL messages / continuation blocks; the writer emits neither (all optional)
        if region.startswith('count of '):
            allowed.add((path, region))
        # h5py stored 'simname' as a fixed-length string (numpy bytes_ from astropy metadata);
        # the writer stores every str attribute as a variable-length string like 'tpcf'/'mode'
        # (both decode to the same str in h5py and in TabCorr.read, tabcorr.py:395-397)
        if region in ("attribute 'simname': datatype", "attribute 'simname': value",
                      "attribute 'simname': header"):
            allowed.add((path, region))
        # the legacy root attribute 'spline' of old Interpolator files is not read by
        # Interpolator.read (interpolator.py:72-96) and not written by Interpolator.write
        if path == '' and region.startswith("attribute 'spline'"):
            allowed.add((path, region))
        # base/free-space/driver addresses: the fixture has a non-trivial free-space state
        if (path, region) == ('', 'superblock.addresses'):
            allowed.add((path, region))
    assert differences <= allowed, sorted(differences - allowed)
    # and the regions that carry the content are all there
    assert sum(1 for _, region in ref if region == 'raw data') == \
        sum(1 for _, region in new if region == 'raw data') > 0
    for key, value in ref.items():
        if key[1] == 'raw data':
            assert new[key] == value, key
