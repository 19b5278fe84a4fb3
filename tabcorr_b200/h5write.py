"""Dependency-free writer for the HDF5 subset TabCorr tables use (the inverse of ``h5mini``).

``TabCorr.write`` / ``Interpolator.write`` of the reference (``tabcorr/tabcorr.py:418-463``,
``tabcorr/interpolator.py:98-122``) go through h5py + astropy; neither exists in this image.  This
module emits the same "classic" structures h5py's defaults produce for such files, byte layout
mirrored from the shipped fixtures (``docs/examples/bolplanck_wp.hdf5``):

* superblock version 0, 8-byte offsets/lengths, group leaf K = 4, internal K = 16,
* version-1 object headers (one chunk, no continuation),
* old-style groups: symbol-table message -> v1 B-tree ``TREE`` (up to two levels) -> ``SNOD``
  nodes -> local ``HEAP`` with the link names,
* contiguous datasets with dataspace v1, fill-value v2 and layout v3 messages,
* datatypes: little-endian integers and IEEE floats, fixed-length strings, compound v1,
  variable-length UTF-8 strings (attributes; stored in one global heap collection ``GCOL``),
* attribute messages version 1.

It offers the small part of h5py's write API that the reference's ``write`` methods use
(``File(fname, 'w'|'w-')``, ``group[path] = value``, ``create_group``, ``.attrs[key] = value``).
Files are validated by reading them back with ``h5mini`` (tests/test_h5write.py); libhdf5 is not
available here to cross-check.  When ``h5py`` is importable and the target is an ``h5py.Group``
the ``write_*`` helpers use it directly.
"""

import os
import struct

import numpy as np

from . import h5mini
from .table import Table

try:
    import h5py as _h5py
except ImportError:  # pragma: no cover
    _h5py = None

_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 4        # symbols per SNOD <= 2 * _LEAF_K
_INTERNAL_K = 16   # children per TREE node <= 2 * _INTERNAL_K
_GCOL_SIZE = 4096


def _pad8(n):
    return (n + 7) & ~7


# ---------------------------------------------------------------------------------------------
# datatype / dataspace messages
# ---------------------------------------------------------------------------------------------
def _dtype_message(dtype):
    """Datatype message body (unpadded) for a numpy dtype."""
    dtype = np.dtype(dtype)
    if dtype.kind in 'iu':
        bits0 = 0x08 if dtype.kind == 'i' else 0x00
        return struct.pack('<BBBBIHH', 0x10, bits0, 0, 0, dtype.itemsize, 0, 8 * dtype.itemsize)
    if dtype.kind == 'b':
        return _dtype_message(np.int8)
    if dtype.kind == 'f':
        if dtype.itemsize == 8:
            return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
        if dtype.itemsize == 4:
            return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        if dtype.itemsize == 2:
            return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 15, 0, 2, 0, 16, 10, 5, 0, 10, 15)
        raise TypeError('unsupported float size {}'.format(dtype.itemsize))
    if dtype.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x01, 0, 0, dtype.itemsize)  # null-padded ASCII
    if dtype.kind == 'V' and dtype.names:
        out = struct.pack('<BBBBI', 0x16, len(dtype.names) & 0xFF, len(dtype.names) >> 8, 0,
                          dtype.itemsize)
        for name in dtype.names:
            member, offset = dtype.fields[name][0], dtype.fields[name][1]
            if member.shape:
                raise TypeError('array members of compound types are not supported')
            raw = name.encode('utf-8') + b'\x00'
            out += raw.ljust(_pad8(len(raw)), b'\x00')
            out += struct.pack('<IB3xII16x', offset, 0, 0, 0)
            out += _dtype_message(member)
        return out
    raise TypeError('cannot store dtype {} in the HDF5 subset'.format(dtype))


_VLEN_STR_DTYPE = (struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) +       # vlen string, UTF-8
                   struct.pack('<BBBBIHH', 0x10, 0, 0, 0, 1, 0, 8))       # base: 1-byte integer


def _dataspace_message(shape):
    if shape == ():
        return struct.pack('<BBB5x', 1, 0, 0)
    body = struct.pack('<BBB5x', 1, len(shape), 1)
    dims = b''.join(struct.pack('<Q', int(n)) for n in shape)
    return body + dims + dims  # current and maximum dimensions


def _message(mtype, body, flags=0):
    body = body.ljust(_pad8(len(body)), b'\x00')
    if len(body) > 0xFFFF:
        raise ValueError('object header message too large ({} bytes)'.format(len(body)))
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _normalise(value):
    """Python/numpy value -> (kind, array) where kind is 'vlen_str' or 'array'."""
    if isinstance(value, str):
        return 'vlen_str', value
    if isinstance(value, (bytes, np.bytes_)):
        raw = bytes(value)
        return 'array', np.array(raw, dtype='S{}'.format(max(len(raw), 1)))
    if isinstance(value, Table):
        value = value.as_array()
    array = np.asarray(value)
    if array.dtype.kind == 'U':
        if array.shape == ():
            return 'vlen_str', str(array)
        width = max(1, max(len(s.encode('utf-8')) for s in array.ravel()))
        array = np.char.encode(array, 'utf-8').astype('S{}'.format(width))
    if array.dtype.kind == 'O':
        raise TypeError('object arrays cannot be stored')
    if array.dtype.kind == 'V' and array.dtype.names:
        # unicode members (decoded gal_type column) go back to fixed-length bytes
        fields = []
        for name in array.dtype.names:
            member = array.dtype.fields[name][0]
            if member.kind == 'U':
                width = max(1, max([len(s.encode('utf-8')) for s in array[name].ravel()] + [1]))
                fields.append((name, 'S{}'.format(width)))
            else:
                fields.append((name, member.newbyteorder('<') if member.kind in 'iuf' else member))
        packed = np.empty(array.shape, dtype=np.dtype(fields))
        for name in array.dtype.names:
            column = array[name]
            packed[name] = np.char.encode(column, 'utf-8') if column.dtype.kind == 'U' else column
        return 'array', packed
    if array.dtype.kind in 'iuf':
        array = array.astype(array.dtype.newbyteorder('<'), copy=False)
    if array.dtype.kind == 'i' and array.dtype.itemsize < 8 and isinstance(value, (int, tuple, list)):
        array = array.astype('<i8')
    return 'array', np.ascontiguousarray(array).reshape(array.shape)  # keeps 0-d scalars 0-d


# ---------------------------------------------------------------------------------------------
# in-memory tree
# ---------------------------------------------------------------------------------------------
class _AttributeDict(dict):
    def __setitem__(self, key, value):
        _normalise(value)  # fail early on unsupported types
        super().__setitem__(str(key), value)


class _Node:
    def __init__(self):
        self.attrs = _AttributeDict()


class Dataset(_Node):
    def __init__(self, value):
        super().__init__()
        kind, array = _normalise(value)
        if kind == 'vlen_str':
            raw = array.encode('utf-8')
            array = np.array(raw, dtype='S{}'.format(max(len(raw), 1)))
        self.array = array

    @property
    def shape(self):
        return self.array.shape

    def __getitem__(self, key):
        return self.array[key]


class Group(_Node):
    """Writable group: ``grp['a/b'] = array`` creates intermediate groups like h5py does."""

    def __init__(self):
        super().__init__()
        self.children = {}

    def create_group(self, name):
        parts = [p for p in str(name).split('/') if p]
        node = self
        for i, part in enumerate(parts):
            if part in node.children:
                if i == len(parts) - 1 or not isinstance(node.children[part], Group):
                    raise ValueError("Unable to create group (name '{}' already exists)".format(name))
            else:
                node.children[part] = Group()
            node = node.children[part]
        return node

    def require_group(self, name):
        node = self
        for part in [p for p in str(name).split('/') if p]:
            node = node.children.setdefault(part, Group())
            if not isinstance(node, Group):
                raise TypeError("'{}' is not a group".format(name))
        return node

    def __setitem__(self, key, value):
        parts = [p for p in str(key).split('/') if p]
        parent = self.require_group('/'.join(parts[:-1])) if len(parts) > 1 else self
        if parts[-1] in parent.children:
            raise OSError("Unable to create link (name '{}' already exists)".format(key))
        parent.children[parts[-1]] = value if isinstance(value, _Node) else Dataset(value)

    def __getitem__(self, key):
        node = self
        for part in [p for p in str(key).split('/') if p]:
            if not isinstance(node, Group) or part not in node.children:
                raise KeyError("Unable to open object '{}'".format(key))
            node = node.children[part]
        return node

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def keys(self):
        return sorted(self.children)


# ---------------------------------------------------------------------------------------------
# serialisation
# ---------------------------------------------------------------------------------------------
class _Image:
    def __init__(self):
        self.buf = bytearray(96)   # superblock + root symbol-table entry, filled in last
        self.gcol_addr = None
        self.gcol_objects = []     # vlen string payloads

    def alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    def reserve(self, size):
        return self.alloc(bytes(size))

    def vlen_ref(self, text):
        raw = text.encode('utf-8')
        if self.gcol_addr is None:
            self.gcol_addr = self.reserve(_GCOL_SIZE)
        self.gcol_objects.append(raw)
        used = 16 + sum(16 + _pad8(len(o)) for o in self.gcol_objects)
        if used + 16 > _GCOL_SIZE:
            raise ValueError('too many variable-length string attributes for one heap collection')
        return struct.pack('<IQI', len(raw), self.gcol_addr, len(self.gcol_objects))

    def finish_gcol(self):
        if self.gcol_addr is None:
            return
        out = b'GCOL' + struct.pack('<B3xQ', 1, _GCOL_SIZE)
        for index, raw in enumerate(self.gcol_objects, start=1):
            out += struct.pack('<HH4xQ', index, 0, len(raw)) + raw.ljust(_pad8(len(raw)), b'\x00')
        free = _GCOL_SIZE - len(out)
        out += struct.pack('<HH4xQ', 0, 0, free)
        self.buf[self.gcol_addr:self.gcol_addr + len(out)] = out


def _attribute_messages(image, attrs):
    out = []
    for name in attrs:  # creation order, like h5py on files without attribute tracking
        kind, value = _normalise(attrs[name])
        if kind == 'vlen_str':
            dtype_msg, shape, data = _VLEN_STR_DTYPE, (), image.vlen_ref(value)
        else:
            dtype_msg, shape, data = _dtype_message(value.dtype), value.shape, value.tobytes()
        space_msg = _dataspace_message(shape)
        raw_name = name.encode('utf-8') + b'\x00'
        body = struct.pack('<BxHHH', 1, len(raw_name), len(dtype_msg), len(space_msg))
        body += raw_name.ljust(_pad8(len(raw_name)), b'\x00')
        body += dtype_msg.ljust(_pad8(len(dtype_msg)), b'\x00')
        body += space_msg.ljust(_pad8(len(space_msg)), b'\x00')
        body += data
        out.append(_message(0x000C, body, flags=4))
    return out


def _object_header(image, messages):
    payload = b''.join(messages)
    header = struct.pack('<BxHII4x', 1, len(messages), 1, len(payload))
    return image.alloc(header + payload)


def _write_dataset(image, node):
    array = node.array
    raw = array.tobytes()
    data_addr = image.alloc(raw) if raw else _UNDEF
    messages = [
        _message(0x0001, _dataspace_message(array.shape)),
        _message(0x0003, _dtype_message(array.dtype), flags=1),
        _message(0x0005, struct.pack('<BBBB', 2, 2, 2, 1) + struct.pack('<I', 0), flags=1),
        _message(0x0008, struct.pack('<BBQQ', 3, 1, data_addr, len(raw))),
    ]
    messages += _attribute_messages(image, node.attrs)
    return _object_header(image, messages)


def _write_group(image, node):
    """Serialise ``node`` (children first); returns (header, btree, heap) addresses."""
    entries = []
    for name in node.children:
        child = node.children[name]
        if isinstance(child, Group):
            header, btree, heap = _write_group(image, child)
            entries.append((name.encode('utf-8'), header, 1, struct.pack('<QQ', btree, heap)))
        else:
            entries.append((name.encode('utf-8'), _write_dataset(image, child), 0, bytes(16)))
    entries.sort(key=lambda e: e[0])

    # local heap: offset 0 holds the empty string the first B-tree key points at
    heap_data = bytearray(8)
    name_offset = []
    for raw_name, _, _, _ in entries:
        name_offset.append(len(heap_data))
        heap_data += (raw_name + b'\x00').ljust(_pad8(len(raw_name) + 1), b'\x00')
    heap_data_addr = image.alloc(bytes(heap_data))
    heap_addr = image.alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), 1, heap_data_addr))

    # symbol-table nodes
    per_snod = 2 * _LEAF_K
    snods = []  # (address, heap offset of the last name)
    for lo in range(0, len(entries), per_snod):
        chunk = entries[lo:lo + per_snod]
        body = b'SNOD' + struct.pack('<BxH', 1, len(chunk))
        for j, (_, header, cache, scratch) in enumerate(chunk):
            body += struct.pack('<QQII', name_offset[lo + j], header, cache, 0) + scratch
        body = body.ljust(8 + per_snod * 40, b'\x00')
        snods.append((image.alloc(body), name_offset[lo + len(chunk) - 1]))

    def tree_level(children, level):
        """Write the TREE nodes of one level over ``children`` [(address, last key)]."""
        per_node = 2 * _INTERNAL_K
        size = 24 + (2 * per_node + 1) * 8
        groups = [children[lo:lo + per_node] for lo in range(0, len(children), per_node)] or [[]]
        addrs = [image.reserve(size) for _ in groups]
        first_key = 0
        nodes = []
        for i, members in enumerate(groups):
            body = b'TREE' + struct.pack('<BBHQQ', 0, level, len(members),
                                         addrs[i - 1] if i > 0 else _UNDEF,
                                         addrs[i + 1] if i + 1 < len(groups) else _UNDEF)
            body += struct.pack('<Q', first_key)
            for address, last_key in members:
                body += struct.pack('<QQ', address, last_key)
                first_key = last_key
            image.buf[addrs[i]:addrs[i] + len(body)] = body
            nodes.append((addrs[i], first_key))
        return nodes

    nodes = tree_level(snods, 0)
    level = 0
    while len(nodes) > 1:
        level += 1
        nodes = tree_level(nodes, level)
    btree_addr = nodes[0][0]

    messages = [_message(0x0011, struct.pack('<QQ', btree_addr, heap_addr))]
    messages += _attribute_messages(image, node.attrs)
    return _object_header(image, messages), btree_addr, heap_addr


def serialise(root):
    """Bytes of an HDF5 file holding the tree under ``root`` (a :class:`Group`)."""
    image = _Image()
    header, btree, heap = _write_group(image, root)
    image.finish_gcol()
    while len(image.buf) % 8:
        image.buf.append(0)
    superblock = h5mini._SIGNATURE + struct.pack(
        '<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
    superblock += struct.pack('<QQQQ', 0, _UNDEF, len(image.buf), _UNDEF)
    superblock += struct.pack('<QQII', 0, header, 1, 0) + struct.pack('<QQ', btree, heap)
    assert len(superblock) == 96
    image.buf[:96] = superblock
    return bytes(image.buf)


class File(Group):
    """``h5write.File(fname, 'w' | 'w-')``: the tree is written when the file is closed."""

    def __init__(self, fname, mode='w-'):
        super().__init__()
        if mode not in ('w', 'w-', 'x'):
            raise ValueError("h5write.File only creates files (mode 'w' or 'w-')")
        self.filename = os.fspath(fname)
        if mode != 'w' and os.path.exists(self.filename):
            raise FileExistsError("Unable to create file (file exists): '{}'".format(self.filename))
        self._open = True

    def close(self):
        if self._open:
            with open(self.filename, 'wb') as f:
                f.write(serialise(self))
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *exc):
        if exc_type is None:
            self.close()
        return False


# ---------------------------------------------------------------------------------------------
# TabCorr / Interpolator layouts
# ---------------------------------------------------------------------------------------------
_ATTR_KEYS = ['tpcf', 'mode', 'simname', 'redshift', 'Num_ptcl_requirement', 'prim_haloprop_key',
              'sec_haloprop_key']


def _fill_tabcorr(halotab, fstream, max_args_size, matrix_dtype):
    """Body of ``TabCorr.write`` (``tabcorr/tabcorr.py:438-463``) on an open group."""
    for key in _ATTR_KEYS:
        fstream.attrs[key] = halotab.attrs[key]
    fstream['tpcf_matrix'] = halotab.tpcf_matrix.astype(matrix_dtype)
    for i, arg in enumerate(halotab.tpcf_args):
        if type(arg) is not np.ndarray or np.prod(arg.shape) < max_args_size:
            fstream['tpcf_args/arg_%d' % i] = arg
    for key in halotab.tpcf_kwargs:
        value = halotab.tpcf_kwargs[key]
        if type(value) is not np.ndarray or np.prod(value.shape) < max_args_size:
            fstream['tpcf_kwargs/' + key] = value
    fstream['tpcf_shape'] = np.asarray(halotab.tpcf_shape, dtype=np.int64)
    gal_type = halotab.gal_type.as_array()
    if _h5py is not None and isinstance(fstream, _h5py.Group):
        kind, gal_type = _normalise(gal_type)
        del kind
    fstream['gal_type'] = gal_type


def write_tabcorr(halotab, fname, overwrite=False, max_args_size=1000000,
                  matrix_dtype=np.float32):
    """``TabCorr.write``: ``fname`` is a file name or an open (h5py or h5write) group."""
    group_types = (Group,) + ((_h5py.Group,) if _h5py is not None else ())
    if isinstance(fname, group_types):
        _fill_tabcorr(halotab, fname, max_args_size, matrix_dtype)
        return
    with File(fname, 'w' if overwrite else 'w-') as fstream:
        _fill_tabcorr(halotab, fstream, max_args_size, matrix_dtype)


def write_interpolator(interpolator, fname, overwrite=False, max_args_size=1000000,
                       matrix_dtype=np.float32):
    """``Interpolator.write`` (``tabcorr/interpolator.py:118-122``).  Unlike the reference, which
    drops them, ``max_args_size`` and ``matrix_dtype`` are honoured."""
    with File(fname, 'w' if overwrite else 'w-') as fstream:
        fstream['param_dict_table'] = interpolator.param_dict_table.as_array()
        for i in range(len(interpolator.param_dict_table)):
            write_tabcorr(interpolator.tabcorr_list[i],
                          fstream.create_group('tabcorr_{}'.format(i)),
                          max_args_size=max_args_size, matrix_dtype=matrix_dtype)
