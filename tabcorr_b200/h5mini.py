"""Dependency-free reader for the HDF5 subset TabCorr tables are written in.

TabCorr stores its tables with h5py + astropy (reference ``tabcorr/tabcorr.py:374-463``,
``tabcorr/interpolator.py:72-122``).  Neither package (nor libhdf5) is available in this image,
so the table-loading half of the hot path (``TabCorr.read``, ``Interpolator.read``,
``database.read``) parses the files itself.  Only the "classic" on-disk structures that h5py's
default settings emit for such files are supported:

* superblock version 0 (8-byte offsets and lengths),
* version-1 object headers including continuation blocks,
* old-style groups (symbol table message -> v1 B-tree ``TREE`` -> ``SNOD`` -> local ``HEAP``),
* contiguous (and compact) dataset layouts, no filters,
* datatypes: fixed-point, IEEE float, fixed-length string, compound (v1-v3), enum,
  variable-length string (through the global heap ``GCOL``),
* attribute messages version 1-3.

Anything else raises :class:`H5FormatError` with a message saying what was found, so that a file
written with newer defaults fails loudly instead of being misread.  When ``h5py`` is importable
the caller may of course use it instead; this module mimics the small part of its API that
``TabCorr.read`` needs (``File``/``Group`` mapping access, ``.attrs``, ``dataset[()]``).
"""

import struct

import numpy as np

__all__ = ['File', 'Group', 'Dataset', 'H5FormatError']

_SIGNATURE = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(OSError):
    """The file uses an HDF5 feature outside the supported subset."""


def _pad8(n):
    return (n + 7) & ~7


class _Reader:
    """Byte-level access to the whole file (tables are at most a few hundred MB)."""

    def __init__(self, buf):
        self.raw = bytes(buf)
        self.buf = memoryview(self.raw)
        if buf[:8] != _SIGNATURE:
            raise H5FormatError('not an HDF5 file (bad signature)')
        version = buf[8]
        if version != 0:
            raise H5FormatError(
                'unsupported HDF5 superblock version {} (only version 0, as written by '
                'h5py with default settings, is supported)'.format(version))
        if buf[13] != 8 or buf[14] != 8:
            raise H5FormatError('only 8-byte offsets/lengths are supported')
        self.base = struct.unpack_from('<Q', buf, 24)[0]
        # root group symbol table entry starts at byte 56
        self.root_header = struct.unpack_from('<Q', buf, 56 + 8)[0]
        self._gcol_cache = {}

    def u8(self, off):
        return self.buf[off]

    def u16(self, off):
        return struct.unpack_from('<H', self.buf, off)[0]

    def u32(self, off):
        return struct.unpack_from('<I', self.buf, off)[0]

    def u64(self, off):
        return struct.unpack_from('<Q', self.buf, off)[0]

    def cstr(self, off):
        end = self.raw.index(b"\x00", off)
        return bytes(self.buf[off:end]).decode('utf-8')

    # ------------------------------------------------------------------ object headers
    def messages(self, addr):
        """Return [(type, flags, body_offset, body_size)] of a version-1 object header."""
        if self.buf[addr:addr + 4] == b'OHDR':
            raise H5FormatError('version-2 object headers are not supported')
        version = self.u8(addr)
        if version != 1:
            raise H5FormatError('unsupported object header version {}'.format(version))
        n_messages = self.u16(addr + 2)
        header_size = self.u32(addr + 8)
        blocks = [(addr + 16, header_size)]
        out = []
        while blocks and len(out) < n_messages:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and len(out) < n_messages:
                mtype, msize, mflags = struct.unpack_from('<HHB', self.buf, off)
                body = off + 8
                if mtype == 0x0010:  # continuation
                    c_off, c_len = struct.unpack_from('<QQ', self.buf, body)
                    blocks.append((c_off + self.base, c_len))
                out.append((mtype, mflags, body, msize))
                off = body + msize
        return out

    # ------------------------------------------------------------------ groups
    def group_links(self, btree_addr, heap_addr):
        """Names -> object header addresses of an old-style group, in B-tree (name) order."""
        if self.buf[heap_addr:heap_addr + 4] != b'HEAP':
            raise H5FormatError('bad local heap signature')
        heap_data = self.u64(heap_addr + 24) + self.base
        links = {}
        self._walk_btree(btree_addr, heap_data, links)
        return links

    def _walk_btree(self, addr, heap_data, links):
        if self.buf[addr:addr + 4] != b'TREE':
            raise H5FormatError('bad B-tree node signature')
        node_type, level, entries = struct.unpack_from('<BBH', self.buf, addr + 4)
        if node_type != 0:
            raise H5FormatError('unexpected B-tree node type {}'.format(node_type))
        off = addr + 24
        for i in range(entries):
            child = self.u64(off + 8 + 16 * i) + self.base
            if level > 0:
                self._walk_btree(child, heap_data, links)
            else:
                self._read_snod(child, heap_data, links)

    def _read_snod(self, addr, heap_data, links):
        if self.buf[addr:addr + 4] != b'SNOD':
            raise H5FormatError('bad symbol table node signature')
        n_symbols = self.u16(addr + 6)
        off = addr + 8
        for i in range(n_symbols):
            name_off, header = struct.unpack_from('<QQ', self.buf, off + 40 * i)
            links[self.cstr(heap_data + name_off)] = header + self.base

    # ------------------------------------------------------------------ datatypes
    def parse_dtype(self, off):
        """Parse a datatype message at ``off``.  Returns (spec, size_of_message)."""
        class_and_version = self.u8(off)
        cls, version = class_and_version & 0x0F, class_and_version >> 4
        bits0, bits1, bits2 = self.buf[off + 1], self.buf[off + 2], self.buf[off + 3]
        size = self.u32(off + 4)
        p = off + 8
        if cls == 0:  # fixed point
            if bits0 & 1:
                raise H5FormatError('big-endian integers are not supported')
            signed = bool(bits0 & 0x08)
            return np.dtype('<{}{}'.format('i' if signed else 'u', size)), 8 + 4
        if cls == 1:  # floating point
            if bits0 & 1:
                raise H5FormatError('big-endian floats are not supported')
            return np.dtype('<f{}'.format(size)), 8 + 12
        if cls == 3:  # fixed-length string
            return np.dtype('S{}'.format(size)), 8
        if cls == 6:  # compound
            n_members = bits0 | (bits1 << 8)
            names, formats, offsets = [], [], []
            for _ in range(n_members):
                name = self.cstr(p)
                if version < 3:
                    p += _pad8(len(name.encode('utf-8')) + 1)
                else:
                    p += len(name.encode('utf-8')) + 1
                if version == 1:
                    m_off = self.u32(p)
                    p += 4 + 1 + 3 + 4 + 4 + 16  # offset, rank, reserved, perm, reserved, dims
                elif version == 2:
                    m_off = self.u32(p)
                    p += 4
                else:
                    n_bytes = 1
                    while size >= (1 << (8 * n_bytes)):
                        n_bytes += 1
                    m_off = int.from_bytes(self.buf[p:p + n_bytes], 'little')
                    p += n_bytes
                m_dtype, m_len = self.parse_dtype(p)
                p += m_len
                names.append(name)
                formats.append(m_dtype)
                offsets.append(m_off)
            if any(isinstance(f, tuple) for f in formats):
                raise H5FormatError('variable-length members of compound types are not supported')
            return np.dtype({'names': names, 'formats': formats, 'offsets': offsets,
                             'itemsize': size}), p - off
        if cls == 8:  # enum: parse the base type, expose values as the base integers
            base, base_len = self.parse_dtype(p)
            n_members = bits0 | (bits1 << 8)
            q = p + base_len
            for _ in range(n_members):
                name = self.cstr(q)
                q += _pad8(len(name) + 1) if version < 3 else len(name) + 1
            q += n_members * base.itemsize
            return base, q - off
        if cls == 9:  # variable length
            kind = bits0 & 0x0F
            base, base_len = self.parse_dtype(p)
            if kind != 1:
                raise H5FormatError('variable-length sequences are not supported (only strings)')
            return ('vlen_str', size), 8 + base_len
        raise H5FormatError('unsupported HDF5 datatype class {}'.format(cls))

    def parse_dataspace(self, off):
        version = self.u8(off)
        rank = self.u8(off + 1)
        flags = self.u8(off + 2)
        if version == 1:
            p = off + 8
        elif version == 2:
            if self.u8(off + 3) == 2:  # null dataspace
                return None
            p = off + 4
        else:
            raise H5FormatError('unsupported dataspace version {}'.format(version))
        del flags
        return tuple(self.u64(p + 8 * i) for i in range(rank))

    def global_heap_object(self, collection_addr, index):
        addr = collection_addr + self.base
        if addr not in self._gcol_cache:
            if self.buf[addr:addr + 4] != b'GCOL':
                raise H5FormatError('bad global heap collection signature')
            size = self.u64(addr + 8)
            objects = {}
            p = addr + 16
            while p + 16 <= addr + size:
                idx = self.u16(p)
                if idx == 0:
                    break
                obj_size = self.u64(p + 8)
                objects[idx] = bytes(self.buf[p + 16:p + 16 + obj_size])
                p += 16 + _pad8(obj_size)
            self._gcol_cache[addr] = objects
        return self._gcol_cache[addr][index]

    def decode(self, dtype, shape, raw):
        """Turn raw bytes into what h5py would hand back for ``dataset[()]``/``attrs[key]``."""
        if isinstance(dtype, tuple):  # variable-length strings
            n = int(np.prod(shape, dtype=np.int64)) if shape else 1
            out = []
            for i in range(n):
                length, coll, idx = struct.unpack_from('<IQI', raw, 16 * i)
                out.append('' if length == 0 else
                           self.global_heap_object(coll, idx)[:length].decode('utf-8'))
            if not shape:
                return out[0]
            return np.array(out, dtype=object).reshape(shape)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        arr = np.frombuffer(raw, dtype=dtype, count=n).copy()
        if not shape:
            value = arr[0]
            if dtype.kind == 'S':
                return bytes(value)
            return value
        return arr.reshape(shape)


class _Attrs(dict):
    """Attribute mapping; iteration order follows the object header, like h5py's creation order
    for files without attribute tracking is name order -- we sort by name to match."""


class _Object:
    def __init__(self, reader, addr, name):
        self._r = reader
        self._addr = addr
        self.name = name
        self._messages = reader.messages(addr)
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            found = {}
            r = self._r
            for mtype, _, body, _ in self._messages:
                if mtype != 0x000C:
                    continue
                version = r.u8(body)
                name_size, dt_size, ds_size = struct.unpack_from('<HHH', r.buf, body + 2)
                if version == 1:
                    p = body + 8
                    name = r.cstr(p)
                    p += _pad8(name_size)
                    dtype, _ = r.parse_dtype(p)
                    p += _pad8(dt_size)
                    shape = r.parse_dataspace(p)
                    p += _pad8(ds_size)
                elif version in (2, 3):
                    p = body + (9 if version == 3 else 8)
                    name = r.cstr(p)
                    p += name_size
                    dtype, _ = r.parse_dtype(p)
                    p += dt_size
                    shape = r.parse_dataspace(p)
                    p += ds_size
                else:
                    raise H5FormatError('unsupported attribute message version {}'.format(version))
                if shape is None:
                    found[name] = None
                    continue
                itemsize = 16 if isinstance(dtype, tuple) else dtype.itemsize
                n = int(np.prod(shape, dtype=np.int64)) if shape else 1
                raw = r.buf[p:p + n * itemsize]
                value = r.decode(dtype, shape, raw)
                if isinstance(value, bytes):
                    # h5py returns fixed-length string attributes as numpy bytes_; TabCorr only
                    # compares/prints them, so hand back str for uniformity with vlen strings.
                    value = value.decode('utf-8')
                found[name] = value
            self._attrs = _Attrs(sorted(found.items()))
        return self._attrs


class Dataset(_Object):
    """Read-only dataset.  ``ds[()]`` returns the full array (or a numpy scalar)."""

    def __init__(self, reader, addr, name):
        super().__init__(reader, addr, name)
        r = reader
        self.shape = None
        self._dtype = None
        self._data = None
        for mtype, _, body, size in self._messages:
            if mtype == 0x0001:
                self.shape = r.parse_dataspace(body)
            elif mtype == 0x0003:
                self._dtype, _ = r.parse_dtype(body)
            elif mtype == 0x000B:
                raise H5FormatError(
                    "dataset '{}' uses a filter pipeline (compression); not supported".format(name))
            elif mtype == 0x0008:
                version = r.u8(body)
                if version != 3:
                    raise H5FormatError('unsupported data layout version {}'.format(version))
                layout_class = r.u8(body + 1)
                if layout_class == 1:  # contiguous
                    address, nbytes = struct.unpack_from('<QQ', r.buf, body + 2)
                    self._data = (None if address == _UNDEF else address + r.base, nbytes)
                elif layout_class == 0:  # compact
                    nbytes = r.u16(body + 2)
                    self._data = (body + 4, nbytes)
                else:
                    raise H5FormatError(
                        "dataset '{}' is chunked; only contiguous/compact layouts are "
                        'supported'.format(name))
        if self._dtype is None or self._data is None:
            raise H5FormatError("object '{}' is not a readable dataset".format(name))

    @property
    def dtype(self):
        return np.dtype(object) if isinstance(self._dtype, tuple) else self._dtype

    def __getitem__(self, key):
        r = self._r
        address, nbytes = self._data
        shape = self.shape if self.shape is not None else ()
        if address is None:
            dtype = self.dtype
            value = np.zeros(shape, dtype=dtype)
        else:
            value = r.decode(self._dtype, shape, r.buf[address:address + nbytes])
        if key == () or key is Ellipsis:
            return value
        return value[key]

    def __len__(self):
        return self.shape[0]


class Group(_Object):
    """Read-only group with mapping access (``grp['a/b']``, ``in``, ``keys()``)."""

    def __init__(self, reader, addr, name):
        super().__init__(reader, addr, name)
        self._links = None
        for mtype, _, body, _ in self._messages:
            if mtype == 0x0011:
                btree, heap = struct.unpack_from('<QQ', reader.buf, body)
                self._links = reader.group_links(btree + reader.base, heap + reader.base)
            elif mtype in (0x0002, 0x0006):
                raise H5FormatError('new-style (link message / fractal heap) groups are not supported')
        if self._links is None:
            raise H5FormatError("object '{}' is not a group".format(name))
        self._links = dict(sorted(self._links.items()))

    def keys(self):
        return self._links.keys()

    def __iter__(self):
        return iter(self._links)

    def __len__(self):
        return len(self._links)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __getitem__(self, key):
        key = str(key)
        node = self
        parts = [p for p in key.split('/') if p]
        for i, part in enumerate(parts):
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError("Unable to open object '{}' (component not found)".format(key))
            addr = node._links[part]
            child_name = (node.name.rstrip('/') + '/' + part)
            node = _open_object(node._r, addr, child_name)
        return node


def _open_object(reader, addr, name):
    types = {m[0] for m in reader.messages(addr)}
    if 0x0011 in types:
        return Group(reader, addr, name)
    if 0x0008 in types:
        return Dataset(reader, addr, name)
    if types & {0x0002, 0x0006}:
        raise H5FormatError('new-style (link message / fractal heap) groups are not supported')
    raise H5FormatError("cannot determine the type of object '{}'".format(name))


class File(Group):
    """``h5mini.File(fname)`` -- read-only stand-in for ``h5py.File(fname, 'r')``."""

    def __init__(self, fname, mode='r'):
        if mode != 'r':
            raise ValueError("h5mini.File is read-only; use tabcorr_b200.h5write to create files")
        with open(fname, 'rb') as f:
            buf = f.read()
        self.filename = str(fname)
        reader = _Reader(buf)
        super().__init__(reader, reader.root_header + reader.base, '/')

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
