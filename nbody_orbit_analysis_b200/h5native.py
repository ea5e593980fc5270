"""Native HDF5 writer / reader for the result files of the orbit-tracking path
(SURVEY.md section 8(f)-2): the layouts of ``track_orbits.py:366-397``,
``track_orbits_onthefly.py:208-252`` and ``postprocessing.py:146-162`` without
h5py / libhdf5 (neither exists in this image).

Scope -- what those layouts need and nothing more: groups, contiguous datasets of
fixed-size numeric types (int / uint 8-64, float 16/32/64), scalar or 1-D numeric
attributes and variable-length UTF-8 string attributes (``attrs['mode']`` must
come back as ``str``, ``postprocessing.py:20``).  The same ``File`` / ``Group`` /
``Dataset`` / ``attrs`` subset of the h5py API as ``h5shim`` (shared classes).

Format (HDF5 File Format Specification, version 0 superblock -- what h5py
writes by default): superblock v0, version-1 object headers, "old style" groups
(symbol-table message -> version-1 B-tree of symbol-table nodes + local heap),
dataspace v1, datatype v1 (classes 0, 1, 9), fill-value v2, contiguous layout
v3, attribute v1, one global heap collection per variable-length string.

Writing strategy: dataset payloads are appended when ``create_dataset`` is called
and never move; ALL metadata (object headers, B-trees, heaps) is written in one
block at ``close()`` and the superblock then points at the new root.  Re-opening
for append ('r+' / 'a') parses the existing tree, appends new payloads and
writes fresh headers for the objects that changed and their ancestors (a new
snapshot group re-writes the root group's name heap and B-tree: ~60 bytes per
snapshot already in the file; the replaced blocks become dead space).  A crash before ``close()`` leaves the previous, consistent
state of the file.

Validation status: written against the specification and read back by the
independent parser below; NOT validated against libhdf5 in this image (there is
none).  ``storage.py`` selects it with ``OA_STORAGE=hdf5``.
"""
import os
import struct

import numpy as np

from . import h5shim

SIG = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INT_K = 4, 16            # library defaults (H5Pset_sym_k / istore_k)
SNOD_ENTRIES = 2 * LEAF_K
NODE_CHILDREN = 2 * INT_K


def _pad8(n):
    return (n + 7) & ~7


# ---------------------------------------------------------------------------
# datatype / dataspace messages
# ---------------------------------------------------------------------------
def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in 'iu':
        bits0 = 0x08 if dt.kind == 'i' else 0x00          # little endian, signed?
        return struct.pack('<BBBBI', 0x10, bits0, 0, 0, dt.itemsize) + \
            struct.pack('<HH', 0, 8 * dt.itemsize)
    if dt.kind == 'f':
        size = dt.itemsize
        exp_bits, man_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
        bias = (1 << (exp_bits - 1)) - 1
        # byte 0: little endian, mantissa normalisation = implied leading 1 (2)
        # byte 1: position of the sign bit
        return struct.pack('<BBBBI', 0x11, 0x20, 8 * size - 1, 0, size) + \
            struct.pack('<HHBBBBI', 0, 8 * size, man_bits, exp_bits, 0, man_bits,
                        bias)
    if dt.kind == 'b':
        return _dtype_msg(np.uint8)
    raise TypeError("no HDF5 encoding for dtype %s in this writer" % dt)


_VLEN_STR_MSG = struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) + \
    struct.pack('<BBBBI', 0x13, 0x10, 0, 0, 1)


def _parse_dtype(buf):
    cls_ver, b0, b1, _b2, size = struct.unpack_from('<BBBBI', buf, 0)
    cls = cls_ver & 0x0F
    if cls == 0:
        if b0 & 0x01:
            raise OSError("big-endian integers are not supported")
        return np.dtype('%s%d' % ('i' if b0 & 0x08 else 'u', size))
    if cls == 1:
        if b0 & 0x01:
            raise OSError("big-endian floats are not supported")
        return np.dtype('f%d' % size)
    if cls == 9 and (b0 & 0x0F) == 1:
        return 'vlen_str'
    if cls == 3:
        return np.dtype('S%d' % size)
    raise OSError("datatype class %d is not supported by this reader" % cls)


def _dspace_msg(shape):
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(
        struct.pack('<Q', int(s)) for s in shape)


def _parse_dspace(buf):
    ver, rank, flags = struct.unpack_from('<BBB', buf, 0)
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
    else:
        raise OSError("dataspace version %d" % ver)
    return tuple(struct.unpack_from('<Q', buf, off + 8 * i)[0]
                 for i in range(rank))


def _message(mtype, data, flags=0):
    data = data + b'\x00' * (_pad8(len(data)) - len(data))
    return struct.pack('<HHB3x', mtype, len(data), flags) + data


def _object_header(messages):
    body = b''.join(messages)
    return struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body)) + body


# ---------------------------------------------------------------------------
# the store: same interface as h5shim._Store
# ---------------------------------------------------------------------------
class _LazyDict(dict):
    """path -> value; a lookup first parses the object's header if the file has
    it and it has not been read yet (objects are read on first access: opening a
    file with hundreds of snapshot groups costs one symbol table, not the whole
    tree)."""

    def __init__(self, store, *a):
        dict.__init__(self, *a)
        self._store = store

    def __contains__(self, path):
        self._store._ensure(path)
        return dict.__contains__(self, path)

    def __getitem__(self, path):
        self._store._ensure(path)
        return dict.__getitem__(self, path)

    def get(self, path, default=None):
        self._store._ensure(path)
        return dict.get(self, path, default)

    def setdefault(self, path, default=None):
        self._store._ensure(path)
        return dict.setdefault(self, path, default)


class _LazySet(set):

    def __init__(self, store, *a):
        set.__init__(self, *a)
        self._store = store

    def __contains__(self, path):
        self._store._ensure(path)
        return set.__contains__(self, path)


class _Store:

    def __init__(self, filename, mode):
        self.filename, self.mode = filename, mode
        exists = os.path.exists(filename)
        if mode in ('r', 'r+') and not exists:
            raise FileNotFoundError(
                "Unable to open file (no such file: %r)" % filename)
        if mode in ('w-', 'x') and exists:
            raise FileExistsError(
                "Unable to create file (file exists: %r)" % filename)
        if mode not in ('r', 'r+', 'w', 'w-', 'x', 'a'):
            raise ValueError("Invalid mode %r" % (mode,))
        self.writable = mode != 'r'
        fresh = mode in ('w', 'w-', 'x') or (mode == 'a' and not exists)
        self.fh = open(filename, 'w+b' if fresh else
                       ('rb' if mode == 'r' else 'r+b'))
        self.groups = _LazySet(self, {'/'})
        self.datasets = _LazyDict(self)   # path -> (dtype, shape, address, nbytes)
        self.attrs = _LazyDict(self, {'/': {}})
        self.dirty = fresh
        self.addr = {}          # path -> object header address (objects on disk)
        self.kids = {'/': set()}   # group -> names of its children
        self._unread = set()    # objects on disk whose header has not been parsed
        self._unlisted = {}     # group -> (B-tree, heap) of a symbol table not read yet
        self.changed = {'/'} if fresh else set()   # objects to (re)write at close
        if fresh:
            self.fh.write(b'\x00' * 96)         # superblock, written at close
            self.end = 96
        else:
            self._parse()

    # ---- reading -----------------------------------------------------------------
    def _at(self, addr, n):
        self.fh.seek(addr)
        buf = self.fh.read(n)
        if len(buf) != n:
            raise OSError("truncated HDF5 file")
        return buf

    def _parse(self):
        head = self._at(0, 96)
        if head[:8] != SIG:
            raise OSError("%r is not an HDF5 file" % self.filename)
        ver = head[8]
        if ver not in (0, 1):
            raise OSError("superblock version %d is not supported" % ver)
        so, sl = head[13], head[14]
        if (so, sl) != (8, 8):
            raise OSError("only 8-byte offsets / lengths are supported")
        self.leaf_k, self.int_k = struct.unpack_from('<HH', head, 16)
        off = 24 + (4 if ver == 1 else 0)
        base, _free, eof, _drv = struct.unpack_from('<QQQQ', head, off)
        if base != 0:
            raise OSError("non-zero base address")
        self.end = eof
        root = off + 32                       # root symbol table entry
        _name, ohdr = struct.unpack_from('<QQ', head, root)
        self._read_object('/', ohdr)

    def _messages(self, addr):
        """(type, data) of every message of a version-1 object header, following
        continuation blocks."""
        ver, _r, nmsg, _ref, size = struct.unpack('<BBHII', self._at(addr, 12))
        if ver != 1:
            raise OSError("object header version %d is not supported" % ver)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            buf = self._at(pos, left)
            p = 0
            while p + 8 <= left and len(out) < nmsg:
                mtype, msize, _fl = struct.unpack_from('<HHB', buf, p)
                data = buf[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                       # continuation
                    blocks.append(struct.unpack('<QQ', data[:16]))
                out.append((mtype, data))
        return out

    def _ensure(self, path):
        """Parse what the file holds about `path` (no-op for paths already read,
        created in this session, or absent): the symbol tables of its ancestors
        and its own object header."""
        if not self._unread and not self._unlisted:
            return
        if not isinstance(path, str) or not path.startswith('/'):
            return
        node = '/'
        for part in [q for q in path.split('/') if q]:
            self._list(node)
            node = h5shim._join(node, part)
            if node in self._unread:
                self._unread.discard(node)
                self._read_object(node, self.addr[node])
            elif not set.__contains__(self.groups, node):
                return                    # a dataset, or nothing: no deeper level

    def _list(self, group):
        """Read a group's symbol table: names and header addresses of its
        children (their headers stay unread until someone asks for them)."""
        where = self._unlisted.pop(group, None)
        if where is None:
            return
        btree, heap = where
        hd = self._heap(heap)
        names = self.kids.setdefault(group, set())
        for name_off, child in self._btree(btree):
            name = hd[name_off:hd.index(b'\x00', name_off)].decode('utf-8')
            cpath = h5shim._join(group, name)
            names.add(name)
            self.addr[cpath] = child
            self._unread.add(cpath)

    def _read_object(self, path, addr):
        self.addr[path] = addr
        msgs = self._messages(addr)
        kinds = {m for m, _ in msgs}
        attrs = dict.setdefault(self.attrs, path, {})
        for mtype, data in msgs:
            if mtype == 0x000C:
                k, v = self._parse_attr(data)
                attrs[k] = v
        if 0x0011 in kinds:                               # group
            set.add(self.groups, path)
            data = dict(msgs)[0x0011]
            self.kids.setdefault(path, set())
            self._unlisted[path] = struct.unpack_from('<QQ', data, 0)
        elif 0x0008 in kinds:                             # dataset
            d = dict(msgs)
            shape = _parse_dspace(d[0x0001])
            dtype = _parse_dtype(d[0x0003])
            lay = d[0x0008]
            if lay[0] != 3:
                raise OSError("data layout version %d is not supported" % lay[0])
            if lay[1] == 1:                               # contiguous
                address, nbytes = struct.unpack_from('<QQ', lay, 2)
                inline = None
            elif lay[1] == 0:                             # compact
                (nbytes,) = struct.unpack_from('<H', lay, 2)
                address, inline = None, bytes(lay[4:4 + nbytes])
            else:
                raise OSError("chunked datasets are not supported by this reader")
            dict.__setitem__(self.datasets, path, (dtype, shape, address, nbytes))
            if inline is not None:
                self._inline = getattr(self, '_inline', {})
                self._inline[path] = inline
        else:
            raise OSError("object at %d is neither an old-style group nor a "
                          "dataset" % addr)

    def _heap(self, addr):
        sig, _ver, size, _free, data = struct.unpack('<4sB3xQQQ', self._at(addr, 32))
        if sig != b'HEAP':
            raise OSError("bad local heap signature")
        return self._at(data, size)

    def _btree(self, addr):
        """(name offset, object header address) of every symbol below a node."""
        sig, ntype, level, used = struct.unpack('<4sBBH', self._at(addr, 8))
        if sig == b'SNOD':
            n = struct.unpack('<H', self._at(addr + 6, 2))[0]
            buf = self._at(addr + 8, 40 * n)
            return [struct.unpack_from('<QQ', buf, 40 * i) for i in range(n)]
        if sig != b'TREE' or ntype != 0:
            raise OSError("bad group B-tree node")
        buf = self._at(addr + 24, 16 * used + 8)
        out = []
        for i in range(used):
            (child,) = struct.unpack_from('<Q', buf, 16 * i + 8)
            out += self._btree(child)
        return out

    def _parse_attr(self, data):
        ver, _r, nlen, tlen, slen = struct.unpack_from('<BBHHH', data, 0)
        if ver != 1:
            raise OSError("attribute message version %d" % ver)
        p = 8
        name = data[p:p + nlen].split(b'\x00')[0].decode('utf-8')
        p += _pad8(nlen)
        dtype = _parse_dtype(data[p:p + tlen])
        p += _pad8(tlen)
        shape = _parse_dspace(data[p:p + slen])
        p += _pad8(slen)
        if isinstance(dtype, str):                        # variable-length string
            length, gcol, index = struct.unpack_from('<IQI', data, p)
            return name, self._global_heap_object(gcol, index)[:length].decode('utf-8')
        count = int(np.prod(shape, dtype=np.int64))
        val = np.frombuffer(data[p:p + count * dtype.itemsize], dtype=dtype).reshape(shape)
        if dtype.kind == 'S':
            return name, (val[()] if shape == () else val.copy())
        return name, (val[()] if shape == () else val.copy())

    def _global_heap_object(self, addr, index):
        sig, ver, size = struct.unpack('<4sB3xQ', self._at(addr, 16))
        if sig != b'GCOL' or ver != 1:
            raise OSError("bad global heap collection")
        buf = self._at(addr, size)
        p = 16
        while p + 16 <= size:
            idx, _ref, osize = struct.unpack_from('<HH4xQ', buf, p)
            if idx == 0:
                break
            if idx == index:
                return buf[p + 16:p + 16 + osize]
            p += 16 + _pad8(osize)
        raise OSError("global heap object %d not found" % index)

    # ---- the h5shim._Store interface -------------------------------------------------
    def exists(self, path):
        return path in self.groups or path in self.datasets

    def children(self, path):
        self._ensure(path)
        self._list(path)
        return sorted(self.kids.get(path, ()))

    def _writable(self, path):
        """`path` changes: it and its ancestors get new object headers at close
        (everything else keeps the header it has on disk)."""
        if not self.writable:
            raise OSError("file is open read-only")
        self.dirty = True
        while True:
            self.changed.add(path)
            if path == '/':
                break
            path = h5shim._norm(os.path.dirname(path))

    def add_group(self, path):
        self._writable(path)
        parent = h5shim._norm(os.path.dirname(path))
        if parent != '/' and parent not in self.groups:
            self.add_group(parent)
        self._list(parent)
        self.kids.setdefault(parent, set()).add(os.path.basename(path))
        self.kids.setdefault(path, set())
        self.groups.add(path)
        self.attrs.setdefault(path, {})

    def add_dataset(self, path, arr):
        self._writable(path)
        parent = h5shim._norm(os.path.dirname(path))
        if parent not in self.groups:
            self.add_group(parent)
        arr = np.asarray(arr)
        if arr.ndim and not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)
        if arr.dtype.kind == 'b':
            arr = arr.astype(np.uint8)
        _dtype_msg(arr.dtype)                             # raises for unsupported types
        if arr.dtype.byteorder == '>':
            arr = arr.astype(arr.dtype.newbyteorder('<'))
        address = UNDEF
        if arr.nbytes:
            address = _pad8(self.end)
            self.fh.seek(address)
            # (the array's own buffer: no intermediate bytes object, and the
            # large write releases the GIL)
            self.fh.write(arr.reshape(-1).view(np.uint8))
            self.end = address + arr.nbytes
        self._list(parent)
        self.kids.setdefault(parent, set()).add(os.path.basename(path))
        self.datasets[path] = (arr.dtype, tuple(arr.shape), address, arr.nbytes)
        self.attrs.setdefault(path, {})

    def set_attr(self, path, key, value):
        self._writable(path)
        if isinstance(value, bytes):
            value = value.decode('utf-8')
        if not isinstance(value, str):
            arr = np.asarray(value)
            if arr.dtype.kind in 'OU':
                raise TypeError("unsupported attribute type %r" % (type(value),))
            _dtype_msg(arr.dtype)
            value = arr[()] if arr.shape == () else arr.copy()
        self.attrs[path][key] = value

    def read(self, path, start=0, stop=None):
        dtype, shape, address, nbytes = self.datasets[path]
        inline = getattr(self, '_inline', {}).get(path)

        def raw(off, n):
            if n == 0:
                return b''
            if inline is not None:
                return inline[off:off + n]
            return self._at(address + off, n)
        if len(shape) == 0:
            return np.frombuffer(raw(0, dtype.itemsize), dtype=dtype)[0]
        if stop is None:
            stop = shape[0]
        row = dtype.itemsize * int(np.prod(shape[1:], dtype=np.int64))
        count = max(stop - start, 0)
        if inline is None and count * row:
            # straight into the result array (one copy out of the page cache)
            out = np.empty((count,) + tuple(shape[1:]), dtype=dtype)
            self.fh.seek(address + start * row)
            view = out.reshape(-1).view(np.uint8)
            got = 0
            while got < view.size:
                k = self.fh.readinto(view[got:])
                if not k:
                    raise OSError("truncated HDF5 file")
                got += k
            return out
        return np.frombuffer(raw(start * row, count * row), dtype=dtype).reshape(
            (count,) + tuple(shape[1:])).copy()

    # ---- writing the metadata block ---------------------------------------------------
    def _emit(self, blob):
        addr = _pad8(self.end)
        self.fh.seek(addr)
        self.fh.write(blob)
        self.end = addr + len(blob)
        return addr

    def _attr_messages(self, path):
        out = []
        for key in sorted(self.attrs.get(path, {})):
            val = self.attrs[path][key]
            name = key.encode('utf-8') + b'\x00'
            if isinstance(val, str):
                raw = val.encode('utf-8')
                body = struct.pack('<HH4xQ', 1, 1, len(raw)) + raw + \
                    b'\x00' * (_pad8(len(raw)) - len(raw))
                used = 16 + len(body)
                size = max(4096, _pad8(used + 16))
                free = size - used
                gcol = self._emit(
                    struct.pack('<4sB3xQ', b'GCOL', 1, size) + body +
                    struct.pack('<HH4xQ', 0, 0, free) + b'\x00' * (free - 16))
                tmsg, smsg = _VLEN_STR_MSG, _dspace_msg(())
                data = struct.pack('<IQI', len(raw), gcol, 1)
            else:
                arr = np.asarray(val)            # (ascontiguousarray would add an axis)
                tmsg, smsg, data = _dtype_msg(arr.dtype), _dspace_msg(arr.shape), \
                    arr.tobytes()

            def padded(b):
                return b + b'\x00' * (_pad8(len(b)) - len(b))
            out.append(_message(0x000C, struct.pack(
                '<BBHHH', 1, 0, len(name), len(tmsg), len(smsg)) +
                padded(name) + padded(tmsg) + padded(smsg) + data))
        return out

    def _write_dataset(self, path):
        if path not in self.changed and path in self.addr:
            return self.addr[path]
        dtype, shape, address, nbytes = self.datasets[path]
        msgs = [_message(0x0001, _dspace_msg(shape)),
                _message(0x0003, _dtype_msg(dtype), flags=1),
                _message(0x0005, struct.pack('<BBBBI', 2, 2, 2, 1, 0)),
                _message(0x0008, struct.pack('<BBQQ', 3, 1, address, nbytes))]
        return self._emit(_object_header(msgs + self._attr_messages(path)))

    def _write_group(self, path):
        if path not in self.changed and path in self.addr:
            return self.addr[path]
        names = self.children(path)
        entries = []                        # (name, object header address, is group)
        for name in names:
            child = h5shim._join(path, name)
            if child not in self.changed and child in self.addr:
                entries.append((name, self.addr[child]))    # as it is on disk
            elif child in self.groups:
                entries.append((name, self._write_group(child)))
            else:
                entries.append((name, self._write_dataset(child)))
        entries.sort(key=lambda e: e[0].encode('utf-8'))
        # local heap: offset 0 holds the empty string (the leftmost B-tree key)
        heap = bytearray(8)
        offs = []
        for name, _ in entries:
            offs.append(len(heap))
            raw = name.encode('utf-8') + b'\x00'
            heap += raw + b'\x00' * (_pad8(len(raw)) - len(raw))
        data_addr = self._emit(bytes(heap))
        heap_addr = self._emit(struct.pack('<4sB3xQQQ', b'HEAP', 0, len(heap), 1,
                                           data_addr))
        # leaves: symbol table nodes of up to 2 LEAF_K entries
        level = []                          # (address, heap offset of the largest name)
        pairs = list(zip(offs, entries))
        for i in range(0, len(pairs), SNOD_ENTRIES):
            chunk = pairs[i:i + SNOD_ENTRIES]
            body = b''.join(struct.pack('<QQII16x', o, addr, 0, 0)
                            for o, (_, addr) in chunk)
            body += b'\x00' * (40 * (SNOD_ENTRIES - len(chunk)))
            snod = self._emit(struct.pack('<4sBBH', b'SNOD', 1, 0, len(chunk)) + body)
            level.append((snod, chunk[-1][0]))
        # B-tree nodes above them, up to 2 INT_K children each; the nodes of a
        # level lie back to back, so the sibling addresses are known up front
        node_bytes = 24 + 8 + 16 * NODE_CHILDREN
        depth = 0
        while True:
            nodes = []
            first = _pad8(self.end)
            count = max(-(-len(level) // NODE_CHILDREN), 1)
            for q in range(count):
                kids = level[q * NODE_CHILDREN:(q + 1) * NODE_CHILDREN]
                first_key = 0 if q == 0 else level[q * NODE_CHILDREN - 1][1]
                body = struct.pack('<Q', first_key) + b''.join(
                    struct.pack('<QQ', addr, key) for addr, key in kids)
                body += b'\x00' * (16 * (NODE_CHILDREN - len(kids)))
                left = first + (q - 1) * node_bytes if q > 0 else UNDEF
                right = first + (q + 1) * node_bytes if q + 1 < count else UNDEF
                node = self._emit(struct.pack('<4sBBHQQ', b'TREE', 0, depth, len(kids),
                                              left, right) + body)
                assert node == first + q * node_bytes
                nodes.append((node, kids[-1][1] if kids else 0))
            level = nodes
            depth += 1
            if len(level) == 1:
                break
        btree = level[0][0]
        msgs = [_message(0x0011, struct.pack('<QQ', btree, heap_addr))]
        ohdr = self._emit(_object_header(msgs + self._attr_messages(path)))
        self._root_cache = (btree, heap_addr)
        return ohdr

    def close(self):
        if self.fh is None:
            return
        if self.writable and self.dirty:
            root = self._write_group('/')
            btree, heap = self._root_cache
            sb = SIG + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INT_K, 0)
            sb += struct.pack('<QQQQ', 0, UNDEF, self.end, UNDEF)
            sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', btree, heap)
            assert len(sb) == 96
            self.fh.seek(0)
            self.fh.write(sb)
        self.fh.flush()
        self.fh.close()
        self.fh = None


class File(h5shim.File):

    def __init__(self, name, mode='r', **kwds):
        self.filename = str(name)
        self.mode = mode
        h5shim.Group.__init__(self, _Store(self.filename, mode), '/')


def is_hdf5(filename):
    try:
        with open(filename, 'rb') as fh:
            return fh.read(8) == SIG
    except OSError:
        return False
