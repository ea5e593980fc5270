"""Minimal stand-in for the subset of the ``h5py`` API that the orbit-tracking
path touches (``File`` / ``Group`` / ``Dataset`` / ``attrs``).

Why it exists: the reference persists its results with h5py
(reference ``orbitanalysis/track_orbits.py:354-397``,
``track_orbits_onthefly.py:208-252``, ``postprocessing.py:10-28,146-162``) but
neither h5py nor libhdf5 exists in this image.  The
storage layer (``storage.py``) uses real h5py when it is importable and this
module otherwise.  The golden-vector generator installs this module as
``sys.modules['h5py']`` so that the *unmodified* reference can run here.

On-disk format: an append-only record log.  ``b'OAH5SHM1'`` magic, then records
``<u32 header_len><header json><payload bytes>``.  Appending a group for one
snapshot therefore costs O(that snapshot), like HDF5, not O(file).
"""
import io
import json
import os
import struct

import numpy as np

_MAGIC = b'OAH5SHM1'


def _norm(path):
    parts = [p for p in path.split('/') if p]
    return '/' + '/'.join(parts)


def _join(base, name):
    if name.startswith('/'):
        return _norm(name)
    return _norm(base + '/' + name)


class _Store:
    """The parsed record log of one open file."""

    def __init__(self, filename, mode):
        self.filename = filename
        self.mode = mode
        exists = os.path.exists(filename)
        if mode == 'r':
            if not exists:
                raise FileNotFoundError(
                    "Unable to open file (no such file: %r)" % filename)
            self.fh = open(filename, 'rb')
            self.writable = False
        elif mode == 'r+':
            if not exists:
                raise FileNotFoundError(
                    "Unable to open file (no such file: %r)" % filename)
            self.fh = open(filename, 'r+b')
            self.writable = True
        elif mode == 'w':
            self.fh = open(filename, 'w+b')
            self.writable = True
        elif mode in ('w-', 'x'):
            if exists:
                raise FileExistsError(
                    "Unable to create file (file exists: %r)" % filename)
            self.fh = open(filename, 'w+b')
            self.writable = True
        elif mode == 'a':
            self.fh = open(filename, 'r+b' if exists else 'w+b')
            self.writable = True
        else:
            raise ValueError("Invalid mode %r" % (mode,))

        self.groups = {'/'}
        self.datasets = {}      # path -> (dtype, shape, offset, nbytes)
        self.attrs = {'/': {}}  # path -> {key: value}
        self._scan()

    def _scan(self):
        fh = self.fh
        fh.seek(0, io.SEEK_END)
        size = fh.tell()
        fh.seek(0)
        if size == 0:
            if self.writable:
                fh.write(_MAGIC)
                fh.flush()
            return
        if fh.read(len(_MAGIC)) != _MAGIC:
            raise OSError("%r is not an orbit-b200 shim container "
                          "(was it written by real h5py?)" % self.filename)
        pos = len(_MAGIC)
        while pos < size:
            fh.seek(pos)
            (hlen,) = struct.unpack('<I', fh.read(4))
            head = json.loads(fh.read(hlen).decode('utf-8'))
            payload = pos + 4 + hlen
            nbytes = head.get('nbytes', 0)
            kind = head['kind']
            if kind == 'group':
                self.groups.add(head['path'])
                self.attrs.setdefault(head['path'], {})
            elif kind == 'dataset':
                self.datasets[head['path']] = (
                    np.dtype(head['dtype']), tuple(head['shape']), payload,
                    nbytes)
                self.attrs.setdefault(head['path'], {})
            elif kind == 'attr':
                if 'str' in head:
                    val = head['str']
                else:
                    fh.seek(payload)
                    val = np.frombuffer(
                        fh.read(nbytes), dtype=np.dtype(head['dtype'])
                    ).reshape(tuple(head['shape'])).copy()
                    if val.shape == ():
                        val = val[()]
                self.attrs.setdefault(head['path'], {})[head['key']] = val
            pos = payload + nbytes

    def _append(self, head, payload=b''):
        if not self.writable:
            raise OSError("file is open read-only")
        head = dict(head)
        head['nbytes'] = len(payload)
        hb = json.dumps(head).encode('utf-8')
        fh = self.fh
        fh.seek(0, io.SEEK_END)
        pos = fh.tell()
        fh.write(struct.pack('<I', len(hb)))
        fh.write(hb)
        fh.write(payload)
        return pos + 4 + len(hb)

    def exists(self, path):
        return path in self.groups or path in self.datasets

    def children(self, path):
        prefix = path.rstrip('/') + '/'
        names = set()
        for p in list(self.groups) + list(self.datasets):
            if p != path and p.startswith(prefix):
                names.add(p[len(prefix):].split('/')[0])
        return sorted(names)   # h5py iterates names alphabetically

    def add_group(self, path):
        parent = _norm(os.path.dirname(path))
        if parent != '/' and parent not in self.groups:
            self.add_group(parent)
        self._append({'kind': 'group', 'path': path})
        self.groups.add(path)
        self.attrs.setdefault(path, {})

    def add_dataset(self, path, arr):
        parent = _norm(os.path.dirname(path))
        if parent not in self.groups:
            self.add_group(parent)
        arr = np.ascontiguousarray(arr)
        if arr.dtype.kind in 'OU':
            raise TypeError("shim container stores numeric arrays only")
        off = self._append(
            {'kind': 'dataset', 'path': path, 'dtype': arr.dtype.str,
             'shape': list(arr.shape)}, arr.tobytes())
        self.datasets[path] = (arr.dtype, tuple(arr.shape), off, arr.nbytes)
        self.attrs.setdefault(path, {})

    def set_attr(self, path, key, value):
        if isinstance(value, (str, bytes)):
            if isinstance(value, bytes):
                value = value.decode('utf-8')
            self._append(
                {'kind': 'attr', 'path': path, 'key': key, 'str': value})
            self.attrs[path][key] = value
            return
        arr = np.asarray(value)
        if arr.dtype.kind in 'OU':
            raise TypeError("unsupported attribute type %r" % (type(value),))
        self._append(
            {'kind': 'attr', 'path': path, 'key': key, 'dtype': arr.dtype.str,
             'shape': list(arr.shape)}, np.ascontiguousarray(arr).tobytes())
        self.attrs[path][key] = arr[()] if arr.shape == () else arr.copy()

    def read(self, path, start=0, stop=None):
        dtype, shape, off, nbytes = self.datasets[path]
        if len(shape) == 0:
            self.fh.seek(off)
            return np.frombuffer(self.fh.read(nbytes), dtype=dtype)[0]
        n0 = shape[0]
        if stop is None:
            stop = n0
        row = dtype.itemsize * int(np.prod(shape[1:], dtype=np.int64))
        self.fh.seek(off + start * row)
        buf = self.fh.read(max(stop - start, 0) * row)
        return np.frombuffer(buf, dtype=dtype).reshape(
            (max(stop - start, 0),) + shape[1:]).copy()

    def close(self):
        if self.fh is not None:
            self.fh.flush()
            self.fh.close()
            self.fh = None


class AttributeManager:

    def __init__(self, store, path):
        self._s, self._p = store, path

    def __getitem__(self, key):
        return self._s.attrs[self._p][key]

    def __setitem__(self, key, value):
        self._s.set_attr(self._p, key, value)

    def __contains__(self, key):
        return key in self._s.attrs[self._p]

    def __iter__(self):
        return iter(sorted(self._s.attrs[self._p]))

    def __len__(self):
        return len(self._s.attrs[self._p])

    def keys(self):
        return sorted(self._s.attrs[self._p])

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def get(self, key, default=None):
        return self._s.attrs[self._p].get(key, default)


class Dataset:

    def __init__(self, store, path):
        self._s, self._p = store, path
        self.dtype, self.shape, _, _ = store.datasets[path]
        self.name = path

    @property
    def attrs(self):
        return AttributeManager(self._s, self._p)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def __len__(self):
        if not self.shape:
            raise TypeError("Attempt to take len() of scalar dataset")
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        a = self._s.read(self._p)
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, key):
        if not self.shape:
            return self._s.read(self._p)[key] if key != () else \
                self._s.read(self._p)
        # fast path: contiguous range of the leading axis
        k0, rest = key, ()
        if isinstance(key, tuple):
            k0, rest = (key[0], key[1:]) if len(key) else (slice(None), ())
        if isinstance(k0, slice) and k0 is not Ellipsis:
            start, stop, step = k0.indices(self.shape[0])
            if step == 1:
                out = self._s.read(self._p, start, max(stop, start))
                return out[(slice(None),) + rest] if rest else out
        return self._s.read(self._p)[key]


class Group:

    def __init__(self, store, path):
        self._s, self._p = store, path
        self.name = path

    @property
    def attrs(self):
        return AttributeManager(self._s, self._p)

    def create_group(self, name):
        path = _join(self._p, name)
        if self._s.exists(path):
            raise ValueError(
                "Unable to create group (name already exists)")
        self._s.add_group(path)
        return Group(self._s, path)

    def require_group(self, name):
        path = _join(self._p, name)
        if path in self._s.groups:
            return Group(self._s, path)
        return self.create_group(name)

    def create_dataset(self, name, shape=None, dtype=None, data=None,
                       **kwds):
        path = _join(self._p, name)
        if self._s.exists(path):
            raise ValueError(
                "Unable to create dataset (name already exists)")
        if data is None:
            data = np.zeros(shape if shape is not None else (), dtype=dtype)
        arr = np.asarray(data)
        if dtype is not None:
            arr = arr.astype(dtype)
        if arr.dtype == object:
            raise TypeError("Object dtype has no native HDF5 equivalent")
        self._s.add_dataset(path, arr)
        return Dataset(self._s, path)

    def __getitem__(self, name):
        path = _join(self._p, name)
        if path in self._s.datasets:
            return Dataset(self._s, path)
        if path in self._s.groups:
            return Group(self._s, path)
        raise KeyError("Unable to open object (object %r doesn't exist)"
                       % name)

    def __contains__(self, name):
        return self._s.exists(_join(self._p, name))

    def keys(self):
        return self._s.children(self._p)

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class File(Group):

    def __init__(self, name, mode='r', **kwds):
        self.filename = str(name)
        self.mode = mode
        super().__init__(_Store(self.filename, mode), '/')

    def close(self):
        self._s.close()

    def flush(self):
        if self._s.fh is not None:
            self._s.fh.flush()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def is_shim_file(filename):
    try:
        with open(filename, 'rb') as fh:
            return fh.read(len(_MAGIC)) == _MAGIC
    except OSError:
        return False
