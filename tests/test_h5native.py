"""The in-repo HDF5 writer / reader (``h5native.py``) on the
CPU: round trips of the result layouts (reference ``track_orbits.py:366-397``,
``postprocessing.py:146-162``), append mode, B-trees of several levels, and the
byte-level structure the HDF5 specification prescribes (superblock v0,
symbol-table nodes sorted by name, B-tree keys = largest name of the left
subtree).  libhdf5 / h5py do not exist in this image: the files are checked
against the specification and by the independent parser, not by libhdf5."""
import struct

import numpy as np
import pytest

from nbody_orbit_analysis_b200 import h5native, h5shim, storage


def _track_like(hf, s, rng):
    g = hf.create_group('snapshot_%03d' % s)
    n = int(rng.integers(0, 50))
    data = {
        'region_offsets': np.sort(rng.integers(0, n + 1, 4)).astype(np.int64),
        'pericenter_IDs': rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64),
        'angles': rng.uniform(0, 6, n).astype(np.float16),
        'halo_IDs': rng.integers(0, 1000, 3).astype(np.int64),
        'region_radii': rng.uniform(0, 1, 3).astype(np.float32),
        'region_positions': rng.uniform(0, 1, (3, 3)),
        'bulk_velocities': rng.normal(0, 1, (3, 3)).astype(np.float32),
    }
    for k, v in data.items():
        g.create_dataset(k, data=v)
    return data


def test_track_layout_round_trip_with_appends(tmp_path):
    f = str(tmp_path / 't.h5')
    rng = np.random.default_rng(0)
    with h5native.File(f, 'w') as hf:
        hf.attrs['mode'] = 'pericentric'
        hf.attrs['box_size'] = 100.0
    exp = {}
    for s in range(40):                       # one open / close per snapshot
        with h5native.File(f, 'r+') as hf:
            exp[s] = _track_like(hf, s, rng)
    with h5native.File(f, 'r') as hf:
        assert hf.attrs['mode'] == 'pericentric'
        assert isinstance(hf.attrs['mode'], str)          # postprocessing.py:20
        assert float(hf.attrs['box_size']) == 100.0
        assert list(hf.keys()) == ['snapshot_%03d' % s for s in range(40)]
        for s, data in exp.items():
            g = hf['snapshot_%03d' % s]
            assert sorted(g.keys()) == sorted(data)
            for k, v in data.items():
                got = g[k][:]
                assert got.dtype == v.dtype and got.shape == v.shape, (s, k)
                assert np.array_equal(got, v), (s, k)
                assert len(g[k]) == len(v)
        assert len(hf['snapshot_003/pericenter_IDs']) == len(exp[3]['pericenter_IDs'])
    # the generic flattening used by every parity test
    tree = storage.tree(f)
    assert str(tree['/__attr__/mode']) == 'pericentric'
    assert np.array_equal(tree['/snapshot_017/angles'], exp[17]['angles'])


@pytest.mark.parametrize('n', [0, 1, 8, 9, 256, 257, 700])
def test_group_btree_structure(tmp_path, n):
    """n children: symbol-table nodes of <= 8 entries in name order, B-tree nodes
    of <= 32 children, key[i+1] = largest name below child i, key[0] = ''."""
    f = str(tmp_path / 'b.h5')
    names = ['d%05d' % ((i * 7919) % 100003) for i in range(n)]
    with h5native.File(f, 'w') as hf:
        for i, name in enumerate(names):
            hf.create_dataset(name, data=np.arange(3, dtype=np.int32) + i)
    raw = open(f, 'rb').read()
    assert raw[:8] == b'\x89HDF\r\n\x1a\n' and raw[8] == 0     # superblock v0
    assert raw[13] == 8 and raw[14] == 8                        # offsets, lengths
    leaf_k, int_k = struct.unpack_from('<HH', raw, 16)
    assert (leaf_k, int_k) == (4, 16)
    base, free, eof, drv = struct.unpack_from('<QQQQ', raw, 24)
    assert base == 0 and eof == len(raw) and free == drv == 2 ** 64 - 1
    _, ohdr, cache, _ = struct.unpack_from('<QQII', raw, 56)
    btree, heap = struct.unpack_from('<QQ', raw, 80)
    assert cache == 1 and raw[ohdr] == 1                         # v1 object header
    assert raw[heap:heap + 4] == b'HEAP'
    hsize, hfree, hdata = struct.unpack_from('<QQQ', raw, heap + 8)
    assert hfree == 1 and raw[hdata:hdata + 8] == b'\x00' * 8    # '' at offset 0

    def name_at(off):
        return raw[hdata + off:raw.index(b'\x00', hdata + off)].decode()
    seen = []

    def walk(addr, lo_key):
        assert raw[addr:addr + 4] == b'TREE'
        ntype, level, used = struct.unpack_from('<BBH', raw, addr + 4)
        assert ntype == 0 and used <= 32
        keys = [struct.unpack_from('<Q', raw, addr + 24 + 16 * i)[0]
                for i in range(used + 1)]
        kids = [struct.unpack_from('<Q', raw, addr + 32 + 16 * i)[0]
                for i in range(used)]
        assert name_at(keys[0]) == lo_key
        for i, kid in enumerate(kids):
            if level > 0:
                walk(kid, name_at(keys[i]))
            else:
                assert raw[kid:kid + 4] == b'SNOD'
                cnt = struct.unpack_from('<H', raw, kid + 6)[0]
                assert 1 <= cnt <= 8
                ents = [name_at(struct.unpack_from('<Q', raw, kid + 8 + 40 * e)[0])
                        for e in range(cnt)]
                assert ents == sorted(ents)
                assert all(name_at(keys[i]) < e for e in ents)
                seen.extend(ents)
            last = seen[-1] if seen else ''
            assert name_at(keys[i + 1]) == last      # largest name of the subtree
        return level
    walk(btree, '')
    assert seen == sorted(names)
    with h5native.File(f, 'r') as hf:
        assert list(hf.keys()) == sorted(names)
        for i in (0, n // 2, n - 1):
            if n:
                assert np.array_equal(hf[names[i]][:], np.arange(3) + i)


def test_dtypes_attributes_and_modes(tmp_path):
    f = str(tmp_path / 'd.h5')
    arrays = {np.dtype(t).name: (np.arange(6) * 3 - 4).astype(t)
              for t in ('i1', 'i2', 'i4', 'i8', 'u1', 'u2', 'u4', 'u8', 'f2',
                        'f4', 'f8')}
    with h5native.File(f, 'w') as hf:
        for k, v in arrays.items():
            hf.create_dataset(k, data=v)
        hf.create_dataset('scalar', data=np.float32(2.5))
        hf.create_dataset('empty', data=np.zeros((0, 3), dtype=np.float64))
        d = hf.create_dataset('flags', data=np.array([True, False]))
        d.attrs['unit'] = 'kpc'
        hf.attrs['vec'] = np.array([1.5, 2.5, 3.5])
        hf.attrs['n'] = np.int64(7)
        with pytest.raises(ValueError):
            hf.create_dataset('int32', data=np.zeros(2))
        with pytest.raises(TypeError):
            hf.create_dataset('text', data=np.array(['a', 'b']))
    with h5native.File(f, 'a') as hf:                  # append to the root
        hf.require_group('extra').create_dataset('x', data=np.arange(4))
    with h5native.File(f, 'r') as hf:
        for k, v in arrays.items():
            assert hf[k].dtype == v.dtype and np.array_equal(hf[k][:], v)
        assert hf['scalar'][()] == np.float32(2.5) and hf['scalar'].shape == ()
        assert hf['empty'].shape == (0, 3)
        assert np.array_equal(hf['flags'][:], [1, 0])
        assert hf['flags'].attrs['unit'] == 'kpc'
        assert np.array_equal(hf.attrs['vec'], [1.5, 2.5, 3.5])
        assert int(hf.attrs['n']) == 7
        assert np.array_equal(hf['extra/x'][1:3], [1, 2])
        with pytest.raises(OSError):
            hf.create_group('nope')
    with pytest.raises(FileNotFoundError):
        h5native.File(str(tmp_path / 'missing.h5'), 'r')
    with pytest.raises(FileExistsError):
        h5native.File(f, 'w-')


def test_storage_opens_either_container(tmp_path):
    """New files are HDF5; a file of the round-1 container (what the golden
    generator's h5py stand-in writes) is still read through ``storage``."""
    a, b = str(tmp_path / 'a.h5'), str(tmp_path / 'b.h5')
    with storage.File(a, 'w') as hf:
        hf.attrs['mode'] = 'apocentric'
        hf.create_dataset('x', data=np.arange(3))
    with h5shim.File(b, 'w') as hf:
        hf.attrs['mode'] = 'apocentric'
        hf.create_dataset('x', data=np.arange(3))
    if storage.BACKEND == 'hdf5':
        assert h5native.is_hdf5(a) and not h5shim.is_shim_file(a)
    assert h5shim.is_shim_file(b)
    ta, tb = storage.tree(a), storage.tree(b)
    assert sorted(ta) == sorted(tb)
    assert all(np.array_equal(ta[k], tb[k]) for k in ta)
    with storage.File(b, 'r+') as hf:                 # appended by its own writer
        hf.create_dataset('y', data=np.arange(2))
    assert h5shim.is_shim_file(b) and '/y' in storage.tree(b)


def test_objects_are_parsed_on_first_access(tmp_path):
    """Opening a file reads the root symbol table only; an object's header is
    parsed when it is first asked for.  Appends after such a partial read keep
    every untouched object (header address unchanged, never parsed), also
    nested groups with attributes, and edits of existing objects land."""
    import numpy as np
    from nbody_orbit_analysis_b200 import h5native, storage
    f = str(tmp_path / 'lazy.h5')
    rng = np.random.default_rng(3)
    expect = {}
    with h5native.File(f, 'w') as hf:
        hf.attrs['mode'] = 'pericentric'
        for s in range(60):
            g = hf.create_group('snapshot_%03d' % s)
            g.attrs['k'] = s
            for name in ('a', 'b', 'c'):
                arr = rng.integers(0, 99, rng.integers(0, 50)).astype(np.int64)
                g.create_dataset(name, data=arr)
                expect['/snapshot_%03d/%s' % (s, name)] = arr
            sub = g.create_group('deep')
            sub.attrs['note'] = 'n%d' % s
            sub.create_dataset('x', data=np.float16([s, 1.5]))
            expect['/snapshot_%03d/deep/x' % s] = np.float16([s, 1.5])

    calls = []
    real = h5native._Store._read_object

    def counting(self, path, addr):
        calls.append(path)
        return real(self, path, addr)
    h5native._Store._read_object = counting
    try:
        with h5native.File(f, 'r') as hf:
            assert calls == ['/']
            assert len(hf.keys()) == 60 and list(hf.keys())[-1] == 'snapshot_059'
            assert calls == ['/']                      # names need no headers
            assert 'snapshot_007' in hf and 'snapshot_999' not in hf
            assert np.array_equal(hf['snapshot_007']['b'][:],
                                  expect['/snapshot_007/b'])
            assert hf['snapshot_007/deep'].attrs['note'] == 'n7'
            assert sorted(calls) == ['/', '/snapshot_007', '/snapshot_007/b',
                                     '/snapshot_007/deep']
        # append: one new group, one dataset into an existing (unread) group, one
        # attribute on an existing (unread) nested group
        del calls[:]
        with h5native.File(f, 'r+') as hf:
            old_addr = dict(hf._s.addr) if hf._s.addr else {}
            g = hf.create_group('snapshot_060')
            g.create_dataset('a', data=np.arange(5))
            expect['/snapshot_060/a'] = np.arange(5)
            hf['snapshot_010'].create_dataset('extra', data=np.float32([1, 2]))
            expect['/snapshot_010/extra'] = np.float32([1, 2])
            hf['snapshot_020/deep'].attrs['more'] = 7
            with pytest.raises(ValueError):
                hf.create_group('snapshot_030')        # exists (found lazily)
            with pytest.raises(ValueError):
                hf['snapshot_010'].create_dataset('a', data=np.arange(2))
            store = hf._s
        touched = {'/', '/snapshot_010', '/snapshot_020', '/snapshot_020/deep',
                   '/snapshot_030'}
        assert set(calls) <= touched | {'/snapshot_010/a'}, sorted(set(calls) - touched)
        # untouched groups keep the header they had
        with h5native.File(f, 'r') as hf:
            hf.keys()
            for s in (0, 5, 30, 59):
                assert hf._s.addr['/snapshot_%03d' % s] == \
                    store.addr['/snapshot_%03d' % s]
    finally:
        h5native._Store._read_object = real
    got = storage.tree(f)
    data = {k: v for k, v in got.items() if '__attr__' not in k}
    assert sorted(data) == sorted(expect)
    for k, v in expect.items():
        assert data[k].dtype == v.dtype and np.array_equal(data[k], v), k
    assert got['/__attr__/mode'] == 'pericentric'
    assert int(got['/__attr__/snapshot_020/deep/more']) == 7
    assert str(got['/__attr__/snapshot_020/deep/note']) == 'n20'
    assert int(got['/__attr__/snapshot_059/k']) == 59
    # reopening costs the same whatever the number of groups
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        with h5native.File(f, 'r') as hf:
            hf['snapshot_059/a'][:]
    assert (time.perf_counter() - t0) / 20 < 0.01
