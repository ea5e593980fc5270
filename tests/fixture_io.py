"""Record / replay helpers for the golden fixtures under ``tests/golden``.

A fixture is one ``.npz``: the exact arrays the two user callbacks returned to
the reference (``in/...``), the flattened result file the unmodified reference
wrote (``out/...``) and a JSON ``meta`` blob with the call arguments.
"""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

_SCALARS = ('box_size', 'redshift', 'H0', 'Omega_m', 'Omega_L', 'Omega_k',
            'masses')
_ARRAYS = ('ids', 'coordinates', 'velocities', 'region_offsets')


class Recorder:
    """Wraps the two callbacks and remembers everything they returned."""

    def __init__(self, regions, load_snapshot_data):
        self._regions, self._load = regions, load_snapshot_data
        self.data = {}

    def regions(self, snapshot_number, halo_ids):
        out = self._regions(snapshot_number, halo_ids)
        key = 'in/s%d/' % int(snapshot_number)
        self.data[key + 'regions_pos'] = np.asarray(out[0])
        self.data[key + 'regions_rad'] = np.asarray(out[1])
        if len(out) > 2 and out[2] is not None:
            self.data[key + 'regions_bulk'] = np.asarray(out[2])
        self.data[key + 'regions_n'] = np.asarray(len(out))
        return out

    def load_snapshot_data(self, snapshot_number, positions, radii):
        snap = self._load(snapshot_number, positions, radii)
        key = 'in/s%d/' % int(snapshot_number)
        for name in _ARRAYS:
            self.data[key + name] = np.asarray(snap[name])
        for name in _SCALARS:
            if name in snap:
                self.data[key + name] = np.asarray(snap[name])
        return snap


class Replay:
    """Callbacks that hand back the recorded arrays of a fixture."""

    def __init__(self, fixture):
        self.fx = fixture

    def regions(self, snapshot_number, halo_ids):
        key = 'in/s%d/' % int(snapshot_number)
        pos = self.fx[key + 'regions_pos']
        rad = self.fx[key + 'regions_rad']
        assert len(pos) == len(np.atleast_1d(halo_ids))
        if int(self.fx[key + 'regions_n']) == 2:
            return pos, rad
        bulk = self.fx[key + 'regions_bulk'] \
            if key + 'regions_bulk' in self.fx else None
        return pos, rad, bulk

    def load_snapshot_data(self, snapshot_number, positions, radii):
        key = 'in/s%d/' % int(snapshot_number)
        snap = {name: self.fx[key + name] for name in _ARRAYS}
        for name in _SCALARS:
            if key + name in self.fx:
                v = self.fx[key + name]
                snap[name] = v if v.ndim else v.item()
        return snap


def save_fixture(name, meta, inputs, outputs):
    arrays = {'meta': np.asarray(json.dumps(meta))}
    arrays.update(inputs)
    for k, v in outputs.items():
        v = np.asarray(v)
        if v.dtype.kind == 'U':
            v = np.asarray(str(v))
        arrays['out' + k] = v
    path = os.path.join(GOLDEN_DIR, name + '.npz')
    np.savez_compressed(path, **arrays)
    return path


def load_fixture(name):
    path = os.path.join(GOLDEN_DIR, name + '.npz')
    with np.load(path, allow_pickle=False) as z:
        fx = {k: z[k] for k in z.files}
    fx['meta'] = json.loads(str(fx['meta']))
    return fx


def expected_tree(fx, prefix='out'):
    return {k[len(prefix):]: v for k, v in fx.items()
            if k.startswith(prefix + '/')}


def list_fixtures(kind=None):
    names = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                   if f.endswith('.npz'))
    if kind is not None:
        names = [n for n in names if n.startswith(kind)]
    return names
