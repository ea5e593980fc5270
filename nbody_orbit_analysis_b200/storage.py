"""Storage backend for the result files of the orbit-tracking path.

The file layout *is* part of the drop-in contract (reference
``track_orbits.py:366-397``, ``track_orbits_onthefly.py:208-252``,
``postprocessing.py:146-162``): HDF5 groups/datasets/attributes with fixed
names and dtypes.  When ``h5py`` is importable it is used unchanged; this image
has neither h5py nor libhdf5, so the default is the
in-repo HDF5 writer / reader ``h5native.py`` (real HDF5 files: superblock v0,
old-style groups, contiguous datasets).  ``OA_STORAGE=shim`` selects the round-1
record container ``h5shim.py``, ``OA_STORAGE=h5py`` insists on h5py.
"""
import os

from . import h5native, h5shim

_choice = os.environ.get('OA_STORAGE', 'auto')
_h5py = None
if _choice in ('auto', 'h5py'):
    try:
        import h5py as _h5py
        if getattr(_h5py, '_MAGIC', None) is not None:
            _h5py = None       # the shim was registered as sys.modules['h5py']
    except ImportError:
        if _choice == 'h5py':
            raise
if _choice == 'shim':
    _backend, BACKEND = h5shim, 'shim'
elif _h5py is not None:
    _backend, BACKEND = _h5py, 'h5py'
else:
    # no h5py / libhdf5 (this image): the in-repo HDF5 writer / reader
    _backend, BACKEND = h5native, 'hdf5'


def File(name, mode='r', **kwds):
    """Open a result file.  New files are written by the selected backend
    (``BACKEND``); an existing file is opened by whatever wrote it -- HDF5
    (h5py when importable, else ``h5native``) or the round-1 record container
    (``h5shim``, still used by the golden-vector generator as the reference's
    h5py stand-in)."""
    name = str(name)
    if mode in ('r', 'r+', 'a') and os.path.exists(name) and \
            os.path.getsize(name) > 0:
        if h5shim.is_shim_file(name):
            return h5shim.File(name, mode, **kwds)
        if h5native.is_hdf5(name):
            return (_h5py or h5native).File(name, mode, **kwds)
    return _backend.File(name, mode, **kwds)


def tree(filename):
    """Flatten a result file into ``{path: ndarray}`` (attributes appear as
    ``'/__attr__<path>/<key>'``).  Used by the parity tests and ``bench.py``."""
    import numpy as np
    out = {}

    def walk(g, prefix):
        for k, v in g.attrs.items():
            out['/__attr__%s/%s' % (prefix, k)] = np.asarray(v)
        for name in g.keys():
            item = g[name]
            path = prefix + '/' + name
            if hasattr(item, 'keys'):
                walk(item, path)
            else:
                out[path] = np.asarray(item[()] if item.shape == ()
                                       else item[:])
    with File(filename, 'r') as hf:
        walk(hf, '')
    return out
