"""Storage backend for the result files of the orbit-tracking path.

The file layout *is* part of the drop-in contract (reference
``track_orbits.py:366-397``, ``track_orbits_onthefly.py:208-252``,
``postprocessing.py:146-162``): HDF5 groups/datasets/attributes with fixed
names and dtypes.  When ``h5py`` is importable it is used unchanged; this image
has neither h5py nor libhdf5, so the fallback is the
API-compatible container in ``h5shim.py``.  Set ``OA_STORAGE=shim`` or
``OA_STORAGE=h5py`` to force one.
"""
import os

_choice = os.environ.get('OA_STORAGE', 'auto')

if _choice == 'shim':
    from . import h5shim as _backend
    BACKEND = 'shim'
else:
    try:
        import h5py as _backend
        BACKEND = 'h5py'
        if getattr(_backend, '_MAGIC', None) is not None:
            BACKEND = 'shim'   # the shim was registered as sys.modules['h5py']
    except ImportError:
        if _choice == 'h5py':
            raise
        from . import h5shim as _backend
        BACKEND = 'shim'

File = _backend.File


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
