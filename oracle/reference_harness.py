"""TEST INFRASTRUCTURE -- never imported by the product path.

Imports the UNMODIFIED reference (``/root/reference/orbitanalysis``) in this
container so that golden vectors can be generated from it
(``tests/golden/make_golden.py``).  The reference imports ``h5py`` and
``pathos`` at module level (reference ``track_orbits.py:3-4``,
``track_orbits_onthefly.py:3``, ``postprocessing.py:2``); neither is installed
here, so two stand-ins are registered first:

* ``h5py``   -> ``nbody_orbit_analysis_b200.h5shim`` (same File/Group/Dataset
  subset, on-disk container of our own);
* ``pathos.multiprocessing.Pool`` -> ``multiprocess.Pool`` (that is what pathos
  re-exports).

``/root/reference`` does not exist on the GPU box; there the pip-installed copy
under ``oracle/_ref`` (``make -C oracle ref``, run by ``__graft_entry__.build()``
in the build container) is imported instead.  Only ``bench.py``'s CPU legs
(``--impl reference``, ``cpu_baseline``) and the golden-vector generator use
this module.
"""
import importlib
import os
import sys
import types

# where the unmodified reference package is importable from: the mounted tree in
# the build container, else the pip-installed copy that `make -C oracle ref`
# vendors into oracle/_ref (git-ignored; it travels with the snapshot to the GPU
# box, where /root/reference does not exist)
_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get('OA_REFERENCE_ROOT', '/root/reference'),
               os.path.join(_HERE, '_ref')]
REFERENCE_ROOT = next((c for c in _CANDIDATES
                       if os.path.isdir(os.path.join(c, 'orbitanalysis'))),
                      _CANDIDATES[0])


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'orbitanalysis'))


def vendored():
    """True when the copy in oracle/_ref (not the mounted tree) is in use."""
    return available() and os.path.abspath(REFERENCE_ROOT) == \
        os.path.join(_HERE, '_ref')


def load_reference():
    """Return the reference package modules as a namespace."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from nbody_orbit_analysis_b200 import h5shim
    try:
        import h5py  # noqa: F401  (real one, if this box has it)
    except ImportError:
        sys.modules['h5py'] = h5shim
    if 'pathos.multiprocessing' not in sys.modules:
        try:
            import pathos.multiprocessing  # noqa: F401
        except ImportError:
            import multiprocess
            pathos = types.ModuleType('pathos')
            pmp = types.ModuleType('pathos.multiprocessing')
            pmp.Pool = multiprocess.Pool
            pathos.multiprocessing = pmp
            sys.modules['pathos'] = pathos
            sys.modules['pathos.multiprocessing'] = pmp
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.track_orbits = importlib.import_module('orbitanalysis.track_orbits')
    ns.onthefly = importlib.import_module(
        'orbitanalysis.track_orbits_onthefly')
    ns.progenitors = importlib.import_module('orbitanalysis.progenitors')
    ns.postprocessing = importlib.import_module('orbitanalysis.postprocessing')
    ns.utils = importlib.import_module('orbitanalysis.utils')
    return ns
