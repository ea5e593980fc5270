"""Host-side scalars of the tracking path (reference ``utils.py``).

The array primitives of the reference's ``utils.py`` (``myin1d``,
``recenter_coordinates``) have no host implementation here: they are fused into
the CUDA kernels (``csrc/oa_track.cu``).  Only the per-snapshot scalar H(z)
stays on the host, as in SURVEY.md section 8(a-5).
"""
import numpy as np


def hubble_parameter(z, H0, Omega_m, Omega_L, Omega_k=0):
    """H(z) = H0 sqrt(Om (1+z)^3 + Ok (1+z)^2 + OL); reference
    ``utils.py:36-39`` (same evaluation order, so the same float64)."""
    zp1 = 1 + z
    radicand = Omega_m * zp1**3 + Omega_k * zp1**2 + Omega_L
    return H0 * np.sqrt(radicand)


def myin1d(a, b):
    """Indices ``k`` with ``a[k] == b`` element-wise, in ``b``'s order
    (reference ``utils.py:4-11``; ``a`` and ``b`` unique, ``b`` a subset of
    ``a``).  Host helper for the SMALL halo-catalogue joins of the
    post-processing (tens to 1e5 halo IDs); particle-sized joins run on the GPU
    (``oa_lookup_sorted``)."""
    a, b = np.asarray(a), np.asarray(b)
    if len(a) == 0 or len(b) == 0:
        return np.zeros(0, dtype=np.int64)
    by_value = np.argsort(a, kind='stable')
    return by_value[np.searchsorted(a, b, sorter=by_value)].astype(np.int64)
