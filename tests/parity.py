"""Comparison helpers shared by the parity tests."""
import numpy as np

# north_star tolerances: rel 1e-6 (fp64 data) / 1e-5 (fp32 data)
RTOL = {'float64': 1e-6, 'float32': 1e-5}
F16_ULP = 2.0 ** -10


def assert_same_array(name, got, exp, exact_float=True, rtol=None):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, '%s: shape %s != %s' % (
        name, got.shape, exp.shape)
    if exp.dtype.kind in 'US':
        assert str(got) == str(exp), name
        return
    assert got.dtype == exp.dtype, '%s: dtype %s != %s' % (
        name, got.dtype, exp.dtype)
    if exp.dtype.kind in 'iub' or exact_float:
        assert np.array_equal(got, exp, equal_nan=exp.dtype.kind == 'f'), \
            '%s differs (%d of %d elements)' % (
                name, int(np.sum(~((got == exp) | (
                    (got != got) & (exp != exp))))), exp.size)
    else:
        tol = rtol if rtol is not None else RTOL.get(exp.dtype.name, 1e-5)
        assert np.allclose(got, exp, rtol=tol, atol=0, equal_nan=True), \
            '%s differs beyond rtol=%g' % (name, tol)


def assert_same_tree(got, exp, exact_float=True, skip=()):
    assert sorted(got) == sorted(exp), \
        'tree keys differ: only got %s / only expected %s' % (
            sorted(set(got) - set(exp)), sorted(set(exp) - set(got)))
    for k in sorted(exp):
        if any(k.endswith(s) for s in skip):
            continue
        assert_same_array(k, got[k], exp[k], exact_float)


def f16_ulps(a, b):
    """Distance in float16 units-in-the-last-place between two f16 arrays
    (NaN vs NaN counts as 0, NaN vs number as inf)."""
    a = np.asarray(a, dtype=np.float16)
    b = np.asarray(b, dtype=np.float16)
    ia = a.view(np.int16).astype(np.int32)
    ib = b.view(np.int16).astype(np.int32)
    d = np.abs(ia - ib).astype(np.float64)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    d[nan_a & nan_b] = 0
    d[nan_a ^ nan_b] = np.inf
    return d
