"""CPU: the hand-written exp / sqrt of csrc/kernel_math.cuh (compiled for the host from the same source as the
device kernels) against numpy's libm, in ulps, over the argument ranges the Gram kernels produce."""
import ctypes

import numpy as np
import pytest

from projected_lmc_b200 import _cabi


def profile(kid, s):
    lib = _cabi.load()
    s = np.ascontiguousarray(s, dtype=np.float64)
    k, dk = np.empty_like(s), np.empty_like(s)
    rc = lib.plmc_kernel_profile_host(kid, s.ctypes.data_as(ctypes.c_void_p), s.size, k.ctypes.data_as(ctypes.c_void_p),
                                      dk.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return k, dk


def reference(kid, s):
    s = s.astype(np.longdouble)
    if kid == 0:
        k = np.exp(-s / 2)
        return k, -k / 2
    r = np.sqrt(np.maximum(s, np.longdouble(1e-30)))
    if kid == 1:
        a = np.sqrt(np.longdouble(5)) * r
        e = np.exp(-a)
        return (1 + a + 5 * r * r / 3) * e, -(np.longdouble(5) / 6) * (1 + a) * e
    if kid == 2:
        a = np.sqrt(np.longdouble(3)) * r
        e = np.exp(-a)
        return (1 + a) * e, -1.5 * e
    k = np.exp(-r)
    return k, np.where(s > 1e-30, -0.5 * k / r, 0.0)


def ulps(got, ref):
    ref64 = ref.astype(np.float64)
    spacing = np.spacing(np.abs(ref64))
    return np.abs(got.astype(np.longdouble) - ref) / spacing


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_kernel_profiles_are_within_two_ulps_of_libm(kid):
    rng = np.random.default_rng(kid)
    s = np.concatenate([
        rng.uniform(0, 4, 200000), rng.uniform(0, 60, 200000), 10.0 ** rng.uniform(-30, 3, 200000),
        np.array([0.0, 1e-30, 1e-300, 1.0, 2.0, 4.0, 1e3, 1417.0]),
    ])
    k, dk = profile(kid, s)
    kr, dkr = reference(kid, s)
    keep = kr > 1e-290                                  # deep underflow is flushed to zero by design
    # exp(-a) turns the rounding of its FP64 argument a = sqrt(c s) (c s: 1/2 ulp, sqrt: <= 1 ulp) into a RELATIVE
    # error of up to ~1.5 ulp(a) = O(a) ulps of the result: common to any float64 evaluation of the formula (the
    # reference above is evaluated in long double).  Beyond that the hand-written code may add 2 ulp.
    a = np.zeros_like(s) if kid == 0 else np.sqrt({1: 5.0, 2: 3.0, 3: 1.0}[kid] * np.maximum(s, 1e-30))
    bound = 3.0 + 2.0 * a   # (1 + a + ..) is rounded at the ulp of [1, 2), the product lands just below 1
    assert (ulps(k[keep], kr[keep]) <= bound[keep]).all()
    keep_d = (np.abs(dkr) > 1e-290)
    assert (ulps(dk[keep_d], dkr[keep_d]) <= bound[keep_d] + 1.0).all()
    small = keep & (a < 1.0)
    assert ulps(k[small], kr[small]).max() <= 4.5
    assert np.all(k[~keep] >= 0) and np.all(k[~keep] < 1e-280)


def test_profile_at_zero_distance_and_monotonicity():
    for kid in range(4):
        k, dk = profile(kid, np.array([0.0]))
        assert abs(k[0] - 1.0) <= 1e-15          # Matern: r = sqrt(1e-30) = 1e-15
        s = np.linspace(0, 50, 5001)
        k, dk = profile(kid, s)
        assert np.all(np.diff(k) <= 0) and np.all(dk[1:] <= 0)
    k, _ = profile(0, np.array([5000.0, 1e300]))
    assert k[0] == 0.0 and k[1] == 0.0


def test_pivot_root_and_reciprocal_of_the_cholesky_leaf_within_one_and_two_ulps():
    """sqrt_and_reciprocal (one coupled Goldschmidt iteration; the pivot step of potrf_leaf_kernel): root <= 1 ulp,
    reciprocal <= 2 ulp over twelve decades; non-positive pivots give NaN like sqrt()."""
    lib = _cabi.load()
    rng = np.random.default_rng(5)
    s = np.concatenate([10.0 ** rng.uniform(-6, 6, 200000), rng.uniform(0.5, 2.0, 100000), [1.0, 4.0, 0.25, 2.0, 1e-300, 1e300]])
    root, inv = np.empty_like(s), np.empty_like(s)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.plmc_sqrt_reciprocal_host(vp(s), s.size, vp(root), vp(inv)) == 0
    sl = s.astype(np.longdouble)
    assert ulps(root, np.sqrt(sl)).max() <= 1.0
    assert ulps(inv, 1 / np.sqrt(sl)).max() <= 2.0
    bad = np.array([0.0, -1.0, np.nan])
    r2, i2 = np.empty_like(bad), np.empty_like(bad)
    assert lib.plmc_sqrt_reciprocal_host(vp(bad), bad.size, vp(r2), vp(i2)) == 0
    assert np.isnan(r2).all() and np.isnan(i2).all()
