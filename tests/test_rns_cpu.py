"""CPU: the arithmetic of the residue-number-system FP64 GEMM (csrc/ozaki2.cu) restated with Python integers.

No GPU: the reconstruction constants come from the library's host-only entry point plmc_rns_constants and are
checked against exact big-integer arithmetic; the device-side tricks (balanced base-256 digits, DP4A weights,
float32 reciprocal reduction, hi/lo split of the INT32 accumulator, two-sum CRT) are emulated step by step with
numpy at the precision the kernels use."""
import ctypes
import math
import random
from fractions import Fraction

import numpy as np
import pytest

from projected_lmc_b200 import _cabi

MAGIC = np.float32(12582912.0)


def constants(moduli, K):
    lib = _cabi.load()
    mods = (ctypes.c_int * moduli)()
    H = (ctypes.c_double * moduli)()
    L = (ctypes.c_double * moduli)()
    ps, bits = ctypes.c_double(0.0), ctypes.c_int(0)
    rc = lib.plmc_rns_constants(moduli, K, mods, H, L, ctypes.byref(ps), ctypes.byref(bits))
    assert rc == 0
    return list(mods), list(H), list(L), ps.value, bits.value


def symmod(x, p):
    r = x % p
    if 2 * r >= p:
        r -= p
    return r


def mod_from_biased(x, p):
    """The float32 reduction of the kernels: x (|x| < 2^22) -> symmetric residue, via MAGIC-biased arithmetic."""
    F = np.float32(MAGIC + np.float32(x))
    f = np.float32(F - MAGIC)
    kk = np.float32(np.float64(f) * np.float64(np.float32(1.0) / np.float32(p)) + np.float64(MAGIC))  # one rounding (fma)
    k = np.float32(kk - MAGIC)
    rb = np.float32(np.float64(F) - np.float64(k) * np.float64(p))                                    # exact
    byte = int(rb.view(np.uint32)) & 0xFF
    return byte - 256 if byte >= 128 else byte


@pytest.mark.parametrize("moduli", [10, 14, 16, 18])
def test_moduli_are_coprime_and_bits_satisfy_the_uniqueness_bound(moduli):
    for K in (128, 640, 16384, 22272, 65408):
        mods, H, L, pscale, bits = constants(moduli, K)
        assert all(2 <= p <= 256 for p in mods)
        for i in range(moduli):
            for j in range(i):
                assert math.gcd(mods[i], mods[j]) == 1
        P = math.prod(mods)
        assert 2 * K * 4 ** bits < P                       # |c'| <= K 4^bits < P / 2: CRT is unique
        assert bits == 61 or 2 * K * 4 ** (bits + 1) >= P  # and bits is the largest such
        assert pscale == float(Fraction(P, 4 ** bits))
    assert constants(16, 16384)[4] == 55 and constants(14, 16384)[4] == 47   # the 7- / 6-slice grades of ozaki.cu


@pytest.mark.parametrize("moduli", [14, 16])
def test_crt_constants_match_exact_arithmetic(moduli):
    mods, H, L, _, _ = constants(moduli, 1024)
    P = math.prod(mods)
    for i, p in enumerate(mods):
        u = pow((P // p) % p, -1, p)
        h = (u << 40) // p
        assert H[i] == h / 2.0 ** 40
        assert L[i] == float(Fraction((u << 40) % p, p)) / 2.0 ** 40 or abs(L[i] - float(Fraction(u, p) - Fraction(h, 2 ** 40))) < 2e-29


def test_balanced_digits_and_dp4a_weights_give_the_residue():
    random.seed(0)
    mods = constants(18, 1024)[0]
    for _ in range(2000):
        q = random.getrandbits(random.choice([3, 17, 40, 55, 61])) * random.choice([-1, 1])
        Q = ((q + 0x8080808080808080) & 0xFFFFFFFFFFFFFFFF) ^ 0x8080808080808080
        digits = [((Q >> (8 * b)) & 0xFF) for b in range(8)]
        digits = [d - 256 if d >= 128 else d for d in digits]
        assert sum(d * 256 ** b for b, d in enumerate(digits)) == q
        for p in mods:
            w = [symmod(256 ** b, p) for b in range(8)]
            assert all(-128 <= x <= 127 for x in w)
            pre = sum(d * x for d, x in zip(digits, w))
            assert abs(pre) < 2 ** 22
            r = mod_from_biased(pre, p)
            assert (r - q) % p == 0 and -128 <= r <= 127 and abs(r) <= p // 2


def test_accumulator_reduction_is_exact_over_its_whole_range():
    random.seed(1)
    mods = constants(18, 1024)[0]
    for p in mods[1:]:
        c16 = symmod(65536, p)
        for _ in range(3000):
            acc = random.randint(-(2 ** 30) + 1, 2 ** 30 - 1)
            x = (acc >> 16) * c16 + (acc & 0xFFFF)
            assert abs(x) < 2 ** 22
            r = mod_from_biased(x, p)
            assert (r - acc) % p == 0 and abs(r) <= p // 2
        # worst cases for the float32 quotient: values whose quotient sits closest to a half-integer
        for x in (2 ** 22 - 1, -(2 ** 22) + 1):
            for dlt in range(-3 * p, 3 * p):
                r = mod_from_biased(x - abs(dlt) if x > 0 else x + abs(dlt), p)
                assert abs(r) <= p // 2


@pytest.mark.parametrize("moduli,K", [(16, 256), (14, 640), (16, 16384)])
def test_emulated_product_matches_exact_rational_product(moduli, K):
    """End to end on a few rows/columns with wildly different scales: every entry of A B^T is recovered with an
    error below  K 2^-bits  relative to  max|row| max|col|  (the bound of the scheme)."""
    rng = np.random.default_rng(moduli * 1000 + K)
    mods, H, L, pscale, bits = constants(moduli, K)
    M = N = 6
    A = rng.standard_normal((M, K)) * np.exp2(rng.integers(-40, 40, size=(M, 1)).astype(float))
    B = rng.standard_normal((N, K)) * np.exp2(rng.integers(-40, 40, size=(N, 1)).astype(float))
    A[1, ::3] *= 1e-9          # wide dynamic range inside a row
    B[2, 5] = 0.0

    def quantise(X):
        ex, Q = [], []
        for row in X:
            e = math.frexp(np.abs(row).max())[1]
            ex.append(e)
            Q.append([int(np.rint(math.ldexp(float(v), bits - e))) for v in row])
        return ex, Q

    ea, QA = quantise(A)
    eb, QB = quantise(B)
    for i in range(M):
        for j in range(N):
            c = sum(a * b for a, b in zip(QA[i], QB[j]))                      # the INT8 GEMMs compute this mod p_i
            assert 2 * abs(c) < math.prod(mods)
            s1 = s2 = 0.0
            for k, p in enumerate(mods):
                r = float(symmod(c, p))
                s1 += r * H[k]                                                   # exact
                s2 += r * L[k]
            v = (s1 - np.rint(s1)) + s2
            v -= np.rint(v)
            got = v * pscale * math.ldexp(1.0, ea[i]) * math.ldexp(1.0, eb[j])
            exact = sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(A[i], B[j]))
            scale = float(np.abs(A[i]).max() * np.abs(B[j]).max())
            assert abs(Fraction(got) - exact) <= Fraction(scale) * K * Fraction(1, 2 ** (bits - 2))
            # and the reconstruction itself is exact to FP64 rounding of the integer product
            ref = float(Fraction(c) * Fraction(2) ** (ea[i] + eb[j] - 2 * bits))
            assert got == pytest.approx(ref, rel=4e-16, abs=0.0) or c == 0
