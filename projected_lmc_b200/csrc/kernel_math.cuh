// Stationary kernel profiles k(s) and dk/ds as functions of the squared scaled distance s (gpytorch 1.11
// RBFKernel / MaternKernel semantics, reached from handle_covar_, projected_lmc.py:151-167):
//   RBF exp(-s/2);  Matern: r = sqrt(max(s, 1e-30)),  nu=5/2 (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r),
//   nu=3/2 (1 + sqrt3 r) exp(-sqrt3 r),  nu=1/2 exp(-r).
//
// The Gram build and the gradient sweep are bound by the FP64 instruction rate of the transform, not by HBM
// (ncu, profiles/r02_ncu_summary.md: 141 instructions per Gram entry with libdevice's exp/sqrt, FP64 pipe 26 %
// active, issue slots 51 %).  exp and sqrt are therefore written out for the argument ranges that occur here:
//   * exp(x), x <= 0:  n = rint(x 32/ln2), r = x - n ln2/32 (two-term Cody-Waite, |r| <= 0.0109),
//     e^r - 1 by a degree-6 polynomial (remainder 3e-18), times the table value 2^(j/32), j = n mod 32, exponent
//     n div 32 added to the exponent field: 13 FP64 instructions, 1 L1-resident load, < 1 ulp (checked against
//     libm on the host by tests/test_kernel_math_cpu.py);
//   * sqrt(s), s >= 1e-30:  MUFU.RSQ64H seed (rsqrt.approx.ftz.f64) and two coupled Newton steps (Goldschmidt form)
//     plus one residual correction: 9 FP64 instructions, <= 1 ulp.
// The same code compiles for the host (seed from float rsqrt) so that the accuracy claims are tested without a GPU.
#pragma once
#include "plmc_common.cuh"

#include <cmath>
#include <cstring>

namespace plmc {

// max(a, b) for non-NaN arguments.  fmax() on doubles has no hardware instruction on sm_100a: it expands to
// DSETP.MAX + SEL + FSEL + a NaN-quieting LOP3 per half (~7 instructions, ncu source page of gram_kernel); a
// compare-and-select is 3.
__host__ __device__ __forceinline__ double dmax(double a, double b) { return (a > b) ? a : b; }

#define PLMC_EXP_TABLE                                                                                                \
    {1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,         \
     1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484,                 \
     1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,               \
     1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,              \
     1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429, 1.718619298122478,                 \
     1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,                 \
     1.9571441241754002}

// device copy in GLOBAL memory, read with __ldg: the index differs between the lanes of a warp, which a __constant__
// bank would serialise; the 256-byte table is two L1 lines
static __device__ const double d_exp_table[32] = PLMC_EXP_TABLE;
static const double h_exp_table[32] = PLMC_EXP_TABLE;

// exp(x) for x <= 0 (x > 0 is outside the contract; the callers have checked their inputs for NaN).  Branch-free:
// the argument is clamped at -707 and results below e^-707 are flushed to 0 by a select (kernel values that
// small are irrelevant next to the noise floor e^-9 on the diagonal).
__host__ __device__ __forceinline__ double exp_nonpos(double x0) {
    const double x = dmax(x0, -707.0);
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: low word of (t + MAGIC) is rint(t)
    const double tm = fma(x, 46.16624130844683, MAGIC);      // x * 32 / ln 2
    const double nd = tm - MAGIC;
#ifdef __CUDA_ARCH__
    const int n = __double2loint(tm);
#else
    long long bits;
    memcpy(&bits, &tm, 8);
    const int n = (int)(unsigned)(bits & 0xFFFFFFFFll);
#endif
    double r = fma(nd, -0.021660849392446835, x);            // ln2/32, leading 36 bits: n * L_hi is exact
    r = fma(nd, -5.145609244655338e-14, r);
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = p * r;                                               // e^r - 1
#ifdef __CUDA_ARCH__
    const double T = __ldg(&d_exp_table[n & 31]);
    const double v = fma(T, p, T);
    const double r2 = __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
#else
    const double T = h_exp_table[n & 31];
    const double v = fma(T, p, T);
    const double r2 = ldexp(v, n >> 5);
#endif
    return (x0 < -707.0) ? 0.0 : r2;
}

// sqrt(s) for 1e-300 < s < 1e300
__host__ __device__ __forceinline__ double sqrt_pos(double s) {
    double y;
#ifdef __CUDA_ARCH__
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
#else
    y = (double)(1.0f / sqrtf((float)s));
    if (!(y > 0.0) || y > 1e30) y = 1.0 / sqrt(s);           // outside float range (host test only)
#endif
    double g = s * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, s);                          // residual: g <- g + d / (2 g)
    return fma(d, h, g);
}

// sqrt(s) and 1/sqrt(s) from ONE coupled Goldschmidt iteration (the Cholesky leaf's pivot step, csrc/potrf_leaf.cuh):
// root to <= 1 ulp, reciprocal to <= 2 ulp (checked on the host by tests/test_kernel_math_cpu.py).  s <= 0 or NaN
// gives NaN for both, like sqrt() of a negative number.
__host__ __device__ __forceinline__ void sqrt_and_reciprocal(double s, double& root, double& inv) {
    double y;
#ifdef __CUDA_ARCH__
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
#else
    y = (double)(1.0f / sqrtf((float)s));
    if (!(y > 0.0) || y > 1e30) y = 1.0 / sqrt(s);           // outside float range (host test only)
#endif
    double g = s * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, s);                          // residual: g <- g + d / (2 g)
    g = fma(d, h, g);
    r = fma(-h, g, 0.5);                                     // h <- 1 / (2 g) for the corrected root
    h = fma(h, r, h);
    const bool ok = s > 0.0;
#ifdef __CUDA_ARCH__
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#else
    const double qnan = NAN;
#endif
    root = ok ? g : qnan;
    inv = ok ? h + h : qnan;
}

template <int KID>
__host__ __device__ __forceinline__ double kernel_value(double s) {
    if (KID == 0) return exp_nonpos(-0.5 * s);
    if (KID == 1) {
        const double u = 5.0 * dmax(s, 1e-30);               // (sqrt5 r)^2
        const double a = sqrt_pos(u);
        return (1.0 + a + u * (1.0 / 3.0)) * exp_nonpos(-a);
    }
    if (KID == 2) {
        const double a = sqrt_pos(3.0 * dmax(s, 1e-30));
        return (1.0 + a) * exp_nonpos(-a);
    }
    return exp_nonpos(-sqrt_pos(dmax(s, 1e-30)));
}

// k(s) and dk/ds
template <int KID>
__host__ __device__ __forceinline__ void kernel_value_grad(double s, double& k, double& dk) {
    if (KID == 0) {
        k = exp_nonpos(-0.5 * s);
        dk = -0.5 * k;
        return;
    }
    if (KID == 1) {
        const double u = 5.0 * dmax(s, 1e-30);
        const double a = sqrt_pos(u);
        const double e = exp_nonpos(-a);
        k = (1.0 + a + u * (1.0 / 3.0)) * e;
        dk = -(5.0 / 6.0) * (1.0 + a) * e;
    } else if (KID == 2) {
        const double a = sqrt_pos(3.0 * dmax(s, 1e-30));
        const double e = exp_nonpos(-a);
        k = (1.0 + a) * e;
        dk = -1.5 * e;
    } else {
        const double r = sqrt_pos(dmax(s, 1e-30));
        k = exp_nonpos(-r);
        dk = (s > 1e-30) ? -0.5 * k / r : 0.0;
    }
}

}  // namespace plmc
