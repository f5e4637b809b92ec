// Stationary kernel profiles k(s) and dk/ds as functions of the squared scaled distance s (gpytorch 1.11
// RBFKernel / MaternKernel semantics, reached from handle_covar_, projected_lmc.py:151-167):
//   RBF exp(-s/2);  Matern: r = sqrt(max(s, 1e-30)),  nu=5/2 (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r),
//   nu=3/2 (1 + sqrt3 r) exp(-sqrt3 r),  nu=1/2 exp(-r).
#pragma once
#include "plmc_common.cuh"

namespace plmc {

template <int KID>
__device__ __forceinline__ double kernel_value(double s) {
    if (KID == 0) return exp(-0.5 * s);
    const double r = sqrt(fmax(s, 1e-30));
    if (KID == 1) {
        const double a = 2.23606797749978969641 * r;  // sqrt(5) r
        return (1.0 + a + (5.0 / 3.0) * r * r) * exp(-a);
    }
    if (KID == 2) {
        const double a = 1.73205080756887729353 * r;
        return (1.0 + a) * exp(-a);
    }
    return exp(-r);
}

// k(s) and dk/ds
template <int KID>
__device__ __forceinline__ void kernel_value_grad(double s, double& k, double& dk) {
    if (KID == 0) {
        k = exp(-0.5 * s);
        dk = -0.5 * k;
        return;
    }
    const double r = sqrt(fmax(s, 1e-30));
    if (KID == 1) {
        const double a = 2.23606797749978969641 * r;
        const double e = exp(-a);
        k = (1.0 + a + (5.0 / 3.0) * r * r) * e;
        dk = -(5.0 / 6.0) * (1.0 + a) * e;
    } else if (KID == 2) {
        const double a = 1.73205080756887729353 * r;
        const double e = exp(-a);
        k = (1.0 + a) * e;
        dk = -1.5 * e;
    } else {
        k = exp(-r);
        dk = (s > 1e-30) ? -0.5 * k / r : 0.0;
    }
}

}  // namespace plmc
