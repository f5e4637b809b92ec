// Roofline denominators measured on the box (the FP64
// tensor peak is not in MEASURED_PEAKS.json, so the library carries its own
// register-resident DMMA / DFMA loops and an HBM copy).  Timing is done by the
// caller with CUDA events on the launch stream.
#include "plmc_common.cuh"

namespace plmc {

// each warp: `iters` x 16 m8n8k4 DMMA (256 FMA each) on 8 independent accumulators
__global__ void __launch_bounds__(1024) peak_dmma_kernel(double* out, long long iters, double seed) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;  // keep the loop alive
}

// each thread: `iters` x 16 independent DFMA
__global__ void __launch_bounds__(1024) peak_dfma_kernel(double* out, long long iters, double seed) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i * 1e-3;
    const double a = 1.0 + seed * 1e-12, b = seed * 1e-9 + threadIdx.x * 1e-12;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                                        long long n2) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) dst[i] = src[i];
}

}  // namespace plmc

extern "C" {

// flops per launch = blocks * (threads/32) * iters * 16 * 512
int plmc_peak_dmma(int blocks, int threads, long long iters, double* scratch, void* stream) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || (threads & 31)) return PLMC_ERR_BADARG;
    plmc::peak_dmma_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters, 1.0);
    PLMC_CHECK_LAUNCH();
    return PLMC_OK;
}

// flops per launch = blocks * threads * iters * 16 * 2
int plmc_peak_dfma(int blocks, int threads, long long iters, double* scratch, void* stream) {
    if (blocks <= 0 || threads <= 0 || threads > 1024) return PLMC_ERR_BADARG;
    plmc::peak_dfma_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters, 1.0);
    PLMC_CHECK_LAUNCH();
    return PLMC_OK;
}

// bytes moved per launch = 2 * 8 * n
int plmc_peak_copy(const double* src, double* dst, long long n, void* stream) {
    if (n <= 0 || (n & 1)) return PLMC_ERR_BADARG;
    plmc::peak_copy_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>((const double2*)src, (double2*)dst, n / 2);
    PLMC_CHECK_LAUNCH();
    return PLMC_OK;
}
}

// ---- do DMMA (tensor pipe) and DFMA (FP64 pipe) overlap? ---------------------------------
// even warps run the DMMA loop, odd warps the DFMA loop, in the same CTA.
namespace plmc {
__global__ void __launch_bounds__(1024) peak_mixed_kernel(double* out, long long iters_mma, long long iters_fma,
                                                          double seed) {
    const int warp = threadIdx.x >> 5;
    double s = 0.0;
    if ((warp & 1) == 0) {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
        double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
        for (long long it = 0; it < iters_mma; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    } else {
        double c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = i * 1e-3;
        const double a = 1.0 + seed * 1e-12, b = seed * 1e-9 + threadIdx.x * 1e-12;
        for (long long it = 0; it < iters_fma; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i];
    }
    if (s == 123.456) out[0] = s;
}
}  // namespace plmc

extern "C" int plmc_peak_mixed(int blocks, int threads, long long iters_mma, long long iters_fma, double* scratch,
                               void* stream) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || (threads & 63)) return PLMC_ERR_BADARG;
    plmc::peak_mixed_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters_mma, iters_fma, 1.0);
    PLMC_CHECK_LAUNCH();
    return PLMC_OK;
}
