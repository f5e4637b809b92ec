// Shared device/host helpers for the projected-LMC sm_100a library.
//
// Conventions used by every kernel in csrc/:
//   * all dense matrices are row-major FP64 with a leading dimension `ld`;
//   * the factorisation layer works on matrices whose order is a multiple of
//     PLMC_TILE (=128): the Gram builder pads K with an identity block, so no
//     kernel in the O(n^3) path needs a bounds check;
//   * batching over latent processes is done with gridDim.z and a batch stride;
//   * nothing here allocates, frees or synchronises the device: all work is
//     stream-ordered on the stream handed in through the C ABI.
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define PLMC_TILE 128

#define PLMC_OK 0
#define PLMC_ERR_BADARG (-1)
#define PLMC_ERR_LAUNCH (-2)

#define PLMC_CHECK_LAUNCH()                              \
    do {                                                 \
        cudaError_t e__ = cudaGetLastError();            \
        if (e__ != cudaSuccess) {                        \
            fprintf(stderr, "libplmc_b200: %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return PLMC_ERR_LAUNCH;                      \
        }                                                \
    } while (0)

namespace plmc {

// host-side launch statistics (read through plmc_stats_get; bench.py reports them)
void note_launch(long long kernels, double gemm_flops = 0.0);

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core (SASS: DMMA.8x8x4).
// lane = 4*g + t :  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed tree), result valid in thread 0.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sh /* >= NT/32 doubles */) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (l < NT / 32) ? sh[l] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

}  // namespace plmc
