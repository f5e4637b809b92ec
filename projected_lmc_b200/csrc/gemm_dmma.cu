#include "gemm_dmma.cuh"

#include <atomic>

namespace plmc {

static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_gemm_launches{0};
static std::atomic<double> g_gemm_flops{0.0};

void note_launch(long long kernels, double gemm_flops) {
    g_launches.fetch_add(kernels, std::memory_order_relaxed);
    if (gemm_flops > 0.0) {
        g_gemm_launches.fetch_add(1, std::memory_order_relaxed);
        double cur = g_gemm_flops.load(std::memory_order_relaxed);
        while (!g_gemm_flops.compare_exchange_weak(cur, cur + gemm_flops, std::memory_order_relaxed)) {
        }
    }
}
void stats_get(long long* launches, long long* gemm_launches, double* gemm_flops) {
    if (launches) *launches = g_launches.load();
    if (gemm_launches) *gemm_launches = g_gemm_launches.load();
    if (gemm_flops) *gemm_flops = g_gemm_flops.load();
}
void stats_reset() {
    g_launches = 0;
    g_gemm_launches = 0;
    g_gemm_flops = 0.0;
}

int gemm_init_attrs() {
    cudaError_t e;
    e = cudaFuncSetAttribute(gemm_dmma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
    if (e != cudaSuccess) return PLMC_ERR_LAUNCH;
    e = cudaFuncSetAttribute(gemm_dmma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
    if (e != cudaSuccess) return PLMC_ERR_LAUNCH;
    e = cudaFuncSetAttribute(gemm_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
    if (e != cudaSuccess) return PLMC_ERR_LAUNCH;
    e = cudaFuncSetAttribute(gemm_dmma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
    if (e != cudaSuccess) return PLMC_ERR_LAUNCH;
    return PLMC_OK;
}

int gemm_launch(bool aKC, bool bKC, const GemmArgs& a, int batch, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || batch <= 0) return PLMC_OK;
    if (a.M % G_BM || a.N % G_BN || a.K % G_BK || a.K <= 0) return PLMC_ERR_BADARG;
    if ((a.lda & 1) || (a.ldb & 1) || (a.ldc & 1)) return PLMC_ERR_BADARG;
    const long long tm = a.M / G_BM, tn = a.N / G_BN;
    long long tiles;
    if (a.lower) {
        if (tm != tn) return PLMC_ERR_BADARG;
        tiles = tm * (tm + 1) / 2;
    } else {
        tiles = tm * tn;
    }
    if (tiles > 2147483647LL || batch > 65535) return PLMC_ERR_BADARG;
    dim3 grid((unsigned)tiles, 1, (unsigned)batch);
    if (aKC && bKC)
        gemm_dmma_kernel<true, true><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
    else if (aKC && !bKC)
        gemm_dmma_kernel<true, false><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
    else if (!aKC && bKC)
        gemm_dmma_kernel<false, true><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
    else
        gemm_dmma_kernel<false, false><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
    PLMC_CHECK_LAUNCH();
    note_launch(1, 2.0 * (double)tiles * G_BM * G_BN * (double)a.K * batch);
    return PLMC_OK;
}

}  // namespace plmc
