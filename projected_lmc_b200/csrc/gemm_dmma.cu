#include "gemm_dmma.cuh"

namespace plmc {

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
    return PLMC_OK;
}

}  // namespace plmc
