#include "gemm_dmma.cuh"

#include <atomic>
#include <cstdlib>

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

template <bool A, bool B, bool T>
static int set_attr() {
    return (cudaFuncSetAttribute(gemm_dmma_kernel<A, B, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 G_SMEM_BYTES) == cudaSuccess &&
            cudaFuncSetAttribute(gemm_dmma_small_kernel<A, B, T, 32, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 S_SMEM_BYTES) == cudaSuccess &&
            cudaFuncSetAttribute(gemm_dmma_small_kernel<A, B, T, 128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 S_SMEM_BYTES) == cudaSuccess)
               ? PLMC_OK
               : PLMC_ERR_LAUNCH;
}

int gemm_init_attrs() {
    int r = 0;
    r |= set_attr<true, true, false>();
    r |= set_attr<true, false, false>();
    r |= set_attr<false, true, false>();
    r |= set_attr<false, false, false>();
    r |= set_attr<true, true, true>();
    r |= set_attr<true, false, true>();
    r |= set_attr<false, true, true>();
    r |= set_attr<false, false, true>();
    return r ? PLMC_ERR_LAUNCH : PLMC_OK;
}

template <bool T>
static void launch_variant(bool aKC, bool bKC, dim3 grid, cudaStream_t st, const GemmArgs& a) {
    if (aKC && bKC)
        gemm_dmma_kernel<true, true, T><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
    else if (aKC && !bKC)
        gemm_dmma_kernel<true, false, T><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
    else if (!aKC && bKC)
        gemm_dmma_kernel<false, true, T><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
    else
        gemm_dmma_kernel<false, false, T><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
}

template <bool T, int TM, int TN>
static void launch_small(bool aKC, bool bKC, dim3 grid, cudaStream_t st, const GemmArgs& a) {
    if (aKC && bKC)
        gemm_dmma_small_kernel<true, true, T, TM, TN><<<grid, S_THREADS, S_SMEM_BYTES, st>>>(a);
    else if (aKC && !bKC)
        gemm_dmma_small_kernel<true, false, T, TM, TN><<<grid, S_THREADS, S_SMEM_BYTES, st>>>(a);
    else if (!aKC && bKC)
        gemm_dmma_small_kernel<false, true, T, TM, TN><<<grid, S_THREADS, S_SMEM_BYTES, st>>>(a);
    else
        gemm_dmma_small_kernel<false, false, T, TM, TN><<<grid, S_THREADS, S_SMEM_BYTES, st>>>(a);
}

// diagnostics: PLMC_SMALL_GEMM=0 in the environment keeps every product on the 128 x 128 kernel
static bool small_gemm_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PLMC_SMALL_GEMM");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
// products of at most this many 128-tiles (over the batch) take the 32 x 128 kernel (PLMC_SMALL_GEMM_TILES)
static long long small_gemm_tiles() {
    static long long v = -1;
    if (v < 0) {
        const char* e = getenv("PLMC_SMALL_GEMM_TILES");
        v = e ? atoll(e) : 296;
        if (v < 0) v = 0;
    }
    return v;
}

// the >48 KB shared-memory opt-in is a per-device function attribute: done lazily, once per device ordinal
static bool g_gemm_attr_done[64];
static int gemm_attrs_for_current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return PLMC_ERR_LAUNCH;
    if (!g_gemm_attr_done[dev]) {
        if (gemm_init_attrs() != PLMC_OK) return PLMC_ERR_LAUNCH;
        g_gemm_attr_done[dev] = true;
    }
    return PLMC_OK;
}

int gemm_launch(bool aKC, bool bKC, const GemmArgs& a, int batch, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || batch <= 0) return PLMC_OK;
    if (int rc = gemm_attrs_for_current_device()) return rc;
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
    // at most two waves of 128-tiles: 32 x 128 CTA tiles (four times the CTAs, two per SM); when C aliases B (M = K =
    // 128: the leaves of the left-sided solves / multiplies) the transposed tiling 128 x 32, whose CTAs read only
    // their own columns of B
    const long long tiles_full = a.lower ? tiles : tm * tn;
    const bool alias_b = (a.C == a.B);
    if (small_gemm_enabled() && tiles_full * batch <= small_gemm_tiles() && (!alias_b || (a.M == 128 && a.C != a.A))) {
        const bool tri = a.triA || a.triB;
        if (!alias_b) {
            dim3 grid_s((unsigned)((a.M / 32) * tn), 1, (unsigned)batch);
            if (tri) launch_small<true, 32, 128>(aKC, bKC, grid_s, st, a);
            else launch_small<false, 32, 128>(aKC, bKC, grid_s, st, a);
        } else {
            dim3 grid_s((unsigned)(tm * (a.N / 32)), 1, (unsigned)batch);
            if (tri) launch_small<true, 128, 32>(aKC, bKC, grid_s, st, a);
            else launch_small<false, 128, 32>(aKC, bKC, grid_s, st, a);
        }
        PLMC_CHECK_LAUNCH();
        note_launch(1, 2.0 * (double)tiles * G_BM * G_BN * (double)a.K * batch);
        return PLMC_OK;
    }
    dim3 grid((unsigned)tiles, 1, (unsigned)batch);
    if (a.triA || a.triB)
        launch_variant<true>(aKC, bKC, grid, st, a);
    else
        launch_variant<false>(aKC, bKC, grid, st, a);
    PLMC_CHECK_LAUNCH();
    note_launch(1, 2.0 * (double)tiles * G_BM * G_BN * (double)a.K * batch);
    return PLMC_OK;
}

}  // namespace plmc
