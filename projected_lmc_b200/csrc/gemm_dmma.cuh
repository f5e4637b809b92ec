// FP64 tensor-core GEMM for sm_100a:  C = alpha * op(A) * op(B) + beta * C
//
// B200 has no tcgen05 kind for f64; FP64 matrix math is the warp-level
// mma.sync.m8n8k4 (SASS DMMA.8x8x4).  This kernel is the single engine behind
// the blocked Cholesky / TRSM / TRTRI / LAUUM of the projected-LMC path
// (reference: the potrf / triangular solves that gpytorch's log_prob runs,
// projected_lmc.py:1201).
//
// Shape contract (guaranteed by the padding the Gram builder applies):
//   M % 128 == 0, N % 128 == 0, K % G_BK == 0, all leading dims % 2 == 0,
//   all base pointers 16-byte aligned.  No bounds checks in the hot loop.
//
// CTA tile 128x128xG_BK, 8 warps (2 x 4), warp tile 64x32 -> 32 DMMA per k4 step
// per warp against 12 LDS.64: the kernel is DMMA-issue bound by construction.
// Operands are staged with 16-byte cp.async into a shared-memory ring; all
// per-thread global/shared addresses are loop invariant (one 64-bit add per chunk
// per k-tile) and the two warps that share an SM sub-partition issue their copies
// at different points of the k-tile so the tensor pipe always has a warp feeding it.
// Shared layouts are padded so that every fragment load is bank-conflict free:
//   k-contiguous operand  -> tile[128][G_BK+4]
//   m/n-contiguous operand-> tile[G_BK][128+4]
#pragma once
#include "plmc_common.cuh"

namespace plmc {

#ifndef PLMC_GEMM_BK
#define PLMC_GEMM_BK 32
#endif
#ifndef PLMC_GEMM_STAGES
#define PLMC_GEMM_STAGES 3
#endif

constexpr int G_BM = 128, G_BN = 128, G_BK = PLMC_GEMM_BK;
constexpr int G_THREADS = 256;
constexpr int G_STAGES = PLMC_GEMM_STAGES;
constexpr int G_LDK = G_BK + 4;      // k-contiguous tile row stride (doubles); (G_BK+4) % 16 == 4
constexpr int G_LDM = G_BM + 4;      // m-contiguous tile row stride
constexpr int G_TILE = (128 * G_LDK > G_BK * G_LDM) ? 128 * G_LDK : G_BK * G_LDM;
constexpr int G_STAGE = 2 * G_TILE;  // A + B
constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE * 8;
constexpr int G_NCH = G_BK / 4;      // 16-byte chunks per thread per operand per k-tile
static_assert(G_BK % 16 == 0 && (G_LDK % 16) == 4, "fragment loads must stay conflict free");
static_assert(G_SMEM_BYTES <= 227 * 1024, "shared memory ring too large");

struct GemmArgs {
    const double* A;
    const double* B;
    double* C;
    long long lda, ldb, ldc;
    long long sA, sB, sC;  // batch strides (elements)
    int M, N, K;
    double alpha, beta;
    int lower;  // 1: only tiles with tile_row >= tile_col are computed (C origin on the diagonal)
    int triA;   // stored A element (row r, col c) relative to A is zero unless c <= r
    int triB;
};

__device__ __forceinline__ void cp_async16s(uint32_t smem_addr, const void* gmem_src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(gmem_src), "r"(src_bytes));
}

// Alternative pipeline synchronisation (per-stage mbarriers instead of one CTA barrier per
// k-tile).  Validated on B200 but not faster (33.8 vs 34.2 TFLOP/s at 8192^3): the CTA
// barrier is not what limits the kernel, so the simpler scheme stays the default.
#ifndef PLMC_GEMM_MBAR
#define PLMC_GEMM_MBAR 0
#endif

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
// arrive on `bar` once all cp.async copies previously issued by this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}

// Per-thread copy plan for one operand: G_NCH chunks per k-tile.
//   KC: element (x, k) at P[(x0+x)*ld + k],  chunk c -> row c / (G_BK/2), col-pair c % (G_BK/2)
//   MC: element (x, k) at P[k*ld + x0+x],    chunk c -> k-row c / 64,     col-pair c % 64
template <bool KC, bool TRI>
struct TileLoader {
    const double* g0;   // running global pointer of chunk 0 (advances by one k-tile per load)
    long long step;     // elements between consecutive chunks of this thread
    long long adv;      // elements per k-tile
    uint32_t soff;      // byte offset of chunk 0 inside a tile
    int r0, c0;         // stored (row, col) of chunk 0 at k-tile 0 (TRI only)
    bool mask;          // TRI only: operand is lower-triangular (stored col > stored row reads as 0)

    __device__ __forceinline__ void init(const double* __restrict__ P, long long ld, int x0, int tid, bool msk) {
        mask = msk;
        if (KC) {
            constexpr int CPR = G_BK / 2;  // chunks per row
            const int row = tid / CPR, ch = tid % CPR;
            g0 = P + (long long)(x0 + row) * ld + ch * 2;
            step = (long long)(G_THREADS / CPR) * ld;
            soff = (uint32_t)(row * G_LDK + ch * 2) * 8u;
            adv = G_BK;
            r0 = x0 + row;
            c0 = ch * 2;
        } else {
            const int krow = tid >> 6, ch = tid & 63;
            g0 = P + (long long)krow * ld + x0 + ch * 2;
            step = 4 * ld;
            soff = (uint32_t)(krow * G_LDM + ch * 2) * 8u;
            adv = (long long)G_BK * ld;
            r0 = krow;
            c0 = x0 + ch * 2;
        }
    }
    // copy k-tile `kt` into the tile at shared address `sbase`, then advance
    __device__ __forceinline__ void load(uint32_t sbase, int kt) {
        constexpr int CPR = G_BK / 2;
#pragma unroll
        for (int i = 0; i < G_NCH; ++i) {
            const uint32_t dst = sbase + soff + (uint32_t)i * (KC ? (G_THREADS / CPR) * G_LDK * 8u : 4u * G_LDM * 8u);
            int bytes = 16;
            if (TRI) {
                if (mask) {
                    const int gr = KC ? r0 + i * (G_THREADS / CPR) : r0 + i * 4 + kt * G_BK;
                    const int gc = KC ? c0 + kt * G_BK : c0;
                    int v = gr - gc + 1;
                    v = v < 0 ? 0 : (v > 2 ? 2 : v);
                    bytes = v * 8;
                }
            }
            cp_async16s(dst, g0 + i * step, bytes);
        }
        g0 += adv;
    }
};

template <bool A_KC, bool B_KC, bool TRI>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_dmma_kernel(const GemmArgs p) {
    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;

    // ---- tile coordinates -------------------------------------------------
    int ti, tj;
    const int tiles_n = p.N / G_BN;
    if (p.lower) {
        const long long b = blockIdx.x;
        int r = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
        while ((long long)(r + 1) * (r + 2) / 2 <= b) ++r;
        while ((long long)r * (r + 1) / 2 > b) --r;
        ti = r;
        tj = (int)(b - (long long)r * (r + 1) / 2);
    } else {
        // grouped raster: bands of 8 tile-rows walk the columns together so a
        // wave of 148 CTAs reuses both operand panels out of L2
        const int tiles_m = p.M / G_BM;
        const int GROUP = 8;
        const int per_group = GROUP * tiles_n;
        const int gid = blockIdx.x / per_group;
        const int first = gid * GROUP;
        const int gsz = min(tiles_m - first, GROUP);
        const int rem = blockIdx.x - gid * per_group;
        ti = first + rem % gsz;
        tj = rem / gsz;
    }
    const int m0 = ti * G_BM, n0 = tj * G_BN;

    const double* __restrict__ A = p.A + (long long)blockIdx.z * p.sA;
    const double* __restrict__ B = p.B + (long long)blockIdx.z * p.sB;
    double* __restrict__ C = p.C + (long long)blockIdx.z * p.sC;

    TileLoader<A_KC, TRI> la;
    TileLoader<B_KC, TRI> lb;
    la.init(A, p.lda, m0, tid, TRI && p.triA);
    lb.init(B, p.ldb, n0, tid, TRI && p.triB);
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(smem);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = p.K / G_BK;

    auto issue = [&](int kt) {
        const uint32_t sb = sm0 + (uint32_t)(kt % G_STAGES) * (G_STAGE * 8u);
        la.load(sb, kt);
        lb.load(sb + G_TILE * 8u, kt);
    };

#if PLMC_GEMM_MBAR
    // Split arrive/wait pipeline: per-stage "full" mbarriers are completed by the cp.async
    // copies of all 256 threads, per-stage "empty" mbarriers by the 8 consumer warps.  No
    // CTA-wide barrier in the loop: a warp only ever waits for data, or (mid-tile, when it
    // refills a stage) for the other warps to have finished the PREVIOUS tile.
    __shared__ __align__(8) unsigned long long full_bar[G_STAGES];
    __shared__ __align__(8) unsigned long long empty_bar[G_STAGES];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&full_bar[s], G_THREADS);
            mbar_init(&empty_bar[s], G_THREADS / 32);
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < G_STAGES - 1; ++s) {
        if (s < nk) {
            issue(s);
            cp_async_mbar_arrive(&full_bar[s]);
        }
    }
    for (int kt = 0; kt < nk; ++kt) {
        mbar_wait(&full_bar[kt % G_STAGES], (kt / G_STAGES) & 1);
        const double* sa = smem + (kt % G_STAGES) * G_STAGE;
        const double* sb = sa + G_TILE;
        const double* pa = A_KC ? sa + (wm * 64 + g) * G_LDK + t : sa + t * G_LDM + wm * 64 + g;
        const double* pb = B_KC ? sb + (wn * 32 + g) * G_LDK + t : sb + t * G_LDM + wn * 32 + g;
        const int nt = kt + G_STAGES - 1;
#pragma unroll
        for (int kk = 0; kk < G_BK / 4; ++kk) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = A_KC ? pa[i * 8 * G_LDK + kk * 4] : pa[kk * 4 * G_LDM + i * 8];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = B_KC ? pb[j * 8 * G_LDK + kk * 4] : pb[kk * 4 * G_LDM + j * 8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            if (kk == (wm ? G_BK / 8 : 0)) {
                if (nt < nk) {
                    const int rn = nt / G_STAGES;
                    if (rn > 0) mbar_wait(&empty_bar[nt % G_STAGES], (rn - 1) & 1);
                    issue(nt);
                    cp_async_mbar_arrive(&full_bar[nt % G_STAGES]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[kt % G_STAGES]);
    }
#else
    // ---- prologue ---------------------------------------------------------
#pragma unroll
    for (int s = 0; s < G_STAGES - 1; ++s) {
        if (s < nk) issue(s);
        cp_async_commit();
    }

    // ---- main loop --------------------------------------------------------
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();
        const double* sa = smem + (kt % G_STAGES) * G_STAGE;
        const double* sb = sa + G_TILE;
        const double* pa = A_KC ? sa + (wm * 64 + g) * G_LDK + t : sa + t * G_LDM + wm * 64 + g;
        const double* pb = B_KC ? sb + (wn * 32 + g) * G_LDK + t : sb + t * G_LDM + wn * 32 + g;
        const int nt = kt + G_STAGES - 1;
#pragma unroll
        for (int kk = 0; kk < G_BK / 4; ++kk) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = A_KC ? pa[i * 8 * G_LDK + kk * 4] : pa[kk * 4 * G_LDM + i * 8];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = B_KC ? pb[j * 8 * G_LDK + kk * 4] : pb[kk * 4 * G_LDM + j * 8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            // the two warps of an SM sub-partition (wm = 0 / 1) refill the ring at
            // different k4 steps, so one of them is always issuing DMMAs
            if (kk == (wm ? G_BK / 8 : 0)) {
                if (nt < nk) issue(nt);
                cp_async_commit();
            }
        }
    }
    cp_async_wait<0>();
#endif

    // ---- epilogue ---------------------------------------------------------
    // beta != 0: the old values of row group i + 1 are already in flight while group i is combined and stored
    // (8 dependent load round trips per tile would otherwise rival the main loop of a K = 128 update)
    const double alpha = p.alpha, beta = p.beta;
    auto cptr = [&](int i, int j) {
        return reinterpret_cast<double2*>(C + (long long)(m0 + wm * 64 + i * 8 + g) * p.ldc + (n0 + wn * 32 + j * 8 + 2 * t));
    };
    if (beta != 0.0) {
        double2 cn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cn[j] = *cptr(0, j);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double2 c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = cn[j];
            if (i + 1 < 8) {
#pragma unroll
                for (int j = 0; j < 4; ++j) cn[j] = *cptr(i + 1, j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double2 o;
                o.x = alpha * acc[i][j][0];
                o.y = alpha * acc[i][j][1];
                o.x += beta * c[j].x;
                o.y += beta * c[j].y;
                *cptr(i, j) = o;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) *cptr(i, j) = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Small products.  The blocked factorisation has a latency-bound chain of tiny GEMMs (the solves and updates inside
// every diagonal 2048-block, the inverses of the 512- and 2048-blocks): 1400 launches per C4 iteration, 3000 per C2
// iteration, most of them a handful of 128-tiles.  With 128 x 128 CTA tiles such a launch occupies 4-60 SMs for the
// 17 us one SM needs for a 128^3 tile on its FP64 tensor pipe.  This variant cuts the same product into 32 x 128 CTA
// tiles (4 warps, warp tile 32 x 32): four times as many SMs, a quarter of the time.  A CTA reads only ITS 32 rows of
// op(A), so C may alias A under the same condition as above (K a single 128-tile).  The transposed tiling, 128 x 32,
// reads only ITS 32 columns of op(B) and all of op(A): it serves the products in which C aliases B (M = K = 128).
// ---------------------------------------------------------------------------------------------------------
constexpr int S_BK = 32, S_THREADS = 128, S_STAGES = 2;
constexpr int S_LDK = S_BK + 4;       // k-contiguous tiles: [rows][36]
template <int TM, int TN>
struct SmallCfg {                     // CTA tile TM x TN: 32 x 128 (C may alias A) or 128 x 32 (C may alias B)
    static constexpr int LDA = TM + 4, LDN = TN + 4;   // m- / n-contiguous tiles: [32 k][TM + 4], [32 k][TN + 4]
    static constexpr int ATILE = (TM * S_LDK > S_BK * LDA) ? TM * S_LDK : S_BK * LDA;
    static constexpr int BTILE = (TN * S_LDK > S_BK * LDN) ? TN * S_LDK : S_BK * LDN;
    static constexpr int STAGE = ATILE + BTILE;
    static constexpr int SMEM_BYTES = S_STAGES * STAGE * 8;
};
constexpr int S_SMEM_BYTES = SmallCfg<32, 128>::SMEM_BYTES;   // = SmallCfg<128, 32>::SMEM_BYTES

// one operand tile of R rows (R = TM for A, TN for B) and 32 k: 16-byte pieces, zero-filled above the diagonal of a
// triangular operand.  KC: element (x, k) at P[(x0+x) ld + k] -> tile[x][36];  MC: at P[k ld + x0+x] -> tile[k][R+4]
template <int R, bool KC, bool TRI>
__device__ __forceinline__ void small_tile_load(uint32_t sdst, const double* __restrict__ P, long long ld, int x0, int k0,
                                                int tid, bool mask) {
    auto piece = [&](uint32_t dst, const double* src, int r, int c) {
        int bytes = 16;
        if (TRI && mask) {
            int v = r - c + 1;
            v = v < 0 ? 0 : (v > 2 ? 2 : v);
            bytes = v * 8;
        }
        cp_async16s(dst, src, bytes);
    };
    if (KC) {
        const int ch = tid & 15, rr = tid >> 4;        // 16 pieces per 32-wide row, 8 rows per pass
#pragma unroll
        for (int i = 0; i < R / 8; ++i) {
            const int row = rr + 8 * i;
            piece(sdst + (uint32_t)(row * S_LDK + ch * 2) * 8u, P + (long long)(x0 + row) * ld + k0 + ch * 2, x0 + row,
                  k0 + ch * 2);
        }
    } else {
        constexpr int PPR = R / 2;                     // pieces per k-row
        constexpr int RPP = S_THREADS / PPR;           // k-rows per pass
        const int ch = tid % PPR, rr = tid / PPR;
#pragma unroll
        for (int i = 0; i < S_BK / RPP; ++i) {
            const int krow = rr + RPP * i;
            piece(sdst + (uint32_t)(krow * (R + 4) + ch * 2) * 8u, P + (long long)(k0 + krow) * ld + x0 + ch * 2, k0 + krow,
                  x0 + ch * 2);
        }
    }
}

template <bool A_KC, bool B_KC, bool TRI, int TM, int TN>
__global__ void __launch_bounds__(S_THREADS, 2) gemm_dmma_small_kernel(const GemmArgs p) {
    using SC = SmallCfg<TM, TN>;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (TM == 128) ? warp * 32 : 0, wn0 = (TN == 128) ? warp * 32 : 0;   // warp tile 32 x 32
    const int tiles_m = p.M / TM;
    const int ti = blockIdx.x % tiles_m, tj = blockIdx.x / tiles_m;
    const int m0 = ti * TM, n0 = tj * TN;
    if (p.lower && (m0 >> 7) < (n0 >> 7)) return;     // only 128-tiles on or below the diagonal (uniform per CTA)

    const double* __restrict__ A = p.A + (long long)blockIdx.z * p.sA;
    const double* __restrict__ B = p.B + (long long)blockIdx.z * p.sB;
    double* __restrict__ C = p.C + (long long)blockIdx.z * p.sC;
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(smem);
    const bool maskA = TRI && p.triA, maskB = TRI && p.triB;

    auto issue = [&](int kt) {
        const uint32_t sa = sm0 + (uint32_t)(kt & 1) * (SC::STAGE * 8u), sb = sa + SC::ATILE * 8u;
        small_tile_load<TM, A_KC, TRI>(sa, A, p.lda, m0, kt * S_BK, tid, maskA);
        small_tile_load<TN, B_KC, TRI>(sb, B, p.ldb, n0, kt * S_BK, tid, maskB);
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = p.K / S_BK;
    issue(0);
    cp_async_commit();
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) issue(kt + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double* sa = smem + (kt & 1) * SC::STAGE;
        const double* sb = sa + SC::ATILE;
        const double* pa = A_KC ? sa + (wm0 + g) * S_LDK + t : sa + t * SC::LDA + wm0 + g;
        const double* pb = B_KC ? sb + (wn0 + g) * S_LDK + t : sb + t * SC::LDN + wn0 + g;
#pragma unroll
        for (int kk = 0; kk < S_BK / 4; ++kk) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = A_KC ? pa[i * 8 * S_LDK + kk * 4] : pa[kk * 4 * SC::LDA + i * 8];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = B_KC ? pb[j * 8 * S_LDK + kk * 4] : pb[kk * 4 * SC::LDN + j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
    cp_async_wait<0>();

    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2* cp =
                reinterpret_cast<double2*>(C + (long long)(m0 + wm0 + i * 8 + g) * p.ldc + (n0 + wn0 + j * 8 + 2 * t));
            double2 o = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
            if (beta != 0.0) {
                const double2 c = *cp;
                o.x += beta * c.x;
                o.y += beta * c.y;
            }
            *cp = o;
        }
}

enum GemmLayout { A_KC_B_KC = 0, A_KC_B_NC = 1, A_MC_B_KC = 2, A_MC_B_NC = 3 };

int gemm_init_attrs();
void stats_get(long long* launches, long long* gemm_launches, double* gemm_flops);
void stats_reset();
// op layouts: aKC -> A(m,k) at A[m*lda+k] else A[k*lda+m];  bKC -> B(k,n) at B[n*ldb+k] else B[k*ldb+n]
int gemm_launch(bool aKC, bool bKC, const GemmArgs& a, int batch, cudaStream_t st);

}  // namespace plmc
