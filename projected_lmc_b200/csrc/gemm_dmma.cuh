// FP64 tensor-core GEMM for sm_100a:  C = alpha * op(A) * op(B) + beta * C
//
// B200 has no tcgen05 kind for f64; FP64 matrix math is the warp-level
// mma.sync.m8n8k4 (SASS DMMA.8x8x4).  This kernel is the single engine behind
// the blocked Cholesky / TRSM / TRTRI / LAUUM of the projected-LMC path
// (reference: the potrf / triangular solves that gpytorch's log_prob runs,
// projected_lmc.py:1201).
//
// Shape contract (guaranteed by the padding the Gram builder applies):
//   M % 128 == 0, N % 128 == 0, K % 16 == 0, all leading dims % 2 == 0,
//   all base pointers 16-byte aligned.  No bounds checks in the hot loop.
//
// CTA tile 128x128x16, 8 warps (2 x 4), warp tile 64x32 -> 32 DMMA per k4 step
// per warp against 12 LDS.64: the kernel is DMMA-issue bound by construction.
// Operands are staged with 16-byte cp.async into a 4-stage shared-memory ring.
// Shared layouts are padded so that every fragment load is bank-conflict free:
//   k-contiguous operand  -> tile[128][16+4]
//   m/n-contiguous operand-> tile[16][128+4]
#pragma once
#include "plmc_common.cuh"

namespace plmc {

constexpr int G_BM = 128, G_BN = 128, G_BK = 16;
constexpr int G_THREADS = 256;
constexpr int G_STAGES = 4;
constexpr int G_LDK = G_BK + 4;    // 20
constexpr int G_LDM = G_BM + 4;    // 132
constexpr int G_TILE = 128 * G_LDK;  // 2560 doubles >= 16*132
constexpr int G_STAGE = 2 * G_TILE;  // A + B
constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE * 8;  // 163840

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

// Load one 128 x 16 operand tile.  KC: element (x, k) at P[(x0+x)*ld + k0+k].
// MC: element (x, k) at P[(k0+k)*ld + x0+x].  `tri` masks stored col > stored row.
template <bool KC>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ P, long long ld, int x0, int k0,
                                          int tri, int tid) {
    if (KC) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * G_THREADS;
            const int row = c >> 3, ch = c & 7;
            const int gr = x0 + row, gc = k0 + ch * 2;
            int bytes = 16;
            if (tri) {
                int v = gr - gc + 1;
                v = v < 0 ? 0 : (v > 2 ? 2 : v);
                bytes = v * 8;
            }
            cp_async16(s + row * G_LDK + ch * 2, P + (long long)gr * ld + gc, bytes);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * G_THREADS;
            const int krow = c >> 6, ch = c & 63;
            const int gr = k0 + krow, gc = x0 + ch * 2;
            int bytes = 16;
            if (tri) {
                int v = gr - gc + 1;
                v = v < 0 ? 0 : (v > 2 ? 2 : v);
                bytes = v * 8;
            }
            cp_async16(s + krow * G_LDM + ch * 2, P + (long long)gr * ld + gc, bytes);
        }
    }
}

template <bool A_KC, bool B_KC>
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

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = p.K / G_BK;

    // ---- prologue ---------------------------------------------------------
#pragma unroll
    for (int s = 0; s < G_STAGES - 1; ++s) {
        if (s < nk) {
            double* sa = smem + s * G_STAGE;
            load_tile<A_KC>(sa, A, p.lda, m0, s * G_BK, p.triA, tid);
            load_tile<B_KC>(sa + G_TILE, B, p.ldb, n0, s * G_BK, p.triB, tid);
        }
        cp_async_commit();
    }

    // ---- main loop --------------------------------------------------------
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<G_STAGES - 2>();
        __syncthreads();
        {
            const int nt = kt + G_STAGES - 1;
            if (nt < nk) {
                double* sa = smem + (nt % G_STAGES) * G_STAGE;
                load_tile<A_KC>(sa, A, p.lda, m0, nt * G_BK, p.triA, tid);
                load_tile<B_KC>(sa + G_TILE, B, p.ldb, n0, nt * G_BK, p.triB, tid);
            }
            cp_async_commit();
        }
        const double* sa = smem + (kt % G_STAGES) * G_STAGE;
        const double* sb = sa + G_TILE;
        const double* pa = A_KC ? sa + (wm * 64 + g) * G_LDK + t : sa + t * G_LDM + wm * 64 + g;
        const double* pb = B_KC ? sb + (wn * 32 + g) * G_LDK + t : sb + t * G_LDM + wn * 32 + g;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = A_KC ? pa[i * 8 * G_LDK + kk * 4] : pa[kk * 4 * G_LDM + i * 8];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = B_KC ? pb[j * 8 * G_LDK + kk * 4] : pb[kk * 4 * G_LDM + j * 8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ---------------------------------------------------------
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + wm * 64 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn * 32 + j * 8 + 2 * t;
            double2* dst = reinterpret_cast<double2*>(C + (long long)row * p.ldc + col);
            double2 o;
            o.x = alpha * acc[i][j][0];
            o.y = alpha * acc[i][j][1];
            if (beta != 0.0) {
                const double2 c = *dst;
                o.x += beta * c.x;
                o.y += beta * c.y;
            }
            *dst = o;
        }
    }
}

enum GemmLayout { A_KC_B_KC = 0, A_KC_B_NC = 1, A_MC_B_KC = 2, A_MC_B_NC = 3 };

int gemm_init_attrs();
void stats_get(long long* launches, long long* gemm_launches, double* gemm_flops);
void stats_reset();
// op layouts: aKC -> A(m,k) at A[m*lda+k] else A[k*lda+m];  bKC -> B(k,n) at B[n*ldb+k] else B[k*ldb+n]
int gemm_launch(bool aKC, bool bKC, const GemmArgs& a, int batch, cudaStream_t st);

}  // namespace plmc
