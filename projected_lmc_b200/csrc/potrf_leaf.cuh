// 128x128 Cholesky leaf: one CTA per batch member, matrix resident in shared memory.
//
// Replaces the inner potrf of torch.linalg.cholesky_ex that gpytorch's
// psd_safe_cholesky runs (reached from projected_lmc.py:1201) and also emits
// inv(L_leaf) into Dinv so that every TRSM leaf of the recursion is a GEMM.
//
// The leaf sits on the critical path of the blocked factorisation (n/128 of them
// run back to back), so it is organised around the FP64 tensor core as well:
//   * right-looking panels of 16 columns: the 16x16 diagonal block is factored
//     by one warp in registers (shuffles), the rows below by one thread per row,
//     and the rank-16 trailing update runs on DMMA.8x8x4 over 8x8 shared-memory
//     tiles (lower tiles only);
//   * inv(L) by recursive doubling (8 -> 16 -> 32 -> 64 -> 128): at each level
//     X21 = -X22 (L21 X11), two triangular b x b products per pair on DMMA, using the
//     free upper triangle as scratch.
// Shared row stride 132 doubles makes row- and column-direction fragment reads
// bank-conflict free.
#pragma once
#include "kernel_math.cuh"

namespace plmc {

constexpr int LEAF = 128;
constexpr int LF_LD = 132;
constexpr int LEAF_THREADS = 256;
constexpr int LF_NB = 16;
constexpr int LEAF_SMEM = LEAF * LF_LD * 8 + 16 + LEAF * 8;   // matrix, failure flag, reciprocals of the diagonal

// sqrt_and_reciprocal (csrc/kernel_math.cuh): pivot root and reciprocal from ONE coupled Goldschmidt iteration, ~70
// cycles of dependent FP64 work.  The library sqrt() followed by a division is ~350 cycles, and it sits on the one
// chain of the whole factorisation that nothing can overlap: the pivot of column c+1 needs the scaled column c (ncu: 53 %
// of the leaf's time was the 16 x 16 diagonal block factorisation, seven of eight warps waiting at its barrier).

__global__ void __launch_bounds__(LEAF_THREADS, 1)
    potrf_leaf_kernel(double* __restrict__ Abase, long long ld, long long sA, double* __restrict__ Dbase,
                      long long sD, int* __restrict__ info, int row_off) {
    extern __shared__ __align__(16) double S[];
    int* fail_sh = reinterpret_cast<int*>(S + LEAF * LF_LD);
    double* rdiag = S + LEAF * LF_LD + 2;            // 1 / L_kk, written by the diagonal-block warp
    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    double* A = Abase + (long long)b * sA;
    double* D = Dbase + (long long)b * sD;

    // ---- load the lower triangle (16-byte chunks; the diagonal chunk may carry one upper element) ----
    for (int idx = tid; idx < LEAF * (LEAF / 2); idx += LEAF_THREADS) {
        const int r = idx >> 6, c = (idx & 63) * 2;
        if (c <= r)
            *reinterpret_cast<double2*>(S + r * LF_LD + c) = *reinterpret_cast<const double2*>(A + (long long)r * ld + c);
    }
    if (tid == 0) *fail_sh = 0;
    __syncthreads();

    // ---- blocked right-looking factorisation ----------------------------------------------------------
    for (int j0 = 0; j0 < LEAF; j0 += LF_NB) {
        // (a1) 16x16 diagonal block: lane r holds row r, column loop with shuffles
        if (warp == 0) {
            double a[LF_NB];
            const int r = lane & (LF_NB - 1);
#pragma unroll
            for (int c = 0; c < LF_NB; ++c) a[c] = (c <= r) ? S[(j0 + r) * LF_LD + j0 + c] : 0.0;
            int bad = 0;
#pragma unroll
            for (int c = 0; c < LF_NB; ++c) {
                const double d = __shfl_sync(0xffffffffu, a[c], c);
                if (!(d > 0.0) && bad == 0) bad = j0 + c + 1;
                double l, inv;
                sqrt_and_reciprocal(d, l, inv);
                if (lane == 0) rdiag[j0 + c] = inv;
                a[c] = (r == c) ? l : a[c] * inv;
#pragma unroll
                for (int cc = c + 1; cc < LF_NB; ++cc) {
                    const double lcc = __shfl_sync(0xffffffffu, a[c], cc);
                    if (r >= cc) a[cc] = fma(-a[c], lcc, a[cc]);
                }
            }
            if (lane < LF_NB) {
#pragma unroll
                for (int c = 0; c < LF_NB; ++c)
                    if (c <= r) S[(j0 + r) * LF_LD + j0 + c] = a[c];
            }
            if (lane == 0 && bad && *fail_sh == 0) *fail_sh = bad;
        }
        __syncthreads();
        const int r0 = j0 + LF_NB;
        if (r0 >= LEAF) break;
        // (a2) rows below the block: x L11^T = a, one thread per row
        if (tid < LEAF - r0) {
            double* row = S + (r0 + tid) * LF_LD + j0;
            double x[LF_NB];
#pragma unroll
            for (int c = 0; c < LF_NB; ++c) x[c] = row[c];
#pragma unroll
            for (int c = 0; c < LF_NB; ++c) {
                const double* lrow = S + (j0 + c) * LF_LD + j0;
                double s = x[c];
#pragma unroll
                for (int k = 0; k < c; ++k) s = fma(-x[k], lrow[k], s);
                x[c] = s * rdiag[j0 + c];
            }
#pragma unroll
            for (int c = 0; c < LF_NB; ++c) row[c] = x[c];
        }
        __syncthreads();
        // (b) trailing update C -= P P^T on DMMA, lower 8x8 tiles, one tile row per warp pass
        const int T = (LEAF - r0) / 8;
        const int ntiles = T * (T + 1) / 2;
        for (int tile = warp; tile < ntiles; tile += LEAF_THREADS / 32) {
            int ti = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
            while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
            while (ti * (ti + 1) / 2 > tile) --ti;
            const int tk = tile - ti * (ti + 1) / 2;
            const int m0 = r0 + 8 * ti, n0 = r0 + 8 * tk;
            double2* cp = reinterpret_cast<double2*>(S + (m0 + g) * LF_LD + n0 + 2 * t);
            double2 c = *cp;
            const double* pa = S + (m0 + g) * LF_LD + j0 + t;
            const double* pb = S + (n0 + g) * LF_LD + j0 + t;
#pragma unroll
            for (int kk = 0; kk < LF_NB / 4; ++kk) dmma884(c.x, c.y, -pa[4 * kk], pb[4 * kk]);
            *cp = c;
        }
        __syncthreads();
    }
    const int fail = *fail_sh;
    if (fail && tid == 0 && info[b] == 0) info[b] = row_off + fail;

    // ---- write back L (lower part only) -----------------------------------------------------------------
    for (int idx = tid; idx < LEAF * LEAF; idx += LEAF_THREADS) {
        const int r = idx >> 7, c = idx & 127;
        if (c <= r) A[(long long)r * ld + c] = S[r * LF_LD + c];
    }
    __syncthreads();

    // ---- inverse: 8x8 diagonal blocks directly, then recursive doubling on DMMA -------------------------
    if (tid < LEAF) {
        const int blk = tid >> 3, j = tid & 7;
        const double* Lb = S + (blk * 8) * LF_LD + blk * 8;
        double l[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) l[i][k] = (k <= i) ? Lb[i * LF_LD + k] : 0.0;
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < i) s = fma(-l[i][k], (k >= j) ? x[k] : 0.0, s);
            x[i] = (i >= j) ? s * rdiag[blk * 8 + i] : 0.0;
        }
        __syncwarp();  // all 8 columns of a block live in one warp: reads above are done
#pragma unroll
        for (int i = 0; i < 8; ++i) S[(blk * 8 + i) * LF_LD + blk * 8 + j] = x[i];   // zeros above the diagonal
    }
    __syncthreads();
    for (int bs = 8; bs < LEAF; bs *= 2) {
        const int tb = bs / 8;                 // 8x8 tiles per block side
        const int pairs = LEAF / (2 * bs);
        const int ntl = pairs * tb * tb;
        // step 1: T = L21 * X11  -> upper-right block of the pair (scratch)
        for (int tile = warp; tile < ntl; tile += LEAF_THREADS / 32) {
            const int p = tile / (tb * tb), rem = tile - p * tb * tb;
            const int it = rem / tb, jt = rem - it * tb;
            const int r = p * 2 * bs;
            double c0 = 0.0, c1 = 0.0;
            // T[i][j] = sum_{k >= j} L21[i][k] X11[k][j]
            const double* pa = S + (r + bs + 8 * it + g) * LF_LD + r + t;
            const double* pb = S + (r + t) * LF_LD + r + 8 * jt + g;
            for (int k0 = 8 * jt; k0 < bs; k0 += 4) dmma884(c0, c1, pa[k0], pb[k0 * LF_LD]);
            *reinterpret_cast<double2*>(S + (r + 8 * it + g) * LF_LD + r + bs + 8 * jt + 2 * t) = make_double2(c0, c1);
        }
        __syncthreads();
        // step 2: X21 = -X22 * T  -> overwrites L21
        for (int tile = warp; tile < ntl; tile += LEAF_THREADS / 32) {
            const int p = tile / (tb * tb), rem = tile - p * tb * tb;
            const int it = rem / tb, jt = rem - it * tb;
            const int r = p * 2 * bs;
            double c0 = 0.0, c1 = 0.0;
            // X21[i][j] = -sum_{k <= i} X22[i][k] T[k][j]
            const double* pa = S + (r + bs + 8 * it + g) * LF_LD + r + bs + t;
            const double* pb = S + (r + t) * LF_LD + r + bs + 8 * jt + g;
            for (int k0 = 0; k0 < 8 * it + 8; k0 += 4) dmma884(c0, c1, -pa[k0], pb[k0 * LF_LD]);
            *reinterpret_cast<double2*>(S + (r + bs + 8 * it + g) * LF_LD + r + 8 * jt + 2 * t) = make_double2(c0, c1);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < LEAF * LEAF; idx += LEAF_THREADS) {
        const int r = idx >> 7, c = idx & 127;
        D[idx] = (c <= r) ? S[r * LF_LD + c] : 0.0;
    }
}

}  // namespace plmc
