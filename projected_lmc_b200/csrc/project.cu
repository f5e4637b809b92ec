// Kernel (1): the Y -> latent projection  TY = T^T Y^T  and its adjoint.
//
// Replaces ProjectedGPModel.project_data (projected_lmc.py:1014-1021), whose
// two matmuls + triangular solve are algebraically TY = T^T Y^T with
// T = projection_matrix() (:1003-1012).  T is tiny and is formed on the host
// (autograd keeps dT -> dH, dM, dnoise there); this file is the HBM-bound part:
// one streaming pass over Y [n, p].
//
// Algorithmic bytes: 8*n*(p+q) forward, the same backward.
#include "plmc_common.cuh"

namespace plmc {

constexpr int PJ_ROWS = 64;   // rows of Y per tile
constexpr int PJ_PC = 32;     // task chunk
constexpr int PJ_QC = 32;     // latent chunk (gridDim.y)
constexpr int PJ_THREADS = 256;

// TY[l, i] = sum_t T[t, l] * Y[i, t]
__global__ void __launch_bounds__(PJ_THREADS) project_fwd_kernel(const double* __restrict__ Y,
                                                                 const double* __restrict__ T,
                                                                 double* __restrict__ TY, long long n, int p, int q,
                                                                 long long ldty) {
    __shared__ double Ys[PJ_ROWS][PJ_PC + 1];
    __shared__ double Ts[PJ_PC][PJ_QC];
    const int tid = threadIdx.x;
    const long long r0 = (long long)blockIdx.x * PJ_ROWS;
    const int q0 = blockIdx.y * PJ_QC;
    const int qc = min(PJ_QC, q - q0);
    const int row = tid & 63, lg = tid >> 6;
    double acc[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) acc[a] = 0.0;

    for (int p0 = 0; p0 < p; p0 += PJ_PC) {
        const int pc = min(PJ_PC, p - p0);
        __syncthreads();
        for (int idx = tid; idx < PJ_ROWS * pc; idx += PJ_THREADS) {
            const int r = idx / pc, c = idx - r * pc;
            const long long gr = r0 + r;
            Ys[r][c] = (gr < n) ? Y[gr * p + p0 + c] : 0.0;
        }
        for (int idx = tid; idx < PJ_PC * PJ_QC; idx += PJ_THREADS) {
            const int t = idx >> 5, l = idx & 31;
            Ts[t][l] = (t < pc && l < qc) ? T[(long long)(p0 + t) * q + q0 + l] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < pc; ++t) {
            const double y = Ys[row][t];
#pragma unroll
            for (int a = 0; a < 8; ++a) acc[a] = fma(y, Ts[t][lg + 4 * a], acc[a]);
        }
    }
    const long long gr = r0 + row;
    if (gr < n) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int l = lg + 4 * a;
            if (l < qc) TY[(long long)(q0 + l) * ldty + gr] = acc[a];
        }
    }
}

// partial[c, t, l] = sum_{i in chunk c} Y[i, t] * G[l, i]
__global__ void __launch_bounds__(PJ_THREADS) project_bwd_kernel(const double* __restrict__ Y,
                                                                 const double* __restrict__ G, long long ldg,
                                                                 double* __restrict__ partial, long long n, int p,
                                                                 int q, long long rows_per_cta) {
    __shared__ double Ys[PJ_ROWS][PJ_PC + 1];
    __shared__ double Gs[PJ_QC][PJ_ROWS + 1];
    const int tid = threadIdx.x;
    const long long rbeg = (long long)blockIdx.x * rows_per_cta;
    const long long rend = min(n, rbeg + rows_per_cta);
    const int q0 = blockIdx.y * PJ_QC;
    const int qc = min(PJ_QC, q - q0);
    const int tt = tid & 31, lg = tid >> 5;
    double* out = partial + (long long)blockIdx.x * p * q;

    for (int p0 = 0; p0 < p; p0 += PJ_PC) {
        const int pc = min(PJ_PC, p - p0);
        double acc[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a] = 0.0;
        for (long long r0 = rbeg; r0 < rend; r0 += PJ_ROWS) {
            __syncthreads();
            for (int idx = tid; idx < PJ_ROWS * pc; idx += PJ_THREADS) {
                const int r = idx / pc, c = idx - r * pc;
                const long long gr = r0 + r;
                Ys[r][c] = (gr < rend) ? Y[gr * p + p0 + c] : 0.0;
            }
            for (int idx = tid; idx < PJ_QC * PJ_ROWS; idx += PJ_THREADS) {
                const int l = idx >> 6, r = idx & 63;
                const long long gr = r0 + r;
                Gs[l][r] = (gr < rend && l < qc) ? G[(long long)(q0 + l) * ldg + gr] : 0.0;
            }
            __syncthreads();
#pragma unroll 4
            for (int r = 0; r < PJ_ROWS; ++r) {
                const double y = Ys[r][tt];
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[a] = fma(y, Gs[lg + 8 * a][r], acc[a]);
            }
        }
        if (tt < pc) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int l = lg + 8 * a;
                if (l < qc) out[(long long)(p0 + tt) * q + q0 + l] = acc[a];
            }
        }
    }
}

// dT[e] = sum_c partial[c, e]   (fixed order -> deterministic)
__global__ void reduce_chunks_kernel(const double* __restrict__ partial, double* __restrict__ out, long long elems,
                                     int chunks) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += partial[(long long)c * elems + e];
    out[e] = s;
}

// ---------------------------------------------------------------------------------------------------------
// Streaming versions (q <= 32): both products on the FP64 tensor cores, Y read exactly once with asynchronous copies
// (16-byte pieces when p is even and Y is 16-byte aligned, 8-byte pieces otherwise).
// A warp owns 32 consecutive rows of Y at a time and walks their columns in chunks of 32: each (row block, column
// chunk) slab is brought into shared memory with cp.async (zero-filled beyond n and p) while the previous slab is
// multiplied -- two slabs per warp in flight, no CTA-wide barrier in the loop.
//   forward   TY[l, i] = sum_t T[t, l] Y[i, t]:   A = slab (rows i, k = t),  B[k][n] = T[t][l] read through L1;
//   backward  dT[t, l] = sum_i Y[i, t] G[l, i]:   A[m][k] = slab^T (m = t, k = i),  B[k][n] = G[l][i] staged beside the slab.
// Slab row stride 36 doubles (= 4 mod 16): 16-byte aligned rows and conflict-free DMMA fragment reads both ways.
// ---------------------------------------------------------------------------------------------------------
constexpr int PS_WARPS = 4;
constexpr int PS_LD = 36;

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, int src_bytes) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes));
}

// 32 rows x 32 columns of Y (rows r0.., columns p0..) as 512 16-byte pieces, 16 per lane; pieces outside the matrix
// are zero-filled (source size 0)
template <bool A16>
__device__ __forceinline__ void slab_async(const double* __restrict__ Y, long long n, int p, long long r0, int p0,
                                           double* slab, int lane) {
    if (A16) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int r = 2 * k + (lane >> 4), c = (lane & 15) * 2;
            const long long gr = r0 + r;
            const bool ok = gr < n && p0 + c < p;       // p is even: a piece is inside or outside as a whole
            cp_async16(slab + r * PS_LD + c, ok ? Y + gr * p + p0 + c : Y, ok ? 16 : 0);
        }
    } else {                                            // odd p (or an unaligned Y): 8-byte pieces, one row per step
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const long long gr = r0 + r;
            const bool ok = gr < n && p0 + lane < p;
            cp_async8(slab + r * PS_LD + lane, ok ? Y + gr * p + p0 + lane : Y, ok ? 8 : 0);
        }
    }
}

template <int QB, bool A16>
__global__ void __launch_bounds__(PS_WARPS * 32) project_fwd_mma_kernel(const double* __restrict__ Y,
                                                                        const double* __restrict__ T,
                                                                        double* __restrict__ TY, long long n, int p,
                                                                        int q, long long ldty) {
    extern __shared__ __align__(16) double psm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    double* slabs = psm + warp * 2 * 32 * PS_LD;
    const long long nblk = (n + 31) / 32;
    const int nch = (p + 31) / 32;
    const long long w0 = (long long)blockIdx.x * PS_WARPS + warp, wstride = (long long)gridDim.x * PS_WARPS;
    if (w0 >= nblk) return;
    // units = (row block, column chunk) in the order they are consumed; unit u+1 is in flight while u is multiplied
    long long blk = w0;
    int ch = 0, buf = 0;
    slab_async<A16>(Y, n, p, blk * 32, 0, slabs, lane);
    cp_async_commit();
    double c[4][QB][2];
    while (blk < nblk) {
        long long nblk_next = blk;
        int nch_next = ch + 1;
        if (nch_next == nch) { nch_next = 0; nblk_next = blk + wstride; }
        if (nblk_next < nblk)
            slab_async<A16>(Y, n, p, nblk_next * 32, nch_next * 32, slabs + (buf ^ 1) * 32 * PS_LD, lane);
        cp_async_commit();
        if (ch == 0) {
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < QB; ++nb) c[mb][nb][0] = c[mb][nb][1] = 0.0;
        }
        cp_async_wait<1>();
        __syncwarp();
        const double* slab = slabs + buf * 32 * PS_LD;
        const int p0 = ch * 32;
        const int kend = min(32, ((p - p0 + 3) >> 2) << 2);
        for (int k0 = 0; k0 < kend; k0 += 4) {
            double a[4], b[QB];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = slab[(mb * 8 + g) * PS_LD + k0 + t];
            const int tt = p0 + k0 + t;
#pragma unroll
            for (int nb = 0; nb < QB; ++nb) {
                const int l = nb * 8 + g;
                b[nb] = (tt < p && l < q) ? __ldg(T + (long long)tt * q + l) : 0.0;
            }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < QB; ++nb) dmma884(c[mb][nb][0], c[mb][nb][1], a[mb], b[nb]);
        }
        __syncwarp();                                   // everyone is done with this buffer before it is refilled
        if (ch == nch - 1) {
            // the consumed buffer is free until the next fetch: turn the accumulator fragments into [latent][row]
            // there, so that every latent's 32 values leave as one 256-byte line
            double* stage = slabs + buf * 32 * PS_LD;
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < QB; ++nb) {
                    stage[(nb * 8 + 2 * t) * PS_LD + mb * 8 + g] = c[mb][nb][0];
                    stage[(nb * 8 + 2 * t + 1) * PS_LD + mb * 8 + g] = c[mb][nb][1];
                }
            __syncwarp();
            const long long gr = blk * 32 + lane;
            if (gr < n) {
                for (int l = 0; l < q; ++l) TY[(long long)l * ldty + gr] = stage[l * PS_LD + lane];
            }
            __syncwarp();
        }
        blk = nblk_next;
        ch = nch_next;
        buf ^= 1;
    }
    cp_async_wait<0>();
}

// partial[cta, t, l] = sum over the rows of this CTA of Y[i, t] G[l, i]
template <int QB, bool A16>
__global__ void __launch_bounds__(PS_WARPS * 32) project_bwd_mma_kernel(const double* __restrict__ Y,
                                                                        const double* __restrict__ G, long long ldg,
                                                                        double* __restrict__ partial, long long n,
                                                                        int p, int q) {
    extern __shared__ __align__(16) double psm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    constexpr int SLOT = (32 + QB * 8) * PS_LD;         // Y slab [32][36] followed by the G block [QB*8][36]
    double* slots = psm + warp * 2 * SLOT;
    double* red = psm + PS_WARPS * 2 * SLOT;            // [PS_WARPS][32][QB*8] per column chunk
    const long long nblk = (n + 31) / 32;
    const long long w0 = (long long)blockIdx.x * PS_WARPS + warp, wstride = (long long)gridDim.x * PS_WARPS;
    double* out = partial + (long long)blockIdx.x * p * q;

    auto fetch = [&](long long blk, int p0, double* slot) {
        slab_async<A16>(Y, n, p, blk * 32, p0, slot, lane);
        double* gs = slot + 32 * PS_LD;
        // G block: QB*8 latents x 32 rows, 8-byte pieces (ldg may be odd), zero-filled outside
#pragma unroll
        for (int k = 0; k < QB * 8; ++k) {
            const long long gr = blk * 32 + lane;
            const bool ok = k < q && gr < n;
            cp_async8(gs + k * PS_LD + lane, ok ? G + (long long)k * ldg + gr : G, ok ? 8 : 0);
        }
    };

    for (int p0 = 0; p0 < p; p0 += 32) {
        double c[4][QB][2];
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < QB; ++nb) c[mb][nb][0] = c[mb][nb][1] = 0.0;
        int buf = 0;
        if (w0 < nblk) fetch(w0, p0, slots);
        cp_async_commit();
        for (long long blk = w0; blk < nblk; blk += wstride) {
            if (blk + wstride < nblk) fetch(blk + wstride, p0, slots + (buf ^ 1) * SLOT);
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            const double* slab = slots + buf * SLOT;
            const double* gs = slab + 32 * PS_LD;
#pragma unroll
            for (int k0 = 0; k0 < 32; k0 += 4) {
                double a[4], b[QB];
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) a[mb] = slab[(k0 + t) * PS_LD + mb * 8 + g];   // A[m = column][k = row]
#pragma unroll
                for (int nb = 0; nb < QB; ++nb) b[nb] = gs[(nb * 8 + g) * PS_LD + k0 + t];     // B[k = row][n = latent]
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < QB; ++nb) dmma884(c[mb][nb][0], c[mb][nb][1], a[mb], b[nb]);
            }
            __syncwarp();
            buf ^= 1;
        }
        cp_async_wait<0>();
        // fixed-order sum over the warps of the CTA
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < QB; ++nb) {
                double* r = red + ((warp * 32 + mb * 8 + g) * (QB * 8) + nb * 8 + 2 * t);
                r[0] = c[mb][nb][0];
                r[1] = c[mb][nb][1];
            }
        __syncthreads();
        for (int idx = threadIdx.x; idx < 32 * QB * 8; idx += PS_WARPS * 32) {
            const int tt = idx / (QB * 8), l = idx - tt * (QB * 8);
            if (p0 + tt < p && l < q) {
                double s2 = 0.0;
#pragma unroll
                for (int w = 0; w < PS_WARPS; ++w) s2 += red[(w * 32 + tt) * (QB * 8) + l];
                out[(long long)(p0 + tt) * q + l] = s2;
            }
        }
        __syncthreads();
    }
}

static inline bool stream_ok(int p, int q) { return q <= 32 && p > 0; }
static inline int stream_ctas(long long n) {
    const long long need = (n + 32 * PS_WARPS - 1) / (32 * PS_WARPS);
    return (int)(need < 296 ? need : 296);
}
static inline size_t fwd_mma_smem() { return (size_t)PS_WARPS * 2 * 32 * PS_LD * 8; }
static inline size_t bwd_mma_smem(int qb) {
    return ((size_t)PS_WARPS * 2 * (32 + qb * 8) * PS_LD + (size_t)PS_WARPS * 32 * qb * 8) * 8;
}

static inline int bwd_chunks(long long n) {
    long long c = (n + PJ_ROWS - 1) / PJ_ROWS;
    return (int)(c < 296 ? c : 296);
}

}  // namespace plmc

using namespace plmc;

extern "C" {

int plmc_project_fwd(const double* Y, const double* T, double* TY, long long n, int p, int q, long long ldty,
                     void* stream) {
    if (!Y || !T || !TY || n <= 0 || p <= 0 || q <= 0 || ldty < n) return PLMC_ERR_BADARG;
    if (stream_ok(p, q)) {
        const bool a16 = (p % 2 == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
        const size_t smem = fwd_mma_smem();
        const long long need = (n + 32 * PS_WARPS - 1) / (32 * PS_WARPS);
        const int ctas = (int)(need < 444 ? need : 444);      // three 74 KB CTAs per SM
        cudaStream_t st = (cudaStream_t)stream;
#define PLMC_PFWD2(QB, A)                                                                                           \
    {                                                                                                               \
        cudaFuncSetAttribute(project_fwd_mma_kernel<QB, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        project_fwd_mma_kernel<QB, A><<<ctas, PS_WARPS * 32, smem, st>>>(Y, T, TY, n, p, q, ldty);                   \
    }
#define PLMC_PFWD(QB)             \
    if (a16) PLMC_PFWD2(QB, true) \
    else PLMC_PFWD2(QB, false)
        switch ((q + 7) / 8) {
            case 1: PLMC_PFWD(1) break;
            case 2: PLMC_PFWD(2) break;
            case 3: PLMC_PFWD(3) break;
            default: PLMC_PFWD(4) break;
        }
#undef PLMC_PFWD
#undef PLMC_PFWD2
        PLMC_CHECK_LAUNCH();
        note_launch(1);
        return PLMC_OK;
    }
    dim3 grid((unsigned)((n + PJ_ROWS - 1) / PJ_ROWS), (q + PJ_QC - 1) / PJ_QC);
    project_fwd_kernel<<<grid, PJ_THREADS, 0, (cudaStream_t)stream>>>(Y, T, TY, n, p, q, ldty);
    PLMC_CHECK_LAUNCH();
    note_launch(1);
    return PLMC_OK;
}

long long plmc_project_bwd_ws(long long n, int p, int q) {
    if (n <= 0 || p <= 0 || q <= 0) return 0;
    const int c = bwd_chunks(n) > stream_ctas(n) ? bwd_chunks(n) : stream_ctas(n);
    return (long long)c * p * q * 8;
}

int plmc_project_bwd(const double* Y, const double* G, long long ldg, double* dT, double* partial, long long n, int p,
                     int q, void* stream) {
    if (!Y || !G || !dT || !partial || n <= 0 || p <= 0 || q <= 0 || ldg < n) return PLMC_ERR_BADARG;
    if (stream_ok(p, q)) {
        const bool a16 = (p % 2 == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
        const int ctas = stream_ctas(n);
        const int qb = (q + 7) / 8;
        const size_t smem = bwd_mma_smem(qb);
        cudaStream_t st = (cudaStream_t)stream;
#define PLMC_PBWD2(QB, A)                                                                                           \
    {                                                                                                               \
        cudaFuncSetAttribute(project_bwd_mma_kernel<QB, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        project_bwd_mma_kernel<QB, A><<<ctas, PS_WARPS * 32, smem, st>>>(Y, G, ldg, partial, n, p, q);               \
    }
#define PLMC_PBWD(QB)             \
    if (a16) PLMC_PBWD2(QB, true) \
    else PLMC_PBWD2(QB, false)
        switch (qb) {
            case 1: PLMC_PBWD(1) break;
            case 2: PLMC_PBWD(2) break;
            case 3: PLMC_PBWD(3) break;
            default: PLMC_PBWD(4) break;
        }
#undef PLMC_PBWD
#undef PLMC_PBWD2
        PLMC_CHECK_LAUNCH();
        const long long elems = (long long)p * q;
        reduce_chunks_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(partial, dT, elems, ctas);
        PLMC_CHECK_LAUNCH();
        note_launch(2);
        return PLMC_OK;
    }
    const int chunks = bwd_chunks(n);
    long long rows = (n + chunks - 1) / chunks;
    rows = ((rows + PJ_ROWS - 1) / PJ_ROWS) * PJ_ROWS;
    const int used = (int)((n + rows - 1) / rows);
    dim3 grid(used, (q + PJ_QC - 1) / PJ_QC);
    project_bwd_kernel<<<grid, PJ_THREADS, 0, (cudaStream_t)stream>>>(Y, G, ldg, partial, n, p, q, rows);
    PLMC_CHECK_LAUNCH();
    note_launch(1);
    const long long elems = (long long)p * q;
    reduce_chunks_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, dT, elems, used);
    PLMC_CHECK_LAUNCH();
    note_launch(1);
    return PLMC_OK;
}
}
