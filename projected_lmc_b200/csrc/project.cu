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
// Streaming versions (the common case: the whole T^T [q, p] fits shared memory, q <= 32, p even).
// One warp owns 32 consecutive rows of Y at a time: it stages a 32 x 32 slab with 16-byte loads (16 lanes cover
// the 256 contiguous bytes of a row chunk; with p = 32 the whole slab is one contiguous 8 KB block), then lane r
// reduces row r against T from shared memory.  No CTA-wide barrier in the loop, Y is read exactly once for all
// latents, TY is written with full 256-byte lines per latent.
// ---------------------------------------------------------------------------------------------------------
constexpr int PS_WARPS = 8;
constexpr int PS_LD = 33;   // slab row stride (doubles): lane r reading column t hits bank (33 r + t) -> conflict free

__device__ __forceinline__ void load_slab(const double* __restrict__ Y, long long n, int p, long long r0, int p0,
                                          int pc, double* slab, int lane) {
    // 32 rows x 32 columns as 512 double2: k-th load of a lane covers row (k*2 + lane/16), column pair lane%16
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const int r = 2 * k + (lane >> 4), c = (lane & 15) * 2;
        double2 v = make_double2(0.0, 0.0);
        const long long gr = r0 + r;
        if (gr < n && c < pc) {
            const double* src = Y + gr * p + p0 + c;
            if (c + 1 < pc) v = *reinterpret_cast<const double2*>(src);
            else v.x = src[0];
        }
        slab[r * PS_LD + c] = v.x;
        slab[r * PS_LD + c + 1] = v.y;
    }
}

// TY[l, i] = sum_t T[t, l] Y[i, t];  Tt in shared memory as [q][p] (latent-major: the inner loop walks t)
__global__ void __launch_bounds__(PS_WARPS * 32) project_fwd_stream_kernel(const double* __restrict__ Y,
                                                                           const double* __restrict__ T,
                                                                           double* __restrict__ TY, long long n, int p,
                                                                           int q, long long ldty) {
    extern __shared__ __align__(16) double psm[];
    double* Tt = psm;                                  // [q][p]
    double* slabs = psm + (size_t)q * p;               // [PS_WARPS][32][PS_LD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < p * q; idx += PS_WARPS * 32) {
        const int t = idx / q, l = idx - t * q;
        Tt[l * p + t] = T[idx];
    }
    __syncthreads();
    double* slab = slabs + warp * 32 * PS_LD;
    const long long nblk = (n + 31) / 32;
    for (long long blk = (long long)blockIdx.x * PS_WARPS + warp; blk < nblk; blk += (long long)gridDim.x * PS_WARPS) {
        const long long r0 = blk * 32;
        double acc[32];
#pragma unroll
        for (int l = 0; l < 32; ++l) acc[l] = 0.0;
        for (int p0 = 0; p0 < p; p0 += 32) {
            const int pc = min(32, p - p0);
            __syncwarp();
            load_slab(Y, n, p, r0, p0, pc, slab, lane);
            __syncwarp();
            const double* yr = slab + lane * PS_LD;
            for (int t = 0; t < pc; ++t) {
                const double y = yr[t];
#pragma unroll
                for (int l = 0; l < 32; ++l)
                    if (l < q) acc[l] = fma(y, Tt[l * p + p0 + t], acc[l]);
            }
        }
        if (r0 + lane < n) {
#pragma unroll
            for (int l = 0; l < 32; ++l)
                if (l < q) TY[(long long)l * ldty + r0 + lane] = acc[l];
        }
    }
}

// partial[c, t, l] = sum over the rows of CTA c of Y[i, t] G[l, i]: lane t of a warp accumulates column p0 + t of
// the slab against the 32 G values of each latent (held one per lane, broadcast by shuffle)
__global__ void __launch_bounds__(PS_WARPS * 32) project_bwd_stream_kernel(const double* __restrict__ Y,
                                                                           const double* __restrict__ G, long long ldg,
                                                                           double* __restrict__ partial, long long n,
                                                                           int p, int q) {
    extern __shared__ __align__(16) double psm[];
    double* red = psm;                                  // [PS_WARPS][q][32] per p-chunk reduction buffer
    double* slabs = psm + (size_t)PS_WARPS * q * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* slab = slabs + warp * 32 * PS_LD;
    const long long nblk = (n + 31) / 32;
    double* out = partial + (long long)blockIdx.x * p * q;
    for (int p0 = 0; p0 < p; p0 += 32) {
        const int pc = min(32, p - p0);
        double acc[32];
#pragma unroll
        for (int l = 0; l < 32; ++l) acc[l] = 0.0;
        for (long long blk = (long long)blockIdx.x * PS_WARPS + warp; blk < nblk;
             blk += (long long)gridDim.x * PS_WARPS) {
            const long long r0 = blk * 32;
            __syncwarp();
            load_slab(Y, n, p, r0, p0, pc, slab, lane);
            double g[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) g[l] = (l < q && r0 + lane < n) ? G[(long long)l * ldg + r0 + lane] : 0.0;
            __syncwarp();
            for (int r = 0; r < 32; ++r) {
                const double y = slab[r * PS_LD + lane];
#pragma unroll
                for (int l = 0; l < 32; ++l)
                    if (l < q) acc[l] = fma(y, __shfl_sync(0xffffffffu, g[l], r), acc[l]);
            }
        }
        // fixed-order sum over the 8 warps of the CTA
#pragma unroll
        for (int l = 0; l < 32; ++l)
            if (l < q) red[(warp * q + l) * 32 + lane] = acc[l];
        __syncthreads();
        for (int idx = threadIdx.x; idx < q * 32; idx += PS_WARPS * 32) {
            const int l = idx >> 5, t = idx & 31;
            if (t < pc) {
                double s2 = 0.0;
                for (int w = 0; w < PS_WARPS; ++w) s2 += red[(w * q + l) * 32 + t];
                out[(long long)(p0 + t) * q + l] = s2;
            }
        }
        __syncthreads();
    }
}

static inline bool stream_ok(int p, int q) {
    return q <= 32 && (p % 2 == 0) && ((size_t)q * p + (size_t)PS_WARPS * 32 * (PS_LD + q)) * 8 <= 200 * 1024;
}
static inline int stream_ctas(long long n) {
    const long long need = (n + 32 * PS_WARPS - 1) / (32 * PS_WARPS);
    return (int)(need < 296 ? need : 296);
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
    if (stream_ok(p, q) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0)) {
        const size_t smem = ((size_t)q * p + (size_t)PS_WARPS * 32 * PS_LD) * 8;
        cudaFuncSetAttribute(project_fwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        project_fwd_stream_kernel<<<stream_ctas(n), PS_WARPS * 32, smem, (cudaStream_t)stream>>>(Y, T, TY, n, p, q, ldty);
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
    if (stream_ok(p, q) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0)) {
        const int ctas = stream_ctas(n);
        const size_t smem = ((size_t)PS_WARPS * q * 32 + (size_t)PS_WARPS * 32 * PS_LD) * 8;
        cudaFuncSetAttribute(project_bwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        project_bwd_stream_kernel<<<ctas, PS_WARPS * 32, smem, (cudaStream_t)stream>>>(Y, G, ldg, partial, n, p, q);
        PLMC_CHECK_LAUNCH();
        const long long elems = (long long)p * q;
        reduce_chunks_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partial, dT, elems, ctas);
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
