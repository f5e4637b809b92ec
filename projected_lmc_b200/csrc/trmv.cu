// Single right-hand-side solves of the training step, done as triangular matrix-vector
// products with the explicit inverse factor (available anyway: trtri precedes lauum):
//     z = L^-1 y,   alpha = L^-T z,   |z|^2,   logdet K = -2 sum log (L^-1)_ii
// Replaces the triangular solve / inv_quad / logdet of gpytorch's log_prob
// (projected_lmc.py:1201).  HBM-bound: the lower triangle of L^-1 is read twice
// (8 * q * n(n+1)/2 bytes each); all reductions are fixed-order (deterministic).
#include "plmc_common.cuh"

namespace plmc {

// z[b, i] = sum_{j <= i} Linv[b, i, j] * y[b, j]        one warp per row
__global__ void __launch_bounds__(256) trmv_n_kernel(const double* __restrict__ Linv, long long ld,
                                                     long long stride, const double* __restrict__ y, long long ldy,
                                                     double* __restrict__ z, long long ldv, long long n) {
    const int b = blockIdx.z;
    const int lane = threadIdx.x & 31;
    // heavy (long) rows first so the tail of the grid is made of short rows
    const long long row = n - 1 - ((long long)blockIdx.x * 8 + (threadIdx.x >> 5));
    if (row < 0) return;
    const double* Lr = Linv + (long long)b * stride + row * ld;
    const double* yb = y + (long long)b * ldy;
    double s0 = 0.0, s1 = 0.0;
    const long long len = row + 1;
    const long long len2 = len & ~1LL;
    for (long long j = 2 * lane; j < len2; j += 64) {
        const double2 a = *reinterpret_cast<const double2*>(Lr + j);
        s0 = fma(a.x, yb[j], s0);
        s1 = fma(a.y, yb[j + 1], s1);
    }
    if (lane == 0 && (len & 1)) s0 = fma(Lr[len - 1], yb[len - 1], s0);
    const double s = warp_sum(s0 + s1);
    if (lane == 0) z[(long long)b * ldv + row] = s;
}

// partial[b, c, j] = sum_{i in chunk c, i >= j} Linv[b, i, j] * z[b, i]     thread per column
__global__ void __launch_bounds__(256) trmv_t_kernel(const double* __restrict__ Linv, long long ld,
                                                     long long stride, const double* __restrict__ z, long long ldv,
                                                     double* __restrict__ partial, long long n, long long npad,
                                                     long long rows_per_chunk) {
    const int b = blockIdx.z;
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long c = blockIdx.y;
    const long long r0 = c * rows_per_chunk;
    const long long r1 = min(n, r0 + rows_per_chunk);
    double* out = partial + ((long long)b * gridDim.y + c) * npad;
    if (j >= n) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const double* Lb = Linv + (long long)b * stride + j;
    const double* zb = z + (long long)b * ldv;
    long long i = max(r0, j);
    for (; i + 3 < r1; i += 4) {
        s0 = fma(Lb[i * ld], zb[i], s0);
        s1 = fma(Lb[(i + 1) * ld], zb[i + 1], s1);
        s2 = fma(Lb[(i + 2) * ld], zb[i + 2], s2);
        s3 = fma(Lb[(i + 3) * ld], zb[i + 3], s3);
    }
    for (; i < r1; ++i) s0 = fma(Lb[i * ld], zb[i], s0);
    out[j] = (s0 + s1) + (s2 + s3);
}

// alpha[b, j] = sum_c partial[b, c, j]
__global__ void __launch_bounds__(256) trmv_t_reduce_kernel(const double* __restrict__ partial, int chunks,
                                                            long long npad, double* __restrict__ alpha,
                                                            long long ldv, long long n) {
    const int b = blockIdx.z;
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += partial[((long long)b * chunks + c) * npad + j];
    alpha[(long long)b * ldv + j] = s;
}

// quad[b] = sum z_i^2 ; logdet[b] = -2 sum log Linv_ii
__global__ void __launch_bounds__(1024) quad_logdet_inv_kernel(const double* __restrict__ z, long long ldv,
                                                                const double* __restrict__ Linv, long long ld,
                                                                long long stride, long long n,
                                                                double* __restrict__ quad,
                                                                double* __restrict__ logdet) {
    __shared__ double sh[32];
    const int b = blockIdx.x;
    const double* Z = z + (long long)b * ldv;
    const double* Lb = Linv + (long long)b * stride;
    double sq = 0.0, sl = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) {
        const double v = Z[i];
        sq += v * v;
        sl += log(Lb[i * ld + i]);
    }
    sq = block_sum<1024>(sq, sh);
    sl = block_sum<1024>(sl, sh);
    if (threadIdx.x == 0) {
        quad[b] = sq;
        logdet[b] = -2.0 * sl;
    }
}

}  // namespace plmc

using namespace plmc;

extern "C" {

long long plmc_trmv_ws(long long npad, int batch) { return npad <= 0 || batch <= 0 ? 0 : 128LL * npad * 8 * batch; }

int plmc_trmv_solve_logdet(const double* Linv, long long ld, long long stride, long long n, long long npad, int batch,
                           const double* y, long long ldy, double* ws, double* z, double* alpha, long long ldv,
                           double* quad, double* logdet, void* stream) {
    if (!Linv || !y || !ws || !z || !alpha || !quad || !logdet || n <= 0 || npad < n || (npad % 128) || ld < npad ||
        (ld & 1) || batch <= 0 || batch > 65535 || ldy < n || ldv < n)
        return PLMC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    trmv_n_kernel<<<dim3((unsigned)((n + 7) / 8), 1, batch), 256, 0, st>>>(Linv, ld, stride, y, ldy, z, ldv, n);
    PLMC_CHECK_LAUNCH();
    const long long rows_per_chunk = npad / 128;      // 128 chunks -> ws = [batch, 128, npad]
    const int chunks = (int)((n + rows_per_chunk - 1) / rows_per_chunk);
    trmv_t_kernel<<<dim3((unsigned)((n + 255) / 256), chunks, batch), 256, 0, st>>>(Linv, ld, stride, z, ldv, ws, n,
                                                                                    npad, rows_per_chunk);
    PLMC_CHECK_LAUNCH();
    trmv_t_reduce_kernel<<<dim3((unsigned)((n + 255) / 256), 1, batch), 256, 0, st>>>(ws, chunks, npad, alpha, ldv, n);
    PLMC_CHECK_LAUNCH();
    quad_logdet_inv_kernel<<<batch, 1024, 0, st>>>(z, ldv, Linv, ld, stride, n, quad, logdet);
    PLMC_CHECK_LAUNCH();
    note_launch(4);
    return PLMC_OK;
}
}
