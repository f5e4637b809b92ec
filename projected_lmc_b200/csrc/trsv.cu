// Single right-hand-side triangular solves with the Cholesky factor (z = L^-1 y, alpha = L^-T z):
// the solve under MultivariateNormal.log_prob / exact prediction (reached from projected_lmc.py:1201 and :1133).
//
// HBM-bound by construction: every 128 x 128 tile of the lower triangle is read exactly once per solve.
// Right-looking block substitution on a vector v of npad entries, one launch per block column j:
//     forward  (L v = y):    v_i -= L[i, j] v_j   for all i > j ;  then v_{j+1} <- Dinv_{j+1} v_{j+1}
//     backward (L^T v = z):  v_i -= L[j, i]^T v_j for all i < j ;  then v_{j-1} <- Dinv_{j-1}^T v_{j-1}
// The CTA that owns block j+1 (j-1) has just applied the last update that block will ever receive, so it
// also applies the explicit inverse of the diagonal leaf (emitted by the potrf leaf): one launch per step
// instead of two.  All reductions are fixed-order (warp shuffles / two-halves sum): bit-reproducible.
#include "plmc_common.cuh"

namespace plmc {

constexpr int TV_B = 128;   // block order = the leaf of the factorisation

// y = T x for a 128 x 128 row-major tile T (row stride ld): warp w owns rows 16 w .. 16 w + 15, a lane 4 columns
__device__ __forceinline__ void tile_matvec_n(const double* __restrict__ T, long long ld, const double* x_sh,
                                              double* y_sh) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double x0 = x_sh[4 * lane], x1 = x_sh[4 * lane + 1], x2 = x_sh[4 * lane + 2], x3 = x_sh[4 * lane + 3];
#pragma unroll 4
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const double2 a = *reinterpret_cast<const double2*>(T + (long long)r * ld + 4 * lane);
        const double2 b = *reinterpret_cast<const double2*>(T + (long long)r * ld + 4 * lane + 2);
        double s = fma(a.x, x0, fma(a.y, x1, fma(b.x, x2, b.y * x3)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y_sh[r] = s;
    }
}

// y = T^T x: thread (c, half) sums 64 rows of column c, the two halves are added in a fixed order
__device__ __forceinline__ void tile_matvec_t(const double* __restrict__ T, long long ld, const double* x_sh,
                                              double* y_sh, double* part_sh) {
    const int c = threadIdx.x & 127, h = threadIdx.x >> 7;
    double s = 0.0;
#pragma unroll 8
    for (int r = h * 64; r < h * 64 + 64; ++r) s = fma(T[(long long)r * ld + c], x_sh[r], s);
    part_sh[h * TV_B + c] = s;
    __syncthreads();
    if (h == 0) y_sh[c] = part_sh[c] + part_sh[TV_B + c];
}

// one step of the block substitution (see the file header).  j < 0 (forward) / j >= nb (backward): leaf only.
template <bool TRANS>
__global__ void __launch_bounds__(256) trsv_step_kernel(const double* __restrict__ Lb, long long ld, long long sL,
                                                        const double* __restrict__ Db, long long sD,
                                                        double* __restrict__ vb, long long sV, int j, int nb) {
    __shared__ double xj[TV_B], acc[TV_B], vi[TV_B], part[2 * TV_B];
    const double* L = Lb + (long long)blockIdx.z * sL;
    const double* D = Db + (long long)blockIdx.z * sD;
    double* v = vb + (long long)blockIdx.z * sV;
    const bool leaf_only = TRANS ? (j >= nb) : (j < 0);
    const int i = TRANS ? (leaf_only ? nb - 1 : (int)blockIdx.x) : (leaf_only ? 0 : j + 1 + (int)blockIdx.x);
    const int next = TRANS ? j - 1 : j + 1;   // the block whose leaf is applied in this launch
    if (threadIdx.x < TV_B) vi[threadIdx.x] = v[(long long)i * TV_B + threadIdx.x];
    if (!leaf_only) {
        if (threadIdx.x < TV_B) xj[threadIdx.x] = v[(long long)j * TV_B + threadIdx.x];
        __syncthreads();
        if (TRANS) tile_matvec_t(L + (long long)j * TV_B * ld + (long long)i * TV_B, ld, xj, acc, part);
        else tile_matvec_n(L + (long long)i * TV_B * ld + (long long)j * TV_B, ld, xj, acc);
        __syncthreads();
        if (threadIdx.x < TV_B) vi[threadIdx.x] -= acc[threadIdx.x];
    }
    __syncthreads();
    if (leaf_only || i == next) {   // v_i is final: apply the explicit inverse of the diagonal leaf
        const double* Di = D + (long long)i * TV_B * TV_B;
        if (TRANS) tile_matvec_t(Di, TV_B, vi, acc, part);
        else tile_matvec_n(Di, TV_B, vi, acc);
        __syncthreads();
        if (threadIdx.x < TV_B) v[(long long)i * TV_B + threadIdx.x] = acc[threadIdx.x];
    } else if (threadIdx.x < TV_B) {
        v[(long long)i * TV_B + threadIdx.x] = vi[threadIdx.x];
    }
}

// v[b, 0:npad] = y[b, 0:n] padded with zeros
__global__ void trsv_pack_kernel(const double* __restrict__ y, long long ldy, double* __restrict__ v, long long sV,
                                 long long n, long long npad) {
    const long long b = blockIdx.z;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npad;
         i += (long long)gridDim.x * blockDim.x)
        v[b * sV + i] = i < n ? y[b * ldy + i] : 0.0;
}

__global__ void trsv_unpack_kernel(const double* __restrict__ v, long long sV, double* __restrict__ out,
                                   long long ldo, long long n) {
    const long long b = blockIdx.z;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[b * ldo + i] = v[b * sV + i];
}

// v (workspace, >= npad doubles per batch member at stride sV) <- y; z <- L^-1 y; alpha <- L^-T z
int trsv_solve(const double* L, long long ld, long long sL, const double* dinv, long long sD, const double* y,
               long long ldy, double* v, long long sV, double* z, double* alpha, long long ldv, long long n,
               long long npad, int batch, cudaStream_t st) {
    const int nb = (int)(npad / TV_B);
    const int gp = (int)((npad + 255) / 256 < 1184 ? (npad + 255) / 256 : 1184);
    trsv_pack_kernel<<<dim3(gp, 1, batch), 256, 0, st>>>(y, ldy, v, sV, n, npad);
    PLMC_CHECK_LAUNCH();
    trsv_step_kernel<false><<<dim3(1, 1, batch), 256, 0, st>>>(L, ld, sL, dinv, sD, v, sV, -1, nb);
    for (int j = 0; j + 1 < nb; ++j)
        trsv_step_kernel<false><<<dim3(nb - 1 - j, 1, batch), 256, 0, st>>>(L, ld, sL, dinv, sD, v, sV, j, nb);
    PLMC_CHECK_LAUNCH();
    trsv_unpack_kernel<<<dim3(gp, 1, batch), 256, 0, st>>>(v, sV, z, ldv, n);
    trsv_step_kernel<true><<<dim3(1, 1, batch), 256, 0, st>>>(L, ld, sL, dinv, sD, v, sV, nb, nb);
    for (int j = nb - 1; j > 0; --j)
        trsv_step_kernel<true><<<dim3(j, 1, batch), 256, 0, st>>>(L, ld, sL, dinv, sD, v, sV, j, nb);
    PLMC_CHECK_LAUNCH();
    trsv_unpack_kernel<<<dim3(gp, 1, batch), 256, 0, st>>>(v, sV, alpha, ldv, n);
    PLMC_CHECK_LAUNCH();
    note_launch(2 * nb + 3);
    return PLMC_OK;
}

}  // namespace plmc
