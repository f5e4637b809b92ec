// FP64 GEMM on the tcgen05 INT8 tensor path, residue-number-system scheme ("Ozaki scheme II").
//
// ozaki.cu splits the operands into s signed 8-bit digit planes and needs s(s+1)/2 INT8 products per FP64
// product (28 for DGEMM-grade 55-bit operands).  Here the operands are scaled to integers once,
//     a'[m,k] = rint(A[m,k] 2^(beta-ea[m])),  b'[n,k] = rint(B[n,k] 2^(beta-eb[n]))   (|a'|, |b'| <= 2^beta),
// and the EXACT integer product c' = sum_k a' b' (|c'| <= K 4^beta < P/2) is recovered from its residues modulo
// nmod pairwise coprime moduli p_i <= 256 (P = prod p_i) by the Chinese remainder theorem:
//     c'/P = symfrac( sum_i r_i u_i / p_i ),   r_i = c' mod p_i,   u_i = (P/p_i)^-1 mod p_i.
// The residues of a' and b' fit signed 8-bit planes, so r_i is ONE plain INT8 GEMM per modulus followed by
// "mod p_i": nmod products instead of s(s+1)/2 -- 16 moduli give beta = 55 bits (K <= 16384), 14 give 47 bits
// (the 7- and 6-slice grades of ozaki.cu), the INT32 accumulators are exact up to K = 131072.
//
// Three stream-ordered steps per GEMM (all batched over the latents):
//   1. residue planes: FP64 operand -> nmod int8 planes, written directly as K-major SWIZZLE_128B UMMA tile
//      images  planes[batch][modulus][k/128][row/128][16 KB]  (+ per-row power-of-two exponents);
//   2. rns_gemm_kernel: persistent tcgen05.mma.kind::i8 GEMM over (batch, modulus, 256x256 tile).  CTA PAIRS
//      (cta_group::2, M = 256): each CTA stages its 128 rows of A and its half of the 256 B rows with bulk
//      async copies (6-stage mbarrier ring), so every operand byte fetched from L2 feeds 256 MACs -- a plain
//      INT8 GEMM with 1-CTA 128x256 tiles needs 96 B/clk/SM at the tensor peak against ~43 B/clk/SM of L2
//      bandwidth per SM; the pair needs 64.  Accumulators: 2 x 256 TMEM columns per CTA (double buffered):
//      the epilogue warps reduce tile t modulo p_i (float reciprocal, exact) and store int8 residues while
//      the tensor core runs tile t+1;
//   3. crt_kernel: residues -> c'/P in two FP64 sums (the leading 40 bits of u_i/p_i are accumulated exactly,
//      the tail in a second sum), wrap to (-1/2, 1/2), scale by P 2^(ea+eb-2 beta), C = alpha AB + beta C.
//
// Replaces the cuBLAS/cuSOLVER products under MultivariateNormal.log_prob (projected_lmc.py:1201) and under
// autograd's cholesky_backward (experiments.py:270) for every GEMM of the recursion with M, N, K >= min_dim.
#include "linalg.cuh"

#include <cmath>
#include <cstring>

namespace plmc {
namespace o2 {

constexpr int MAXMOD = 18;
constexpr int IMG = 16384;                    // one plane of one (128-row, 128-byte-k) tile
constexpr int BK = 128;                       // bytes (= int8 elements) of K per pipeline stage
constexpr int EPI_WARPS = 8;                  // two per TMEM lane quarter, 128 accumulator columns each
constexpr int THREADS = 64 + 32 * EPI_WARPS;  // warp 0 producer, warp 1 MMA issuer / relay + TMEM alloc, 2.. epilogue
constexpr int RTILE = 65536;                  // residue output tile: 256 rows x 256 bytes
constexpr float MAGIC = 12582912.0f;          // 1.5 * 2^23: float(MAGIC + x) has integer bits 0x4B400000 + x

static const int kModuli[MAXMOD] = {256, 255, 253, 251, 247, 241, 239, 233, 229,
                                    227, 223, 217, 211, 199, 197, 193, 191, 181};

struct ModTab {
    int p[MAXMOD];
    float pf[MAXMOD];
    float inv[MAXMOD];
    int c16[MAXMOD];   // 2^16 mod p, symmetric
    int wlo[MAXMOD];   // 256^b mod p (symmetric, int8) for byte b = 0..3 of the balanced digits
    int whi[MAXMOD];   // b = 4..7
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  RELAXED on purpose: a
// release at cluster scope compiles to MEMBAR.ALL.GPU (+ the waiter's acquire to CCTL.IVALL), ~1 us per k-stage in
// the relay -- measured: the CTA-pair kernel ran at half the rate of the single-CTA one.  No generic-proxy data
// is published through these barriers: what they order is async-proxy work that has already completed when the
// arrive is issued (the bulk copy behind the local full barrier; tcgen05.ld behind tcgen05.wait::ld), exactly as
// when a 2-SM TMA load signals the leader's barrier directly.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar), "r"(rank)
        : "memory");
}
// Spin on a phase parity.  A protocol bug must not hang the GPU: after ~4 s of polling the kernel traps.
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        if (CLUSTER)
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(ok)
                : "r"(bar), "r"(parity)
                : "memory");
        else
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(ok)
                : "r"(bar), "r"(parity)
                : "memory");
        if (ok) return;
        if ((spins & 0xFFFF) == 0xFFFF) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major) | SBO>>4 [32,46) = 1024 B between 8-row groups
//   | version = 1 [46,48) | layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor for kind::i8: S32 accumulate, signed 8-bit K-major A and B
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    if (CG == 2)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(da), "l"(db), "r"(idesc), "r"(accum)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
            "l"(da), "l"(db), "r"(idesc), "r"(accum)
            : "memory");
}
// arrive on `bar` (in both CTAs of the pair when CG == 2) once every MMA issued so far by this thread has retired
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 2)
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
            "h"((uint16_t)3)
            : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                     : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// x mod p (symmetric residue, low byte of the result) for an integer |x| < 2^22 held as the float MAGIC + x
__device__ __forceinline__ uint32_t mod_from_biased(float F, float inv, float pf) {
    const float f = F - MAGIC;                    // exact
    const float kk = fmaf(f, inv, MAGIC);         // MAGIC + rint(x / p): one rounding, to an integer
    const float k = kk - MAGIC;                   // exact
    return __float_as_uint(fmaf(-k, pf, F));      // MAGIC + (x - k p), exact; low byte = residue
}
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}

// ------------------------------------------------------------------------------------------
// step 1: residue planes
//   KC operand: element (x, k) at P[x*ld + k]   ;   MC operand: element (x, k) at P[k*ld + x]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int exp_for(double amax) {
    if (!(amax > 0.0)) return 0;
    int e;
    frexp(amax, &e);   // amax = f 2^e, f in [0.5, 1)  ->  amax 2^-e < 1
    return max(e, -1022);
}

// 16 consecutive k of row x: balanced base-256 digits of q = rint(v 2^(beta - e)) (one 64-bit add and xor, as in
// ozaki.cu), then per modulus  r = sum_b digit_b (256^b mod p)  with two DP4A and the float reduction.
struct Digits16 {
    uint32_t lo[16], hi[16];
    __device__ __forceinline__ void set(const double (&v)[16], int sh) {
        const double m1 = __longlong_as_double((long long)(1023 + sh / 2) << 52);   // 2^sh split: |sh| < 1100
        const double m2 = __longlong_as_double((long long)(1023 + sh - sh / 2) << 52);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const unsigned long long q =
                ((unsigned long long)__double2ll_rn(v[u] * m1 * m2) + 0x8080808080808080ull) ^ 0x8080808080808080ull;
            lo[u] = (uint32_t)q;
            hi[u] = (uint32_t)(q >> 32);
        }
    }
    __device__ __forceinline__ uint4 residues(int wl, int wh, float inv, float pf) const {
        uint32_t b[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            int acc = __dp4a((int)lo[u], wl, 0x4B400000);
            acc = __dp4a((int)hi[u], wh, acc);
            b[u] = mod_from_biased(__int_as_float(acc), inv, pf);
        }
        return make_uint4(pack4(b[0], b[1], b[2], b[3]), pack4(b[4], b[5], b[6], b[7]), pack4(b[8], b[9], b[10], b[11]),
                          pack4(b[12], b[13], b[14], b[15]));
    }
};

// byte offset of the 16-byte chunk holding k16..k16+15 of row x inside one modulus plane
__device__ __forceinline__ long long chunk_offset(int x, int k16, int nxt) {
    const int kt = k16 >> 7, c = (k16 >> 4) & 7, xt = x >> 7, r = x & 127;
    return ((long long)kt * nxt + xt) * IMG + r * 128 + ((c ^ (r & 7)) << 4);
}

// KC operand: one warp per row; the row is read twice (absmax, then conversion; the second read hits L2).
// tri = 1: the operand is the LOWER triangle of a square matrix stored in place (entries with k > x are not part of it,
// whatever the buffer holds there): they are ignored by the row maximum and written as zeros up to the end of the
// row's 256-aligned diagonal block -- the triangular product (GemmArgs2::tri) never reads k-tiles beyond it.
__global__ void __launch_bounds__(256) residue_kc_kernel(const double* __restrict__ Pb, long long ld, long long sP,
                                                         int X, int K, int beta, int nmod,
                                                         int8_t* __restrict__ planes_b, long long sPlb, long long sPlm,
                                                         int nxt, int* __restrict__ ex_b, const ModTab T, int tri) {
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (x >= X) return;
    const double* row = Pb + (long long)blockIdx.z * sP + (long long)x * ld;
    int8_t* planes = planes_b + (long long)blockIdx.z * sPlb;
    const bool al16 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    const int kend = tri ? min(K, ((x >> 8) + 1) << 8) : K;      // columns converted
    const int klast = tri ? x : K - 1;                           // last column that belongs to the operand
    auto load16 = [&](int k16, double (&v)[16]) {
        if (al16) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double2 a = *reinterpret_cast<const double2*>(row + k16 + 2 * u);
                v[2 * u] = a.x;
                v[2 * u + 1] = a.y;
            }
        } else {
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = row[k16 + u];
        }
        if (tri && k16 + 15 > klast) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (k16 + u > klast) v[u] = 0.0;
        }
    };
    // Row maximum: only its binary exponent is needed, and for non-negative doubles the high word orders like the
    // magnitude, so the reduction runs on 32-bit integers (fmax on doubles costs ~7 instructions on sm_100a).
    int hmax = 0;
    for (int k16 = lane * 16; k16 <= klast; k16 += 512) {
        double v[16];
        load16(k16, v);
#pragma unroll
        for (int u = 0; u < 16; ++u) hmax = max(hmax, __double2hiint(v[u]) & 0x7fffffff);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hmax = max(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
    // frexp exponent of the maximum (|x| 2^-e in [0.5, 1)); all-zero (or all-subnormal) row: any exponent works
    const int be = hmax >> 20;
    const int e = (hmax == 0) ? 0 : max(be - 1022, -1022);
    if (lane == 0) ex_b[(long long)blockIdx.z * X + x] = e;
    for (int k16 = lane * 16; k16 < kend; k16 += 512) {
        double v[16];
        if (k16 <= klast) {
            load16(k16, v);
        } else {
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = 0.0;
        }
        Digits16 dg;
        dg.set(v, beta - e);
        int8_t* dst = planes + chunk_offset(x, k16, nxt);
        for (int i = 0; i < nmod; ++i)
            *reinterpret_cast<uint4*>(dst + (long long)i * sPlm) = dg.residues(T.wlo[i], T.whi[i], T.inv[i], T.pf[i]);
    }
}

// MC operand, pass 1: exponent of the column-wise absmax over k (integer atomicMax over k-slabs: order independent).
// tri = 2: the operand is the LOWER triangle of a square matrix stored in place, indexed (x = column, k = row):
// entries with k < x are not part of it.
__global__ void __launch_bounds__(256) absmax_mc_kernel(const double* __restrict__ Pb, long long ld, long long sP,
                                                        int X, int K, int* __restrict__ ex_b, int tri) {
    __shared__ int red[4][64];
    const int xl = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int x = blockIdx.x * 64 + xl;
    const double* P = Pb + (long long)blockIdx.z * sP;
    int* ex = ex_b + (long long)blockIdx.z * X;
    const int k0 = blockIdx.y * 256, k1 = min(K, k0 + 256);
    if (tri && k1 <= blockIdx.x * 64) return;   // the whole slab lies above the diagonal
    int emax = -2000000000;
    if (x < X) {
        for (int k = k0 + grp; k < k1; k += 4) {
            if (tri && k < x) continue;
            const long long bits = __double_as_longlong(P[(long long)k * ld + x]);
            const int be = (int)((bits >> 52) & 0x7FF);
            const int e = be ? be - 1022 : -1022;
            if (bits << 1) emax = max(emax, e);
        }
    }
    red[grp][xl] = emax;
    __syncthreads();
    if (grp == 0 && x < X) {
        emax = max(max(red[0][xl], red[1][xl]), max(red[2][xl], red[3][xl]));
        if (emax > -2000000000) atomicMax(ex + x, emax);
    }
}

// MC operand, pass 2: a (128 k) x (32 x) slab is turned through shared memory (row stride 145, 16-element groups
// 18 apart: both the transposing stores and the per-thread reads are bank-conflict free); 8 consecutive lanes then
// write the 128-byte image line of one row.  tri = 2: entries with k < x become zeros; slabs entirely above the
// 256-aligned diagonal block of their columns are skipped (the triangular product never reads them).
__global__ void __launch_bounds__(256) residue_mc_kernel(const double* __restrict__ Pb, long long ld, long long sP,
                                                         int X, int K, int beta, int nmod,
                                                         int8_t* __restrict__ planes_b, long long sPlb, long long sPlm,
                                                         int nxt, const int* __restrict__ ex_b, const ModTab T,
                                                         int tri) {
    __shared__ double tile[32 * 145];
    const double* P = Pb + (long long)blockIdx.z * sP;
    int8_t* planes = planes_b + (long long)blockIdx.z * sPlb;
    const int* ex = ex_b + (long long)blockIdx.z * X;
    const int x0 = blockIdx.x * 32, k0 = blockIdx.y * 128;
    if (tri && k0 + 128 <= ((x0 >> 8) << 8)) return;
    {
        const int xl = threadIdx.x & 31, kr = threadIdx.x >> 5;   // 8 k-rows of 32 x per pass
        double t[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = k0 + kr + 8 * j;
            t[j] = (x0 + xl < X && !(tri && k < x0 + xl)) ? P[(long long)k * ld + x0 + xl] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = kr + 8 * j;
            tile[xl * 145 + (k >> 4) * 18 + (k & 15)] = t[j];
        }
    }
    __syncthreads();
    const int kc = threadIdx.x & 7, xl = threadIdx.x >> 3;
    const int x = x0 + xl;
    if (x >= X) return;
    const int e = max(ex[x], -1022);   // all-zero column: any exponent works
    double v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = tile[xl * 145 + kc * 18 + j];
    Digits16 dg;
    dg.set(v, beta - e);
    int8_t* dst = planes + chunk_offset(x, k0 + kc * 16, nxt);
    for (int i = 0; i < nmod; ++i)
        *reinterpret_cast<uint4*>(dst + (long long)i * sPlm) = dg.residues(T.wlo[i], T.whi[i], T.inv[i], T.pf[i]);
}

// ------------------------------------------------------------------------------------------
// step 2: the INT8 GEMM modulo p_i
// ------------------------------------------------------------------------------------------
struct GemmArgs2 {
    const int8_t* PA;        // planes of op(A): [batch][modulus][K/128][nxtA][16 KB]
    const int8_t* PB;        // planes of op(B): [batch][modulus][K/128][nxtB][16 KB]
    long long sPAb, sPBb;    // batch strides (bytes)
    long long sPAm, sPBm;    // modulus strides (bytes)
    int nxtA, nxtB;          // 128-row tiles per k-tile in each buffer (even)
    int nkt;                 // K / 128
    uint8_t* R;              // residues of the product: [batch][modulus][slot][256 rows][256 bytes]
    long long sRb, sRm;
    int tn256;               // 256-column tiles per row of R slots (full mode)
    int tiles_m, tiles_n;    // work tiles: (128 CG) x 256
    int m128, n128;          // valid 128-row / 128-column blocks of the product
    int lower;               // only blocks on or below the diagonal are produced
    int nmod;
    int per;                 // work tiles per (batch member, modulus)
    int total;               // per * nmod * batch members of this launch
    // triangular operand (square, lower, K = its order): only the k-tiles that can be nonzero are visited.
    //   1: op(A)[m][k] = 0 for k > m (X B)        -> k-tiles [0, end of the row tile's 256-block)
    //   2: op(A)[m][k] = 0 for k < m (X^T B, X^T X lower) -> k-tiles from the start of the row tile's 256-block
    //   3: op(B)[k][n] = 0 for k < n (B X)        -> k-tiles from the start of the column tile's 256-block
    //   4: op(B)[k][n] = 0 for k > n (B X^T)      -> k-tiles [0, end of the column tile's 256-block)
    int tri;
    ModTab T;
};

__host__ __device__ inline long long r_slot(int tm256, int tn256, int ntn256, int lower) {
    return lower ? (long long)tm256 * (tm256 + 1) / 2 + tn256 : (long long)tm256 * ntn256 + tn256;
}

struct WorkTile { int bz, mod, tm, tn, kt0, kt1; };
template <int CG>
__device__ __forceinline__ WorkTile decode(const GemmArgs2& p, int w) {
    WorkTile t;
    const int per_b = p.per * p.nmod;
    t.bz = w / per_b;
    const int rem = w - t.bz * per_b;
    t.mod = rem / p.per;
    const int idx = rem - t.mod * p.per;
    if (p.lower) {
        if (CG == 2) {   // 256 x 256 tiles: tn <= tm
            int tm = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
            while (tm * (tm + 1) / 2 > idx) --tm;
            while ((tm + 1) * (tm + 2) / 2 <= idx) ++tm;
            t.tm = tm;
            t.tn = idx - tm * (tm + 1) / 2;
        } else {         // 128 x 256 tiles: rows 2j and 2j+1 own the column tiles 0..j
            int j = (int)((sqrtf(4.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
            while (j * (j + 1) > idx) --j;
            while ((j + 1) * (j + 2) <= idx) ++j;
            const int r = idx - j * (j + 1);
            t.tm = 2 * j + r / (j + 1);
            t.tn = r % (j + 1);
        }
    } else {
        // grouped raster: bands of 8 tile-rows walk the columns together (operand panels shared out of L2)
        const int GROUP = 8;
        const int per_group = GROUP * p.tiles_n;
        const int gid = idx / per_group;
        const int first = gid * GROUP;
        const int gsz = min(p.tiles_m - first, GROUP);
        const int r = idx - gid * per_group;
        t.tm = first + r % gsz;
        t.tn = r / gsz;
    }
    t.kt0 = 0;
    t.kt1 = p.nkt;
    if (p.tri == 1) t.kt1 = min(p.nkt, 2 * ((t.tm * CG) >> 1) + 2);
    else if (p.tri == 2) t.kt0 = 2 * ((t.tm * CG) >> 1);
    else if (p.tri == 3) t.kt0 = 2 * t.tn;
    else if (p.tri == 4) t.kt1 = min(p.nkt, 2 * t.tn + 2);
    return t;
}

template <int CG>
struct Cfg {
    static constexpr int NST = (CG == 2) ? 6 : 4;
    static constexpr int B_BYTES = (CG == 2) ? IMG : 2 * IMG;
    static constexpr int STAGE = IMG + B_BYTES;
    static constexpr int SMEM = NST * STAGE + 1024;
};

template <int CG>
__global__ void __launch_bounds__(THREADS, 1) rns_gemm_kernel(const GemmArgs2 p) {
    using C = Cfg<CG>;
    extern __shared__ __align__(1024) uint8_t o2_smem[];
    __shared__ __align__(8) unsigned long long full_bar[C::NST], empty_bar[C::NST], peer_full[C::NST], acc_full[2],
        acc_empty[2];
    __shared__ uint32_t tmem_base_sh;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CG, nclusters = gridDim.x / CG;
    const uint32_t smem0 = (smem_u32(o2_smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NST; ++i) {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
            mbar_init(smem_u32(&peer_full[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&acc_full[i]), 1);
            mbar_init(smem_u32(&acc_empty[i]), EPI_WARPS * CG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;

    if (warp == 0) {
        // ===== copy producer: this CTA's 128 rows of A and its share of the 256 rows of B, one k-tile per stage =====
        if (elect_one()) {
            int it = 0;
            for (int w = cluster_id; w < p.total; w += nclusters) {
                const WorkTile t = decode<CG>(p, w);
                const int xa = t.tm * CG + (int)rank;
                const int xb = 2 * t.tn + ((CG == 2) ? (int)rank : 0);
                const int8_t* ga = p.PA + (long long)t.bz * p.sPAb + (long long)t.mod * p.sPAm + (long long)xa * IMG;
                const int8_t* gb = p.PB + (long long)t.bz * p.sPBb + (long long)t.mod * p.sPBm + (long long)xb * IMG;
                for (int kt = t.kt0; kt < t.kt1; ++kt, ++it) {
                    const int st = it % C::NST, round = it / C::NST;
                    if (round > 0) mbar_wait<false>(smem_u32(&empty_bar[st]), (round - 1) & 1);
                    const uint32_t fb = smem_u32(&full_bar[st]);
                    mbar_expect_tx(fb, (uint32_t)C::STAGE);
                    const uint32_t sa = smem0 + (uint32_t)st * C::STAGE;
                    bulk_load(sa, ga + (long long)kt * p.nxtA * IMG, IMG, fb);
                    bulk_load(sa + IMG, gb + (long long)kt * p.nxtB * IMG, C::B_BYTES, fb);
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===== MMA issuer (leader CTA of the pair) =====
        const uint32_t idesc = umma_idesc_i8(128 * CG, 256);
        int it = 0, nt = 0;
        for (int w = cluster_id; w < p.total; w += nclusters, ++nt) {
            const WorkTile t = decode<CG>(p, w);
            const int buf = nt & 1, use = nt >> 1;
            if (use > 0) {   // the epilogue warps (of both CTAs) must have drained this accumulator
                mbar_wait<false>(smem_u32(&acc_empty[buf]), (use - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            for (int kt = t.kt0; kt < t.kt1; ++kt, ++it) {
                const int st = it % C::NST, round = it / C::NST;
                mbar_wait<false>(smem_u32(&full_bar[st]), round & 1);
                if (CG == 2) mbar_wait<false>(smem_u32(&peer_full[st]), round & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t sa = smem0 + (uint32_t)st * C::STAGE;
                    const uint32_t sb = sa + IMG;
#pragma unroll
                    for (int kk = 0; kk < BK / 32; ++kk)
                        umma_i8<CG>(tmem_base + (uint32_t)buf * 256, umma_desc_sw128(sa + kk * 32),
                                    umma_desc_sw128(sb + kk * 32), idesc, (kt > t.kt0 || kk > 0) ? 1u : 0u);
                    umma_commit<CG>(smem_u32(&empty_bar[st]));   // frees the stage (in both CTAs) when the MMAs retire
                    if (kt == t.kt1 - 1) umma_commit<CG>(smem_u32(&acc_full[buf]));
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== relay (second CTA of the pair): tell the leader when this CTA's stage has landed =====
        int it = 0;
        for (int w = cluster_id; w < p.total; w += nclusters) {
            const WorkTile t = decode<CG>(p, w);
            for (int kt = t.kt0; kt < t.kt1; ++kt, ++it) {
                const int st = it % C::NST, round = it / C::NST;
                mbar_wait<false>(smem_u32(&full_bar[st]), round & 1);
                if (lane == 0) mbar_arrive_remote(smem_u32(&peer_full[st]), 0);
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> mod p -> int8 residues (this CTA's 128 rows, 128 columns per warp) =====
        const int quad = warp & 3;            // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;     // which 128 of the 256 columns
        int nt = 0;
        for (int w = cluster_id; w < p.total; w += nclusters, ++nt) {
            const WorkTile t = decode<CG>(p, w);
            const int buf = nt & 1, use = nt >> 1;
            mbar_wait<false>(smem_u32(&acc_full[buf]), use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int rt = t.tm * CG + (int)rank;      // 128-row block of the product
            const int ct = 2 * t.tn + half;            // 128-column block
            if (rt < p.m128 && ct < p.n128 && !(p.lower && ct > rt)) {
                const int pi = p.T.p[t.mod];
                const int c16 = p.T.c16[t.mod];
                const float inv = p.T.inv[t.mod], pf = p.T.pf[t.mod];
                uint8_t* dst = p.R + (long long)t.bz * p.sRb + (long long)t.mod * p.sRm +
                               r_slot(rt >> 1, t.tn, p.tn256, p.lower) * RTILE +
                               (long long)((rt & 1) * 128 + quad * 32 + lane) * 256 + 128 * half;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256 + 128 * half);
#pragma unroll 1
                for (int c4 = 0; c4 < 4; ++c4) {
                    uint32_t r[32];
                    tmem_ld32(taddr + 32 * c4, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    uint32_t o[8];
                    if (pi == 256) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = pack4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            uint32_t b[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int acc = (int)r[4 * j + u];
                                // acc = hi 2^16 + lo:  x = hi (2^16 mod p) + lo, |x| < 2^22, as the float MAGIC + x
                                const int xi = (acc >> 16) * c16 + ((acc & 0xFFFF) | 0x4B400000);
                                b[u] = mod_from_biased(__int_as_float(xi), inv, pf);
                            }
                            o[j] = pack4(b[0], b[1], b[2], b[3]);
                        }
                    }
                    *reinterpret_cast<uint4*>(dst + 32 * c4) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(dst + 32 * c4 + 16) = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
                else mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    if (warp == 1) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// step 3: Chinese remainder reconstruction + FP64 update of C
// ------------------------------------------------------------------------------------------
struct CrtArgs {
    const uint8_t* R;
    long long sRb, sRm;
    int tn256;
    double* C;
    long long ldc, sC;
    const int* ea;     // [batch][M] exponents of the rows of op(A)
    const int* eb;     // [batch][N] exponents of the columns of op(B)
    int M, N, lower, nmod;
    double alpha, beta;
    double pscale;     // P 2^(-2 beta_bits)
    double H[MAXMOD];  // leading 40 bits of u_i / p_i
    double L[MAXMOD];  // u_i / p_i - H_i
};

__device__ __forceinline__ double pow2i(int e) {   // 2^e for e in [-1022, 1023]
    return __longlong_as_double((long long)(e + 1023) << 52);
}

// One CTA per 32 rows of a 256 x 256 residue tile (gridDim.y = 8): a warp covers one row per pass (32 lanes x 8
// columns) and 4 rows in all.  Every residue load of a row (one 8-byte word per modulus) and the old C values are
// issued before the first use, so a warp keeps nmod + 4 independent requests in flight: the kernel is
// HBM-bound (16 B of residues + 16 B of C per element), not latency-bound.
__global__ void __launch_bounds__(256) crt_kernel(const CrtArgs a) {
    int tm, tn;
    if (a.lower) {
        const int idx = blockIdx.x;
        tm = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while (tm * (tm + 1) / 2 > idx) --tm;
        while ((tm + 1) * (tm + 2) / 2 <= idx) ++tm;
        tn = idx - tm * (tm + 1) / 2;
    } else {
        tm = blockIdx.x / a.tn256;
        tn = blockIdx.x - tm * a.tn256;
    }
    const int bz = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gc0 = tn * 256 + 8 * lane;
    if (gc0 >= a.N) return;
    const uint8_t* tile = a.R + (long long)bz * a.sRb + r_slot(tm, tn, a.tn256, a.lower) * RTILE + 8 * lane;
    double fb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) fb[j] = pow2i(max(a.eb[(long long)bz * a.N + gc0 + j], -1022));
    double* Cb = a.C + (long long)bz * a.sC;
    const bool rmw = (a.beta != 0.0);
#pragma unroll 1
    for (int rr = 0; rr < 4; ++rr) {
        const int row = blockIdx.y * 32 + warp * 4 + rr;
        const int gr = tm * 256 + row;
        if (gr >= a.M) break;
        if (a.lower && (gc0 >> 7) > (gr >> 7)) continue;
        const uint8_t* src = tile + row * 256;
        uint2 v[MAXMOD];
#pragma unroll
        for (int i = 0; i < MAXMOD; ++i)
            if (i < a.nmod) v[i] = __ldcs(reinterpret_cast<const uint2*>(src + (long long)i * a.sRm));
        double* crow = Cb + (long long)gr * a.ldc + gc0;
        double2 old[4];
        if (rmw) {
#pragma unroll
            for (int j = 0; j < 4; ++j) old[j] = *reinterpret_cast<const double2*>(crow + 2 * j);
        }
        double s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.0;
#pragma unroll
        for (int i = 0; i < MAXMOD; ++i) {
            if (i < a.nmod) {
                const double h = a.H[i], l = a.L[i];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = (int)(int8_t)(((j < 4) ? v[i].x : v[i].y) >> (8 * (j & 3)));
                    // exact int -> double: 2^52 + 2^31 + r is representable, subtract the bias
                    const double d = __hiloint2double(0x43300000, r ^ 0x80000000) - 4503601774854144.0;
                    s1[j] = fma(d, h, s1[j]);   // exact: 8-bit r times 40-bit h, at most 18 terms
                    s2[j] = fma(d, l, s2[j]);
                }
            }
        }
        const double sa = a.alpha * a.pscale * pow2i(max(a.ea[(long long)bz * a.M + gr], -1022));
        double out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double v2 = (s1[j] - rint(s1[j])) + s2[j];   // c'/P modulo 1
            v2 -= rint(v2);                               // in [-1/2, 1/2]
            out[j] = v2 * sa * fb[j];
        }
        if (rmw) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                out[2 * j] = fma(a.beta, old[j].x, out[2 * j]);
                out[2 * j + 1] = fma(a.beta, old[j].y, out[2 * j + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(crow + j) = make_double2(out[j], out[j + 1]);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int symmod(long long x, int p) {
    long long r = x % p;
    if (r < 0) r += p;
    if (2 * r >= p) r -= p;   // [-p/2, p/2): -128 for p = 256
    return (int)r;
}

static const ModTab& mod_table() {
    static ModTab T;
    static bool done = false;   // idempotent initialisation of immutable data (benign if raced)
    if (!done) {
        ModTab t;
        for (int i = 0; i < MAXMOD; ++i) {
            const int p = kModuli[i];
            t.p[i] = p;
            t.pf[i] = (float)p;
            t.inv[i] = 1.0f / (float)p;
            t.c16[i] = symmod(65536, p);
            unsigned lo = 0, hi = 0;
            long long w = 1;   // 256^b mod p
            for (int b = 0; b < 8; ++b) {
                const unsigned byte = (unsigned)(symmod(w, p) & 0xFF);
                if (b < 4) lo |= byte << (8 * b);
                else hi |= byte << (8 * (b - 4));
                w = (w * 256) % p;
            }
            t.wlo[i] = (int)lo;
            t.whi[i] = (int)hi;
        }
        T = t;
        done = true;
    }
    return T;
}

static long double log2_P(int nmod) {
    long double s = 0;
    for (int i = 0; i < nmod; ++i) s += log2l((long double)kModuli[i]);
    return s;
}

// operand bits for a product of inner dimension K with nmod moduli: 2 K 4^beta < P
int rns_bits(int nmod, int K) {
    const long double b = (log2_P(nmod) - 1.0L - log2l((long double)K) - 1e-6L) / 2.0L;
    int beta = (int)floorl(b);
    if (beta > 61) beta = 61;
    return beta;
}

static void crt_constants(int nmod, int beta, CrtArgs& a) {
    long double P = 1;
    for (int i = 0; i < nmod; ++i) P *= (long double)kModuli[i];
    a.pscale = (double)ldexpl(P, -2 * beta);
    for (int i = 0; i < nmod; ++i) {
        const int p = kModuli[i];
        long long m = 1;   // (P / p_i) mod p_i
        for (int j = 0; j < nmod; ++j)
            if (j != i) m = (m * (kModuli[j] % p)) % p;
        long long u = 0;   // m^-1 mod p
        for (long long c = 1; c < p; ++c)
            if ((m * c) % p == 1) { u = c; break; }
        const long long num = u << 40;
        a.H[i] = ldexp((double)(num / p), -40);
        a.L[i] = ldexp((double)(num % p) / (double)p, -40);
    }
}

struct DevAttr { bool done; int sms; int clusters; };
static DevAttr g_dev[64];   // per-device one-time function attributes (written once per device, idempotent)

static int dev_attrs(DevAttr** out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return PLMC_ERR_LAUNCH;
    DevAttr& d = g_dev[dev];
    if (!d.done) {
        if (cudaFuncSetAttribute(rns_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM) !=
                cudaSuccess ||
            cudaFuncSetAttribute(rns_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM) !=
                cudaSuccess)
            return PLMC_ERR_LAUNCH;
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        d.sms = sms;
        // CTA pairs that can be co-resident (one CTA per SM, both SMs of a TPC)
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(sms - (sms & 1));
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = Cfg<2>::SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, rns_gemm_kernel<2>, &cfg) != cudaSuccess || ncl <= 0) ncl = 0;
        cudaGetLastError();
        d.clusters = ncl < sms / 2 ? ncl : sms / 2;
        d.done = true;
    }
    *out = &d;
    return PLMC_OK;
}

static long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

// scratch of ONE batch member: planes of both operands, residue tiles of the product, exponents
long long rns_ws_bytes(int M, int N, int K, int nmod, bool same_operand, bool lower) {
    const long long Mp = round_up(M, 256), Np = round_up(N, 256);
    const long long a = (long long)nmod * K * Mp, b = same_operand ? 0 : (long long)nmod * K * Np;
    const long long tm = Mp / 256, tn = Np / 256;
    const long long slots = lower ? tm * (tm + 1) / 2 : tm * tn;
    return a + b + (long long)nmod * slots * RTILE + round_up(4LL * (M + N), 1024) + 1024;
}

// tri: GemmArgs2::tri (0 for a plain product); the triangular operand is converted with the matching mask
static int rns_gemm_fit(bool aKC, bool bKC, const double* A, long long lda, long long sA, const double* B, long long ldb,
                        long long sB, double* Cm, long long ldc, long long sC, int M, int N, int K, double alpha,
                        double beta, int lower, int nmod, bool same_operand, int batch, uint8_t* w0, long long avail,
                        int flags, cudaStream_t st, int tri = 0) {
    DevAttr* da = nullptr;
    if (int rc = dev_attrs(&da)) return rc;
    const ModTab& T = mod_table();
    const int bits = rns_bits(nmod, K);
    const long long Mp = round_up(M, 256), Np = round_up(N, 256);
    const long long bytesA = (long long)nmod * K * Mp, bytesB = same_operand ? 0 : (long long)nmod * K * Np;
    const int tm256 = (int)(Mp / 256), tn256 = (int)(Np / 256);
    const long long slots = lower ? (long long)tm256 * (tm256 + 1) / 2 : (long long)tm256 * tn256;
    const long long bytesR = (long long)nmod * slots * RTILE;
    const long long need1 = bytesA + bytesB + bytesR + round_up(4LL * (M + N), 1024);
    const int bc_max = (int)((avail / need1) < batch ? (avail / need1) : batch);
    if (bc_max < 1) return PLMC_ERR_BADARG;
    const int nkt = K / BK;
    const int cg = (!(flags & 1) && da->clusters > 0) ? 2 : 1;   // flags bit 0: single-CTA kernel

    for (int b0 = 0; b0 < batch; b0 += bc_max) {
        const int bc = (batch - b0) < bc_max ? (batch - b0) : bc_max;
        int8_t* pa = (int8_t*)w0;
        int8_t* pb = same_operand ? pa : (int8_t*)(w0 + bytesA * bc);
        uint8_t* R = w0 + (bytesA + bytesB) * bc;
        int* ea = (int*)(R + bytesR * bc);
        int* eb = same_operand ? ea : ea + (long long)bc * M;
        auto planes = [&](bool kc, const double* P, long long ld, long long sP, int X, long long Xp, int8_t* pl,
                          long long bytes, int* ex, bool masked) {
            const long long sPlm = (long long)K * Xp;
            const int nxt = (int)(Xp / 128);
            if (kc) {
                residue_kc_kernel<<<dim3((X + 7) / 8, 1, bc), 256, 0, st>>>(P, ld, sP, X, K, bits, nmod, pl, bytes, sPlm,
                                                                           nxt, ex, T, masked ? 1 : 0);
            } else {
                cudaMemsetAsync(ex, 0x80, sizeof(int) * (size_t)X * bc, st);
                absmax_mc_kernel<<<dim3((X + 63) / 64, (K + 255) / 256, bc), 256, 0, st>>>(P, ld, sP, X, K, ex,
                                                                                          masked ? 2 : 0);
                residue_mc_kernel<<<dim3((X + 31) / 32, K / 128, bc), 256, 0, st>>>(P, ld, sP, X, K, bits, nmod, pl,
                                                                                   bytes, sPlm, nxt, ex, T,
                                                                                   masked ? 2 : 0);
            }
        };
        // tri 1: A = X (KC, zero for k > m); tri 2: A = X^T (MC, zero for k < m); tri 3: B = X (MC, zero for k < n);
        // tri 4: B = X^T, i.e. planes (n, k) = X[n][k] (KC, zero for k > n)
        planes(aKC, A + (long long)b0 * sA, lda, sA, M, Mp, pa, bytesA, ea, tri == 1 || tri == 2);
        if (!same_operand) planes(bKC, B + (long long)b0 * sB, ldb, sB, N, Np, pb, bytesB, eb, tri == 3 || tri == 4);
        PLMC_CHECK_LAUNCH();

        GemmArgs2 g;
        g.PA = pa; g.PB = pb;
        g.sPAb = bytesA; g.sPBb = same_operand ? bytesA : bytesB;
        g.sPAm = (long long)K * Mp; g.sPBm = same_operand ? g.sPAm : (long long)K * Np;
        g.nxtA = (int)(Mp / 128); g.nxtB = same_operand ? g.nxtA : (int)(Np / 128);
        g.nkt = nkt;
        g.R = R; g.sRb = bytesR; g.sRm = slots * RTILE;
        g.tn256 = tn256;
        g.tiles_m = (cg == 2) ? tm256 : (int)(Mp / 128);
        g.tiles_n = tn256;
        g.m128 = M / 128; g.n128 = N / 128;
        g.lower = lower; g.nmod = nmod;
        g.tri = tri;
        long long per;
        if (!lower) per = (long long)g.tiles_m * g.tiles_n;
        else if (cg == 2) per = slots;
        else per = (long long)tm256 * (tm256 + 1);   // rows 2j, 2j+1 own j+1 column tiles each
        if (per * nmod * bc > 2000000000LL) return PLMC_ERR_BADARG;
        g.per = (int)per;
        g.total = (int)(per * nmod * bc);
        g.T = T;
        if (cg == 2) {
            const int ncl = g.total < da->clusters ? g.total : da->clusters;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(2 * ncl);
            cfg.blockDim = dim3(THREADS);
            cfg.dynamicSmemBytes = Cfg<2>::SMEM;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (cudaLaunchKernelEx(&cfg, rns_gemm_kernel<2>, g) != cudaSuccess) {
                cudaGetLastError();
                return PLMC_ERR_LAUNCH;
            }
        } else {
            const int ctas = g.total < da->sms ? g.total : da->sms;
            rns_gemm_kernel<1><<<ctas, THREADS, Cfg<1>::SMEM, st>>>(g);
        }
        PLMC_CHECK_LAUNCH();

        CrtArgs c;
        c.R = R; c.sRb = bytesR; c.sRm = slots * RTILE; c.tn256 = tn256;
        c.C = Cm + (long long)b0 * sC; c.ldc = ldc; c.sC = sC;
        c.ea = ea; c.eb = eb;
        c.M = M; c.N = N; c.lower = lower; c.nmod = nmod;
        c.alpha = alpha; c.beta = beta;
        crt_constants(nmod, bits, c);
        crt_kernel<<<dim3((unsigned)slots, 8, bc), 256, 0, st>>>(c);
        PLMC_CHECK_LAUNCH();
        note_launch(same_operand ? 3 : 4);
    }
    return PLMC_OK;
}


// ------------------------------------------------------------------------------------------
// roofline denominator: the INT8 tensor pipe with shared-memory-resident operands (no global traffic).
// Every CTA (pair) fills one stage with pseudo-random bytes once and issues `iters` x 4 tcgen05.mma.kind::i8
// (M = 128 CG, N = 256, K = 32) into its two TMEM accumulators.
// ------------------------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(128, 1) peak_i8_kernel(long long iters) {
    using C = Cfg<CG>;
    extern __shared__ __align__(1024) uint8_t o2_smem[];
    __shared__ __align__(8) unsigned long long done_bar;
    __shared__ uint32_t tmem_base_sh;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const uint32_t smem0 = (smem_u32(o2_smem) + 1023u) & ~1023u;
    uint32_t* words = reinterpret_cast<uint32_t*>(o2_smem + (smem0 - smem_u32(o2_smem)));
    for (int i = threadIdx.x; i < C::STAGE / 4; i += 128) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        words[i] = h;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;
    if (warp == 1 && rank == 0) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_i8(128 * CG, 256);
            for (long long it = 0; it < iters; ++it) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_i8<CG>(tmem_base + (uint32_t)(it & 1) * 256, umma_desc_sw128(smem0 + kk * 32),
                                umma_desc_sw128(smem0 + IMG + kk * 32), idesc, (it > 1 || kk > 0) ? 1u : 0u);
            }
            umma_commit<CG>(smem_u32(&done_bar));
        }
        __syncwarp();
    }
    mbar_wait<false>(smem_u32(&done_bar), 0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    if (warp == 1) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

static int peak_i8(long long iters, int cta_group, double* ops_host, cudaStream_t st) {
    DevAttr* da = nullptr;
    if (int rc = dev_attrs(&da)) return rc;
    static bool attr[PLMC_MAX_DEVICES];
    const int dev = current_device();
    if (dev < 0) return PLMC_ERR_LAUNCH;
    if (!attr[dev]) {
        if (cudaFuncSetAttribute(peak_i8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM) !=
                cudaSuccess ||
            cudaFuncSetAttribute(peak_i8_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM) !=
                cudaSuccess)
            return PLMC_ERR_LAUNCH;
        attr[dev] = true;
    }
    int ctas;
    if (cta_group == 2 && da->clusters > 0) {
        ctas = 2 * da->clusters;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(ctas);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = Cfg<2>::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, peak_i8_kernel<2>, iters) != cudaSuccess) {
            cudaGetLastError();
            return PLMC_ERR_LAUNCH;
        }
    } else {
        ctas = da->sms;
        peak_i8_kernel<1><<<ctas, 128, Cfg<1>::SMEM, st>>>(iters);
    }
    PLMC_CHECK_LAUNCH();
    if (ops_host) *ops_host = (double)ctas * (double)iters * 4.0 * 2.0 * 128.0 * 256.0 * 32.0;
    return PLMC_OK;
}

}  // namespace o2

int rns_bits(int nmod, int K) { return o2::rns_bits(nmod, K); }
long long rns_ws_bytes(int M, int N, int K, int nmod, bool same_operand, bool lower) {
    return o2::rns_ws_bytes(M, N, K, nmod, same_operand, lower);
}

// Products with a LOWER-triangular square operand X (order n, stored in place: whatever lies above its diagonal is
// ignored) as ONE launch set: planes of both operands, the INT8 product over the k-tiles that can be nonzero only
// (GemmArgs2::tri), reconstruction.  The recursions of csrc/linalg.cu cut such a product into ~n/512 leaf products
// with K = 512 and updates of every size down to that; here every output tile runs its full inner dimension in
// one pass at the rate of a large GEMM, and the flop count is the triangular one (plus the 256-blocks on the diagonal).
//   mode 1: C[m x n] = alpha B[m x n] X            mode 2: C[n x m] = alpha X B[n x m]
//   mode 3: C[n x m] = alpha X^T B[n x m]          mode 4: C[n x n] (lower blocks) = alpha X^T X
//   mode 5: C[m x n] = alpha B[m x n] X^T
// C may alias B (modes 1-3) or X (mode 4): the operands are read into planes before anything is written.
// Returns 1 (nothing launched) when the scratch does not hold one batch member.
int rns_trmm(int mode, const double* X, long long ldx, long long sX, const double* B, long long ldb, long long sB,
             double* C, long long ldc, long long sC, int n, int m, double alpha, double beta, int nmod, int batch,
             void* ws, long long ws_bytes, int flags, cudaStream_t st) {
    if (mode < 1 || mode > 5 || nmod < 4 || nmod > o2::MAXMOD || (n % 128) || n <= 0 || batch < 1) return PLMC_ERR_BADARG;
    if (mode != 4 && ((m % 128) || m <= 0)) return PLMC_ERR_BADARG;
    if (n > 65536 - 128) return 1;
    uint8_t* w0 = (uint8_t*)(((uintptr_t)ws + 1023) / 1024 * 1024);
    const long long avail = ws_bytes - (long long)(w0 - (uint8_t*)ws);
    const int M = (mode == 1 || mode == 5) ? m : n, N = (mode == 1 || mode == 4 || mode == 5) ? n : m;
    if (o2::rns_ws_bytes(M, N, n, nmod, mode == 4, mode == 4) - 1024 > avail) return 1;
    switch (mode) {
        case 1:   // A = B_in (m x n, KC), op(B)[k][j] = X[k][j] (MC, zero for k < j)
            return o2::rns_gemm_fit(true, false, B, ldb, sB, X, ldx, sX, C, ldc, sC, m, n, n, alpha, beta, 0, nmod, false,
                                    batch, w0, avail, flags, st, 3);
        case 2:   // A = X (KC, zero for k > i), op(B)[k][j] = B_in[k][j] (MC)
            return o2::rns_gemm_fit(true, false, X, ldx, sX, B, ldb, sB, C, ldc, sC, n, m, n, alpha, beta, 0, nmod, false,
                                    batch, w0, avail, flags, st, 1);
        case 3:   // A = X^T (MC, zero for k < i), op(B) = B_in (MC)
            return o2::rns_gemm_fit(false, false, X, ldx, sX, B, ldb, sB, C, ldc, sC, n, m, n, alpha, beta, 0, nmod, false,
                                    batch, w0, avail, flags, st, 2);
        case 5:   // A = B_in (m x n, KC), op(B)[k][j] = X[j][k]: the KC planes of X (zero for k > j)
            return o2::rns_gemm_fit(true, true, B, ldb, sB, X, ldx, sX, C, ldc, sC, m, n, n, alpha, beta, 0, nmod, false,
                                    batch, w0, avail, flags, st, 4);
        default:  // X^T X, lower blocks: both operands are the MC planes of X
            return o2::rns_gemm_fit(false, false, X, ldx, sX, X, ldx, sX, C, ldc, sC, n, n, n, alpha, beta, 1, nmod, true,
                                    batch, w0, avail, flags, st, 2);
    }
}

// C[b] = alpha op(A[b]) op(B[b]) + beta C[b].  A product whose planes and residues do not fit the scratch is
// split along N, then M (the SYRK into two SYRKs and a GEMM); batch members are processed in passes.
int rns_gemm(bool aKC, bool bKC, const double* A, long long lda, long long sA, const double* B, long long ldb,
             long long sB, double* C, long long ldc, long long sC, int M, int N, int K, double alpha, double beta,
             int lower, int nmod, bool same_operand, int batch, void* ws, long long ws_bytes, int flags,
             cudaStream_t st) {
    if (nmod < 4 || nmod > o2::MAXMOD || (M % 128) || (N % 128) || (K % 128) || M <= 0 || N <= 0 || K <= 0 ||
        batch < 1)
        return PLMC_ERR_BADARG;
    if ((same_operand || lower) && (M != N)) return PLMC_ERR_BADARG;
    uint8_t* w0 = (uint8_t*)(((uintptr_t)ws + 1023) / 1024 * 1024);
    const long long avail = ws_bytes - (long long)(w0 - (uint8_t*)ws);
    // INT32 accumulators (and the hi/lo reduction of the epilogue) are exact for K 2^14 < 2^30
    const bool k_ok = K <= 65536 - 128;
    if (k_ok && o2::rns_ws_bytes(M, N, K, nmod, same_operand, lower != 0) - 1024 <= avail)
        return o2::rns_gemm_fit(aKC, bKC, A, lda, sA, B, ldb, sB, C, ldc, sC, M, N, K, alpha, beta, lower, nmod,
                                same_operand, batch, w0, avail, flags, st);
    if (!k_ok) {   // accumulate over halves of K (beta = 1 on the second half)
        const int k1 = (int)o2::round_up(K / 2, 128);
        const double* A2 = aKC ? A + k1 : A + (long long)k1 * lda;
        const double* B2 = bKC ? B + k1 : B + (long long)k1 * ldb;
        int rc = rns_gemm(aKC, bKC, A, lda, sA, B, ldb, sB, C, ldc, sC, M, N, k1, alpha, beta, lower, nmod,
                          same_operand, batch, ws, ws_bytes, flags, st);
        if (rc) return rc;
        return rns_gemm(aKC, bKC, A2, lda, sA, B2, ldb, sB, C, ldc, sC, M, N, K - k1, alpha, 1.0, lower, nmod,
                        same_operand, batch, ws, ws_bytes, flags, st);
    }
    auto a_rows = [&](int r) { return aKC ? A + (long long)r * lda : A + r; };
    auto b_cols = [&](int c) { return bKC ? B + (long long)c * ldb : B + c; };
    if (lower) {
        if (M <= 256) return PLMC_ERR_BADARG;
        const int n1 = (int)o2::round_up(M / 2, 256), n2 = M - n1;
        int rc = rns_gemm(aKC, bKC, A, lda, sA, B, ldb, sB, C, ldc, sC, n1, n1, K, alpha, beta, 1, nmod, same_operand,
                          batch, ws, ws_bytes, flags, st);
        if (rc) return rc;
        rc = rns_gemm(aKC, bKC, a_rows(n1), lda, sA, B, ldb, sB, C + (long long)n1 * ldc, ldc, sC, n2, n1, K, alpha,
                      beta, 0, nmod, false, batch, ws, ws_bytes, flags, st);
        if (rc) return rc;
        return rns_gemm(aKC, bKC, a_rows(n1), lda, sA, b_cols(n1), ldb, sB, C + (long long)n1 * ldc + n1, ldc, sC, n2,
                        n2, K, alpha, beta, 1, nmod, same_operand, batch, ws, ws_bytes, flags, st);
    }
    if (N >= M && N > 256) {
        const int n1 = (int)o2::round_up(N / 2, 256);
        int rc = rns_gemm(aKC, bKC, A, lda, sA, B, ldb, sB, C, ldc, sC, M, n1, K, alpha, beta, 0, nmod, false, batch,
                          ws, ws_bytes, flags, st);
        if (rc) return rc;
        return rns_gemm(aKC, bKC, A, lda, sA, b_cols(n1), ldb, sB, C + n1, ldc, sC, M, N - n1, K, alpha, beta, 0, nmod,
                        false, batch, ws, ws_bytes, flags, st);
    }
    if (M > 256) {
        const int m1 = (int)o2::round_up(M / 2, 256);
        int rc = rns_gemm(aKC, bKC, A, lda, sA, B, ldb, sB, C, ldc, sC, m1, N, K, alpha, beta, 0, nmod, false, batch,
                          ws, ws_bytes, flags, st);
        if (rc) return rc;
        return rns_gemm(aKC, bKC, a_rows(m1), lda, sA, B, ldb, sB, C + (long long)m1 * ldc, ldc, sC, M - m1, N, K,
                        alpha, beta, 0, nmod, false, batch, ws, ws_bytes, flags, st);
    }
    if (K > 128) {   // last resort: accumulate over halves of K
        const int k1 = (int)o2::round_up(K / 2, 128);
        auto a_k = [&](int k) { return aKC ? A + k : A + (long long)k * lda; };
        auto b_k = [&](int k) { return bKC ? B + k : B + (long long)k * ldb; };
        int rc = rns_gemm(aKC, bKC, A, lda, sA, B, ldb, sB, C, ldc, sC, M, N, k1, alpha, beta, 0, nmod, false, batch, ws,
                          ws_bytes, flags, st);
        if (rc) return rc;
        return rns_gemm(aKC, bKC, a_k(k1), lda, sA, b_k(k1), ldb, sB, C, ldc, sC, M, N, K - k1, alpha, 1.0, 0, nmod,
                        false, batch, ws, ws_bytes, flags, st);
    }
    return PLMC_ERR_BADARG;
}

}  // namespace plmc

extern "C" {

int plmc_rns_bits(int moduli, int K) {
    if (moduli < 4 || moduli > plmc::o2::MAXMOD || K <= 0) return PLMC_ERR_BADARG;
    return plmc::rns_bits(moduli, K);
}

long long plmc_rns_ws_bytes(int M, int N, int K, int moduli, int same_operand, int lower) {
    return plmc::rns_ws_bytes(M, N, K, moduli, same_operand != 0, lower != 0);
}

/* host only: the reconstruction constants the library uses for `moduli` moduli and inner dimension K */
int plmc_rns_constants(int moduli, int K, int* moduli_out, double* H, double* L, double* pscale, int* bits) {
    if (moduli < 4 || moduli > plmc::o2::MAXMOD || K <= 0 || !H || !L || !pscale || !bits) return PLMC_ERR_BADARG;
    plmc::o2::CrtArgs a;
    *bits = plmc::rns_bits(moduli, K);
    plmc::o2::crt_constants(moduli, *bits, a);
    for (int i = 0; i < moduli; ++i) {
        H[i] = a.H[i];
        L[i] = a.L[i];
        if (moduli_out) moduli_out[i] = plmc::o2::kModuli[i];
    }
    *pscale = a.pscale;
    return PLMC_OK;
}

int plmc_peak_i8(long long iters, int cta_group, void* scratch, double* ops_host, void* stream) {
    (void)scratch;
    if (iters <= 0 || (cta_group != 1 && cta_group != 2)) return PLMC_ERR_BADARG;
    return plmc::o2::peak_i8(iters, cta_group, ops_host, (cudaStream_t)stream);
}

int plmc_rns_gemm(int layout, const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
                  int M, int N, int K, double alpha, double beta, int lower, int moduli, int same_operand, void* ws,
                  long long ws_bytes, int flags, void* stream) {
    if (!A || !B || !C || !ws) return PLMC_ERR_BADARG;
    const bool aKC = !(layout & 2), bKC = !(layout & 1);
    return plmc::rns_gemm(aKC, bKC, A, lda, 0, B, ldb, 0, C, ldc, 0, M, N, K, alpha, beta, lower, moduli,
                          same_operand != 0, 1, ws, ws_bytes, flags, (cudaStream_t)stream);
}
}
