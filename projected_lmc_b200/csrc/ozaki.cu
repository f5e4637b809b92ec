// FP64 GEMM through the 5th-generation tensor cores (tcgen05 / TMEM / bulk async copies):
// Ozaki-style error-free slicing of the FP64 operands into signed 8-bit planes,
// exact INT8 x INT8 -> INT32 products on tcgen05.mma.kind::i8, FP64 recombination.
//
// Why: B200 has no FP64 kind on tcgen05 and its DMMA/DFMA datapath tops out at
// 37 TFLOP/s (measured; the two share one unit, see plmc_peak_mixed).  The INT8
// tensor path is two orders of magnitude wider, so a product of two FP64 matrices
// split into s 8-bit planes costs s(s+1)/2 INT8 GEMMs and still wins.
//
//   x = A[m,k] * 2^-(ea[m]+1)  (|x| < 1/2),  Q = rint(x * 2^(8s-1))  (a 64-bit integer, exact up to the rounding)
//   Q = sum_i a_i 256^(s-1-i) with BALANCED base-256 digits a_i in [-128, 127] (a_0 in [-65, 65]):
//   A[m,k] = 2^(ea[m]+1) * sum_i a_i 2^-(7+8i) + O(2^-8s)     (same for columns of B)
//   C[m,n] += alpha * 2^(ea[m]+eb[n]-12) * sum_{g<s} 2^-8g * D_g[m,n],
//   D_g = sum_{i+j=g} A_i B_j^T   (INT32, exact for K-chunks <= 16384)
// Balanced digits use all 8 bits of a signed plane, so s planes carry 8s-1 bits (s = 6: 47 bits) where
// the usual sign-magnitude truncation carries 7s.
//
// Data layout ("tile images"): the slicers write the planes of an operand directly in the
// order and byte pattern the tensor core wants in shared memory:
//     planes[batch][x / 128][k / 32][plane][4096 B]
// where one 4096-byte image is the canonical K-major SWIZZLE_32B UMMA tile of 128 rows x 32 bytes
// (row r, 16-byte chunk c at  r*32 + ((c ^ ((r >> 2) & 1)) * 16)).  A stage of the pipeline is then
// ONE contiguous 28 KB bulk copy for A plus s 2 KB copies for B (a 64-row B tile is one half of an
// image), issued as cp.async.bulk with mbarrier completion: full 128-byte lines per request.  (The
// first version used 4-D TMA tensor maps with 32-byte inner boxes and was bound by the
// L1TEX->XBAR request rate, ncu: l1tex__m_l1tex2xbar_req_cycles_active 83 %.)
//
// Kernel: persistent, one CTA per SM walking 128 x 64 output tiles; warp 0 = copy producer (runs ahead
// into the next tile), warp 1 = MMA issuer (one elected thread), warps 2..9 = epilogue (two per TMEM lane
// quarter).  s accumulators (one per diagonal g, 64 TMEM columns each, 448 of 512 columns at s = 7).  All pairs (i, j) of one A plane i are a
// single wide MMA: their accumulators g = i..s-1 are consecutive TMEM column blocks and the
// B planes consecutive shared-memory row blocks, so A_i is read from shared memory once.
#include "linalg.cuh"

namespace plmc {

constexpr int OZ_BM = 128, OZ_BN = 64, OZ_BK = 32;   // BK in int8 elements = bytes
constexpr int OZ_SMAX = 7;                           // TMEM: 7 accumulators x 64 columns = 448 <= 512
constexpr int OZ_STAGES = 5;                         // ring depth (4 at s = 7: 227 KB of shared memory per CTA)
constexpr int OZ_EPI_WARPS = 8;                      // two per TMEM lane quarter, 32 columns each
constexpr int OZ_STG_BYTES = OZ_EPI_WARPS * 4096;    // epilogue staging: one 32 x 16 FP64 block per epilogue warp
constexpr int OZ_THREADS = 64 + 32 * OZ_EPI_WARPS;   // warp0 producer, warp1 MMA + TMEM alloc, warps 2.. epilogue
constexpr int OZ_KCHUNK = 16384;                     // INT32 exactness: 7 * 16384 * 128^2 < 2^31
constexpr int OZ_IMG = OZ_BM * OZ_BK;                // 4096 B: one plane of one (128-row, 32-k) tile
constexpr int OZ_B_PLANE = OZ_BN * OZ_BK;            // 2048 B: half an image

__host__ __device__ inline int oz_stage_bytes(int s) { return s * (OZ_IMG + OZ_B_PLANE); }
__host__ __device__ inline int oz_stages(int s) { return s >= 7 ? 4 : OZ_STAGES; }
__host__ __device__ inline int oz_smem_bytes(int s) { return oz_stages(s) * oz_stage_bytes(s) + OZ_STG_BYTES + 1024; }

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init_(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS UBLKCP)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
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
// K-major, SWIZZLE_32B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_32B=6 [61,64)
// rows are 32 B apart, 8-row groups 256 B apart (SBO); LBO is unused for swizzled K-major (1).
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::i8, S32 accumulate, K-major A and B
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// exact int32 -> double without the (slow) I2F.F64 pipe: 2^52 + 2^31 + r is representable, subtract the bias
__device__ __forceinline__ double i32_to_f64(uint32_t r) {
    return __hiloint2double(0x43300000, (int)(r ^ 0x80000000u)) - 4503601774854144.0;
}

// byte offset of element (x, k) of plane i inside the tile-image layout of one batch member
__device__ __forceinline__ long long img_offset(int x, int k, int i, int s, int nkt) {
    const int tile = x >> 7, r = x & 127, kt = k >> 5, kk = k & 31;
    const int c = kk >> 4, b = kk & 15;
    return ((long long)((long long)tile * nkt + kt) * s + i) * OZ_IMG + r * 32 + ((c ^ ((r >> 2) & 1)) << 4) + b;
}

struct OzArgs {
    double* C;
    long long ldc, sC;      // leading dimension and batch stride of C (elements)
    const int8_t* PA;       // tile images of op(A): [batch][M/128][K/32][s][4096]
    const int8_t* PB;       // tile images of op(B): [batch][N/128][K/32][s][4096]
    long long sPA, sPB;     // batch strides (bytes)
    const int* ea;          // [batch][M] row exponents of op(A)
    const int* eb;          // [batch][N] column exponents of op(B)
    int M, N, K;            // M % 128 == 0, N % 128 == 0, K % 32 == 0
    int s;                  // planes (1..7)
    double alpha;
    double beta;            // C = alpha * A B + beta * C   (C is not read when beta == 0)
    int lower;              // M == N: only tiles that touch the lower triangle (C origin on the diagonal)
    int per_member;         // output tiles of one batch member
    int total;              // per_member * batch members of this launch
    long long* dbg;         // diagnostics (plmc_ozaki_debug): clock64 stamps of CTA 0, 8 per tile, or NULL
};
#define OZ_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && nt < 64) p.dbg[nt * 8 + (slot)] = clock64(); } while (0)

// work item w -> (batch member, 128-row block, 64-column block)
struct OzTile { int bz, tm, tn; };
__device__ __forceinline__ OzTile oz_decode(const OzArgs& p, int w) {
    OzTile t;
    t.bz = w / p.per_member;
    const int idx = w - t.bz * p.per_member;
    if (p.lower) {
        // row block tm owns the 2 tm + 2 column blocks tn <= 2 tm + 1: idx in [tm (tm+1), (tm+1)(tm+2))
        int tm = (int)((sqrtf(4.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while (tm * (tm + 1) > idx) --tm;
        while ((tm + 1) * (tm + 2) <= idx) ++tm;
        t.tm = tm;
        t.tn = idx - tm * (tm + 1);
    } else {
        // grouped raster: bands of 8 tile-rows walk the columns together, so the CTAs in flight share
        // 8 A row-blocks and ~18 B column-blocks out of L2 instead of streaming all of B
        const int tiles_m = p.M / OZ_BM, tiles_n = p.N / OZ_BN;
        const int GROUP = 8;
        const int per_group = GROUP * tiles_n;
        const int gid = idx / per_group;
        const int first = gid * GROUP;
        const int gsz = min(tiles_m - first, GROUP);
        const int rem = idx - gid * per_group;
        t.tm = first + rem % gsz;
        t.tn = rem / gsz;
    }
    return t;
}

// Persistent: one CTA per SM walks the work items blockIdx.x, blockIdx.x + gridDim.x, ... ; every item costs
// the same (same K), so the static round-robin is balanced to one tile.  The copy ring, the accumulator
// hand-over and TMEM live across items: the producer is already filling the ring for the next tile while
// the epilogue warps drain the accumulators of the current one.
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(const OzArgs p) {
    extern __shared__ __align__(1024) uint8_t oz_smem[];
    __shared__ __align__(8) unsigned long long full_bar[OZ_STAGES], empty_bar[OZ_STAGES], acc_full, acc_empty;
    __shared__ uint32_t tmem_base_sh;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = p.s;
    const int stage_bytes = oz_stage_bytes(s);
    const int nst = oz_stages(s);
    const uint32_t smem0 = (smem_u32(oz_smem) + 1023u) & ~1023u;
    uint8_t* stg_base = oz_smem + (smem0 - smem_u32(oz_smem)) + nst * stage_bytes;   // 16 KB after the ring
    const int nkt = p.K / OZ_BK;
    const int kt_per_chunk = OZ_KCHUNK / OZ_BK;
    const int nchunks = (nkt + kt_per_chunk - 1) / kt_per_chunk;

    if (threadIdx.x == 0) {
        for (int i = 0; i < OZ_STAGES; ++i) {
            mbar_init_(smem_u32(&full_bar[i]), 1);
            mbar_init_(smem_u32(&empty_bar[i]), 1);
        }
        mbar_init_(smem_u32(&acc_full), 1);
        mbar_init_(smem_u32(&acc_empty), OZ_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_sh))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;

    if (warp == 0) {
        // ===== copy producer: one 28 KB bulk copy (all A planes) + s 2 KB copies (B planes) per stage =====
        if (elect_one()) {
            int it = 0;   // k-steps issued so far (ring position across work items)
            int nt = 0;
            for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++nt) {
                const OzTile t = oz_decode(p, w);
                const int8_t* ga = p.PA + (long long)t.bz * p.sPA + (long long)t.tm * nkt * s * OZ_IMG;
                // the 64-row B tile tn is half (tn & 1) of the image rows of 128-row tile tn >> 1
                const int8_t* gb = p.PB + (long long)t.bz * p.sPB + (long long)(t.tn >> 1) * nkt * s * OZ_IMG +
                                   (t.tn & 1) * OZ_B_PLANE;
                for (int kt = 0; kt < nkt; ++kt, ++it) {
                    const int st = it % nst, round = it / nst;
                    if (round > 0) mbar_wait_(smem_u32(&empty_bar[st]), (round - 1) & 1);
                    const uint32_t fb = smem_u32(&full_bar[st]);
                    if (kt == 0) OZ_STAMP(5);
                    if (kt == nkt - 1) OZ_STAMP(6);
                    mbar_expect_tx_(fb, (uint32_t)stage_bytes);
                    const uint32_t sa = smem0 + (uint32_t)st * stage_bytes;
                    const uint32_t sb = sa + (uint32_t)s * OZ_IMG;
                    bulk_load(sa, ga + (long long)kt * s * OZ_IMG, (uint32_t)(s * OZ_IMG), fb);
                    const int8_t* gbk = gb + (long long)kt * s * OZ_IMG;
                    for (int j = 0; j < s; ++j) bulk_load(sb + j * OZ_B_PLANE, gbk + j * OZ_IMG, OZ_B_PLANE, fb);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int it = 0, phase = 0;   // ring position; accumulation phases (one per work item and K-chunk) so far
        int nt = 0;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++nt) {
            for (int ch = 0; ch < nchunks; ++ch, ++phase) {
                if (phase > 0) {   // accumulators must have been drained by the epilogue warps
                    mbar_wait_(smem_u32(&acc_empty), (phase - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                if (lane == 0) OZ_STAMP(0);
                const int kt0 = ch * kt_per_chunk, kt1 = min(nkt, kt0 + kt_per_chunk);
                for (int kt = kt0; kt < kt1; ++kt, ++it) {
                    const int st = it % nst, round = it / nst;
                    mbar_wait_(smem_u32(&full_bar[st]), round & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t sa = smem0 + (uint32_t)st * stage_bytes;
                        const uint32_t sb = sa + (uint32_t)s * OZ_IMG;
                        // Plane A_i meets B_0 .. B_{s-1-i}; their products belong to the accumulators of the
                        // diagonals g = i .. s-1, which are CONSECUTIVE 64-column blocks of TMEM, and the B
                        // planes are consecutive 64-row blocks of shared memory.  So all pairs of one i are a
                        // single wide MMA (N = 64 (s-i), split at the instruction limit N = 256): A_i is read
                        // from shared memory once instead of s-i times.
                        for (int i = 0; i < s; ++i) {
                            const uint64_t da = umma_desc_sw32(sa + i * OZ_IMG);
                            const int nplanes = s - i;
                            const uint32_t accum = (kt > kt0 || i > 0) ? 1u : 0u;   // i = 0 touches every accumulator
                            const int np1 = nplanes < 4 ? nplanes : 4;
                            umma_i8(tmem_base + (uint32_t)i * OZ_BN, da, umma_desc_sw32(sb),
                                    umma_idesc_i8(OZ_BM, OZ_BN * np1), accum);
                            if (nplanes > 4)
                                umma_i8(tmem_base + (uint32_t)(i + 4) * OZ_BN, da, umma_desc_sw32(sb + 4 * OZ_B_PLANE),
                                        umma_idesc_i8(OZ_BM, OZ_BN * (nplanes - 4)), accum);
                        }
                        umma_commit(smem_u32(&empty_bar[st]));            // frees the smem stage when the MMAs retire
                        if (kt == kt1 - 1) {
                            umma_commit(smem_u32(&acc_full));  // accumulators complete for this chunk
                            OZ_STAMP(1);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> FP64 recombination -> shared-memory transpose -> C =====
        // A TMEM lane is an output ROW, so each thread recombines 16 columns of its own row; the 32 x 16
        // block of the warp is then turned through a swizzled staging tile so that global loads (beta) and
        // stores run along rows: 16 lanes x 8 B = one full 128-byte line per half warp.
        const int quad = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;          // which 32 of the 64 columns
        double* stg = reinterpret_cast<double*>(stg_base + (warp - 2) * 4096);
        const int rsub = lane >> 4, col = lane & 15;
        int phase = 0;
        int nt = 0;
        for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++nt) {
            const OzTile t = oz_decode(p, w);
            const int m0 = t.tm * OZ_BM, n0 = t.tn * OZ_BN;
            const int row = m0 + quad * 32 + lane;
            const double sa = p.alpha * scalbn(1.0, max(p.ea[(long long)t.bz * p.M + row], -1022) - 12);
            const int* ebz = p.eb + (long long)t.bz * p.N + n0;
            double* cw = p.C + (long long)t.bz * p.sC + (long long)(m0 + quad * 32) * p.ldc + n0;
            for (int ch = 0; ch < nchunks; ++ch, ++phase) {
                mbar_wait_(smem_u32(&acc_full), phase & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (warp == 2 && lane == 0) OZ_STAMP(2);
                const double beta = (ch > 0) ? 1.0 : p.beta;   // later K-chunks accumulate onto the first
                // phase A: drain TMEM.  v[b][j] = sum_g 2^-8g D_g[row][8 b + j], this warp's 32 columns of the row kept
                // in registers, so the accumulators go back to the MMA warp after ~3 us and the read-modify-
                // write of C below overlaps the next tile's main loop.
                double v[OZ_BN / 16][8];
#pragma unroll
                for (int b = 0; b < OZ_BN / 16; ++b) {
                    uint32_t r[OZ_SMAX][8];
#pragma unroll
                    for (int gI = 0; gI < OZ_SMAX; ++gI)
                        if (gI < s)
                            tmem_ld8(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(gI * OZ_BN + 32 * half + 8 * b),
                                     r[gI]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[b][j] = 0.0;
#pragma unroll
                    for (int gI = OZ_SMAX - 1; gI >= 0; --gI)
                        if (gI < s) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[b][j] = fma(v[b][j], 0.00390625, i32_to_f64(r[gI][j]));
                        }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_(smem_u32(&acc_empty));
                if (warp == 2 && lane == 0) OZ_STAMP(3);

                // phase B: 16 columns at a time through the staging tile
#pragma unroll
                for (int q4 = 0; q4 < OZ_BN / 32; ++q4) {
                    const int c0 = 32 * half + 16 * q4;
                    double oldv[16];
                    if (beta != 0.0) {   // 16 independent coalesced loads in flight
#pragma unroll
                        for (int i = 0; i < 16; ++i) oldv[i] = cw[(long long)(2 * i + rsub) * p.ldc + c0 + col];
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c)   // 16-byte chunk c of row `lane` lands at chunk c ^ (lane & 7)
                        *reinterpret_cast<double2*>(stg + lane * 16 + ((c ^ (lane & 7)) << 1)) =
                            make_double2(sa * v[2 * q4 + (c >> 2)][2 * (c & 3)], sa * v[2 * q4 + (c >> 2)][2 * (c & 3) + 1]);
                    __syncwarp();
                    const double fcol = scalbn(1.0, max(ebz[c0 + col], -1022));
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int rr = 2 * i + rsub;
                        double out = stg[rr * 16 + ((((col >> 1) ^ (rr & 7)) << 1) | (col & 1))] * fcol;
                        if (beta != 0.0) out = fma(beta, oldv[i], out);
                        cw[(long long)rr * p.ldc + c0 + col] = out;
                    }
                    __syncwarp();
                }
                if (warp == 2 && lane == 0) OZ_STAMP(4);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// slicing: FP64 -> s signed 8-bit planes in tile-image order, per-row power-of-two scaling
//   KC operand: element (x, k) at P[x*ld + k]   ;   MC operand: element (x, k) at P[k*ld + x]
// ex[x] = exponent with |P(x,:)| * 2^-ex < 1
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int exp_for(double amax) {
    if (!(amax > 0.0)) return 0;
    int e;
    frexp(amax, &e);     // amax = f * 2^e, f in [0.5, 1)  ->  amax * 2^-e < 1
    return max(e, -1022);
}

// 4 consecutive k (k4 % 4 == 0) of row x with row exponent e (|v| * 2^-e < 1): one 32-bit store per plane.
// Q = rint(v * 2^(8s-2-e)); adding 0x80 to every byte position turns the balanced base-256 digits of Q
// into the plain bytes of Q + 0x80..80, and x - 128 (mod 256) is x ^ 0x80: so the digits of one element
// cost one 64-bit add and xor, and PRMT gathers byte (s-1-i) of the four elements into the word of plane i.
__device__ __forceinline__ void slice4_store(const double (&v)[4], int e, int x, int k4, int s, int nkt,
                                             int8_t* planes) {
    int8_t* dst = planes + img_offset(x, k4, 0, s, nkt);
    const int sh = 8 * s - 2 - e;                                   // |sh| < 1100: split, 2^sh may overflow
    const double m1 = __longlong_as_double((long long)(1023 + sh / 2) << 52);        // 2^(sh/2), exponent field
    const double m2 = __longlong_as_double((long long)(1023 + sh - sh / 2) << 52);
    const unsigned long long bias = 0x8080808080808080ull >> (8 * (8 - s));
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const unsigned long long q =
            ((unsigned long long)__double2ll_rn(v[u] * m1 * m2) + bias) ^ 0x8080808080808080ull;
        lo[u] = (uint32_t)q;
        hi[u] = (uint32_t)(q >> 32);
    }
    for (int i = 0; i < s; ++i) {
        const int b = s - 1 - i;                                    // byte of Q that is the digit of plane i
        const bool h = b >= 4;
        const uint32_t sel = (uint32_t)(b & 3) | ((uint32_t)(4 + (b & 3)) << 4);
        const uint32_t t01 = __byte_perm(h ? hi[0] : lo[0], h ? hi[1] : lo[1], sel);
        const uint32_t t23 = __byte_perm(h ? hi[2] : lo[2], h ? hi[3] : lo[3], sel);
        *reinterpret_cast<uint32_t*>(dst + (long long)i * OZ_IMG) = __byte_perm(t01, t23, 0x5410);
    }
}

// one CTA per row (KC): coalesced along k
__global__ void __launch_bounds__(256) slice_kc_row_kernel(const double* __restrict__ Pb, long long ld, long long sP,
                                                       int X, int K, int s, int8_t* __restrict__ planes_b,
                                                       long long sPl, int* __restrict__ ex_b) {
    __shared__ double red[8];
    __shared__ int e_sh;
    const int x = blockIdx.x;
    const double* P = Pb + (long long)blockIdx.z * sP;
    int8_t* planes = planes_b + (long long)blockIdx.z * sPl;
    int* ex = ex_b + (long long)blockIdx.z * X;
    const double* row = P + (long long)x * ld;
    double amax = 0.0;
    for (int k = threadIdx.x; k < K; k += 256) amax = fmax(amax, fabs(row[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = red[0];
        for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
        e_sh = exp_for(m);
        ex[x] = e_sh;
    }
    __syncthreads();
    const int e = e_sh;
    const int nkt = K / OZ_BK;
    for (int k4 = threadIdx.x * 4; k4 < K; k4 += 1024) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = row[k4 + u];
        slice4_store(v, e, x, k4, s, nkt, planes);
    }
}

// KC operand: one WARP per row (8 rows per CTA, no block-level synchronisation).  Rows of up to
// 128 * NCH elements stay in registers between the absmax pass and the slicing pass; longer rows are
// read twice (the second time from L1/L2).
template <int NCH>
__global__ void __launch_bounds__(256) slice_kc_kernel(const double* __restrict__ Pb, long long ld, long long sP,
                                                       int X, int K, int s, int8_t* __restrict__ planes_b,
                                                       long long sPl, int* __restrict__ ex_b) {
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (x >= X) return;
    const double* row = Pb + (long long)blockIdx.z * sP + (long long)x * ld;
    int8_t* planes = planes_b + (long long)blockIdx.z * sPl;
    const int nkt = K / OZ_BK;
    const bool al16 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    auto load4 = [&](int k4, double (&v)[4]) {
        if (al16) {
            const double2 a = *reinterpret_cast<const double2*>(row + k4);
            const double2 b = *reinterpret_cast<const double2*>(row + k4 + 2);
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = row[k4 + u];
        }
    };
    double amax = 0.0;
    if (NCH > 0) {
        double v[NCH > 0 ? NCH : 1][4];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int k4 = c * 128 + lane * 4;
            if (k4 < K) {
                load4(k4, v[c]);
#pragma unroll
                for (int u = 0; u < 4; ++u) amax = fmax(amax, fabs(v[c][u]));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const int e = exp_for(amax);
        if (lane == 0) ex_b[(long long)blockIdx.z * X + x] = e;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int k4 = c * 128 + lane * 4;
            if (k4 < K) slice4_store(v[c], e, x, k4, s, nkt, planes);
        }
    } else {
        for (int k4 = lane * 4; k4 < K; k4 += 128) {
            double v[4];
            load4(k4, v);
#pragma unroll
            for (int u = 0; u < 4; ++u) amax = fmax(amax, fabs(v[u]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const int e = exp_for(amax);
        if (lane == 0) ex_b[(long long)blockIdx.z * X + x] = e;
        for (int k4 = lane * 4; k4 < K; k4 += 128) {
            double v[4];
            load4(k4, v);
            slice4_store(v, e, x, k4, s, nkt, planes);
        }
    }
}

// MC operand, pass 1: exponent of the column-wise absmax over k.  The frexp exponent is monotone in
// |x|, so the maximum of the per-element exponents is taken with an integer atomicMax over k-slabs
// (order independent -> deterministic).  ex must be pre-set to a very negative value (memset 0x80).
__global__ void __launch_bounds__(256) absmax_mc_kernel(const double* __restrict__ Pb, long long ld, long long sP,
                                                        int X, int K, int* __restrict__ ex_b) {
    __shared__ int red[4][64];
    const int xl = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int x = blockIdx.x * 64 + xl;
    const double* P = Pb + (long long)blockIdx.z * sP;
    int* ex = ex_b + (long long)blockIdx.z * X;
    const int k0 = blockIdx.y * 256, k1 = min(K, k0 + 256);
    int emax = -2000000000;
    if (x < X) {
        for (int k = k0 + grp; k < k1; k += 4) {
            const long long bits = __double_as_longlong(P[(long long)k * ld + x]);
            const int be = (int)((bits >> 52) & 0x7FF);          // biased exponent; 0 for zero / subnormal
            const int e = be ? be - 1022 : -1022;                   // frexp exponent (|x| * 2^-e in [0.5, 1))
            if (bits << 1) emax = max(emax, e);                     // skip exact zeros
        }
    }
    red[grp][xl] = emax;
    __syncthreads();
    if (grp == 0 && x < X) {
        emax = max(max(red[0][xl], red[1][xl]), max(red[2][xl], red[3][xl]));
        if (emax > -2000000000) atomicMax(ex + x, emax);
    }
}

// MC operand, pass 2: 32(k) x 128(x) tiles transposed through shared memory; one CTA writes one whole
// 4096-byte image per plane (8 lanes cover the 32 bytes of a row, 4 rows per warp store)
__global__ void __launch_bounds__(256) slice_mc_kernel(const double* __restrict__ Pb, long long ld, long long sP,
                                                       int X, int K, int s, int8_t* __restrict__ planes_b,
                                                       long long sPl, const int* __restrict__ ex_b) {
    __shared__ double tile[32][129];
    const double* P = Pb + (long long)blockIdx.z * sP;
    int8_t* planes = planes_b + (long long)blockIdx.z * sPl;
    const int* ex = ex_b + (long long)blockIdx.z * X;
    const int x0 = blockIdx.x * 128, k0 = blockIdx.y * 32;
    {
        const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;   // 2 k-rows of 128 x per pass
        double t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = P[(long long)(k0 + 2 * i + ty) * ld + x0 + tx];
#pragma unroll
        for (int i = 0; i < 16; ++i) tile[2 * i + ty][tx] = t[i];
    }
    __syncthreads();
    const int kl = (threadIdx.x & 7) * 4;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int xl = pass * 32 + (threadIdx.x >> 3);
        const int e = max(ex[x0 + xl], -1022);   // all-zero column: any exponent works
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = tile[kl + u][xl];
        slice4_store(v, e, x0 + xl, k0 + kl, s, K / OZ_BK, planes);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static bool g_oz_attr[PLMC_MAX_DEVICES];   // per device: opt-in shared memory of the kernel set, SM count read
static int g_oz_sms_dev[PLMC_MAX_DEVICES];
static long long* g_oz_dbg = nullptr;      // diagnostics only (plmc_ozaki_debug)

// scratch for ONE batch member (tile images of both operands + exponents)
long long ozaki_ws_bytes(int M, int N, int K, int s, bool same_operand) {
    const long long a = (long long)s * M * K, b = same_operand ? 0 : (long long)s * N * K;
    const long long pad = 1024;
    return ((a + pad - 1) / pad) * pad + ((b + pad - 1) / pad) * pad + ((4LL * (M + N) + pad - 1) / pad) * pad + pad;
}

// C[b] = alpha * op(A[b]) op(B[b]) + beta * C[b]  for b < batch (as many members per pass as the scratch holds).
// aKC / bKC as in gemm_dmma (bKC: B(k,n) at B[n*ldb+k]).  same_operand: op(B)^T == op(A) (SYRK).
int ozaki_gemm(bool aKC, bool bKC, const double* A, long long lda, long long sA, const double* B, long long ldb,
               long long sB, double* C, long long ldc, long long sC, int M, int N, int K, double alpha, double beta,
               int lower, int s, bool same_operand, int batch, void* ws, long long ws_bytes, cudaStream_t st) {
    if (s < 1 || s > OZ_SMAX || (M % OZ_BM) || (N % OZ_BM) || (K % 32) || batch < 1) return PLMC_ERR_BADARG;
    if ((same_operand || lower) && (M != N)) return PLMC_ERR_BADARG;
    const long long pad = 1024;
    uint8_t* w0 = (uint8_t*)(((uintptr_t)ws + pad - 1) / pad * pad);
    const long long avail = ws_bytes - (long long)(w0 - (uint8_t*)ws);
    // ozaki_ws_bytes carries one extra `pad` so that a scratch of exactly that size still fits after
    // the base pointer has been rounded up to 1024 bytes
    const long long need1 = ozaki_ws_bytes(M, N, K, s, same_operand) - pad;
    if (avail < need1) return PLMC_ERR_BADARG;
    const int bc_max = (int)((avail / need1) < batch ? (avail / need1) : batch);
    const long long bytesA = (long long)s * M * K;                       // multiple of 4096
    const long long bytesB = same_operand ? 0 : (long long)s * N * K;
    const int dev = current_device();
    if (dev < 0) return PLMC_ERR_LAUNCH;
    if (!g_oz_attr[dev]) {
        int smem_max = 0;
        for (int t = 1; t <= OZ_SMAX; ++t) smem_max = oz_smem_bytes(t) > smem_max ? oz_smem_bytes(t) : smem_max;
        if (cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max) !=
            cudaSuccess)
            return PLMC_ERR_LAUNCH;
        int sms = 0;
        g_oz_sms_dev[dev] =
            (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) ? sms : 148;
        g_oz_attr[dev] = true;
    }
    const int g_oz_sms = g_oz_sms_dev[dev];
    const int smem = oz_smem_bytes(s);

    for (int b0 = 0; b0 < batch; b0 += bc_max) {
        const int bc = (batch - b0) < bc_max ? (batch - b0) : bc_max;
        // scratch of this pass: A images [bc][...] | B images [bc][...] | ea [bc][M] | eb [bc][N]
        int8_t* pa = (int8_t*)w0;
        int8_t* pb = same_operand ? pa : (int8_t*)(w0 + bytesA * bc);
        int* ea = (int*)(w0 + (bytesA + bytesB) * bc);
        int* eb = same_operand ? ea : ea + (long long)bc * M;
        auto slice = [&](bool kc, const double* P, long long ld, long long sP, int X, int8_t* planes,
                         long long sPl, int* ex) {
            if (kc) {
                const dim3 g((X + 7) / 8, 1, bc);
                if (K <= 512) slice_kc_kernel<4><<<g, 256, 0, st>>>(P, ld, sP, X, K, s, planes, sPl, ex);
                else if (K <= 1024) slice_kc_kernel<8><<<g, 256, 0, st>>>(P, ld, sP, X, K, s, planes, sPl, ex);
                else slice_kc_row_kernel<<<dim3(X, 1, bc), 256, 0, st>>>(P, ld, sP, X, K, s, planes, sPl, ex);
            } else {
                cudaMemsetAsync(ex, 0x80, sizeof(int) * (size_t)X * bc, st);
                absmax_mc_kernel<<<dim3((X + 63) / 64, (K + 255) / 256, bc), 256, 0, st>>>(P, ld, sP, X, K, ex);
                slice_mc_kernel<<<dim3(X / 128, K / 32, bc), 256, 0, st>>>(P, ld, sP, X, K, s, planes, sPl, ex);
            }
        };
        slice(aKC, A + (long long)b0 * sA, lda, sA, M, pa, bytesA, ea);
        if (!same_operand) slice(bKC, B + (long long)b0 * sB, ldb, sB, N, pb, bytesB, eb);
        PLMC_CHECK_LAUNCH();
        OzArgs p;
        p.C = C + (long long)b0 * sC; p.ldc = ldc; p.sC = sC;
        p.PA = pa; p.PB = pb; p.sPA = bytesA; p.sPB = same_operand ? bytesA : bytesB;
        p.ea = ea; p.eb = eb;
        p.M = M; p.N = N; p.K = K; p.s = s;
        p.alpha = alpha; p.beta = beta; p.lower = lower;
        const long long tiles_m = M / OZ_BM;
        const long long per = lower ? tiles_m * (tiles_m + 1) : tiles_m * (N / OZ_BN);
        if (per * bc > 2000000000LL) return PLMC_ERR_BADARG;
        p.per_member = (int)per; p.total = (int)(per * bc);
        p.dbg = g_oz_dbg;
        const int ctas = p.total < g_oz_sms ? p.total : g_oz_sms;
        ozaki_gemm_kernel<<<ctas, OZ_THREADS, smem, st>>>(p);
        PLMC_CHECK_LAUNCH();
        note_launch(3);
    }
    return PLMC_OK;
}

}  // namespace plmc

extern "C" {

int plmc_ozaki_debug(long long* stamps) {
    plmc::g_oz_dbg = stamps;
    return PLMC_OK;
}

long long plmc_ozaki_ws_bytes(int M, int N, int K, int slices, int same_operand) {
    return plmc::ozaki_ws_bytes(M, N, K, slices, same_operand != 0);
}

int plmc_ozaki_gemm(int layout, const double* A, long long lda, const double* B, long long ldb, double* C,
                    long long ldc, int M, int N, int K, double alpha, double beta, int lower, int slices,
                    int same_operand, void* ws, long long ws_bytes, void* stream) {
    if (!A || !B || !C || !ws) return PLMC_ERR_BADARG;
    const bool aKC = !(layout & 2), bKC = !(layout & 1);
    return plmc::ozaki_gemm(aKC, bKC, A, lda, 0, B, ldb, 0, C, ldc, 0, M, N, K, alpha, beta, lower, slices,
                            same_operand != 0, 1, ws, ws_bytes, (cudaStream_t)stream);
}
}
