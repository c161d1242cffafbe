// 3xTF32 dense projection on the 5th-generation tensor cores (tcgen05 + TMEM), fp32 in / fp32 out.
//
//     C[M, N] = act( A[M, K] . W[N, K]^T + bias )          A, W row-major (K contiguous) == nn.Linear layout
//
// This is the "dense per-relation feature projection" of the north-star: the batched, observation-side GEMMs of a
// BPTT window (x = relu(W_aggr xin + b), [pv | pg] = W_x x + b, dx = [dgi | dv] W_dx, d_xin = dpre W_aggr; M = T*N
// ~ 10^5 rows, K, N <= 288) and the sweep's fc_src / fc_dst / res_fc.  1e-5 fp32 parity rules out single-pass TF32
// (2e-4 error), so every operand is split  x = hi + lo  (hi = x rounded to TF32, lo = x - hi rounded to
// TF32) and  hi*hi + hi*lo + lo*hi  is accumulated in the fp32 TMEM accumulator: error ~2^-22 per product, unbiased.
//
// Structure (one CTA per SM, persistent over 128-row tiles of A; no TMA descriptors needed because the split has to
// pass through the CUDA cores anyway):
//   warps 0-11 producers : three groups of 128 threads on alternate chunks (48 KB of loads in flight per SM): coalesced
//                          16-byte global loads of the A tile chunk (128 rows x 32 floats), hi/lo split in registers,
//                          st.shared into the canonical K-major SWIZZLE_128B layout, fence.proxy.async, mbarrier
//                          arrive (full[stage])
//   warp  16   MMA issuer: one thread issues 12 tcgen05.mma.kind::tf32 (M=128, N, K=8) per chunk and commits to
//                          empty[stage] / tmem_full[acc]; this warp also owns tcgen05.alloc / dealloc
//   warps 12-15 epilogue : tcgen05.ld 32x32b (each warp its own TMEM lane quadrant), bias + ReLU, 64-byte row stores
//   W (both halves) is split once per CTA and stays resident in shared memory; accumulators are double buffered in
//   TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {
namespace tc {

// NGRP producer groups of 128 threads take alternate chunks: one group holds a 16 KB chunk in registers between its
// loads and its stores, and 16 KB in flight per SM is far below what HBM latency x bandwidth asks for (~35 KB)
constexpr int BM = 128, BK = 32, NPROD = 128, NGRP = 3, NTHREADS = (4 * NGRP + 5) * 32;
constexpr int EPI_WARP0 = 4 * NGRP, MMA_WARP = 4 * NGRP + 4;          // EPI_WARP0 % 4 == 0: warp % 4 = its TMEM lane quadrant

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 bytes, 8-row swizzle
// atoms 1024 bytes apart (SBO), version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                      // descriptor version
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}
// byte offset of (row r, 16-byte chunk c) inside a K-major SWIZZLE_128B tile (Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void split_store(char* hi_tile, char* lo_tile, int r, int c, float4 x) {
    // hi = x rounded to nearest TF32 (10 explicit mantissa bits), lo = (x - hi) rounded to nearest TF32: rounding
    // (instead of the hardware's truncation) keeps the residual unbiased, so it does not grow linearly with K
    auto rn = [](float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); };
    float4 h, l;
    h.x = rn(x.x); l.x = rn(x.x - h.x);
    h.y = rn(x.y); l.y = rn(x.y - h.y);
    h.z = rn(x.z); l.z = rn(x.z - h.z);
    h.w = rn(x.w); l.w = rn(x.w - h.w);
    const uint32_t off = sw128(r, c);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
}

struct Args {
    const float* A; const float* W; const float* bias; float* C;
    long long lda, ldw, ldc, M;
    int N, K, relu, stages, tmem_cols, nbuf;      // nbuf accumulator buffers, each {main (hi*hi) | cross (hi*lo + lo*hi)}
};

__global__ void __launch_bounds__(NTHREADS, 1) tf32x3_gemm_kernel(const Args a) {
    extern __shared__ __align__(1024) char smem_raw[];
    char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);    // SWIZZLE_128B tiles need 1024-B alignment
    const int N = a.N, K = a.K, KC = K / BK, S = a.stages;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    // ---- shared memory carve-up (every tile 1024-byte aligned)
    const uint32_t w_tile = (uint32_t)N * 128;                    // one K-chunk of W: N rows x 128 B
    char* sWhi = smem;
    char* sWlo = sWhi + (size_t)KC * w_tile;
    char* sA = sWlo + (size_t)KC * w_tile;                        // S stages x {hi 16 KB, lo 16 KB}
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * 2 * BM * 128);
    uint64_t* full = bars;            // [S]
    uint64_t* empty = bars + S;       // [S]
    uint64_t* tfull = bars + 2 * S;   // [2]
    uint64_t* tempty = tfull + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, NPROD); mbar_init(empty + s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 128); }   // a.nbuf of them are used
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // W: split once, resident for every tile of this CTA
    for (int i = threadIdx.x; i < N * (K / 4); i += NTHREADS) {
        const int n = i / (K / 4), c4 = i - n * (K / 4);            // c4: 16-byte chunk along K
        const float4 x = __ldg(reinterpret_cast<const float4*>(a.W + (size_t)n * a.ldw) + c4);
        const int kc = c4 / 8, c = c4 % 8;
        split_store(sWhi + (size_t)kc * w_tile, sWlo + (size_t)kc * w_tile, n, c, x);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long n_tiles = (a.M + BM - 1) / BM;

    if (warp < EPI_WARP0) {
        // ================================ producers ================================
        const int grp = warp >> 2, t = threadIdx.x & 127;           // group, thread within the group
        const int c = t & 7, r0 = t >> 3;                           // chunk within the 128-byte row, first row
        const long long my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const uint32_t n_it = (uint32_t)my_tiles * (uint32_t)KC;    // chunks of this CTA, in the order the MMA warp eats them
        // chunk `it` = (tile it / KC, kc = it % KC) lives in stage it % S; this group takes it = grp, grp + ngrp, ...
        // No more groups than stages: a group waits for the MMAs of chunk it - S by PARITY, which is only unambiguous
        // if that barrier cannot be two phases behind — i.e. if the group's previous chunk it - ngrp is not older than
        // it - S (the MMAs complete in order).
        const uint32_t ngrp = S < NGRP ? (uint32_t)S : (uint32_t)NGRP;
        uint32_t tl = (uint32_t)grp / (uint32_t)KC, kc = (uint32_t)grp % (uint32_t)KC;
        const uint32_t d_tl = ngrp / (uint32_t)KC, d_kc = ngrp % (uint32_t)KC;
        for (uint32_t it = grp; (uint32_t)grp < ngrp && it < n_it; it += ngrp) {
            const uint32_t s = it % (uint32_t)S, ph = (it / (uint32_t)S) & 1u;
            const long long m0 = ((long long)blockIdx.x + (long long)tl * gridDim.x) * BM;
            const float* src = a.A + (m0 + r0) * a.lda + (size_t)kc * BK;
            float4 x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {                           // issue the global loads before waiting for the slot
                x[j] = m0 + r0 + 16 * j < a.M ? __ldg(reinterpret_cast<const float4*>(src + (size_t)(16 * j) * a.lda) + c)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(empty + s, ph ^ 1u);
            char* hi = sA + (size_t)s * 2 * BM * 128;
            char* lo = hi + BM * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) split_store(hi, lo, r0 + 16 * j, c, x[j]);
            fence_proxy_async();
            mbar_arrive(full + s);
            tl += d_tl; kc += d_kc;
            if (kc >= (uint32_t)KC) { kc -= (uint32_t)KC; ++tl; }
        }
    } else if (warp == MMA_WARP) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint32_t it = 0, tc_ = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc_) {
                const int acc = tc_ % a.nbuf;
                mbar_wait(tempty + acc, ((tc_ / a.nbuf) & 1) ^ 1);
                tc_fence_after();
                // The tensor core adds into the fp32 accumulator with truncation, an error that grows with the number
                // of accumulating MMAs and the magnitude of the sum: keep the large hi*hi terms (K/8 MMAs) apart from
                // the 2^-11-times smaller cross terms (2K/8 MMAs) and add the two accumulators in the epilogue (RN).
                const uint32_t d = tmem_base + (uint32_t)acc * (uint32_t)(2 * N), dx = d + (uint32_t)N;
                for (int kc = 0; kc < KC; ++kc, ++it) {
                    const int s = it % S;
                    mbar_wait(full + s, (it / S) & 1);
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(sA + (size_t)s * 2 * BM * 128), alo = ahi + BM * 128;
                    const uint32_t whi = smem_u32(sWhi + (size_t)kc * w_tile), wlo = smem_u32(sWlo + (size_t)kc * w_tile);
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint32_t o = ks * 32;                 // 8 tf32 = 32 bytes along K inside the swizzle atom
                        umma_tf32(d, make_desc(ahi + o), make_desc(whi + o), idesc, (kc | ks) != 0);
                        umma_tf32(dx, make_desc(ahi + o), make_desc(wlo + o), idesc, (kc | ks) != 0);
                        umma_tf32(dx, make_desc(alo + o), make_desc(whi + o), idesc, 1);
                    }
                    umma_commit(empty + s);                         // smem slot free once these MMAs have read it
                }
                umma_commit(tfull + acc);                           // accumulator ready for the epilogue
            }
        }
    } else {
        // ================================ epilogue (4 warps) ================================
        const int q = warp - EPI_WARP0;                             // TMEM lane quadrant of this warp (warp % 4)
        uint32_t tc_ = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc_) {
            const int acc = tc_ % a.nbuf;
            mbar_wait(tfull + acc, (tc_ / a.nbuf) & 1);
            tc_fence_after();
            const long long row = tile * BM + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * (uint32_t)(2 * N);
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr + (uint32_t)c0));
                uint32_t u[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(taddr + (uint32_t)(N + c0)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < a.M) {
                    float* out = a.C + row * a.ldc + c0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 o;
                        o.x = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                        o.y = __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]);
                        o.z = __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]);
                        o.w = __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]);
                        if (a.bias != nullptr) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + j));
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4*>(out + j) = o;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty + acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Weight-gradient products of the BPTT window:  C[Mo, No] = A[R, Mo]^T . B[R, No]   (R = T*N ~ 10^5 rows: the
// REDUCTION dimension is the slow one of both operands).  Same 3xTF32 tcgen05 pipeline; the producers transpose while
// they split: lane = row r of a 32-row chunk, so the 4-byte st.shared of one warp land in 32 different banks of the
// K-major SWIZZLE_128B tile (K = r).  R is cut into slices of 256 rows (8 chunks = 32 accumulating MMAs per TMEM
// accumulator, the same accumulation depth as the forward projections — the tensor core adds into fp32 with
// truncation); every (128-row tile of Mo, slice) work item writes one partial tile and a second kernel adds the
// partials in a fixed order (deterministic, round-to-nearest).
constexpr int SLICE = 256;

struct ArgsTN {
    const float* A; const float* B; float* ws;
    long long lda, ldb, R;
    int Mo, No, stages, tmem_cols, nbuf, n_mtiles, a_vec, b_vec, depth, stage_stg_bytes;
    long long n_items;
};

// 4 consecutive columns of one row, zero outside the matrix.  vec: rows are 16-byte aligned and ncols % 4 == 0, so a
// float4 is either fully inside or fully outside; the load itself is unconditional (clamped address) so that all the
// loads of a chunk are in flight together.
__device__ __forceinline__ float4 load4(const float* base, long long row, long long ld, int col, int ncols, long long nrows, int vec) {
    if (vec) {
        const bool ok = row < nrows && col < ncols;
        const long long r = row < nrows ? row : nrows - 1;
        const int c = col < ncols ? col : 0;
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + r * ld + c));
        return ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const long long r = row < nrows ? row : nrows - 1;
    const float* p = base + r * ld;
    float4 v;
    v.x = __ldg(p + (col < ncols ? col : 0));
    v.y = __ldg(p + (col + 1 < ncols ? col + 1 : 0));
    v.z = __ldg(p + (col + 2 < ncols ? col + 2 : 0));
    v.w = __ldg(p + (col + 3 < ncols ? col + 3 : 0));
    const bool okr = row < nrows;
    v.x = (okr && col < ncols) ? v.x : 0.f;
    v.y = (okr && col + 1 < ncols) ? v.y : 0.f;
    v.z = (okr && col + 2 < ncols) ? v.z : 0.f;
    v.w = (okr && col + 3 < ncols) ? v.w : 0.f;
    return v;
}

// element (tile row m, k = r) of a K-major SWIZZLE_128B tile
__device__ __forceinline__ void split_store_t(char* hi_tile, char* lo_tile, int m, int r, float x) {
    const float h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    const float d = x - h;
    const float l = __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xFFFFE000u);
    const uint32_t off = sw128(m, r >> 2) + (uint32_t)((r & 3) << 2);
    *reinterpret_cast<float*>(hi_tile + off) = h;
    *reinterpret_cast<float*>(lo_tile + off) = l;
}

// 2 x 8 producer warps | 4 epilogue | 1 MMA.  The two producer groups take alternate chunks (one UMMA stage each): a
// chunk costs a producer warp ~500 dependent-ish instructions, and two warps per scheduler cannot hide their latency.
constexpr int TN_PROD_WARPS = 8, TN_GROUPS = 2, TN_EPI_WARP0 = TN_PROD_WARPS * TN_GROUPS, TN_MMA_WARP = TN_EPI_WARP0 + 4;
constexpr int TN_THREADS = (TN_MMA_WARP + 1) * 32;

// Asynchronous copy of 4 consecutive columns of one row into a private 16-byte staging slot, zero-filled outside the
// matrix (cp.async src-size 0): the global -> shared traffic of several chunks is in flight per thread without
// holding registers.
__device__ __forceinline__ void cp_async_row4(uint32_t dst, const float* base, long long row, long long ld, int col, int ncols,
                                              long long nrows, int vec) {
    const long long r = row < nrows ? row : nrows - 1;
    if (vec) {
        const int ok = (row < nrows && col < ncols) ? 16 : 0;
        const float* src = base + r * ld + (col < ncols ? col : 0);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok) : "memory");
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ok = (row < nrows && col + e < ncols) ? 4 : 0;
            const float* src = base + r * ld + (col + e < ncols ? col + e : 0);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4 * e), "l"(src), "r"(ok) : "memory");
        }
    }
}

// stage the (item, kc) chunk: slots [0,4) = this thread's A columns, [4, 4+nb4) = its B columns
__device__ __forceinline__ void tn_stage(const ArgsTN& a, uint32_t stg, long long item, int kc, int warp, int lane, int nb4) {
    const int mt = (int)(item % a.n_mtiles);
    const long long row = (item / a.n_mtiles) * SLICE + kc * BK + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        cp_async_row4(stg + (uint32_t)(j * TN_PROD_WARPS * 32) * 16, a.A, row, a.lda, mt * BM + warp * 16 + 4 * j, a.Mo, a.R, a.a_vec);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (j < nb4)
            cp_async_row4(stg + (uint32_t)((4 + j) * TN_PROD_WARPS * 32) * 16, a.B, row, a.ldb, 4 * (warp + TN_PROD_WARPS * j), a.No,
                          a.R, a.b_vec);
}

// ---- producer loop for 16-byte-aligned operands (the shapes of the BPTT window) --------------------------------------
// Every shared-memory access below is explicit ld/st.shared on 32-bit addresses with immediate offsets: through generic
// pointers the compiler re-derived the shared window (S2R SR_CgaCtaId) in front of each access and spilled.
template <int IMM>
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(IMM));
    return v;
}
// hi / lo TF32 halves of x to [bh + IMM_H] / [bl + IMM_L]
template <int IMM_H, int IMM_L>
__device__ __forceinline__ void st_hl(uint32_t bh, uint32_t bl, float x) {
    const float h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    const float l = __uint_as_float((__float_as_uint(x - h) + 0x1000u) & 0xFFFFE000u);
    asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(bh), "f"(h), "n"(IMM_H));
    asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(bl), "f"(l), "n"(IMM_L));
}
// A tile rows m = 16 warp + 4 J + e of k-column `lane`: K-major SWIZZLE_128B offset = (m >> 3) * 1024 + (m & 7) * 128 +
// ((lane >> 2) ^ (m & 7)) * 16 + (lane & 3) * 4; (m & 7) = (4 J + e) & 7 is a compile-time constant, so tA[k] = stage
// base + 2048 warp + the per-thread swizzle term for k, and the rest is an immediate.  lo tile = hi tile + 16 KB.
template <int J>
__device__ __forceinline__ void tn_put_a(const float4& x, const uint32_t (&tA)[8]) {
    constexpr int m = 4 * J;
#define UBS_TN_A(E, V) st_hl<((m + E) >> 3) * 1024 + ((m + E) & 7) * 128, ((m + E) >> 3) * 1024 + ((m + E) & 7) * 128 + BM * 128>(tA[(m + E) & 7], tA[(m + E) & 7], V)
    UBS_TN_A(0, x.x); UBS_TN_A(1, x.y); UBS_TN_A(2, x.z); UBS_TN_A(3, x.w);
#undef UBS_TN_A
}
// B tile rows n = 4 (warp + 8 J) + e: 4 KB further per J; tBh[e] / tBl[e] carry everything else
template <int J>
__device__ __forceinline__ void tn_put_b(const float4& x, const uint32_t (&tBh)[4], const uint32_t (&tBl)[4]) {
    st_hl<J * 4096, J * 4096>(tBh[0], tBl[0], x.x);
    st_hl<J * 4096, J * 4096>(tBh[1], tBl[1], x.y);
    st_hl<J * 4096, J * 4096>(tBh[2], tBl[2], x.z);
    st_hl<J * 4096, J * 4096>(tBh[3], tBl[3], x.w);
}

struct TnCursor { uint32_t slot, kc, item, mt, sl; };        // position of a chunk stream: staging slot, (item, kc)

// Coalesced 16-byte cp.async of chunk q into its staging slot: warp w copies rows w, w + 8, w + 16, w + 24 — 512
// contiguous bytes of the A row (lane = float4 column) and the No floats of the B row (lanes < No / 4).  Rows past R
// are zero-filled (src-size 0), A columns past Mo are skipped (never read back).
__device__ __forceinline__ void tn_stage_vec(const ArgsTN& a, const TnCursor& q, uint32_t slot_addr, uint32_t PW, int warp, int lane) {
    const long long row0 = (long long)q.sl * SLICE + q.kc * BK;
    const int left = (int)(a.R - row0 < BK ? a.R - row0 : BK);      // rows of this chunk inside the matrix (may be <= 0)
    const uint32_t dst = slot_addr + ((uint32_t)warp * PW + 4u * (uint32_t)lane) * 4u;
    const bool a_on = (int)q.mt * BM + 4 * lane < a.Mo, b_on = 4 * lane < a.No;
    const float* pa = a.A + (row0 + warp) * a.lda + ((int)q.mt * BM + 4 * lane);
    const float* pb = a.B + (row0 + warp) * a.ldb + 4 * lane;
    const long long sa = 8 * a.lda, sb = 8 * a.ldb;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool ok = warp + 8 * i < left;
        const int sz = ok ? 16 : 0;
        const float* srca = pa + i * sa;
        const float* srcb = pb + i * sb;
        srca = ok ? srca : a.A;                                     // nothing is read when sz == 0; keep the address valid
        srcb = ok ? srcb : a.B;
        if (a_on)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)i * 8u * PW * 4u), "l"(srca), "r"(sz) : "memory");
        if (b_on)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)i * 8u * PW * 4u + 512u), "l"(srcb), "r"(sz) : "memory");
    }
}

__device__ __forceinline__ void tn_cursor_next(TnCursor& q, uint32_t D, uint32_t n_mt) {
    constexpr int KC = SLICE / BK;
    q.slot = q.slot + 1 == D ? 0 : q.slot + 1;
    q.kc += TN_GROUPS;
    if (q.kc >= KC) { q.kc -= KC; q.item += gridDim.x; q.sl = q.item / n_mt; q.mt = q.item - q.sl * n_mt; }
}

// grp: producer group (chunks grp, grp + TN_GROUPS, ... of the CTA's chunk order; TN_GROUPS == stages == 2, so a group
// owns one UMMA stage); warp: 0..7 within the group.  Per chunk: wait for the staged rows, read them back transposed
// (thread = row `lane`; warp w = A columns 16 w .. 16 w + 15 and every 8th float4 column of B), refill the slot, then
// split and store — ~7 instructions per element (5 to split, 2 stores) plus ~100 per chunk.
__device__ __forceinline__ void tn_produce_vec(const ArgsTN& a, uint32_t sT, uint32_t sStage, uint64_t* full, uint64_t* empty,
                                               uint32_t stage_bytes, uint32_t b_tile, int grp, int warp, int lane) {
    constexpr int KC = SLICE / BK;
    static_assert(KC % TN_GROUPS == 0 && TN_GROUPS == 2, "a group keeps its kc residue and owns UMMA stage grp (stages == 2)");
    const int No = a.No;
    const uint32_t D = (uint32_t)a.depth;                           // staging slots of ONE group
    const int nb4 = (No / 4 - warp + TN_PROD_WARPS - 1) / TN_PROD_WARPS;          // <= 4 (No <= 128)
    const uint32_t n_items = (uint32_t)a.n_items, n_mt = (uint32_t)a.n_mtiles;
    const uint32_t my_items = n_items > blockIdx.x ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t n_chunks = my_items * (KC / TN_GROUPS);          // chunks of this group
    // staging slot: 32 rows x (128 A columns | No B columns | 4 pad) floats; the row pitch is 4 or 20 words mod 32, so
    // the transposed float4 read-back (8 lanes = 8 rows per wavefront) is conflict-free
    const uint32_t PW = 128u + (uint32_t)No + 4u;
    const uint32_t stg_stage = (uint32_t)a.stage_stg_bytes;
    const uint32_t stg_base = sStage + (uint32_t)grp * D * stg_stage;
    const uint32_t rd_off = (uint32_t)lane * PW * 4u + (uint32_t)warp * 64u;      // row lane, A column 16 warp
    const uint32_t c = (uint32_t)lane >> 2, lane_off = ((uint32_t)lane & 3u) << 2;
    uint32_t preA[8], preB[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) preA[k] = ((c ^ (uint32_t)k) << 4) + lane_off + (uint32_t)warp * 2048u;
    const uint32_t kb0 = 4u * ((uint32_t)warp & 1u);
#pragma unroll
    for (int e = 0; e < 4; ++e)
        preB[e] = ((c ^ (kb0 + e)) << 4) + lane_off + (kb0 + e) * 128u + ((uint32_t)warp >> 1) * 1024u + 2u * BM * 128u;
    const uint32_t bar_id = 1u + (uint32_t)grp;

    TnCursor qs, qc;                                                // staging / consuming stream
    qs.slot = 0; qs.kc = (uint32_t)grp; qs.item = blockIdx.x; qs.sl = qs.item / n_mt; qs.mt = qs.item - qs.sl * n_mt;
    qc = qs;
    for (uint32_t g = 0; g < D; ++g) {                              // prologue: D chunks in flight
        if (g < n_chunks) { tn_stage_vec(a, qs, stg_base + qs.slot * stg_stage, PW, warp, lane); tn_cursor_next(qs, D, n_mt); }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const uint32_t st = sT + (uint32_t)grp * stage_bytes;           // this group's UMMA stage {A hi | A lo | B hi | B lo}
    uint32_t tA[8], tBh[4], tBl[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) tA[k] = st + preA[k];
#pragma unroll
    for (int e = 0; e < 4; ++e) { tBh[e] = st + preB[e]; tBl[e] = tBh[e] + b_tile; }
    for (uint32_t g = 0; g < n_chunks; ++g) {
        if (D >= 3) asm volatile("cp.async.wait_group 2;" ::: "memory");            // chunk g of this thread has landed ...
        else if (D == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");                 // ... and everybody else's part of it
        const int v = (a.Mo - (int)qc.mt * BM - warp * 16) >> 2;                    // valid float4 A columns of this warp
        const int na = v < 0 ? 0 : (v > 4 ? 4 : v);
        const uint32_t rd = stg_base + qc.slot * stg_stage + rd_off;
        const uint32_t rdb = rd + 512u - 48u * (uint32_t)warp;                      // B column 4 warp: (128 + 4 warp) * 4 bytes
        // A columns past Mo hold stale staging data: read them anyway (in range), they are not stored below
        const float4 xa0 = lds128<0>(rd), xa1 = lds128<16>(rd), xa2 = lds128<32>(rd), xa3 = lds128<48>(rd);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 xb0 = z4, xb1 = z4, xb2 = z4, xb3 = z4;
        if (nb4 > 0) xb0 = lds128<0>(rdb);
        if (nb4 > 1) xb1 = lds128<128>(rdb);
        if (nb4 > 2) xb2 = lds128<256>(rdb);
        if (nb4 > 3) xb3 = lds128<384>(rdb);
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");                 // the slot has been read: refill it
        if (g + D < n_chunks) { tn_stage_vec(a, qs, stg_base + qs.slot * stg_stage, PW, warp, lane); tn_cursor_next(qs, D, n_mt); }
        asm volatile("cp.async.commit_group;" ::: "memory");
        tn_cursor_next(qc, D, n_mt);
        mbar_wait(empty + grp, (g & 1u) ^ 1u);                                      // stage grp, its g-th use
        if (na > 0) tn_put_a<0>(xa0, tA);
        if (na > 1) tn_put_a<1>(xa1, tA);
        if (na > 2) tn_put_a<2>(xa2, tA);
        if (na > 3) tn_put_a<3>(xa3, tA);
        if (nb4 > 0) tn_put_b<0>(xb0, tBh, tBl);
        if (nb4 > 1) tn_put_b<1>(xb1, tBh, tBl);
        if (nb4 > 2) tn_put_b<2>(xb2, tBh, tBl);
        if (nb4 > 3) tn_put_b<3>(xb3, tBh, tBl);
        fence_proxy_async();
        mbar_arrive(full + grp);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(TN_THREADS, 1) tf32x3_gemm_tn_kernel(const ArgsTN a) {
    extern __shared__ __align__(1024) char smem_raw[];
    char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int No = a.No, S = a.stages;
    constexpr int KC = SLICE / BK;                                  // chunks per work item
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t b_tile = (uint32_t)No * 128;
    const uint32_t stage_bytes = 2 * BM * 128 + 2 * b_tile;        // {A hi, A lo, B hi, B lo}
    char* sT = smem;
    char* sStage = sT + (size_t)S * stage_bytes;                   // cp.async staging: depth x slots x 256 x 16 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + (size_t)TN_GROUPS * a.depth * a.stage_stg_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, TN_PROD_WARPS * 32); mbar_init(empty + s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TN_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TN_EPI_WARP0 && a.a_vec && a.b_vec) {
        tn_produce_vec(a, smem_u32(sT), smem_u32(sStage), full, empty, stage_bytes, b_tile, warp / TN_PROD_WARPS, warp % TN_PROD_WARPS, lane);
    } else if (warp < TN_PROD_WARPS) {
        // ================================ producers (transposing split), any alignment ================================
        // warp w: 16 columns of the A tile and every 8th float4 column of the B tile; lane = row of the chunk.
        // cp.async keeps `depth` chunks in flight per thread in private staging slots (no registers held).
        const int nb4 = (No / 4 - warp + TN_PROD_WARPS - 1) / TN_PROD_WARPS;          // <= 4 (No <= 128)
        const int D = a.depth;                                                       // chunks in flight per thread
        const long long my_items = a.n_items > blockIdx.x ? (a.n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const long long n_chunks = my_items * KC;
        const uint32_t slots = 4 + (uint32_t)((No / 4 + TN_PROD_WARPS - 1) / TN_PROD_WARPS);
        const uint32_t stg_stage = slots * TN_PROD_WARPS * 32 * 16;
        const uint32_t stg0 = smem_u32(sStage) + (uint32_t)threadIdx.x * 16;
        for (int g = 0; g < D; ++g) {                                                // prologue: D chunks in flight
            if (g < n_chunks) tn_stage(a, stg0 + (uint32_t)g * stg_stage, blockIdx.x + (g / KC) * gridDim.x, g % KC, warp, lane, nb4);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (long long g = 0; g < n_chunks; ++g) {
            const uint32_t it = (uint32_t)g;
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            // chunk g has landed once at most D-1 younger groups are pending
            if (D == 6) asm volatile("cp.async.wait_group 5;" ::: "memory");
            else if (D == 5) asm volatile("cp.async.wait_group 4;" ::: "memory");
            else if (D == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
            else if (D == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else if (D == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            const char* stg = sStage + (size_t)(g % D) * stg_stage + (size_t)threadIdx.x * 16;
            float4 xa[4], xb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) xa[j] = *reinterpret_cast<const float4*>(stg + (size_t)(j * TN_PROD_WARPS * 32) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < nb4) xb[j] = *reinterpret_cast<const float4*>(stg + (size_t)((4 + j) * TN_PROD_WARPS * 32) * 16);
            // the slot is free again: refill it with chunk g + D
            if (g + D < n_chunks)
                tn_stage(a, stg0 + (uint32_t)(g % D) * stg_stage, blockIdx.x + ((g + D) / KC) * gridDim.x, (int)((g + D) % KC), warp, lane, nb4);
            asm volatile("cp.async.commit_group;" ::: "memory");
            mbar_wait(empty + s, ph ^ 1);
            char* ahi = sT + (size_t)s * stage_bytes;
            char* alo = ahi + BM * 128;
            char* bhi = alo + BM * 128;
            char* blo = bhi + b_tile;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = warp * 16 + 4 * j;
                split_store_t(ahi, alo, m, lane, xa[j].x);
                split_store_t(ahi, alo, m + 1, lane, xa[j].y);
                split_store_t(ahi, alo, m + 2, lane, xa[j].z);
                split_store_t(ahi, alo, m + 3, lane, xa[j].w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nb4) {
                    const int n = 4 * (warp + TN_PROD_WARPS * j);
                    split_store_t(bhi, blo, n, lane, xb[j].x);
                    split_store_t(bhi, blo, n + 1, lane, xb[j].y);
                    split_store_t(bhi, blo, n + 2, lane, xb[j].z);
                    split_store_t(bhi, blo, n + 3, lane, xb[j].w);
                }
            }
            fence_proxy_async();
            mbar_arrive(full + s);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (warp < TN_EPI_WARP0) {
        // second producer group: idle on the any-alignment path
    } else if (warp == TN_MMA_WARP) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(No >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint32_t it = 0, tc_ = 0;
            for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x, ++tc_) {
                const int acc = tc_ % a.nbuf;
                mbar_wait(tempty + acc, ((tc_ / a.nbuf) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)acc * (uint32_t)(2 * No), dx = d + (uint32_t)No;
                for (int kc = 0; kc < KC; ++kc, ++it) {
                    const int s = it % S;
                    mbar_wait(full + s, (it / S) & 1);
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(sT + (size_t)s * stage_bytes), alo = ahi + BM * 128;
                    const uint32_t bhi = alo + BM * 128, blo = bhi + b_tile;
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint32_t o = ks * 32;
                        umma_tf32(d, make_desc(ahi + o), make_desc(bhi + o), idesc, (kc | ks) != 0);
                        umma_tf32(dx, make_desc(ahi + o), make_desc(blo + o), idesc, (kc | ks) != 0);
                        umma_tf32(dx, make_desc(alo + o), make_desc(bhi + o), idesc, 1);
                    }
                    umma_commit(empty + s);
                }
                umma_commit(tfull + acc);
            }
        }
    } else {
        // ================================ epilogue (warps 8..11): partial tile -> workspace ================================
        const int q = warp - TN_EPI_WARP0;             // == warp % 4: the TMEM lane quadrant this warp may read
        uint32_t tc_ = 0;
        for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x, ++tc_) {
            const int acc = tc_ % a.nbuf;
            mbar_wait(tfull + acc, (tc_ / a.nbuf) & 1);
            tc_fence_after();
            const int mt = (int)(item % a.n_mtiles);
            const long long sl = item / a.n_mtiles;
            const int m = mt * BM + q * 32 + lane;                  // row of C
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * (uint32_t)(2 * No);
            for (int c0 = 0; c0 < No; c0 += 16) {
                uint32_t v[16], u[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(taddr + (uint32_t)(No + c0)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < a.Mo) {
                    float* out = a.ws + ((size_t)sl * a.Mo + m) * No + c0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 o;
                        o.x = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                        o.y = __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]);
                        o.z = __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]);
                        o.w = __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]);
                        *reinterpret_cast<float4*>(out + j) = o;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty + acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TN_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
    }
}

// C[m, n] = sum over slices of ws[sl, m, n] in a fixed order.  A block owns 32 consecutive outputs; its 8 warps take
// the slices w, w + 8, ... (16 independent loads in flight per thread, 128-byte coalesced), then the 8 partial sums are
// added in warp order: hundreds of slices cost a few dependent load rounds instead of one thread walking all of them.
constexpr int RED_OUT = 32, RED_WARPS = 8;
__global__ void __launch_bounds__(RED_OUT * RED_WARPS) tn_reduce_kernel(const float* __restrict__ ws, long long n_slices, int P, int No,
                                                                        float* __restrict__ C, long long ldc) {
    __shared__ float part[RED_WARPS][RED_OUT];
    const int lane = threadIdx.x % RED_OUT, w = threadIdx.x / RED_OUT;
    const int i = blockIdx.x * RED_OUT + lane;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    if (i < P) {
        long long sl = w;
        for (; sl + 15 * RED_WARPS < n_slices; sl += 16 * RED_WARPS) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += ws[(sl + j * RED_WARPS) * P + i];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (sl + j * RED_WARPS < n_slices) acc[j] += ws[(sl + j * RED_WARPS) * P + i];
    }
#pragma unroll
    for (int h = 8; h > 0; h >>= 1)
#pragma unroll
        for (int j = 0; j < h; ++j) acc[j] += acc[j + h];
    part[w][lane] = acc[0];
    __syncthreads();
    if (w == 0 && i < P) {
        float t = part[0][lane];
#pragma unroll
        for (int k = 1; k < RED_WARPS; ++k) t += part[k][lane];
        C[(long long)(i / No) * ldc + (i % No)] = t;
    }
}

}  // namespace tc
}  // namespace ubs

extern "C" UBS_API int ubs_tf32x3_gemm(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                                       float* C, int64_t ldc, int64_t M, int N, int K, int relu, void* stream) {
    using namespace ubs::tc;
    UBS_REQUIRE(A && W && C && M >= 0, "ubs_tf32x3_gemm: NULL argument");
    UBS_REQUIRE(K >= 32 && K % 32 == 0, "ubs_tf32x3_gemm: K must be a multiple of 32 (got %d)", K);
    UBS_REQUIRE(N >= 16 && N % 16 == 0 && N <= 256, "ubs_tf32x3_gemm: N must be a multiple of 16, <= 256 (got %d)", N);
    UBS_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0 && lda >= K && ldw >= K && ldc >= N, "ubs_tf32x3_gemm: bad leading dimensions");
    UBS_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)C % 16) == 0 &&
                (bias == nullptr || ((uintptr_t)bias % 16) == 0), "ubs_tf32x3_gemm: pointers must be 16-byte aligned");
    if (M == 0) return 0;
    const size_t w_bytes = 2 * (size_t)(K / BK) * N * 128;
    const size_t stage_bytes = 2 * BM * 128;
    const size_t fixed = w_bytes + 256 + 1024;                       // barriers + alignment slack
    int stages = (int)((227 * 1024 - fixed) / stage_bytes);
    if (stages > 6) stages = 6;
    if (stages < 2) { ubs::set_error("ubs_tf32x3_gemm: W (%d x %d) does not fit shared memory next to two A stages", N, K); return 3; }
    const int nbuf = 4 * N <= 512 ? 2 : 1;          // double-buffered accumulators when they fit the 512 TMEM columns
    int cols = 32;
    while (cols < 2 * N * nbuf) cols *= 2;
    Args a{};
    a.A = A; a.W = W; a.bias = bias; a.C = C; a.lda = lda; a.ldw = ldw; a.ldc = ldc; a.M = M;
    a.N = N; a.K = K; a.relu = relu; a.stages = stages; a.tmem_cols = cols; a.nbuf = nbuf;
    const size_t smem = w_bytes + stages * stage_bytes + 256 + 1024;
    UBS_OPT_IN_SMEM(tf32x3_gemm_kernel, "ubs_tf32x3_gemm");
    const long long n_tiles = (M + BM - 1) / BM;
    const int grid = (int)(n_tiles < ubs::kNumSMs ? n_tiles : ubs::kNumSMs);
    tf32x3_gemm_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_tf32x3_gemm");
}

extern "C" UBS_API int64_t ubs_tf32x3_gemm_tn_workspace(int64_t R, int Mo, int No) {
    const int64_t n_slices = (R + ubs::tc::SLICE - 1) / ubs::tc::SLICE;
    return n_slices * (int64_t)Mo * No;
}

extern "C" UBS_API int ubs_tf32x3_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                                          float* workspace, int64_t R, int Mo, int No, void* stream) {
    using namespace ubs::tc;
    UBS_REQUIRE(A && B && C && workspace && R >= 1, "ubs_tf32x3_gemm_tn: NULL argument");
    UBS_REQUIRE(Mo >= 1 && lda >= Mo, "ubs_tf32x3_gemm_tn: bad Mo / lda");
    UBS_REQUIRE(No >= 16 && No % 16 == 0 && No <= 128 && ldb >= No && ldc >= No, "ubs_tf32x3_gemm_tn: No must be a multiple of 16, <= 128 (got %d)", No);
    UBS_REQUIRE(((uintptr_t)workspace % 16) == 0, "ubs_tf32x3_gemm_tn: workspace must be 16-byte aligned");
    ArgsTN a{};
    a.A = A; a.B = B; a.ws = workspace; a.lda = lda; a.ldb = ldb; a.R = R; a.Mo = Mo; a.No = No;
    a.a_vec = (lda % 4 == 0 && Mo % 4 == 0 && ((uintptr_t)A % 16) == 0) ? 1 : 0;
    a.b_vec = (ldb % 4 == 0 && ((uintptr_t)B % 16) == 0) ? 1 : 0;
    static const bool generic = [] { const char* e = getenv("UBS_TN_GENERIC"); return e && e[0] == '1'; }();   // debugging: any-alignment path
    if (generic) a.a_vec = a.b_vec = 0;
    a.n_mtiles = (Mo + BM - 1) / BM;
    const long long n_slices = (R + SLICE - 1) / SLICE;
    a.n_items = n_slices * a.n_mtiles;
    const size_t stage_bytes = 2 * BM * 128 + 2 * (size_t)No * 128;           // UMMA operand tiles of one chunk
    // staging slot of one chunk: thread-private 16-byte slots (any-alignment path) or 32 padded rows (vector path)
    const size_t stg_priv = (size_t)(4 + (No / 4 + TN_PROD_WARPS - 1) / TN_PROD_WARPS) * TN_PROD_WARPS * 32 * 16;
    const size_t stg_rows = (size_t)BK * (BM + No + 4) * 4;
    const size_t stg_bytes = stg_priv > stg_rows ? stg_priv : stg_rows;
    const size_t budget = 227 * 1024 - 256 - 1024;
    const int stages = TN_GROUPS;                                              // one UMMA stage per producer group (the MMAs of a chunk are short)
    int depth = (int)((budget - stages * stage_bytes) / stg_bytes) / TN_GROUPS;    // staging slots per producer group
    if (depth > 3) depth = 3;
    UBS_REQUIRE(depth >= 1, "ubs_tf32x3_gemm_tn: tiles do not fit shared memory");
    a.stages = stages; a.depth = depth; a.stage_stg_bytes = (int)stg_bytes;
    a.nbuf = 4 * No <= 512 ? 2 : 1;
    int cols = 32;
    while (cols < 2 * No * a.nbuf) cols *= 2;
    a.tmem_cols = cols;
    const size_t smem = stages * stage_bytes + TN_GROUPS * depth * stg_bytes + 256 + 1024;
    UBS_OPT_IN_SMEM(tf32x3_gemm_tn_kernel, "ubs_tf32x3_gemm_tn");
    const int grid = (int)(a.n_items < ubs::kNumSMs ? a.n_items : ubs::kNumSMs);
    tf32x3_gemm_tn_kernel<<<grid, TN_THREADS, smem, (cudaStream_t)stream>>>(a);
    if (int rc = ubs::check_launch("ubs_tf32x3_gemm_tn")) return rc;
    const int P = Mo * No;
    tn_reduce_kernel<<<(P + RED_OUT - 1) / RED_OUT, RED_OUT * RED_WARPS, 0, (cudaStream_t)stream>>>(workspace, n_slices, P, No, C, ldc);
    return ubs::check_launch("ubs_tf32x3_gemm_tn(reduce)");
}
