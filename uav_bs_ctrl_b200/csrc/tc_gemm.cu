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
//   warps 0-3  producers : coalesced 16-byte global loads of the A tile chunk (128 rows x 32 floats), hi/lo split in
//                          registers, st.shared into the canonical K-major SWIZZLE_128B layout, fence.proxy.async,
//                          mbarrier arrive (full[stage])
//   warp  8    MMA issuer: one thread issues 12 tcgen05.mma.kind::tf32 (M=128, N, K=8) per chunk and commits to
//                          empty[stage] / tmem_full[acc]; this warp also owns tcgen05.alloc / dealloc
//   warps 4-7  epilogue  : tcgen05.ld 32x32b (each warp its own TMEM lane quadrant), bias + ReLU, 64-byte row stores
//   W (both halves) is split once per CTA and stays resident in shared memory; accumulators are double buffered in
//   TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {
namespace tc {

constexpr int BM = 128, BK = 32, NPROD = 128, NTHREADS = 288;

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
    if (warp == 8) {
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

    if (warp < 4) {
        // ================================ producers ================================
        const int t = threadIdx.x;                                  // 0..127
        const int c = t & 7, r0 = t >> 3;                           // chunk within the 128-byte row, first row
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long m0 = tile * BM;
            for (int kc = 0; kc < KC; ++kc, ++it) {
                const int s = it % S;
                const uint32_t ph = (it / S) & 1;
                float4 x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {                       // issue the global loads before waiting for the slot
                    const long long row = m0 + r0 + 16 * j;
                    x[j] = row < a.M ? __ldg(reinterpret_cast<const float4*>(a.A + row * a.lda + (size_t)kc * BK) + c)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                mbar_wait(empty + s, ph ^ 1);
                char* hi = sA + (size_t)s * 2 * BM * 128;
                char* lo = hi + BM * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) split_store(hi, lo, r0 + 16 * j, c, x[j]);
                fence_proxy_async();
                mbar_arrive(full + s);
            }
        }
    } else if (warp == 8) {
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
        // ================================ epilogue (warps 4..7) ================================
        const int q = warp - 4;                                     // TMEM lane quadrant of this warp (warp % 4)
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
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
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
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(tf32x3_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    const long long n_tiles = (M + BM - 1) / BM;
    const int grid = (int)(n_tiles < ubs::kNumSMs ? n_tiles : ubs::kNumSMs);
    tf32x3_gemm_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_tf32x3_gemm");
}
