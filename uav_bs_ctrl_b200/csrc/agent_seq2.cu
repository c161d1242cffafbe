// Persistent recurrent sequence kernels with the recurrent weights RESIDENT IN SHARED MEMORY (forward + backward).
//
// Same math as agent_step.cu (reference TarMAC.forward + GRUCell, algos/madrqn/agents/gnn_agents.py:248-271), but
// split along the only true dependency of the BPTT window: everything that depends on the observation alone
//     x  = relu(W_aggr [x_gt ‖ x_ubs] + b)         pv = W_vsq[:, :H] x + b_vsq         pg = W_ih[:, :H] x + b_ih
// is computed for ALL T·N rows by batched library GEMMs before this kernel (the encoder does not depend on h,
// gnn_agents.py:53), and the Q head / every x-path gradient / every parameter gradient after it.  What is left in the
// time loop is the part that really is sequential,
//     vsq = pv + W_vsq[:, H:] h        attention over the env block        gi = pg + W_ih[:, H:] c
//     gh  = W_hh h + b_hh              GRU gates -> h'
// 30.7 k of the 57.9 k MACs per agent row, whose weights (123 KB at H = M = 64, K = 16) stay in shared memory for the
// whole sequence: no per-step weight traffic at all, and the hidden state never leaves the SM.
// A CTA owns a tile of <= 16 agent rows (whole envs; envs never exchange data) for all timesteps.
#include "common.cuh"
#include "../../include/ubs_gnn.h"
#include <stdlib.h>

namespace ubs {
namespace seq2 {

#ifndef UBS_SEQ2_NT
#define UBS_SEQ2_NT 512
#endif
constexpr int R = 16, RP = 20, NT = UBS_SEQ2_NT;   // 16 warps per CTA: the only latency hiding a 1-CTA-per-SM kernel has

struct Dims {
    int H, M, K, U, flags;
    __host__ __device__ bool tarmac() const { return flags & UBS_STEP_TARMAC; }
    __host__ __device__ int V() const { return M + 2 * K; }
    __host__ __device__ int Vp() const { return (V() + 3) & ~3; }
    __host__ __device__ int rows_per_tile() const { return tarmac() ? (R / U) * U : R; }
};

// out[j*RP + r] = init + sum_{k<Kd} W[k*ldw + j] * A[k*RP + r],  W and A in SHARED memory (feature-major / K-major).
// init: row-major GLOBAL tile g_init[(row0 + r) * ld_init + j] (pv / pg; prefetched into registers before the k loop)
// and/or a per-column bias.  mode 0 store, 2 accumulate into out.  Thread tile 4 columns x 4 rows; split-K fills
// the CTA for narrow outputs.  All threads call; ends with __syncthreads().
__device__ __forceinline__ void gemm_s(const float* W, int ldw, const float* A, int Kd, float* out, int Nout,
                                       const float* __restrict__ g_init, int64_t row0, int n_valid, int64_t ld_init,
                                       const float* bias, int mode, float* scratch) {
    const int ntc = Nout >> 2, tiles = ntc * 4;
    int ksplit = 1;
    while (ksplit < 8 && tiles * ksplit * 2 <= NT && Kd >= ksplit * 16) ksplit *= 2;
    const int kchunk = (((Kd + ksplit - 1) / ksplit) + 3) & ~3;
    for (int base = 0; base < tiles * ksplit; base += NT) {
        const int t = base + threadIdx.x;
        const bool active = t < tiles * ksplit;
        const int ks = active ? t / tiles : 0, tile = active ? t - ks * tiles : 0;
        const int ct = tile % ntc, rt = tile / ntc;
        float acc[4][4];                                  // [row][col]
        float4 ini[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) ini[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && ks == 0 && g_init != nullptr) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (4 * rt + r < n_valid)
                    ini[r] = __ldg(reinterpret_cast<const float4*>(g_init + (row0 + 4 * rt + r) * ld_init + 4 * ct));
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
        if (active) {
            const int k0 = ks * kchunk, k1 = min(Kd, k0 + kchunk);
            const float* wp = W + 4 * ct;
            const float* ap = A + 4 * rt;
#pragma unroll 4
            for (int k = k0; k < k1; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(wp + k * ldw);
                const float4 x = *reinterpret_cast<const float4*>(ap + k * RP);
                acc[0][0] = fmaf(x.x, w.x, acc[0][0]); acc[0][1] = fmaf(x.x, w.y, acc[0][1]);
                acc[0][2] = fmaf(x.x, w.z, acc[0][2]); acc[0][3] = fmaf(x.x, w.w, acc[0][3]);
                acc[1][0] = fmaf(x.y, w.x, acc[1][0]); acc[1][1] = fmaf(x.y, w.y, acc[1][1]);
                acc[1][2] = fmaf(x.y, w.z, acc[1][2]); acc[1][3] = fmaf(x.y, w.w, acc[1][3]);
                acc[2][0] = fmaf(x.z, w.x, acc[2][0]); acc[2][1] = fmaf(x.z, w.y, acc[2][1]);
                acc[2][2] = fmaf(x.z, w.z, acc[2][2]); acc[2][3] = fmaf(x.z, w.w, acc[2][3]);
                acc[3][0] = fmaf(x.w, w.x, acc[3][0]); acc[3][1] = fmaf(x.w, w.y, acc[3][1]);
                acc[3][2] = fmaf(x.w, w.z, acc[3][2]); acc[3][3] = fmaf(x.w, w.w, acc[3][3]);
            }
        }
        if (ksplit > 1) {
            if (active && ks > 0) {
                float4* sp = reinterpret_cast<float4*>(scratch + ((ks - 1) * tiles + tile) * 16);
#pragma unroll
                for (int r = 0; r < 4; ++r) sp[r] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            }
            __syncthreads();
            if (active && ks == 0) {
                for (int s = 1; s < ksplit; ++s) {
                    const float4* sp = reinterpret_cast<const float4*>(scratch + ((s - 1) * tiles + tile) * 16);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 v = sp[r];
                        acc[r][0] += v.x; acc[r][1] += v.y; acc[r][2] += v.z; acc[r][3] += v.w;
                    }
                }
            }
        }
        if (active && ks == 0) {
            const float ic[4][4] = {{ini[0].x, ini[0].y, ini[0].z, ini[0].w}, {ini[1].x, ini[1].y, ini[1].z, ini[1].w},
                                    {ini[2].x, ini[2].y, ini[2].z, ini[2].w}, {ini[3].x, ini[3].y, ini[3].z, ini[3].w}};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = 4 * ct + c;
                const float b = bias ? bias[j] : 0.f;
                float4* op = reinterpret_cast<float4*>(out + j * RP + 4 * rt);
                float4 v = make_float4(acc[0][c] + ic[0][c] + b, acc[1][c] + ic[1][c] + b, acc[2][c] + ic[2][c] + b,
                                       acc[3][c] + ic[3][c] + b);
                if (mode == 2) { const float4 o = *op; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                *op = v;
            }
        }
    }
    __syncthreads();
}

// Row-major global tile <-> feature-major shared tile.  A warp owns whole rows (no integer division; every global
// access is a contiguous 128-byte row segment).
__device__ __forceinline__ void load_tile(const float* __restrict__ g, int64_t row0, int n_valid, int F, int64_t ld, float* s) {
    const int lane = threadIdx.x & 31;
    for (int r = threadIdx.x >> 5; r < R; r += NT / 32) {
        const float* gr = g + (row0 + r) * ld;
        for (int f = lane; f < F; f += 32) s[f * RP + r] = r < n_valid ? __ldg(gr + f) : 0.f;
    }
}
__device__ __forceinline__ void store_tile(float* __restrict__ g, int64_t row0, int n_valid, int F, int64_t ld, const float* s) {
    const int lane = threadIdx.x & 31;
    for (int r = threadIdx.x >> 5; r < n_valid; r += NT / 32) {
        float* gr = g + (row0 + r) * ld;
        for (int f = lane; f < F; f += 32) gr[f] = s[f * RP + r];
    }
}
__device__ __forceinline__ void copy_to_smem(float* dst, const float* __restrict__ src, int n) {
    for (int i = threadIdx.x * 4; i < n; i += NT * 4)
        *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
}

struct Args {
    Dims d;
    // resident weights, K-major: forward  wt_vsq_h [H][Vp], wt_ih_c [M][3H], wt_hh [H][3H], b_hh [3H]
    //                            backward w_hh [3H][H], w_ih_c [3H][M]
    const float *w0, *w1, *w2, *b_hh;
    const float* pv; const float* pg; const float* h0; const uint32_t* mask;
    float* h_out; float* sv_vsq; float* sv_alpha; float* sv_c; float* sv_gate;
    // backward
    const float* dhq; float* st_dgi; float* st_dgh; float* st_dvsq; float* d_h0;
    int64_t ld_pv, ld_pg, ld_st;       // row strides (floats) of pv / pg and of the three stash pointers
    int64_t N; int T;
};

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(NT) seq2_fwd_kernel(const Args a) {
    extern __shared__ __align__(16) float sm[];
    const Dims d = a.d;
    const int H = d.H, H3 = 3 * H, M = d.M, K = d.K, U = d.U, Vp = d.Vp();
    const bool tm = d.tarmac();
    int o = 0;
    auto take = [&](int n) { float* p = sm + o; o += (n + 3) & ~3; return p; };
    float* wVH = take(tm ? H * Vp : 0);
    float* wIC = take(tm ? M * H3 : 0);
    float* wHH = take(H * H3);
    float* bHH = take(H3);
    float* sH = take(H * RP);
    float* sVSQ = take(tm ? Vp * RP : 0);
    float* sC = take(tm ? M * RP : 0);
    float* sGI = take(H3 * RP);
    float* sGH = take(H3 * RP);
    float* sAl = take(tm ? U * RP : 0);
    float* scratch = take(NT * 16);
    if (tm) { copy_to_smem(wVH, a.w0, H * Vp); copy_to_smem(wIC, a.w1, M * H3); }
    copy_to_smem(wHH, a.w2, H * H3);
    copy_to_smem(bHH, a.b_hh, H3);

    const int rpt = d.rows_per_tile();
    const int64_t row0 = (int64_t)blockIdx.x * rpt;
    const int n_valid = (int)min((int64_t)rpt, a.N - row0);
    const int64_t n = a.N;
    const bool training = a.sv_gate != nullptr;
    const float scale = tm ? 1.0f / (float)K : 0.f;
    load_tile(a.h0, row0, n_valid, H, H, sH);
    __syncthreads();

    for (int t = 0; t < a.T; ++t) {
        const float* pg = a.pg + (size_t)t * n * a.ld_pg;
        if (tm) {
            gemm_s(wVH, Vp, sH, H, sVSQ, Vp, a.pv + (size_t)t * n * a.ld_pv, row0, n_valid, a.ld_pv, nullptr, 0, scratch);
            const uint32_t* mk = a.mask + (size_t)t * n;
            for (int p = threadIdx.x; p < R * U; p += NT) {
                const int r = p / U, i = p - r * U;
                float e = -CUDART_INF_F;
                if (r < n_valid && ((__ldg(mk + row0 + r) >> i) & 1u)) {
                    const int src = (r / U) * U + i;
                    float acc = 0.f;
                    for (int kk = 0; kk < K; ++kk)
                        acc = fmaf(sVSQ[(M + kk) * RP + src], sVSQ[(M + K + kk) * RP + r], acc);
                    e = acc * scale;
                }
                sAl[i * RP + r] = e;
            }
            __syncthreads();
            if (threadIdx.x < R) {
                const int r = threadIdx.x;
                float mx = -CUDART_INF_F;
                for (int i = 0; i < U; ++i) mx = fmaxf(mx, sAl[i * RP + r]);
                float den = 0.f;
                for (int i = 0; i < U; ++i) {
                    const float e = sAl[i * RP + r];
                    const float p = e == -CUDART_INF_F ? 0.f : expf(e - mx);
                    sAl[i * RP + r] = p;
                    den += p;
                }
                const float inv = den > 0.f ? 1.0f / den : 0.f;
                for (int i = 0; i < U; ++i) sAl[i * RP + r] *= inv;
            }
            __syncthreads();
            for (int p = threadIdx.x; p < M * R; p += NT) {
                const int m = p / R, r = p - m * R;
                const int b0 = (r / U) * U;
                float acc = 0.f;
                if (r < n_valid)
                    for (int i = 0; i < U; ++i) acc = fmaf(sAl[i * RP + r], sVSQ[m * RP + b0 + i], acc);
                sC[m * RP + r] = acc;
            }
            __syncthreads();
            gemm_s(wIC, H3, sC, M, sGI, H3, pg, row0, n_valid, a.ld_pg, nullptr, 0, scratch);
        } else {
            load_tile(pg, row0, n_valid, H3, a.ld_pg, sGI);
        }
        gemm_s(wHH, H3, sH, H, sGH, H3, nullptr, 0, 0, 0, bHH, 0, scratch);
        if (training) {
            if (tm) {
                store_tile(a.sv_vsq + (size_t)t * n * Vp, row0, n_valid, Vp, Vp, sVSQ);
                store_tile(a.sv_alpha + (size_t)t * n * U, row0, n_valid, U, U, sAl);
                store_tile(a.sv_c + (size_t)t * n * M, row0, n_valid, M, M, sC);
            }
        }
        // gates, row-major thread mapping so that the global stores are coalesced
        float* hout = a.h_out + (size_t)t * n * H;
        float* gt = training ? a.sv_gate + (size_t)t * n * 4 * H : nullptr;
        for (int r = threadIdx.x >> 5; r < R; r += NT / 32)
        for (int ch = threadIdx.x & 31; ch < H; ch += 32) {
            const float rr = sigmoidf_(sGI[ch * RP + r] + sGH[ch * RP + r]);
            const float zz = sigmoidf_(sGI[(H + ch) * RP + r] + sGH[(H + ch) * RP + r]);
            const float ghn = sGH[(2 * H + ch) * RP + r];
            const float nn = tanhf(fmaf(rr, ghn, sGI[(2 * H + ch) * RP + r]));
            const float hp = sH[ch * RP + r];
            const float hn = fmaf(zz, hp - nn, nn);
            if (r < n_valid) {
                hout[(row0 + r) * H + ch] = hn;
                if (training) {
                    float* g = gt + (row0 + r) * 4 * H;
                    g[ch] = rr; g[H + ch] = zz; g[2 * H + ch] = nn; g[3 * H + ch] = ghn;
                }
            }
            sGI[ch * RP + r] = hn;                 // stage h' (sH is still being read by other threads)
        }
        __syncthreads();
        for (int p = threadIdx.x; p < H * RP; p += NT) sH[p] = sGI[p];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ backward
// Per step t (T-1 .. 0): dh = carry + dhq_t ; gates' -> dgi, dgh ; carry = dh*z + dgh W_hh ; dc = dgi W_ih[:, H:] ;
// attention' (dc, alpha, vsq) -> dvsq.  dgi / dgh / dvsq are stashed for the batched GEMMs that follow the kernel.
__global__ void __launch_bounds__(NT) seq2_bwd_kernel(const Args a) {
    extern __shared__ __align__(16) float sm[];
    const Dims d = a.d;
    const int H = d.H, H3 = 3 * H, M = d.M, K = d.K, U = d.U, Vp = d.Vp();
    const bool tm = d.tarmac();
    int o = 0;
    auto take = [&](int n) { float* p = sm + o; o += (n + 3) & ~3; return p; };
    float* wHH = take(H3 * H);                 // [3H][H]  K-major for dgh W_hh
    float* wIC = take(tm ? H3 * M : 0);        // [3H][M]  K-major for dgi W_ih[:, H:]
    float* sDH = take(H * RP);
    float* sDGI = take(H3 * RP);
    float* sDGH = take(H3 * RP);
    float* sDC = take(tm ? M * RP : 0);
    float* sVSQ = take(tm ? Vp * RP : 0);
    float* sDVSQ = take(tm ? Vp * RP : 0);
    float* sAl = take(tm ? U * RP : 0);
    float* sDS = take(tm ? U * RP : 0);
    float* scratch = take(NT * 16);
    copy_to_smem(wHH, a.w0, H3 * H);
    if (tm) copy_to_smem(wIC, a.w1, H3 * M);

    const int rpt = d.rows_per_tile();
    const int64_t row0 = (int64_t)blockIdx.x * rpt;
    const int n_valid = (int)min((int64_t)rpt, a.N - row0);
    const int64_t n = a.N;
    const float scale = tm ? 1.0f / (float)K : 0.f;
    for (int p = threadIdx.x; p < H * RP; p += NT) sDH[p] = 0.f;
    __syncthreads();

    for (int t = a.T - 1; t >= 0; --t) {
        const float* gt = a.sv_gate + (size_t)t * n * 4 * H;
        const float* hprev = t > 0 ? a.h_out + (size_t)(t - 1) * n * H : a.h0;
        const float* dhq = a.dhq + (size_t)t * n * H;
        float* gdgi = a.st_dgi + (size_t)t * n * a.ld_st;
        float* gdgh = a.st_dgh + (size_t)t * n * a.ld_st;
        for (int r = threadIdx.x >> 5; r < R; r += NT / 32)
        for (int ch = threadIdx.x & 31; ch < H; ch += 32) {
            float dr = 0.f, dz = 0.f, dn = 0.f, dnr = 0.f, dir = 0.f;
            if (r < n_valid) {
                const float* g = gt + (row0 + r) * 4 * H;
                const float rr = __ldg(g + ch), zz = __ldg(g + H + ch), nn = __ldg(g + 2 * H + ch), ghn = __ldg(g + H3 + ch);
                const float gv = sDH[ch * RP + r] + __ldg(dhq + (row0 + r) * H + ch);
                dn = gv * (1.0f - zz) * (1.0f - nn * nn);
                dz = gv * (__ldg(hprev + (row0 + r) * H + ch) - nn) * zz * (1.0f - zz);
                dr = dn * ghn * rr * (1.0f - rr);
                dnr = dn * rr;
                dir = gv * zz;
                float* o1 = gdgi + (row0 + r) * a.ld_st;
                float* o2 = gdgh + (row0 + r) * a.ld_st;
                o1[ch] = dr; o1[H + ch] = dz; o1[2 * H + ch] = dn;
                o2[ch] = dr; o2[H + ch] = dz; o2[2 * H + ch] = dnr;
            }
            sDGI[ch * RP + r] = dr; sDGI[(H + ch) * RP + r] = dz; sDGI[(2 * H + ch) * RP + r] = dn;
            sDGH[ch * RP + r] = dr; sDGH[(H + ch) * RP + r] = dz; sDGH[(2 * H + ch) * RP + r] = dnr;
            sDH[ch * RP + r] = dir;
        }
        __syncthreads();
        gemm_s(wHH, H, sDGH, H3, sDH, H, nullptr, 0, 0, 0, nullptr, 2, scratch);          // carry = dh*z + dgh W_hh
        if (tm) {
            gemm_s(wIC, M, sDGI, H3, sDC, M, nullptr, 0, 0, 0, nullptr, 0, scratch);      // dc = dgi W_ih[:, H:]
            load_tile(a.sv_vsq + (size_t)t * n * Vp, row0, n_valid, Vp, Vp, sVSQ);
            load_tile(a.sv_alpha + (size_t)t * n * U, row0, n_valid, U, U, sAl);
            __syncthreads();
            for (int p = threadIdx.x; p < R * U; p += NT) {
                const int r = p / U, i = p - r * U;
                const int src = (r / U) * U + i;
                float acc = 0.f;
                if (r < n_valid)
                    for (int m = 0; m < M; ++m) acc = fmaf(sDC[m * RP + r], sVSQ[m * RP + src], acc);
                sDS[i * RP + r] = acc;
            }
            __syncthreads();
            if (threadIdx.x < R) {
                const int r = threadIdx.x;
                float tot = 0.f;
                for (int i = 0; i < U; ++i) tot = fmaf(sAl[i * RP + r], sDS[i * RP + r], tot);
                for (int i = 0; i < U; ++i) sDS[i * RP + r] = sAl[i * RP + r] * (sDS[i * RP + r] - tot);
            }
            __syncthreads();
            for (int p = threadIdx.x; p < Vp * R; p += NT) {
                const int f = p / R, r = p - f * R;
                const int b0 = (r / U) * U, li = r - b0;
                float acc = 0.f;
                if (r >= n_valid) {
                } else if (f < M) {
                    for (int j = 0; j < U; ++j) acc = fmaf(sAl[li * RP + b0 + j], sDC[f * RP + b0 + j], acc);
                } else if (f < M + K) {
                    for (int j = 0; j < U; ++j) acc = fmaf(sDS[li * RP + b0 + j], sVSQ[(f + K) * RP + b0 + j], acc);
                    acc *= scale;
                } else if (f < M + 2 * K) {
                    for (int i = 0; i < U; ++i) acc = fmaf(sDS[i * RP + r], sVSQ[(f - K) * RP + b0 + i], acc);
                    acc *= scale;
                }
                sDVSQ[f * RP + r] = acc;
            }
            __syncthreads();
            store_tile(a.st_dvsq + (size_t)t * n * a.ld_st, row0, n_valid, Vp, a.ld_st, sDVSQ);
        }
        __syncthreads();
    }
    if (a.d_h0 != nullptr) store_tile(a.d_h0, row0, n_valid, H, H, sDH);
}

// ================================================================================================ tensor-core path
// The same window kernels with every dense product on the tensor cores.  A CTA still owns 16 agent rows — exactly the M
// of one mma.sync.m16n8k8 tile — so the per-step products h W_hh^T (16 x H x 3H), h W_vsq_h^T (16 x H x Vp) and
// c W_ih_c^T (16 x M x 3H) become 8-column warp tiles whose accumulators stay in registers; fp32 accuracy comes from
// the 3xTF32 split (a = a_hi + a_lo, b = b_hi + b_lo; a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32 accumulate: ~2^-21),
// done on the fly in registers because hi / lo copies of the resident weights (2 x 123 KB) would not fit one SM.
//   * warp w < H/8 ("gate warp") owns the 8 channels [8w, 8w+8) of all three gates: its gh and gi accumulators and the
//     hidden state of its (row, channel) elements never leave registers — the GRU gates run on the accumulator
//     fragments directly, h' goes to shared memory only as the next step's A operand;
//   * warps 8..15 ("comm warps") produce vsq = pv + h W_vsq_h^T, run the block attention and hand c to the gate warps,
//     concurrently with the gate warps' gh product (both only need h);
//   * weights sit K-major with the row stride padded to 8 (mod 32) floats and activations row-major with stride
//     4 (mod 32): every fragment load is bank-conflict free.
// tcgen05 is the wrong tool here: its smallest tile (M = 64) would need the transposed product with both operands
// pre-split in shared memory, and every dependent layer would pay a TMEM -> register -> shared-memory round trip; the
// chain of three dependent 16-row products per timestep is latency-, not throughput-bound.
namespace mma {
#ifdef UBS_SEQ2_TRACE
__device__ long long ubs_trace[8 * 16 * 8];
#endif

constexpr int NW = 16, NTM = NW * 32, NGATE = 8;

__device__ __forceinline__ void split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;                       // explicit truncation: the MMA must see exactly hi
    lo = __float_as_uint(v - __uint_as_float(hi)) & 0xffffe000u;
}
__device__ __forceinline__ void mma8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__host__ __device__ inline int pad8(int n) { return n + ((8 - n % 32) + 32) % 32; }     // == 8 (mod 32)
__host__ __device__ inline int pad4(int n) { return n + ((4 - n % 32) + 32) % 32; }     // == 4 (mod 32)

// acc[j] (+)= A[16 x KD] . W[KD x 8] at columns n0[j] for NT column tiles of one warp.  A comes PRE-SPLIT (two
// row-major shared arrays hi / lo, stride LDA: the producer splits each activation once instead of every consumer warp
// every k-step), W is fp32 K-major (stride LDW) and split on the fly.  main / corr keep the hi.hi and the cross terms
// apart.  All strides / trip counts are compile-time: every fragment load is [base + immediate].
template <int NT, int KD, int LDA, int LDW>
__device__ __forceinline__ void warp_gemm(const float* Ahi, const float* Alo, const float* W, const int (&n0)[NT],
                                          int ntiles, float (&main_)[NT][4], float (&corr)[NT][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const uint32_t* ah_p = reinterpret_cast<const uint32_t*>(Ahi) + g * LDA + c;
    const uint32_t* al_p = reinterpret_cast<const uint32_t*>(Alo) + g * LDA + c;
    const float* wp = W + c * LDW + g;
    // HMMA latency is several issue slots: no two MMAs of a k-step may depend on each other.  The second cross term has
    // its own accumulator, and a warp with a single tile alternates accumulator sets between even and odd k-steps.
    constexpr int SETS = NT == 1 ? 2 : 1;
    float c2[SETS][NT][4], m2[NT][4], c1b[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            m2[j][q] = 0.f; c1b[j][q] = 0.f;
#pragma unroll
            for (int s_ = 0; s_ < SETS; ++s_) c2[s_][j][q] = 0.f;
        }
#pragma unroll
    for (int k0 = 0; k0 < KD; k0 += 8) {
        const uint32_t ah[4] = {ah_p[k0], ah_p[k0 + 8 * LDA], ah_p[k0 + 4], ah_p[k0 + 8 * LDA + 4]};
        const uint32_t al[4] = {al_p[k0], al_p[k0 + 8 * LDA], al_p[k0 + 4], al_p[k0 + 8 * LDA + 4]};
        const bool odd = SETS == 2 && ((k0 >> 3) & 1);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            if (j < ntiles) {
                uint32_t bh0, bl0, bh1, bl1;
                split(wp[k0 * LDW + n0[j]], bh0, bl0);
                split(wp[(k0 + 4) * LDW + n0[j]], bh1, bl1);
                if (odd) {
                    mma8(c1b[j], al, bh0, bh1);
                    mma8(c2[SETS - 1][j], ah, bl0, bl1);
                    mma8(m2[j], ah, bh0, bh1);
                } else {
                    mma8(corr[j], al, bh0, bh1);
                    mma8(c2[0][j], ah, bl0, bl1);
                    mma8(main_[j], ah, bh0, bh1);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float cc = corr[j][q] + c2[0][j][q];
            if (SETS == 2) { cc += c1b[j][q] + c2[SETS - 1][j][q]; main_[j][q] += m2[j][q]; }
            corr[j][q] = cc;
        }
}

__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// gates: 1 / (1 + 2^(-x log2 e)) with the hardware exp2 / reciprocal (error ~1e-7, far inside the parity budget);
// tanh(x) = 2 sigmoid(2x) - 1
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// branch-free (MUFU.EX2 + MUFU.RCP, no slow paths: twelve of them interleave freely); saturates correctly for |x| large
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.0f, fast_sigmoid(2.0f * x), -1.0f); }

template <int H, int M, int VP>
struct Lay {                                      // shared-memory layout (float offsets), all compile-time
    static constexpr bool TM = M > 0;
    static constexpr int H3 = 3 * H;
    static constexpr int ldV = VP + ((8 - VP % 32) + 32) % 32, ld3 = H3 + ((8 - H3 % 32) + 32) % 32;
    static constexpr int ldh = H + ((4 - H % 32) + 32) % 32, ldc = M + ((4 - M % 32) + 32) % 32;
    static constexpr int ldv = VP + ((4 - VP % 32) + 32) % 32;
    static constexpr int wVH = 0, wIC = wVH + (TM ? H * ldV : 0), wHH = wIC + (TM ? M * ld3 : 0), bHH = wHH + H * ld3;
    static constexpr int sHh = bHH + H3, sHl = sHh + R * ldh, sCh = sHl + R * ldh, sCl = sCh + (TM ? R * ldc : 0);
    static constexpr int sVSQ = sCl + (TM ? R * ldc : 0), sAl = sVSQ + (TM ? R * ldv : 0), sMk = sAl + (TM ? R * 16 : 0);
    static constexpr int total = sMk + 2 * R;
};

__device__ __forceinline__ void copy_padded(float* dst, int ld_dst, const float* __restrict__ src, int rows, int cols) {
    const int c4 = cols >> 2;
    for (int i = threadIdx.x; i < rows * c4; i += NTM) {
        const int r = i / c4, q = i - r * c4;
        *reinterpret_cast<float4*>(dst + r * ld_dst + 4 * q) = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * cols + 4 * q));
    }
}

__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Step schedule (the dependent chain is  h -> vsq -> attention -> c -> gi -> gates -> h'; gh only needs h):
//   S1  warps 0 .. Vp/8-1 : one vsq column tile each (pv + h W_vsq_h^T) -> shared memory
//   S2  comm warps        : attention -> c            ||   gate warps : gh = h W_hh^T + b_hh   (off the critical path)
//   S3  gate warps        : gi = pg + c W_ih_c^T, GRU gates on the fragments, h' -> shared (pre-split) / global
//                                                     ||   comm warps : prefetch pv / masks of step t + 1
// Producer -> consumer hand-offs use bar.arrive / bar.sync pairs (ids 2, 3) so that producers never wait.
template <int H, int M, int VP>
__global__ void __launch_bounds__(NTM, 1) seq2_fwd_mma_kernel(const Args a) {
    extern __shared__ __align__(16) float sm[];
    using L = Lay<H, M, VP>;
    constexpr bool tm = L::TM;
    constexpr int H3 = 3 * H, Vp = VP;
    const int K = a.d.K, U = a.d.U;
    float* wVH = sm + L::wVH; float* wIC = sm + L::wIC; float* wHH = sm + L::wHH; float* bHH = sm + L::bHH;
    float* sHh = sm + L::sHh; float* sHl = sm + L::sHl; float* sCh = sm + L::sCh; float* sCl = sm + L::sCl;
    float* sVSQ = sm + L::sVSQ; float* sAl = sm + L::sAl;
    uint32_t* sMk = reinterpret_cast<uint32_t*>(sm + L::sMk);              // [2][R] talk masks, double buffered
    if (tm) { copy_padded(wVH, L::ldV, a.w0, H, Vp); copy_padded(wIC, L::ld3, a.w1, M, H3); }
    copy_padded(wHH, L::ld3, a.w2, H, H3);
    for (int i = threadIdx.x; i < H3; i += NTM) bHH[i] = __ldg(a.b_hh + i);

    const int rpt = a.d.rows_per_tile();
    const int64_t row0 = (int64_t)blockIdx.x * rpt;
    const int n_valid = (int)min((int64_t)rpt, a.N - row0);
    const int64_t n = a.N;
    const bool training = a.sv_gate != nullptr;
    const float scale = tm ? 1.0f / (float)K : 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const bool gate_warp = warp < NGATE && warp < H / 8;
    const bool comm_warp = warp >= NGATE;
    const int ct = threadIdx.x - NGATE * 32;                               // comm thread index
    constexpr int NCT = (NW - NGATE) * 32;
    const int r0 = g, r1 = g + 8;                                          // fragment rows of this lane
    const bool v0 = r0 < n_valid, v1 = r1 < n_valid;
    const int chb = 8 * warp + 2 * c;                                      // gate warps: first of this lane's 2 channels
    constexpr int vtiles = tm ? Vp / 8 : 0;
    static_assert(vtiles <= NW, "one vsq column tile per warp");
    const bool vsq_warp = warp < vtiles;

    for (int i = threadIdx.x; i < R * L::ldh; i += NTM) {
        const int r = i / L::ldh, f = i - r * L::ldh;
        const float v = (r < n_valid && f < H) ? __ldg(a.h0 + (row0 + r) * H + f) : 0.f;
        uint32_t hi, lo;
        split(v, hi, lo);
        sHh[i] = __uint_as_float(hi); sHl[i] = __uint_as_float(lo);
    }
    float hreg[4] = {0.f, 0.f, 0.f, 0.f};                                  // h[r0][chb], h[r0][chb+1], h[r1][chb], h[r1][chb+1]
    if (gate_warp) {
        if (v0) { hreg[0] = __ldg(a.h0 + (row0 + r0) * H + chb); hreg[1] = __ldg(a.h0 + (row0 + r0) * H + chb + 1); }
        if (v1) { hreg[2] = __ldg(a.h0 + (row0 + r1) * H + chb); hreg[3] = __ldg(a.h0 + (row0 + r1) * H + chb + 1); }
    }
    float pvr[1][4] = {{0.f, 0.f, 0.f, 0.f}};                              // vsq warps: pv fragment of the coming step
    float pgr[3][4];                                                       // gate warps: pg fragments of the coming step
    // per-thread global addresses of the fragments this lane reads / writes every step (advanced by a stride per step)
    const size_t o0 = (size_t)(row0 + r0), o1 = (size_t)(row0 + r1);
    const float* pv0 = tm ? a.pv + o0 * a.ld_pv + 8 * warp + 2 * c : nullptr;
    const float* pv1 = tm ? a.pv + o1 * a.ld_pv + 8 * warp + 2 * c : nullptr;
    const size_t st_pv = (size_t)n * a.ld_pv, st_pg = (size_t)n * a.ld_pg;
    const float* pg0 = a.pg + o0 * a.ld_pg + chb;
    const float* pg1 = a.pg + o1 * a.ld_pg + chb;
    auto load_pv = [&](int t) {
        if (vsq_warp) {
            pvr[0][0] = pvr[0][1] = pvr[0][2] = pvr[0][3] = 0.f;
            if (v0) { const float2 t2 = __ldg(reinterpret_cast<const float2*>(pv0 + t * st_pv)); pvr[0][0] = t2.x; pvr[0][1] = t2.y; }
            if (v1) { const float2 t2 = __ldg(reinterpret_cast<const float2*>(pv1 + t * st_pv)); pvr[0][2] = t2.x; pvr[0][3] = t2.y; }
        }
    };
    auto load_pg = [&](int t) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            pgr[j][0] = pgr[j][1] = pgr[j][2] = pgr[j][3] = 0.f;
            if (v0) { const float2 t2 = __ldg(reinterpret_cast<const float2*>(pg0 + t * st_pg + j * H)); pgr[j][0] = t2.x; pgr[j][1] = t2.y; }
            if (v1) { const float2 t2 = __ldg(reinterpret_cast<const float2*>(pg1 + t * st_pg + j * H)); pgr[j][2] = t2.x; pgr[j][3] = t2.y; }
        }
    };
    auto load_mask = [&](int t) {
        if (tm && ct >= 0 && ct < R) sMk[(t & 1) * R + ct] = ct < n_valid ? __ldg(a.mask + (size_t)t * n + row0 + ct) : 0u;
    };
    if (tm) load_pv(0);
    if (gate_warp) load_pg(0);
    load_mask(0);
    __syncthreads();

#ifdef UBS_SEQ2_TRACE
#define TR(slot) do { if (blockIdx.x == 0 && lane == 0 && t < 8) ubs_trace[(t * 16 + warp) * 8 + (slot)] = clock64(); } while (0)
#else
#define TR(slot) do { } while (0)
#endif
    for (int t = 0; t < a.T; ++t) {
        TR(0);
        // ---- S1: vsq = pv + h W_vsq_h^T, one column tile per warp
        if (tm && vsq_warp) {
            float corr[1][4] = {{0.f, 0.f, 0.f, 0.f}};
            const int vn0[1] = {8 * warp};
            warp_gemm<1, H, L::ldh, L::ldV>(sHh, sHl, wVH, vn0, 1, pvr, corr);
            const int col = vn0[0] + 2 * c;
            const float2 lo2 = make_float2(pvr[0][0] + corr[0][0], pvr[0][1] + corr[0][1]);
            const float2 hi2 = make_float2(pvr[0][2] + corr[0][2], pvr[0][3] + corr[0][3]);
            *reinterpret_cast<float2*>(sVSQ + r0 * L::ldv + col) = lo2;
            *reinterpret_cast<float2*>(sVSQ + r1 * L::ldv + col) = hi2;
            if (training) {
                float* sv = a.sv_vsq + (size_t)t * n * Vp + col;
                if (v0) *reinterpret_cast<float2*>(sv + o0 * Vp) = lo2;
                if (v1) *reinterpret_cast<float2*>(sv + o1 * Vp) = hi2;
            }
        }
        float gh[3][4], gi[3][4];
        TR(1);
        if (!comm_warp) {
            if (tm) bar_arrive(2, NTM);                     // vsq tile handed to the comm warps; go on with gh
            if (gate_warp) {
                // ---- S2 (gate warps): gh = h W_hh^T + b_hh
                float corr[3][4];
                const int n0[3] = {8 * warp, H + 8 * warp, 2 * H + 8 * warp};
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float b0 = bHH[n0[j] + 2 * c], b1 = bHH[n0[j] + 2 * c + 1];
                    gh[j][0] = b0; gh[j][1] = b1; gh[j][2] = b0; gh[j][3] = b1;
                    corr[j][0] = corr[j][1] = corr[j][2] = corr[j][3] = 0.f;
                }
                warp_gemm<3, H, L::ldh, L::ld3>(sHh, sHl, wHH, n0, 3, gh, corr);
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) { gh[j][q] += corr[j][q]; gi[j][q] = pgr[j][q]; }
            }
            TR(2);
            if (tm) bar_sync(3, NTM);                       // c is ready
            TR(3);
        } else if (tm) {
            bar_sync(2, NTM);                               // every vsq tile is in shared memory
            TR(2);
            // ---- S2 (comm warps): block attention (TarMAC.forward: u_dot_v / key_size, edge_softmax, u_mul_e + sum).
            // 16 lanes per row, lane i < U scores source i; the softmax runs on shuffles inside the 16-lane group.
            {
                const int r = ct >> 4, i = ct & 15;
                float e = -CUDART_INF_F;
                if (r < n_valid && i < U && ((sMk[(t & 1) * R + r] >> i) & 1u)) {
                    const float* sp = sVSQ + ((r / U) * U + i) * L::ldv + M;      // signature of the source
                    const float* qp = sVSQ + r * L::ldv + M + K;                  // query of the destination
                    float acc = 0.f;
                    int kk = 0;
                    if ((K & 3) == 0) {
                        for (; kk < K; kk += 4) {
                            const float4 s4 = *reinterpret_cast<const float4*>(sp + kk);
                            const float4 q4 = *reinterpret_cast<const float4*>(qp + kk);
                            acc = fmaf(s4.x, q4.x, acc); acc = fmaf(s4.y, q4.y, acc);
                            acc = fmaf(s4.z, q4.z, acc); acc = fmaf(s4.w, q4.w, acc);
                        }
                    }
                    for (; kk < K; ++kk) acc = fmaf(sp[kk], qp[kk], acc);
                    e = acc * scale;
                }
                float mx = e;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o, 16));
                const float pr = e == -CUDART_INF_F ? 0.f : __expf(e - mx);
                float den = pr;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o, 16);
                const float al = den > 0.f ? __fdividef(pr, den) : 0.f;
                if (i < U) {
                    sAl[r * 16 + i] = al;
                    if (training && r < n_valid) a.sv_alpha[((size_t)t * n + row0 + r) * U + i] = al;
                }
            }
            TR(3);
            bar_sync(1, NCT);
            TR(4);
            // c[r][m..m+3] = sum_i alpha[r][i] v[b0 + i][m..m+3]
            for (int p = ct; p < R * (M / 4); p += NCT) {
                const int r = p / (M / 4 > 0 ? M / 4 : 1), m = 4 * (p - r * (M / 4));
                const int b0 = (r / U) * U;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < n_valid) {
#pragma unroll 4
                    for (int i = 0; i < U; ++i) {
                        const float al = sAl[r * 16 + i];
                        const float4 v4 = *reinterpret_cast<const float4*>(sVSQ + (b0 + i) * L::ldv + m);
                        acc.x = fmaf(al, v4.x, acc.x); acc.y = fmaf(al, v4.y, acc.y);
                        acc.z = fmaf(al, v4.z, acc.z); acc.w = fmaf(al, v4.w, acc.w);
                    }
                }
                uint32_t hi[4], lo[4];
                split(acc.x, hi[0], lo[0]); split(acc.y, hi[1], lo[1]); split(acc.z, hi[2], lo[2]); split(acc.w, hi[3], lo[3]);
                *reinterpret_cast<uint4*>(sCh + r * L::ldc + m) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(sCl + r * L::ldc + m) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                if (training && r < n_valid) *reinterpret_cast<float4*>(a.sv_c + ((size_t)t * n + row0 + r) * M + m) = acc;
            }
            TR(5);
            bar_arrive(3, NTM);                             // c handed to the gate warps
            load_mask(t + 1 < a.T ? t + 1 : t);             // S3 shadow: next step's masks
        }
        if (tm && t + 1 < a.T) load_pv(t + 1);              // in flight during S3 (vsq warps)
        if (gate_warp) {
            // ---- S3: gi = pg + c W_ih_c^T, then the GRU gates on the accumulator fragments
            if (tm) {
                float corr[3][4];
                const int n0[3] = {8 * warp, H + 8 * warp, 2 * H + 8 * warp};
#pragma unroll
                for (int j = 0; j < 3; ++j) corr[j][0] = corr[j][1] = corr[j][2] = corr[j][3] = 0.f;
                warp_gemm<3, (M > 0 ? M : 8), L::ldc, L::ld3>(sCh, sCl, wIC, n0, 3, gi, corr);
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) gi[j][q] += corr[j][q];
            }
            TR(4);
            if (t + 1 < a.T) load_pg(t + 1);                // next step's initialisers fly during the gates + barrier + S1/S2
            float hn[4], rr[4], zz[4], nn[4];
            uint32_t hh[4], hl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { rr[q] = fast_sigmoid(gi[0][q] + gh[0][q]); zz[q] = fast_sigmoid(gi[1][q] + gh[1][q]); }
#pragma unroll
            for (int q = 0; q < 4; ++q) nn[q] = fast_tanh(fmaf(rr[q], gh[2][q], gi[2][q]));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                hn[q] = fmaf(zz[q], hreg[q] - nn[q], nn[q]);
                hreg[q] = hn[q];
                split(hn[q], hh[q], hl[q]);
            }
            TR(5);
            *reinterpret_cast<uint2*>(sHh + r0 * L::ldh + chb) = make_uint2(hh[0], hh[1]);
            *reinterpret_cast<uint2*>(sHh + r1 * L::ldh + chb) = make_uint2(hh[2], hh[3]);
            *reinterpret_cast<uint2*>(sHl + r0 * L::ldh + chb) = make_uint2(hl[0], hl[1]);
            *reinterpret_cast<uint2*>(sHl + r1 * L::ldh + chb) = make_uint2(hl[2], hl[3]);
            float* hout = a.h_out + (size_t)t * n * H + chb;
            if (v0) *reinterpret_cast<float2*>(hout + o0 * H) = make_float2(hn[0], hn[1]);
            if (v1) *reinterpret_cast<float2*>(hout + o1 * H) = make_float2(hn[2], hn[3]);
            if (training) {
                float* gt = a.sv_gate + (size_t)t * n * 4 * H + chb;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 0 ? v0 : v1) {
                        float* gp = gt + (half == 0 ? o0 : o1) * 4 * H;
                        const int q = 2 * half;
                        *reinterpret_cast<float2*>(gp) = make_float2(rr[q], rr[q + 1]);
                        *reinterpret_cast<float2*>(gp + H) = make_float2(zz[q], zz[q + 1]);
                        *reinterpret_cast<float2*>(gp + 2 * H) = make_float2(nn[q], nn[q + 1]);
                        *reinterpret_cast<float2*>(gp + 3 * H) = make_float2(gh[2][q], gh[2][q + 1]);
                    }
                }
            }
        }
        TR(6);
        __syncthreads();                                    // h' (and the next masks) are in shared memory
        TR(7);
    }
}

// configurations with a compiled instance: (H, M, Vp)
template <int H, int M, int VP>
static int launch_fwd_mma(const Args& a, cudaStream_t st) {
    static const cudaError_t rc_attr = cudaFuncSetAttribute(seq2_fwd_mma_kernel<H, M, VP>,
                                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (rc_attr != cudaSuccess) { set_error("ubs_agent_seq2_fwd: %s", cudaGetErrorString(rc_attr)); return 1; }
    const int rpt = a.d.rows_per_tile();
    const size_t smem = (size_t)Lay<H, M, VP>::total * sizeof(float);
    seq2_fwd_mma_kernel<H, M, VP><<<(unsigned)((a.N + rpt - 1) / rpt), NTM, smem, st>>>(a);
    return check_launch("ubs_agent_seq2_fwd(mma)");
}
// ------------------------------------------------------------------------------------------------ backward (mma)
// Reverse-time walk with both products of a step on the tensor cores:
//     carry' = dh z + dgh W_hh      (16 x 3H x H)        warps 0 .. H/8-1, one 8-column tile each, 3H/8 k-steps
//     dc     = dgi W_ih[:, H:]      (16 x 3H x M)        warps 8 .. 8+M/8-1
// The warps of the first product own dh: the carry accumulators are the fragments the gate derivatives are computed
// on, so dh never leaves registers.  dgi / dgh go to shared memory once (fp32, row-major) where they are the A operand
// of both products AND the staging area from which all threads write the stash rows to global memory with coalesced
// 16-byte stores.  Weights are kept in FRAGMENT ORDER ([k-step][tile][lane][2]): one conflict-free LDS.64 per MMA
// B operand.  The attention backward runs one warp per destination row (lanes = source x quarter of the message).
template <int H, int M, int VP>
struct LayB {
    static constexpr bool TM = M > 0;
    static constexpr int H3 = 3 * H, KS = H3 / 8;                     // k-steps of both products
    static constexpr int ldg = H3 + ((4 - H3 % 32) + 32) % 32;        // dgi / dgh rows: == 4 (mod 32)
    static constexpr int ldc = M + 4, ldv = VP + 4;
    static constexpr int wHH = 0, wIC = wHH + KS * (H / 8) * 64, sDGI = wIC + (TM ? KS * (M / 8) * 64 : 0);
    static constexpr int sDGH = sDGI + R * ldg, sDC = sDGH + R * ldg, sVSQ = sDC + (TM ? R * ldc : 0);
    static constexpr int sAl = sVSQ + (TM ? R * ldv : 0), sDS = sAl + (TM ? R * 16 : 0), total = sDS + (TM ? R * 16 : 0);
};

// W (K x N, K-major, global) -> fragment order in shared memory: dst[((ks * NT + j) * 32 + lane) * 2 + {0,1}] =
// W[8 ks + c (+4)][8 j + g]  with g = lane / 4, c = lane % 4.
__device__ __forceinline__ void load_frag_order(float* dst, const float* __restrict__ W, int Kd, int N) {
    const int NT = N / 8, total = (Kd / 8) * NT * 64;
    for (int i = threadIdx.x; i < total; i += NTM) {
        const int e = i & 1, lane = (i >> 1) & 31, j = (i >> 6) % NT, ks = (i >> 6) / NT;
        dst[i] = __ldg(W + (size_t)(8 * ks + (lane & 3) + 4 * e) * N + 8 * j + (lane >> 2));
    }
}

// one 8-column tile: acc (+)= A[16 x 8 KS] (shared fp32 row-major, split on the fly) . Wf (fragment order, tile j of NT)
template <int KS, int LDA, int NT>
__device__ __forceinline__ void tile_gemm_f(const float* A, const float* Wf, int j, float (&acc)[4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const float* a0p = A + g * LDA + c;
    const float2* wp = reinterpret_cast<const float2*>(Wf) + j * 32 + lane;
    float m[2][4], c1[2][4], c2[2][4];
#pragma unroll
    for (int s_ = 0; s_ < 2; ++s_)
#pragma unroll
        for (int q = 0; q < 4; ++q) { m[s_][q] = 0.f; c1[s_][q] = 0.f; c2[s_][q] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
        split(a0p[8 * ks], ah[0], al[0]);
        split(a0p[8 * ks + 8 * LDA], ah[1], al[1]);
        split(a0p[8 * ks + 4], ah[2], al[2]);
        split(a0p[8 * ks + 8 * LDA + 4], ah[3], al[3]);
        const float2 w = wp[ks * NT * 32];
        split(w.x, bh0, bl0);
        split(w.y, bh1, bl1);
        mma8(c1[ks & 1], al, bh0, bh1);
        mma8(c2[ks & 1], ah, bl0, bl1);
        mma8(m[ks & 1], ah, bh0, bh1);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] += (m[0][q] + m[1][q]) + ((c1[0][q] + c1[1][q]) + (c2[0][q] + c2[1][q]));
}

template <int H, int M, int VP>
__global__ void __launch_bounds__(NTM, 1) seq2_bwd_mma_kernel(const Args a) {
    extern __shared__ __align__(16) float sm[];
    using L = LayB<H, M, VP>;
    constexpr bool tm = L::TM;
    constexpr int H3 = 3 * H, Vp = VP, Md = M > 0 ? M : 8, Vq = VP > 0 ? VP / 4 : 1;
    const int K = a.d.K, U = a.d.U;
    float* wHH = sm + L::wHH; float* wIC = sm + L::wIC; float* sDGI = sm + L::sDGI; float* sDGH = sm + L::sDGH;
    float* sDC = sm + L::sDC; float* sVSQ = sm + L::sVSQ; float* sAl = sm + L::sAl; float* sDS = sm + L::sDS;
    load_frag_order(wHH, a.w0, H3, H);
    if (tm) load_frag_order(wIC, a.w1, H3, M);

    const int rpt = a.d.rows_per_tile();
    const int64_t row0 = (int64_t)blockIdx.x * rpt;
    const int n_valid = (int)min((int64_t)rpt, a.N - row0);
    const int64_t n = a.N;
    const float scale = tm ? 1.0f / (float)K : 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const bool own = warp < H / 8;                                   // carry tile + dh owner
    const bool dcw = tm && warp >= 8 && warp - 8 < Md / 8;           // dc tile
    const int r0 = g, r1 = g + 8;
    const bool v0 = r0 < n_valid, v1 = r1 < n_valid;
    const int chb = 8 * warp + 2 * c;
    const size_t o0 = (size_t)(row0 + r0), o1 = (size_t)(row0 + r1);

    float carry[4] = {0.f, 0.f, 0.f, 0.f};                          // dh flowing into step t: (r0,chb) (r0,chb+1) (r1,chb) (r1,chb+1)
    // prefetched per-step inputs of the owners: gates r z n ghn, previous hidden state, dhq
    float2 pr_[2][4], ph[2], pq[2];
    auto prefetch = [&](int t) {
        const float* gt = a.sv_gate + (size_t)t * n * 4 * H + chb;
        const float* hp = (t > 0 ? a.h_out + (size_t)(t - 1) * n * H : a.h0) + chb;
        const float* dq = a.dhq + (size_t)t * n * H + chb;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const bool v = hf == 0 ? v0 : v1;
            const size_t o = hf == 0 ? o0 : o1;
#pragma unroll
            for (int q = 0; q < 4; ++q) pr_[hf][q] = v ? __ldg(reinterpret_cast<const float2*>(gt + o * 4 * H + q * H)) : make_float2(0.f, 0.f);
            ph[hf] = v ? __ldg(reinterpret_cast<const float2*>(hp + o * H)) : make_float2(0.f, 0.f);
            pq[hf] = v ? __ldg(reinterpret_cast<const float2*>(dq + o * H)) : make_float2(0.f, 0.f);
        }
    };
    if (own) prefetch(a.T - 1);
    __syncthreads();

    for (int t = a.T - 1; t >= 0; --t) {
        // ---- E1: gate derivatives on the carry fragments -> dgi / dgh (shared, fp32) ; carry <- dh z
        if (own) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int r = hf == 0 ? r0 : r1;
                float dgi_[3][2], dnr[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float rr = e ? pr_[hf][0].y : pr_[hf][0].x, zz = e ? pr_[hf][1].y : pr_[hf][1].x;
                    const float nn = e ? pr_[hf][2].y : pr_[hf][2].x, ghn = e ? pr_[hf][3].y : pr_[hf][3].x;
                    const float hp = e ? ph[hf].y : ph[hf].x, dq = e ? pq[hf].y : pq[hf].x;
                    const float gv = carry[2 * hf + e] + dq;
                    const float dn = gv * (1.0f - zz) * (1.0f - nn * nn);
                    const float dz = gv * (hp - nn) * zz * (1.0f - zz);
                    const float dr = dn * ghn * rr * (1.0f - rr);
                    dgi_[0][e] = dr; dgi_[1][e] = dz; dgi_[2][e] = dn; dnr[e] = dn * rr;
                    carry[2 * hf + e] = gv * zz;
                }
                float* gi_ = sDGI + r * L::ldg + chb;
                float* gh_ = sDGH + r * L::ldg + chb;
                *reinterpret_cast<float2*>(gi_) = make_float2(dgi_[0][0], dgi_[0][1]);
                *reinterpret_cast<float2*>(gi_ + H) = make_float2(dgi_[1][0], dgi_[1][1]);
                *reinterpret_cast<float2*>(gi_ + 2 * H) = make_float2(dgi_[2][0], dgi_[2][1]);
                *reinterpret_cast<float2*>(gh_) = make_float2(dgi_[0][0], dgi_[0][1]);
                *reinterpret_cast<float2*>(gh_ + H) = make_float2(dgi_[1][0], dgi_[1][1]);
                *reinterpret_cast<float2*>(gh_ + 2 * H) = make_float2(dnr[0], dnr[1]);
            }
            if (t > 0) prefetch(t - 1);                  // in flight during the products and the attention backward
        }
        if (tm) {                                        // vsq / alpha of this step -> shared memory (used after the products)
            const float* gv_ = a.sv_vsq + (size_t)t * n * Vp;
            for (int i = threadIdx.x; i < R * Vq; i += NTM) {
                const int r = i / Vq, q = i - r * Vq;
                const float4 v = r < n_valid ? __ldg(reinterpret_cast<const float4*>(gv_ + (row0 + r) * Vp) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(sVSQ + r * L::ldv + 4 * q) = v;
            }
            const float* ga = a.sv_alpha + (size_t)t * n * U;
            for (int i = threadIdx.x; i < R * 16; i += NTM) {
                const int r = i >> 4, j = i & 15;
                sAl[i] = (r < n_valid && j < U) ? __ldg(ga + (row0 + r) * U + j) : 0.f;
            }
        }
        __syncthreads();                                 // dgi / dgh (and vsq / alpha) are in shared memory
        {   // stash rows -> global, coalesced (every thread: 16-byte pieces of the 16 x 3H tiles)
            float* gdgi = a.st_dgi + (size_t)t * n * a.ld_st;
            float* gdgh = a.st_dgh + (size_t)t * n * a.ld_st;
            constexpr int Q = H3 / 4;
            for (int i = threadIdx.x; i < 2 * R * Q; i += NTM) {
                const int arr = i / (R * Q), rem = i - arr * (R * Q), r = rem / Q, q = rem - r * Q;
                if (r < n_valid) {
                    const float4 v = *reinterpret_cast<const float4*>((arr ? sDGH : sDGI) + r * L::ldg + 4 * q);
                    *reinterpret_cast<float4*>((arr ? gdgh : gdgi) + (row0 + r) * a.ld_st + 4 * q) = v;
                }
            }
        }
        if (own) {
            tile_gemm_f<L::KS, L::ldg, H / 8>(sDGH, wHH, warp, carry);                       // carry' = dh z + dgh W_hh
        } else if (dcw) {
            float dc[4] = {0.f, 0.f, 0.f, 0.f};
            tile_gemm_f<L::KS, L::ldg, Md / 8>(sDGI, wIC, warp - 8, dc);                      // dc = dgi W_ih[:, H:]
            const int col = 8 * (warp - 8) + 2 * c;
            *reinterpret_cast<float2*>(sDC + r0 * L::ldc + col) = make_float2(dc[0], dc[1]);
            *reinterpret_cast<float2*>(sDC + r1 * L::ldc + col) = make_float2(dc[2], dc[3]);
        }
        if (tm) {
            __syncthreads();                             // dc is in shared memory
            // ---- attention backward, one warp per destination row: lanes = (source i, quarter of the message)
            {
                const int r = warp, b0 = (r / U) * U;
                const int Up = U <= 8 ? 8 : 16, parts = 32 / Up;       // lanes per source
                const int i = lane / parts, part = lane - i * parts;
                float acc = 0.f;
                if (r < n_valid && i < U) {
                    const float* dcp = sDC + r * L::ldc;
                    const float* vp = sVSQ + (b0 + i) * L::ldv;
                    for (int m = 4 * part; m < M; m += 4 * parts) {
                        const float4 d4 = *reinterpret_cast<const float4*>(dcp + m);
                        const float4 v4 = *reinterpret_cast<const float4*>(vp + m);
                        acc = fmaf(d4.x, v4.x, acc); acc = fmaf(d4.y, v4.y, acc);
                        acc = fmaf(d4.z, v4.z, acc); acc = fmaf(d4.w, v4.w, acc);
                    }
                }
                for (int o = parts >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                const float al = (r < n_valid && i < U) ? sAl[r * 16 + i] : 0.f;
                float tot = part == 0 ? al * acc : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                if (part == 0 && i < 16) sDS[r * 16 + i] = al * (acc - tot);
            }
            __syncthreads();
            float* gdv = a.st_dvsq + (size_t)t * n * a.ld_st;
            for (int p = threadIdx.x; p < R * Vp; p += NTM) {
                const int r = p / (Vp > 0 ? Vp : 1), f = p - r * Vp;
                const int b0 = (r / U) * U, li = r - b0;
                float acc = 0.f;
                if (r < n_valid) {
                    if (f < M) {
                        for (int j = 0; j < U; ++j) acc = fmaf(sAl[(b0 + j) * 16 + li], sDC[(b0 + j) * L::ldc + f], acc);
                    } else if (f < M + K) {
                        for (int j = 0; j < U; ++j) acc = fmaf(sDS[(b0 + j) * 16 + li], sVSQ[(b0 + j) * L::ldv + f + K], acc);
                        acc *= scale;
                    } else if (f < M + 2 * K) {
                        for (int i = 0; i < U; ++i) acc = fmaf(sDS[r * 16 + i], sVSQ[(b0 + i) * L::ldv + f - K], acc);
                        acc *= scale;
                    }
                    gdv[(row0 + r) * a.ld_st + f] = acc;
                }
            }
        }
        __syncthreads();                                 // shared tiles are free for the next step
    }
    if (a.d_h0 != nullptr && own) {
        if (v0) *reinterpret_cast<float2*>(a.d_h0 + o0 * H + chb) = make_float2(carry[0], carry[1]);
        if (v1) *reinterpret_cast<float2*>(a.d_h0 + o1 * H + chb) = make_float2(carry[2], carry[3]);
    }
}

template <int H, int M, int VP>
static int launch_bwd_mma(const Args& a, cudaStream_t st) {
    static const cudaError_t rc_attr = cudaFuncSetAttribute(seq2_bwd_mma_kernel<H, M, VP>,
                                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (rc_attr != cudaSuccess) { set_error("ubs_agent_seq2_bwd: %s", cudaGetErrorString(rc_attr)); return 1; }
    const int rpt = a.d.rows_per_tile();
    const size_t smem = (size_t)LayB<H, M, VP>::total * sizeof(float);
    seq2_bwd_mma_kernel<H, M, VP><<<(unsigned)((a.N + rpt - 1) / rpt), NTM, smem, st>>>(a);
    return check_launch("ubs_agent_seq2_bwd(mma)");
}
static int launch_bwd(const Args& a, cudaStream_t st);

#ifdef UBS_SEQ2_TRACE
}  // namespace mma
}  // namespace seq2
}  // namespace ubs
extern "C" UBS_API int ubs_seq2_trace_read(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, ubs::seq2::mma::ubs_trace, sizeof(long long) * 8 * 16 * 8);
}
namespace ubs {
namespace seq2 {
namespace mma {
#endif
static bool fwd_supported(const Dims& d) {
    if (d.tarmac()) return (d.H == 64 || d.H == 32) && d.M == 64 && d.Vp() == 96 && d.U <= 16 && d.K <= 32;
    return d.H == 64 || d.H == 32;
}
static int launch_fwd(const Args& a, cudaStream_t st) {
    if (a.d.tarmac()) return a.d.H == 64 ? launch_fwd_mma<64, 64, 96>(a, st) : launch_fwd_mma<32, 64, 96>(a, st);
    return a.d.H == 64 ? launch_fwd_mma<64, 0, 0>(a, st) : launch_fwd_mma<32, 0, 0>(a, st);
}
static int launch_bwd(const Args& a, cudaStream_t st) {
    if (a.d.tarmac()) return a.d.H == 64 ? launch_bwd_mma<64, 64, 96>(a, st) : launch_bwd_mma<32, 64, 96>(a, st);
    return a.d.H == 64 ? launch_bwd_mma<64, 0, 0>(a, st) : launch_bwd_mma<32, 0, 0>(a, st);
}

}  // namespace mma

static size_t fwd_smem(const Dims& d) {
    const int H = d.H, H3 = 3 * H, Vp = d.Vp();
    size_t f = (size_t)H * H3 + H3 + (size_t)(H + 2 * H3) * RP + NT * 16;
    if (d.tarmac()) f += (size_t)H * Vp + (size_t)d.M * H3 + (size_t)(Vp + d.M + d.U) * RP + 16;
    return f * sizeof(float);
}
static size_t bwd_smem(const Dims& d) {
    const int H = d.H, H3 = 3 * H, Vp = d.Vp();
    size_t f = (size_t)H3 * H + (size_t)(H + 2 * H3) * RP + NT * 16;
    if (d.tarmac()) f += (size_t)H3 * d.M + (size_t)(d.M + 2 * Vp + 2 * d.U) * RP + 16;
    return f * sizeof(float);
}
static int check(const char* fn, const Dims& d) {
    if (d.H < 16 || d.H % 4 || d.H > 256) { set_error("%s: hidden size %d unsupported", fn, d.H); return 2; }
    if (d.tarmac() && (d.U < 1 || d.U > R || d.M % 4 || d.M < 4 || d.K < 1)) { set_error("%s: TarMAC shape unsupported", fn); return 2; }
    return 0;
}

}  // namespace seq2
}  // namespace ubs

extern "C" UBS_API int64_t ubs_agent_seq2_smem_bytes(int H, int M, int K, int U, int flags, int backward) {
    ubs::seq2::Dims d{H, (flags & UBS_STEP_TARMAC) ? M : 0, (flags & UBS_STEP_TARMAC) ? K : 0,
                      (flags & UBS_STEP_TARMAC) ? U : 1, flags};
    return (int64_t)(backward ? ubs::seq2::bwd_smem(d) : ubs::seq2::fwd_smem(d));
}

extern "C" UBS_API int ubs_agent_seq2_fwd(int H, int M, int K, int U, int flags, const float* wt_vsq_h,
                                          const float* wt_ih_c, const float* wt_hh, const float* b_hh, const float* pv,
                                          const float* pg, const float* h0, const uint32_t* mask, float* h_out,
                                          float* sv_vsq, float* sv_alpha, float* sv_c, float* sv_gate, int64_t ld_pv,
                                          int64_t ld_pg, int64_t n_rows, int n_steps, void* stream) {
    using namespace ubs::seq2;
    Args a{};
    a.d = Dims{H, (flags & UBS_STEP_TARMAC) ? M : 0, (flags & UBS_STEP_TARMAC) ? K : 0, (flags & UBS_STEP_TARMAC) ? U : 1, flags};
    if (int rc = check("ubs_agent_seq2_fwd", a.d)) return rc;
    UBS_REQUIRE(wt_hh && b_hh && pg && h0 && h_out, "ubs_agent_seq2_fwd: NULL argument");
    UBS_REQUIRE(!a.d.tarmac() || (wt_vsq_h && wt_ih_c && pv && mask), "ubs_agent_seq2_fwd: TarMAC arguments missing");
    UBS_REQUIRE(!a.d.tarmac() || n_rows % a.d.U == 0, "ubs_agent_seq2_fwd: n_rows must be a multiple of agents per env");
    UBS_REQUIRE(sv_gate == nullptr || !a.d.tarmac() || (sv_vsq && sv_alpha && sv_c), "ubs_agent_seq2_fwd: incomplete save buffers");
    if (n_rows == 0 || n_steps == 0) return 0;
    const size_t smem = fwd_smem(a.d);
    if (smem > 227 * 1024) { ubs::set_error("ubs_agent_seq2_fwd: weights do not fit shared memory (%zu B)", smem); return 3; }
    a.w0 = wt_vsq_h; a.w1 = wt_ih_c; a.w2 = wt_hh; a.b_hh = b_hh; a.pv = pv; a.pg = pg; a.h0 = h0; a.mask = mask;
    a.h_out = h_out; a.sv_vsq = sv_vsq; a.sv_alpha = sv_alpha; a.sv_c = sv_c; a.sv_gate = sv_gate;
    a.N = n_rows; a.T = n_steps; a.ld_pv = ld_pv; a.ld_pg = ld_pg;
    UBS_REQUIRE(ld_pg >= 3 * H && ld_pg % 4 == 0 && (!a.d.tarmac() || (ld_pv >= a.d.Vp() && ld_pv % 4 == 0)), "ubs_agent_seq2_fwd: bad leading dimensions");
    UBS_REQUIRE(((uintptr_t)pg % 16) == 0 && ((uintptr_t)pv % 16) == 0, "ubs_agent_seq2_fwd: pv / pg must be 16-byte aligned");
    const int rpt = a.d.rows_per_tile();
    static const bool use_mma = [] { const char* e = getenv("UBS_SEQ2_MMA"); return !(e && e[0] == '0'); }();
    // tensor-core window kernel (mma.sync 3xTF32) for the compiled (H, M, K) instances; UBS_SEQ2_MMA=0 keeps the FP32
    // kernel for A/B measurements
    if (use_mma && mma::fwd_supported(a.d) && ld_pg % 2 == 0 && ld_pv % 2 == 0) return mma::launch_fwd(a, (cudaStream_t)stream);
    UBS_OPT_IN_SMEM(seq2_fwd_kernel, "ubs_agent_seq2_fwd");
    seq2_fwd_kernel<<<(unsigned)((n_rows + rpt - 1) / rpt), NT, smem, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_agent_seq2_fwd");
}

extern "C" UBS_API int ubs_agent_seq2_bwd(int H, int M, int K, int U, int flags, const float* w_hh, const float* w_ih_c,
                                          const float* h0, const float* h_out, const float* sv_vsq,
                                          const float* sv_alpha, const float* sv_gate, const float* dhq, float* st_dgi,
                                          float* st_dgh, float* st_dvsq, float* d_h0, int64_t ld_stash, int64_t n_rows,
                                          int n_steps, void* stream) {
    using namespace ubs::seq2;
    Args a{};
    a.d = Dims{H, (flags & UBS_STEP_TARMAC) ? M : 0, (flags & UBS_STEP_TARMAC) ? K : 0, (flags & UBS_STEP_TARMAC) ? U : 1, flags};
    if (int rc = check("ubs_agent_seq2_bwd", a.d)) return rc;
    UBS_REQUIRE(w_hh && h0 && h_out && sv_gate && dhq && st_dgi && st_dgh, "ubs_agent_seq2_bwd: NULL argument");
    UBS_REQUIRE(!a.d.tarmac() || (w_ih_c && sv_vsq && sv_alpha && st_dvsq), "ubs_agent_seq2_bwd: TarMAC arguments missing");
    if (n_rows == 0 || n_steps == 0) return 0;
    const size_t smem = bwd_smem(a.d);
    if (smem > 227 * 1024) { ubs::set_error("ubs_agent_seq2_bwd: weights do not fit shared memory (%zu B)", smem); return 3; }
    a.w0 = w_hh; a.w1 = w_ih_c; a.h0 = h0; a.h_out = const_cast<float*>(h_out);
    a.sv_vsq = const_cast<float*>(sv_vsq); a.sv_alpha = const_cast<float*>(sv_alpha); a.sv_gate = const_cast<float*>(sv_gate);
    a.dhq = dhq; a.st_dgi = st_dgi; a.st_dgh = st_dgh; a.st_dvsq = st_dvsq; a.d_h0 = d_h0; a.N = n_rows; a.T = n_steps;
    a.ld_st = ld_stash;
    UBS_REQUIRE(ld_stash >= 3 * H, "ubs_agent_seq2_bwd: bad stash leading dimension");
    const int rpt = a.d.rows_per_tile();
    static const bool use_mma = [] { const char* e = getenv("UBS_SEQ2_MMA"); return !(e && e[0] == '0'); }();
    if (use_mma && mma::fwd_supported(a.d) && ld_stash % 4 == 0 && ((uintptr_t)st_dgi % 16) == 0 && ((uintptr_t)st_dgh % 16) == 0 &&
        ((uintptr_t)sv_vsq % 16) == 0)
        return mma::launch_bwd(a, (cudaStream_t)stream);
    UBS_OPT_IN_SMEM(seq2_bwd_kernel, "ubs_agent_seq2_bwd");
    seq2_bwd_kernel<<<(unsigned)((n_rows + rpt - 1) / rpt), NT, smem, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_agent_seq2_bwd");
}
