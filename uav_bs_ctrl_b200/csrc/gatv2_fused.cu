// Fused GATv2 relation kernels (forward + backward) for small source/destination feature widths.
//
// Replaces what the reference delegates to dglnn.GATv2Conv (call sites
// algos/madrqn/agents/gnn_agents.py:92-97,103-104; algos/drqn/agents/gnn_agents.py:17-18,27), i.e. DGL's
// fc_src/fc_dst GEMMs + gSDDMM(u_add_v) + leaky_relu + (e*attn).sum + edge_softmax (5 kernels) + gSpMM(u_mul_e,sum)
// + res_fc + ReLU, with ONE pass over the edges.
//
// Design notes:
//  * F_s <= 4 and F_d <= 2 in the env configs, so el = W_src x_u + b_src is never materialised: it is recomputed
//    in registers from the 8..16-byte source row.  The aggregation is linear in x, so the per-edge weighted sum
//    runs on the RAW source row (F_s values per head) and is projected once per destination:
//        ft[v,k,:] = W_src[k] (sum_e alpha[e,k] x_e) + b_src[k]
//  * forward: a group of GS lanes owns one destination, one EDGE per lane; the H score channels are looped with
//    the weights broadcast from shared memory; online softmax per lane, merged across the group at the end.
//    Star layout => consecutive lanes read consecutive 16-byte rows: fully coalesced, no index array at all.
//  * backward: one warp per destination, lanes own CPL = H/32 channels so every parameter-gradient accumulator
//    lives in a register for the whole kernel; per-CTA partials go to a workspace and are summed by a second
//    kernel in a fixed order (deterministic; no atomics on the parameter gradients).
#include "common.cuh"
#include "../../include/ubs_gnn.h"
#include <stdlib.h>

namespace ubs {

struct GatArgs {
    const float* x_src; const float* x_dst; const int* indptr; const int* src_idx;
    const float* W_src; const float* b_src; const float* W_dst; const float* b_dst;
    const float* attn; const float* W_res; const float* b_res;
    float* out; float* smax; float* ssum;
    // backward only
    const float* grad_out; const float* out_in; const float* smax_in; const float* ssum_in;
    float* partial; float* grad_x_src; float* grad_x_dst;
    int n_dst; int F_d; int D; float slope; int flags;
    // strided segments (one per timestep of a sequence arena): destination v belongs to segment v / n_dst_seg whose
    // x_src / x_dst / indptr / src_idx start seg * st_* elements after the base pointers (indptr is segment-local).
    // Outputs are indexed by the global destination id with row strides ld_out / ld_gout.
    int n_dst_seg; long long st_xsrc, st_xdst, st_ip, st_sidx; int ld_out, ld_gout;
    // optional per-(edge slot, head) raw attention scores: written by a training forward, read by the backward instead of
    // recomputing them (16 B per edge at 4 heads instead of ~5 H FMAs); slot e of segment s lives at (s * st_score + e) * heads
    float* score; const float* score_in; long long st_score;
};

template <int FS>
__device__ __forceinline__ void load_row(const float* __restrict__ p, size_t row, float (&x)[FS]) {
    if constexpr (FS == 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + row);
        x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else if constexpr (FS == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p) + row);
        x[0] = t.x; x[1] = t.y;
    } else {
#pragma unroll
        for (int f = 0; f < FS; ++f) x[f] = __ldg(p + row * FS + f);
    }
}

// ------------------------------------------------------------------------------------------------ forward
// EPL = edges per lane per pass: the weight / destination operands of a channel (one LDS.128 + one LDS.64) are reused
// for EPL edges, which moves the loop from shared-memory-issue bound (2 LDS : 5 FFMA) towards FFMA bound.  A lane
// still visits its edges in the same order (beg+li, +GS, +2GS, ...), so the result is bit-identical for every EPL.
// HS = head split: a destination is HS work items, each owning HEADS / HS heads (heads are independent: own softmax,
// own output channels), so a launch that cannot fill the SMs (the act step: 2 048 destinations) gets HS x the warps,
// each with 1 / HS of the channel loop.  Results are bit-identical for every HS.
template <int FS, int HEADS, int GS, int EPL, int HS>
__global__ void __launch_bounds__(256) gatv2_fwd_kernel(const GatArgs a) {
    constexpr int HPW = HEADS / HS;        // heads per work item
    constexpr int GPW = 32 / GS;           // destination groups per warp
    constexpr int GPB = 256 / GS;          // groups per block
    const int D = a.D, H = HEADS * D;
    extern __shared__ float4 smem4[];
    float4* wA = smem4;                    // [H] W_src row (zero padded to 4)
    float4* wR = wA + H;                   // [H] {W_res[0], W_res[1], b_res, b_src}
    float4* wD = wR + H;                   // [H] {W_dst[0], W_dst[1], b_src + b_dst, attn}
    float* cbase = reinterpret_cast<float*>(wD + H);
    const int cstride = 2 * H + 2;         // +2: groups of one warp hit different banks
    const int tid = threadIdx.x;
    for (int ch = tid; ch < H; ch += 256) {
        float w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int f = 0; f < FS; ++f) w[f] = a.W_src[ch * FS + f];
        wA[ch] = make_float4(w[0], w[1], w[2], w[3]);
        const float bs = a.b_src ? a.b_src[ch] : 0.f;
        const float bd = a.b_dst ? a.b_dst[ch] : 0.f;
        float r0 = 0.f, r1 = 0.f, rb = 0.f;
        if (a.flags & UBS_GAT_RESIDUAL) {
            r0 = a.W_res[ch * a.F_d];
            r1 = a.F_d > 1 ? a.W_res[ch * a.F_d + 1] : 0.f;
            rb = a.b_res ? a.b_res[ch] : 0.f;
        }
        wR[ch] = make_float4(r0, r1, rb, bs);
        // leaky_relu(z) = (1+s)/2 z + (1-s)/2 |z|: the |z| part costs one FFMA per channel (free |.| operand
        // modifier), the linear part folds into per-head constants (hP below)
        wD[ch] = make_float4(a.W_dst[ch * a.F_d], a.F_d > 1 ? a.W_dst[ch * a.F_d + 1] : 0.f, bs + bd,
                             0.5f * (1.0f - a.slope) * a.attn[ch]);
    }
    __syncthreads();
    // hP[k] = (1+s)/2 * sum_d attn[k,d] * {W_src[k,d,0..3] | W_dst[k,d,0..1], b_src+b_dst}: 8 floats per head
    float* hP = cbase + GPB * cstride;
    if (tid < HEADS * 8) {
        const int k = tid / 8, j = tid % 8;
        float acc = 0.f;
        for (int d0 = 0; d0 < D; ++d0) {
            const int ch = k * D + d0;
            const float at = a.attn[ch];
            float w = 0.f;
            if (j < 4) w = j == 0 ? wA[ch].x : j == 1 ? wA[ch].y : j == 2 ? wA[ch].z : wA[ch].w;
            else if (j == 4) w = wD[ch].x;
            else if (j == 5) w = wD[ch].y;
            else if (j == 6) w = wD[ch].z;
            acc = fmaf(at, w, acc);
        }
        hP[tid] = 0.5f * (1.0f + a.slope) * acc;
    }
    __syncthreads();

    const int li = tid % GS;                               // lane inside the group
    const int grp_in_block = tid / GS;
    float2* cbuf = reinterpret_cast<float2*>(cbase + grp_in_block * cstride);
    const int warp_first = (blockIdx.x * GPB) + (tid / 32) * GPW;
    const int stride = gridDim.x * GPB;
    const float slope = a.slope;
    const bool relu = a.flags & UBS_GAT_RELU, has_res = a.flags & UBS_GAT_RESIDUAL;

    const int n_items = a.n_dst * HS;
    for (int base = warp_first; base < n_items; base += stride) {
        const int item = base + (tid % 32) / GS;
        const int v = item / HS, k0 = (item - v * HS) * HPW;       // destination, first head of this work item
        const bool active = item < n_items;
        int beg = 0, end = 0;
        float xv0 = 0.f, xv1 = 0.f;
        const int seg = active ? v / a.n_dst_seg : 0;
        const int vl = v - seg * a.n_dst_seg;
        const float* xsrc = a.x_src + seg * a.st_xsrc;
        const int* sidx = a.src_idx ? a.src_idx + seg * a.st_sidx : nullptr;
        if (active) {
            const int* ip = a.indptr + seg * a.st_ip;
            const float* xd = a.x_dst + seg * a.st_xdst + (size_t)vl * a.F_d;
            beg = __ldg(ip + vl);
            end = __ldg(ip + vl + 1);
            xv0 = __ldg(xd);
            xv1 = a.F_d > 1 ? __ldg(xd + 1) : 0.f;
        }
        // destination part of the score pre-activation, shared by all edges of v
        for (int ch = k0 * D + li; ch < (k0 + HPW) * D; ch += GS) {
            const float4 w = wD[ch];
            cbuf[ch] = make_float2(fmaf(w.y, xv1, fmaf(w.x, xv0, w.z)), w.w);
        }
        __syncwarp();
        float lin[HPW];
#pragma unroll
        for (int k = 0; k < HPW; ++k)
            lin[k] = fmaf(hP[(k0 + k) * 8 + 5], xv1, fmaf(hP[(k0 + k) * 8 + 4], xv0, hP[(k0 + k) * 8 + 6]));

        float m[HPW], l[HPW], acc[HPW][FS];
#pragma unroll
        for (int k = 0; k < HPW; ++k) {
            m[k] = -CUDART_INF_F; l[k] = 0.f;
#pragma unroll
            for (int f = 0; f < FS; ++f) acc[k][f] = 0.f;
        }
        for (int e = beg + li; e < end; e += GS * EPL) {
            float x[EPL][FS];
            float sk[EPL][HPW];
            bool ok[EPL];
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                const int ej = e + j * GS;
                ok[j] = ej < end;
                if (ok[j]) {
                    const size_t u = sidx ? (size_t)__ldg(sidx + ej) : (size_t)ej;
                    load_row<FS>(xsrc, u, x[j]);
                } else {
#pragma unroll
                    for (int f = 0; f < FS; ++f) x[j][f] = 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < HPW; ++k) {
                const float4* wk = wA + (k0 + k) * D;
                const float2* ck = cbuf + (k0 + k) * D;
                // linear part of the score: (1+s)/2 * attn_k . (W_src x + W_dst x_v + b)
                const float4 pk = *reinterpret_cast<const float4*>(hP + (k0 + k) * 8);
                float s[EPL];
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    s[j] = fmaf(pk.x, x[j][0], lin[k]);
                    if constexpr (FS > 1) s[j] = fmaf(pk.y, x[j][1], s[j]);
                    if constexpr (FS > 2) s[j] = fmaf(pk.z, x[j][2], s[j]);
                    if constexpr (FS > 3) s[j] = fmaf(pk.w, x[j][3], s[j]);
                }
#pragma unroll 8
                for (int d = 0; d < D; ++d) {
                    const float4 w = wk[d];
                    const float2 c = ck[d];
#pragma unroll
                    for (int j = 0; j < EPL; ++j) {
                        float z = c.x;
                        z = fmaf(w.x, x[j][0], z);
                        if constexpr (FS > 1) z = fmaf(w.y, x[j][1], z);
                        if constexpr (FS > 2) z = fmaf(w.z, x[j][2], z);
                        if constexpr (FS > 3) z = fmaf(w.w, x[j][3], z);
                        s[j] = fmaf(c.y, fabsf(z), s[j]);          // (1-s)/2 * attn * |z|
                    }
                }
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    sk[j][k] = s[j];
                    if (ok[j]) {
                        const float mn = fmaxf(m[k], s[j]);
                        const float sc = __expf(m[k] - mn);
                        const float p = __expf(s[j] - mn);
                        l[k] = fmaf(l[k], sc, p);
#pragma unroll
                        for (int f = 0; f < FS; ++f) acc[k][f] = fmaf(acc[k][f], sc, p * x[j][f]);
                        m[k] = mn;
                    }
                }
            }
            if (a.score != nullptr) {
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    if (ok[j]) {
                        float* sp = a.score + ((size_t)seg * a.st_score + (e + j * GS)) * HEADS + k0;
                        if constexpr (HPW == 4) *reinterpret_cast<float4*>(sp) = make_float4(sk[j][0], sk[j][1], sk[j][2], sk[j][3]);
                        else if constexpr (HPW == 2) *reinterpret_cast<float2*>(sp) = make_float2(sk[j][0], sk[j][1]);
                        else {
#pragma unroll
                            for (int k = 0; k < HPW; ++k) sp[k] = sk[j][k];
                        }
                    }
                }
            }
        }
        // merge the per-lane online-softmax states of the group
#pragma unroll
        for (int k = 0; k < HPW; ++k) {
            const float M = group_max<GS>(m[k]);
            const float sc = (m[k] == -CUDART_INF_F) ? 0.f : __expf(m[k] - M);
            l[k] = group_sum<GS>(l[k] * sc);
#pragma unroll
            for (int f = 0; f < FS; ++f) acc[k][f] = group_sum<GS>(acc[k][f] * sc);
            m[k] = M;
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < HPW; ++k) {
                const float inv = l[k] > 0.f ? 1.0f / l[k] : 0.f;
                for (int d = li; d < D; d += GS) {
                    const int ch = (k0 + k) * D + d;
                    const float4 w = wA[ch];
                    const float4 r = wR[ch];
                    float o = 0.f;
                    if (l[k] > 0.f) {
                        float t = w.x * acc[k][0];
                        if constexpr (FS > 1) t = fmaf(w.y, acc[k][1], t);
                        if constexpr (FS > 2) t = fmaf(w.z, acc[k][2], t);
                        if constexpr (FS > 3) t = fmaf(w.w, acc[k][3], t);
                        o = fmaf(t, inv, r.w);
                    }
                    if (has_res) o += fmaf(r.y, xv1, fmaf(r.x, xv0, r.z));
                    if (relu) o = fmaxf(o, 0.f);
                    a.out[(size_t)v * a.ld_out + ch] = o;
                }
                if (li == 0 && a.smax != nullptr) {
                    a.smax[(size_t)v * HEADS + k0 + k] = l[k] > 0.f ? m[k] : 0.f;
                    a.ssum[(size_t)v * HEADS + k0 + k] = l[k];
                }
            }
        }
        __syncwarp();   // cbuf is rewritten by the next destination
    }
}

// ------------------------------------------------------------------------------------------------ backward
__host__ __device__ inline int gat_param_count(int H, int FS, int FD) { return H * (FS + 2 * FD + 4); }

template <int FS, int CPL, int HEADS>
__global__ void __launch_bounds__(128) gatv2_bwd_kernel(const GatArgs a) {
    constexpr int H = 32 * CPL, D = H / HEADS, LPH = 32 / HEADS;
    static_assert(D % CPL == 0, "a lane's channels must stay inside one head");
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int head = (lane * CPL) / D;
    const int FD = a.F_d;
    const bool relu = a.flags & UBS_GAT_RELU, has_res = a.flags & UBS_GAT_RESIDUAL;
    const float slope = a.slope;
    __shared__ float4 xs[4][32];
    __shared__ float red[H * (4 + 2 * 2 + 4)];

    float ws[CPL][FS], wd[CPL][2], wr[CPL][2], bsum[CPL], bs[CPL], brr[CPL], at[CPL];
    float g_ws[CPL][FS], g_wd[CPL][2], g_wr[CPL][2], g_bsd[CPL], g_bsm[CPL], g_br[CPL], g_at[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const int ch = lane * CPL + j;
#pragma unroll
        for (int f = 0; f < FS; ++f) { ws[j][f] = a.W_src[ch * FS + f]; g_ws[j][f] = 0.f; }
        wd[j][0] = a.W_dst[ch * FD]; wd[j][1] = FD > 1 ? a.W_dst[ch * FD + 1] : 0.f;
        wr[j][0] = has_res ? a.W_res[ch * FD] : 0.f;
        wr[j][1] = (has_res && FD > 1) ? a.W_res[ch * FD + 1] : 0.f;
        brr[j] = (has_res && a.b_res) ? a.b_res[ch] : 0.f;
        bs[j] = a.b_src ? a.b_src[ch] : 0.f;
        bsum[j] = bs[j] + (a.b_dst ? a.b_dst[ch] : 0.f);
        at[j] = a.attn[ch];
        g_wd[j][0] = g_wd[j][1] = g_wr[j][0] = g_wr[j][1] = 0.f;
        g_bsd[j] = g_bsm[j] = g_br[j] = g_at[j] = 0.f;
    }

    const int total_warps = gridDim.x * 4;
    for (int v = blockIdx.x * 4 + warp; v < a.n_dst; v += total_warps) {
        const int seg = v / a.n_dst_seg, vl = v - seg * a.n_dst_seg;
        const float* xsrc = a.x_src + seg * a.st_xsrc;
        const int* sidx = a.src_idx ? a.src_idx + seg * a.st_sidx : nullptr;
        const int* ip = a.indptr + seg * a.st_ip;
        const float* xd = a.x_dst + seg * a.st_xdst + (size_t)vl * FD;
        const int beg = __ldg(ip + vl), end = __ldg(ip + vl + 1);
        const float xv0 = __ldg(xd);
        const float xv1 = FD > 1 ? __ldg(xd + 1) : 0.f;
        float gp[CPL], ft[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int chj = lane * CPL + j;
            const float go = __ldg(a.grad_out + (size_t)v * a.ld_gout + chj), oo = __ldg(a.out_in + (size_t)v * a.ld_out + chj);
            gp[j] = (relu && !(oo > 0.f)) ? 0.f : go;
            const float res = has_res ? fmaf(wr[j][1], xv1, fmaf(wr[j][0], xv0, brr[j])) : 0.f;
            ft[j] = oo - res;                                   // only used where gp != 0
            g_wr[j][0] = fmaf(gp[j], xv0, g_wr[j][0]);
            g_wr[j][1] = fmaf(gp[j], xv1, g_wr[j][1]);
            g_br[j] += gp[j];
        }
        float sdz[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) sdz[j] = 0.f;
        if (end > beg) {                                        // warp-uniform
            float dotp = 0.f, qk = 0.f, pk[FS];
#pragma unroll
            for (int f = 0; f < FS; ++f) pk[f] = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                dotp = fmaf(gp[j], ft[j], dotp);
                qk = fmaf(gp[j], bs[j], qk);
#pragma unroll
                for (int f = 0; f < FS; ++f) pk[f] = fmaf(gp[j], ws[j][f], pk[f]);
            }
            dotp = group_sum<LPH>(dotp);
            qk = group_sum<LPH>(qk);
#pragma unroll
            for (int f = 0; f < FS; ++f) pk[f] = group_sum<LPH>(pk[f]);
            const float Mx = __ldg(a.smax_in + (size_t)v * HEADS + head);
            const float invL = 1.0f / __ldg(a.ssum_in + (size_t)v * HEADS + head);
            float c[CPL], abar[FS];
#pragma unroll
            for (int j = 0; j < CPL; ++j) c[j] = fmaf(wd[j][1], xv1, fmaf(wd[j][0], xv0, bsum[j]));
#pragma unroll
            for (int f = 0; f < FS; ++f) abar[f] = 0.f;

            for (int e0 = beg; e0 < end; e0 += 32) {
                const int cnt = min(32, end - e0);
                int u_lane = 0;
                if (lane < cnt) {
                    u_lane = sidx ? __ldg(sidx + e0 + lane) : (e0 + lane);
                    float xr[FS];
                    load_row<FS>(xsrc, (size_t)u_lane, xr);
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    t.x = xr[0];
                    if constexpr (FS > 1) t.y = xr[1];
                    if constexpr (FS > 2) t.z = xr[2];
                    if constexpr (FS > 3) t.w = xr[3];
                    xs[warp][lane] = t;
                }
                __syncwarp();
                for (int i = 0; i < cnt; ++i) {
                    const float4 t = xs[warp][i];
                    const float x[4] = {t.x, t.y, t.z, t.w};
                    float z[CPL], y[CPL], sp = 0.f;
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        float zz = c[j];
#pragma unroll
                        for (int f = 0; f < FS; ++f) zz = fmaf(ws[j][f], x[f], zz);
                        z[j] = zz;
                        y[j] = fmaxf(zz, slope * zz);
                        sp = fmaf(at[j], y[j], sp);
                    }
                    const float s = group_sum<LPH>(sp);
                    const float alpha = __expf(s - Mx) * invL;               // the forward's exp (same alpha in both directions)
                    float da = qk;
#pragma unroll
                    for (int f = 0; f < FS; ++f) da = fmaf(pk[f], x[f], da);
                    const float ds = alpha * (da - dotp);
#pragma unroll
                    for (int f = 0; f < FS; ++f) abar[f] = fmaf(alpha, x[f], abar[f]);
                    float gxs[FS];
#pragma unroll
                    for (int f = 0; f < FS; ++f) gxs[f] = 0.f;
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        g_at[j] = fmaf(ds, y[j], g_at[j]);
                        const float dz = ds * at[j] * (z[j] > 0.f ? 1.0f : slope);
#pragma unroll
                        for (int f = 0; f < FS; ++f) g_ws[j][f] = fmaf(dz, x[f], g_ws[j][f]);
                        sdz[j] += dz;
                        if (a.grad_x_src != nullptr) {
                            const float t2 = fmaf(alpha, gp[j], dz);
#pragma unroll
                            for (int f = 0; f < FS; ++f) gxs[f] = fmaf(t2, ws[j][f], gxs[f]);
                        }
                    }
                    if (a.grad_x_src != nullptr) {               // warp-uniform
                        const int u = __shfl_sync(0xffffffffu, u_lane, i);
#pragma unroll
                        for (int f = 0; f < FS; ++f) {
                            const float tot = warp_sum(gxs[f]);
                            if (lane == 0) {
                                float* gx = a.grad_x_src + seg * a.st_xsrc + (size_t)u * FS + f;
                                if (a.src_idx) atomicAdd(gx, tot);
                                else *gx = tot;
                            }
                        }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
#pragma unroll
                for (int f = 0; f < FS; ++f) g_ws[j][f] = fmaf(gp[j], abar[f], g_ws[j][f]);   // message path
                g_bsm[j] += gp[j];                                                            // sum_e alpha = 1
                g_wd[j][0] = fmaf(sdz[j], xv0, g_wd[j][0]);
                g_wd[j][1] = fmaf(sdz[j], xv1, g_wd[j][1]);
                g_bsd[j] += sdz[j];
            }
        }
        if (a.grad_x_dst != nullptr) {
            float t0 = 0.f, t1 = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                t0 += sdz[j] * wd[j][0] + gp[j] * wr[j][0];
                t1 += sdz[j] * wd[j][1] + gp[j] * wr[j][1];
            }
            t0 = warp_sum(t0); t1 = warp_sum(t1);
            if (lane == 0) {
                float* gd = a.grad_x_dst + seg * a.st_xdst + (size_t)vl * FD;
                gd[0] = t0;
                if (FD > 1) gd[1] = t1;
            }
        }
    }

    // ---- CTA reduction in a fixed warp order, then one partial row per CTA
    const int oWs = 0, obs = H * FS, oWd = obs + H, obd = oWd + H * FD, oat = obd + H, oWr = oat + H,
              obr = oWr + H * FD, P = obr + H;
    for (int w = 0; w < 4; ++w) {
        if (warp == w) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int ch = lane * CPL + j;
                auto put = [&](int idx, float val) { red[idx] = (w == 0) ? val : red[idx] + val; };
#pragma unroll
                for (int f = 0; f < FS; ++f) put(oWs + ch * FS + f, g_ws[j][f]);
                put(obs + ch, g_bsd[j] + g_bsm[j]);
                put(oWd + ch * FD, g_wd[j][0]);
                if (FD > 1) put(oWd + ch * FD + 1, g_wd[j][1]);
                put(obd + ch, g_bsd[j]);
                put(oat + ch, g_at[j]);
                put(oWr + ch * FD, g_wr[j][0]);
                if (FD > 1) put(oWr + ch * FD + 1, g_wr[j][1]);
                put(obr + ch, g_br[j]);
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < P; i += 128) a.partial[(size_t)blockIdx.x * P + i] = red[i];
}

// ---- backward for leaf observations (grad_x_src == NULL: every env config), two passes per chunk of 32 edges ----------
// The kernel above lets every lane walk every edge of its destination: the per-edge scalar work (score reduction by
// shuffles, expf, the softmax gradient) is repeated by all lanes of a head and sits on the dependent path of the
// channel loop.  Here a chunk of 32 edges is processed twice:
//   pass 1, lanes own EDGES (like the forward): score of every head with the weights broadcast from shared memory,
//           alpha = exp(s - max) / sum, ds = alpha (<g', el> - <g', ft>) -> shared memory; sum_e alpha x_e in registers;
//   pass 2, lanes own CHANNELS: only the parameter-gradient FMAs — z recomputed from the 16-byte source row,
//           g_attn += ds y,  t = ds lrelu'(z),  sum_e t x_e,  sum_e t — no shuffles, no expf, no per-head redundancy.
// Same per-CTA partials + fixed-order reduction as above (deterministic).
template <int FS, int CPL, int HEADS>
__global__ void __launch_bounds__(128) gatv2_bwd2_kernel(const GatArgs a) {
    constexpr int H = 32 * CPL, D = H / HEADS, LPH = 32 / HEADS;
    static_assert(D % CPL == 0, "a lane's channels must stay inside one head");
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int head = (lane * CPL) / D;
    const int FD = a.F_d;
    const bool relu = a.flags & UBS_GAT_RELU, has_res = a.flags & UBS_GAT_RESIDUAL;
    const float slope = a.slope;
    __shared__ float4 wA[H];                   // W_src rows (zero padded to 4)
    __shared__ float2 cb[4][H];                // per warp: {W_dst x_v + b_src + b_dst, (1-s)/2 attn} per channel
    __shared__ float hP[HEADS * 8];            // (1+s)/2 sum_d attn {W_src[0..3] | W_dst[0..1], b_src + b_dst}
    __shared__ float4 xs[4][32];               // source rows of the chunk
    __shared__ float dsS[4][HEADS][33];        // softmax-gradient of the chunk's (edge, head) scores (+1: bank skew)
    __shared__ float4 hc[4][HEADS][2];         // per (warp, head): {<g',ft>, <g',b_src>, <g',W_src[:,0..1]>}, {.., [:,2..3], max, 1/sum}
    __shared__ float red[H * (4 + 2 * 2 + 4)];

    float ws[CPL][FS], wd[CPL][2], wr[CPL][2], bsum[CPL], bs[CPL], brr[CPL], at[CPL];
    float g_ws[CPL][FS], g_wd[CPL][2], g_wr[CPL][2], g_bsd[CPL], g_bsm[CPL], g_br[CPL], g_at[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const int ch = lane * CPL + j;
#pragma unroll
        for (int f = 0; f < FS; ++f) { ws[j][f] = a.W_src[ch * FS + f]; g_ws[j][f] = 0.f; }
        wd[j][0] = a.W_dst[ch * FD]; wd[j][1] = FD > 1 ? a.W_dst[ch * FD + 1] : 0.f;
        wr[j][0] = has_res ? a.W_res[ch * FD] : 0.f;
        wr[j][1] = (has_res && FD > 1) ? a.W_res[ch * FD + 1] : 0.f;
        brr[j] = (has_res && a.b_res) ? a.b_res[ch] : 0.f;
        bs[j] = a.b_src ? a.b_src[ch] : 0.f;
        bsum[j] = bs[j] + (a.b_dst ? a.b_dst[ch] : 0.f);
        at[j] = a.attn[ch];
        g_wd[j][0] = g_wd[j][1] = g_wr[j][0] = g_wr[j][1] = 0.f;
        g_bsd[j] = g_bsm[j] = g_br[j] = g_at[j] = 0.f;
    }
    for (int ch = threadIdx.x; ch < H; ch += 128) {
        float w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int f = 0; f < FS; ++f) w[f] = a.W_src[ch * FS + f];
        wA[ch] = make_float4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    if (threadIdx.x < HEADS * 8) {
        const int k = threadIdx.x / 8, j = threadIdx.x % 8;
        float acc = 0.f;
        for (int d0 = 0; d0 < D; ++d0) {
            const int ch = k * D + d0;
            float w = 0.f;
            if (j < 4) w = j == 0 ? wA[ch].x : j == 1 ? wA[ch].y : j == 2 ? wA[ch].z : wA[ch].w;
            else if (j == 4) w = a.W_dst[ch * FD];
            else if (j == 5) w = FD > 1 ? a.W_dst[ch * FD + 1] : 0.f;
            else if (j == 6) w = (a.b_src ? a.b_src[ch] : 0.f) + (a.b_dst ? a.b_dst[ch] : 0.f);
            acc = fmaf(a.attn[ch], w, acc);
        }
        hP[threadIdx.x] = 0.5f * (1.0f + slope) * acc;
    }
    __syncthreads();

    const int total_warps = gridDim.x * 4;
    for (int v = blockIdx.x * 4 + warp; v < a.n_dst; v += total_warps) {
        const int seg = v / a.n_dst_seg, vl = v - seg * a.n_dst_seg;
        const float* xsrc = a.x_src + seg * a.st_xsrc;
        const int* sidx = a.src_idx ? a.src_idx + seg * a.st_sidx : nullptr;
        const int* ip = a.indptr + seg * a.st_ip;
        const float* xd = a.x_dst + seg * a.st_xdst + (size_t)vl * FD;
        const int beg = __ldg(ip + vl), end = __ldg(ip + vl + 1);
        const float xv0 = __ldg(xd);
        const float xv1 = FD > 1 ? __ldg(xd + 1) : 0.f;
        float gp[CPL], ft[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int chj = lane * CPL + j;
            const float go = __ldg(a.grad_out + (size_t)v * a.ld_gout + chj), oo = __ldg(a.out_in + (size_t)v * a.ld_out + chj);
            gp[j] = (relu && !(oo > 0.f)) ? 0.f : go;
            const float res = has_res ? fmaf(wr[j][1], xv1, fmaf(wr[j][0], xv0, brr[j])) : 0.f;
            ft[j] = oo - res;                                   // only used where gp != 0
            g_wr[j][0] = fmaf(gp[j], xv0, g_wr[j][0]);
            g_wr[j][1] = fmaf(gp[j], xv1, g_wr[j][1]);
            g_br[j] += gp[j];
        }
        float sdz[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) sdz[j] = 0.f;
        if (end > beg) {                                        // warp-uniform
            float dotp = 0.f, qk = 0.f, pk[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                dotp = fmaf(gp[j], ft[j], dotp);
                qk = fmaf(gp[j], bs[j], qk);
#pragma unroll
                for (int f = 0; f < FS; ++f) pk[f] = fmaf(gp[j], ws[j][f], pk[f]);
            }
            dotp = group_sum<LPH>(dotp);
            qk = group_sum<LPH>(qk);
#pragma unroll
            for (int f = 0; f < FS; ++f) pk[f] = group_sum<LPH>(pk[f]);
            float c[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                c[j] = fmaf(wd[j][1], xv1, fmaf(wd[j][0], xv0, bsum[j]));
                cb[warp][lane * CPL + j] = make_float2(c[j], 0.5f * (1.0f - slope) * at[j]);
            }
            if (lane % LPH == 0) {
                hc[warp][head][0] = make_float4(dotp, qk, pk[0], pk[1]);
                hc[warp][head][1] = make_float4(pk[2], pk[3], __ldg(a.smax_in + (size_t)v * HEADS + head),
                                                1.0f / __ldg(a.ssum_in + (size_t)v * HEADS + head));
            }
            __syncwarp();
            float abar[HEADS][FS], tws[CPL][FS], st[CPL];
#pragma unroll
            for (int k = 0; k < HEADS; ++k)
#pragma unroll
                for (int f = 0; f < FS; ++f) abar[k][f] = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                st[j] = 0.f;
#pragma unroll
                for (int f = 0; f < FS; ++f) tws[j][f] = 0.f;
            }
            for (int e0 = beg; e0 < end; e0 += 32) {
                const int cnt = min(32, end - e0);
                // ---- pass 1: lane = edge
                float x[4] = {0.f, 0.f, 0.f, 0.f};
                if (lane < cnt) {
                    const size_t u = sidx ? (size_t)__ldg(sidx + e0 + lane) : (size_t)(e0 + lane);
                    float xr[FS];
                    load_row<FS>(xsrc, u, xr);
#pragma unroll
                    for (int f = 0; f < FS; ++f) x[f] = xr[f];
                }
                xs[warp][lane] = make_float4(x[0], x[1], x[2], x[3]);
                float sv[HEADS] = {};
                if (a.score_in != nullptr && lane < cnt) {       // scores saved by the forward: no recomputation
                    const float* sp = a.score_in + ((size_t)seg * a.st_score + (e0 + lane)) * HEADS;
                    if constexpr (HEADS == 4) {
                        const float4 t4 = __ldg(reinterpret_cast<const float4*>(sp));
                        sv[0] = t4.x; sv[1] = t4.y; sv[2] = t4.z; sv[3] = t4.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < HEADS; ++k) sv[k] = __ldg(sp + k);
                    }
                }
#pragma unroll
                for (int k = 0; k < HEADS; ++k) {
                    const float4 h0 = hc[warp][k][0], h1 = hc[warp][k][1];
                    float s;
                    if (a.score_in != nullptr) {
                        s = sv[k];
                    } else {
                    const float4 pl = *reinterpret_cast<const float4*>(hP + k * 8);
                    s = fmaf(hP[k * 8 + 5], xv1, fmaf(hP[k * 8 + 4], xv0, hP[k * 8 + 6]));
                    s = fmaf(pl.x, x[0], s);
                    if constexpr (FS > 1) s = fmaf(pl.y, x[1], s);
                    if constexpr (FS > 2) s = fmaf(pl.z, x[2], s);
                    if constexpr (FS > 3) s = fmaf(pl.w, x[3], s);
                    const float4* wk = wA + k * D;
                    const float2* ck = cb[warp] + k * D;
#pragma unroll 8
                    for (int d = 0; d < D; ++d) {
                        const float4 w = wk[d];
                        const float2 cc = ck[d];
                        float z = cc.x;
                        z = fmaf(w.x, x[0], z);
                        if constexpr (FS > 1) z = fmaf(w.y, x[1], z);
                        if constexpr (FS > 2) z = fmaf(w.z, x[2], z);
                        if constexpr (FS > 3) z = fmaf(w.w, x[3], z);
                        s = fmaf(cc.y, fabsf(z), s);
                    }
                    }
                    const float alpha = lane < cnt ? __expf(s - h1.z) * h1.w : 0.f;
                    float da = h0.y;
                    da = fmaf(h0.z, x[0], da);
                    if constexpr (FS > 1) da = fmaf(h0.w, x[1], da);
                    if constexpr (FS > 2) da = fmaf(h1.x, x[2], da);
                    if constexpr (FS > 3) da = fmaf(h1.y, x[3], da);
                    dsS[warp][k][lane] = alpha * (da - h0.x);
#pragma unroll
                    for (int f = 0; f < FS; ++f) abar[k][f] = fmaf(alpha, x[f], abar[k][f]);
                }
                __syncwarp();
                // ---- pass 2: lane = CPL channels
                const float* dsh = dsS[warp][head];
#pragma unroll 2
                for (int i = 0; i < cnt; ++i) {
                    const float4 t4 = xs[warp][i];
                    const float xx[4] = {t4.x, t4.y, t4.z, t4.w};
                    const float dsv = dsh[i];
                    const float dsl = dsv * slope;
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        float z = c[j];
#pragma unroll
                        for (int f = 0; f < FS; ++f) z = fmaf(ws[j][f], xx[f], z);
                        const float y = fmaxf(z, slope * z);
                        g_at[j] = fmaf(dsv, y, g_at[j]);
                        const float t = z > 0.f ? dsv : dsl;
#pragma unroll
                        for (int f = 0; f < FS; ++f) tws[j][f] = fmaf(t, xx[f], tws[j][f]);
                        st[j] += t;
                    }
                }
                __syncwarp();
            }
            // sum_e alpha x_e of this lane's head
            float ab[FS];
#pragma unroll
            for (int f = 0; f < FS; ++f) ab[f] = 0.f;
#pragma unroll
            for (int k = 0; k < HEADS; ++k)
#pragma unroll
                for (int f = 0; f < FS; ++f) {
                    const float tot = warp_sum(abar[k][f]);
                    if (k == head) ab[f] = tot;
                }
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                sdz[j] = at[j] * st[j];
#pragma unroll
                for (int f = 0; f < FS; ++f) g_ws[j][f] += fmaf(at[j], tws[j][f], gp[j] * ab[f]);   // score path + message path
                g_bsm[j] += gp[j];                                                                 // sum_e alpha = 1
                g_wd[j][0] = fmaf(sdz[j], xv0, g_wd[j][0]);
                g_wd[j][1] = fmaf(sdz[j], xv1, g_wd[j][1]);
                g_bsd[j] += sdz[j];
            }
        }
        if (a.grad_x_dst != nullptr) {
            float t0 = 0.f, t1 = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                t0 += sdz[j] * wd[j][0] + gp[j] * wr[j][0];
                t1 += sdz[j] * wd[j][1] + gp[j] * wr[j][1];
            }
            t0 = warp_sum(t0); t1 = warp_sum(t1);
            if (lane == 0) {
                float* gd = a.grad_x_dst + seg * a.st_xdst + (size_t)vl * FD;
                gd[0] = t0;
                if (FD > 1) gd[1] = t1;
            }
        }
    }

    // ---- CTA reduction in a fixed warp order, then one partial row per CTA
    const int oWs = 0, obs = H * FS, oWd = obs + H, obd = oWd + H * FD, oat = obd + H, oWr = oat + H,
              obr = oWr + H * FD, P = obr + H;
    for (int w = 0; w < 4; ++w) {
        if (warp == w) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int ch = lane * CPL + j;
                auto put = [&](int idx, float val) { red[idx] = (w == 0) ? val : red[idx] + val; };
#pragma unroll
                for (int f = 0; f < FS; ++f) put(oWs + ch * FS + f, g_ws[j][f]);
                put(obs + ch, g_bsd[j] + g_bsm[j]);
                put(oWd + ch * FD, g_wd[j][0]);
                if (FD > 1) put(oWd + ch * FD + 1, g_wd[j][1]);
                put(obd + ch, g_bsd[j]);
                put(oat + ch, g_at[j]);
                put(oWr + ch * FD, g_wr[j][0]);
                if (FD > 1) put(oWr + ch * FD + 1, g_wr[j][1]);
                put(obr + ch, g_br[j]);
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < P; i += 128) a.partial[(size_t)blockIdx.x * P + i] = red[i];
}

// out[i] = sum_p partial[p*P + i], fixed order: 32 slices of parts per column (8 independent loads in flight per thread:
// a few dependent rounds for ~600 partials), combined in slice order.
constexpr int RED_SLICES = 32;
__global__ void __launch_bounds__(32 * RED_SLICES) reduce_partials_kernel(const float* __restrict__ partial, int nparts, int P,
                                                                          float* __restrict__ out) {
    __shared__ float sm[RED_SLICES][32];
    const int col = blockIdx.x * 32 + threadIdx.x % 32, slice = threadIdx.x / 32;
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    if (col < P) {
        for (int p = slice; p < nparts; p += 8 * RED_SLICES) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (p + u * RED_SLICES < nparts) acc[u] += partial[(size_t)(p + u * RED_SLICES) * P + col];
        }
    }
    sm[slice][threadIdx.x % 32] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    __syncthreads();
    if (slice == 0 && col < P) {
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < RED_SLICES; ++s) t += sm[s][threadIdx.x];
        out[col] = t;
    }
}

static int bwd_grid(int64_t n_dst) {
    int64_t need = (n_dst + 3) / 4;
    int64_t cap = (int64_t)kNumSMs * 4;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

template <int FS, int HEADS>
static int launch_fwd(const GatArgs& a, int64_t n_edges, cudaStream_t st) {
    const int H = HEADS * a.D;
    // lanes per destination from the mean in-degree: 8 (<= 12 edges), 16 (<= 96: deg 80 = 5 full passes), else 32
    const int64_t nd = a.n_dst > 0 ? a.n_dst : 1;
    int gs = n_edges <= 12 * nd ? 8 : (n_edges <= 96 * nd ? 16 : 32);
    while (gs < 32 && nd * gs / 32 < (int64_t)kNumSMs * 8) gs *= 2;   // small launches (act step): parallelism first
    // two edges per lane per pass once a lane has at least two passes of work (halves the shared-memory operand
    // traffic per edge); a launch that cannot fill the SMs keeps the lanes instead
    int epl = (gs < 32 || n_edges >= 64 * nd) && n_edges >= 2 * gs * nd ? 2 : 1;
    if (epl == 2 && gs > 8 && nd * (gs / 2) / 32 >= (int64_t)kNumSMs * 8 && n_edges <= 96 * nd) gs /= 2;
    // A/B overrides (UBS_GAT_GS / _EPL / _HS), read once per process
    static const int env_gs = [] { const char* e = getenv("UBS_GAT_GS"); return e ? atoi(e) : 0; }();
    static const int env_epl = [] { const char* e = getenv("UBS_GAT_EPL"); return e ? atoi(e) : 0; }();
    static const int env_hs = [] { const char* e = getenv("UBS_GAT_HS"); return e ? atoi(e) : 0; }();
    if (env_gs == 8 || env_gs == 16 || env_gs == 32) gs = env_gs;
    if (env_epl == 1 || env_epl == 2) epl = env_epl;
    // head split for launches that cannot fill the SMs even with 32 lanes per destination (the act step)
    int hs = (HEADS % 2 == 0 && gs == 32 && nd < (int64_t)kNumSMs * 32) ? 2 : 1;
    if (env_hs == 1 || (env_hs == 2 && HEADS % 2 == 0 && gs == 32)) hs = env_hs;
    const int gpb = 256 / gs;
    int64_t need = ((int64_t)a.n_dst * hs + gpb - 1) / gpb;
    int64_t cap = (int64_t)kNumSMs * 4;                            // persistent: up to 4 resident CTAs per SM
    const int grid = (int)(need < cap ? (need > 0 ? need : 1) : cap);
    const size_t smem = (size_t)H * 3 * sizeof(float4) + (size_t)gpb * (2 * H + 2) * sizeof(float) + HEADS * 8 * sizeof(float);
    if constexpr (HEADS % 2 == 0) {
        if (hs == 2) {
            if (epl == 2) gatv2_fwd_kernel<FS, HEADS, 32, 2, 2><<<grid, 256, smem, st>>>(a);
            else gatv2_fwd_kernel<FS, HEADS, 32, 1, 2><<<grid, 256, smem, st>>>(a);
            return check_launch("ubs_gatv2_fwd");
        }
    }
    if (epl == 2) {
        if (gs == 8) gatv2_fwd_kernel<FS, HEADS, 8, 2, 1><<<grid, 256, smem, st>>>(a);
        else if (gs == 16) gatv2_fwd_kernel<FS, HEADS, 16, 2, 1><<<grid, 256, smem, st>>>(a);
        else gatv2_fwd_kernel<FS, HEADS, 32, 2, 1><<<grid, 256, smem, st>>>(a);
    } else {
        if (gs == 8) gatv2_fwd_kernel<FS, HEADS, 8, 1, 1><<<grid, 256, smem, st>>>(a);
        else if (gs == 16) gatv2_fwd_kernel<FS, HEADS, 16, 1, 1><<<grid, 256, smem, st>>>(a);
        else gatv2_fwd_kernel<FS, HEADS, 32, 1, 1><<<grid, 256, smem, st>>>(a);
    }
    return check_launch("ubs_gatv2_fwd");
}

template <int FS, int HEADS>
static int launch_bwd(const GatArgs& a, int H, cudaStream_t st) {
    const int grid = bwd_grid(a.n_dst);
    static const bool two_pass = [] { const char* e = getenv("UBS_GAT_BWD2"); return !(e && e[0] == '0'); }();
    if (a.grad_x_src == nullptr && two_pass) {             // observations are leaves: the two-pass kernel
        switch (H / 32) {
            case 1: gatv2_bwd2_kernel<FS, 1, HEADS><<<grid, 128, 0, st>>>(a); break;
            case 2: gatv2_bwd2_kernel<FS, 2, HEADS><<<grid, 128, 0, st>>>(a); break;
            case 4: gatv2_bwd2_kernel<FS, 4, HEADS><<<grid, 128, 0, st>>>(a); break;
            default: set_error("ubs_gatv2_bwd: H=%d unsupported (32, 64, 128)", H); return 2;
        }
        return check_launch("ubs_gatv2_bwd");
    }
    switch (H / 32) {
        case 1: gatv2_bwd_kernel<FS, 1, HEADS><<<grid, 128, 0, st>>>(a); break;
        case 2: gatv2_bwd_kernel<FS, 2, HEADS><<<grid, 128, 0, st>>>(a); break;
        case 4: gatv2_bwd_kernel<FS, 4, HEADS><<<grid, 128, 0, st>>>(a); break;
        default: set_error("ubs_gatv2_bwd: H=%d unsupported (32, 64, 128)", H); return 2;
    }
    return check_launch("ubs_gatv2_bwd");
}

static int check_shape(const char* fn, int F_s, int F_d, int heads, int D, float slope) {
    const int H = heads * D;
    if (F_s < 1 || F_s > 4 || F_d < 1 || F_d > 2) { set_error("%s: fused path needs F_s<=4, F_d<=2 (got %d,%d)", fn, F_s, F_d); return 2; }
    if (!(heads == 1 || heads == 2 || heads == 4 || heads == 8)) { set_error("%s: heads must be 1,2,4,8 (got %d)", fn, heads); return 2; }
    if (!(H == 32 || H == 64 || H == 128)) { set_error("%s: heads*D must be 32, 64 or 128 (got %d)", fn, H); return 2; }
    if (!(slope >= 0.f && slope <= 1.f)) { set_error("%s: negative_slope must be in [0,1] (got %g)", fn, slope); return 2; }
    return 0;
}

}  // namespace ubs

#define UBS_DISPATCH_FS_HEADS(FN, ...)                                                      \
    switch (F_s * 16 + heads) {                                                             \
        case 1 * 16 + 1: rc = FN<1, 1>(__VA_ARGS__); break;                                 \
        case 1 * 16 + 2: rc = FN<1, 2>(__VA_ARGS__); break;                                 \
        case 1 * 16 + 4: rc = FN<1, 4>(__VA_ARGS__); break;                                 \
        case 1 * 16 + 8: rc = FN<1, 8>(__VA_ARGS__); break;                                 \
        case 2 * 16 + 1: rc = FN<2, 1>(__VA_ARGS__); break;                                 \
        case 2 * 16 + 2: rc = FN<2, 2>(__VA_ARGS__); break;                                 \
        case 2 * 16 + 4: rc = FN<2, 4>(__VA_ARGS__); break;                                 \
        case 2 * 16 + 8: rc = FN<2, 8>(__VA_ARGS__); break;                                 \
        case 3 * 16 + 1: rc = FN<3, 1>(__VA_ARGS__); break;                                 \
        case 3 * 16 + 2: rc = FN<3, 2>(__VA_ARGS__); break;                                 \
        case 3 * 16 + 4: rc = FN<3, 4>(__VA_ARGS__); break;                                 \
        case 3 * 16 + 8: rc = FN<3, 8>(__VA_ARGS__); break;                                 \
        case 4 * 16 + 1: rc = FN<4, 1>(__VA_ARGS__); break;                                 \
        case 4 * 16 + 2: rc = FN<4, 2>(__VA_ARGS__); break;                                 \
        case 4 * 16 + 4: rc = FN<4, 4>(__VA_ARGS__); break;                                 \
        case 4 * 16 + 8: rc = FN<4, 8>(__VA_ARGS__); break;                                 \
        default: ubs::set_error("unsupported (F_s, heads)"); rc = 2;                        \
    }

extern "C" UBS_API int ubs_gatv2_fwd(const float* x_src, const float* x_dst, const int32_t* indptr, const int32_t* src_idx,
                             const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                             const float* attn, const float* W_res, const float* b_res,
                             float* out, float* smax, float* ssum, int64_t n_dst, int64_t n_edges,
                             int F_s, int F_d, int heads, int D, float negative_slope, int flags, void* stream) {
    return ubs_gatv2_seg_fwd(x_src, x_dst, indptr, src_idx, W_src, b_src, W_dst, b_dst, attn, W_res, b_res, out, smax,
                             ssum, 1, n_dst, n_edges, 0, 0, 0, 0, heads * D, F_s, F_d, heads, D, negative_slope, flags,
                             stream);
}

extern "C" UBS_API int ubs_gatv2_seg_fwd(const float* x_src, const float* x_dst, const int32_t* indptr,
                                 const int32_t* src_idx, const float* W_src, const float* b_src, const float* W_dst,
                                 const float* b_dst, const float* attn, const float* W_res, const float* b_res,
                                 float* out, float* smax, float* ssum, int64_t n_seg, int64_t n_dst_seg,
                                 int64_t n_edges, int64_t st_xsrc, int64_t st_xdst, int64_t st_ip, int64_t st_sidx,
                                 int64_t ld_out, int F_s, int F_d, int heads, int D, float negative_slope, int flags,
                                 void* stream) {
    return ubs_gatv2_seg_fwd_scores(x_src, x_dst, indptr, src_idx, W_src, b_src, W_dst, b_dst, attn, W_res, b_res, out, smax,
                                    ssum, nullptr, 0, n_seg, n_dst_seg, n_edges, st_xsrc, st_xdst, st_ip, st_sidx, ld_out, F_s,
                                    F_d, heads, D, negative_slope, flags, stream);
}

extern "C" UBS_API int ubs_gatv2_seg_fwd_scores(const float* x_src, const float* x_dst, const int32_t* indptr,
                                 const int32_t* src_idx, const float* W_src, const float* b_src, const float* W_dst,
                                 const float* b_dst, const float* attn, const float* W_res, const float* b_res,
                                 float* out, float* smax, float* ssum, float* scores, int64_t st_scores, int64_t n_seg,
                                 int64_t n_dst_seg, int64_t n_edges, int64_t st_xsrc, int64_t st_xdst, int64_t st_ip,
                                 int64_t st_sidx, int64_t ld_out, int F_s, int F_d, int heads, int D, float negative_slope,
                                 int flags, void* stream) {
    const int64_t n_dst = n_seg * n_dst_seg;
    if (int rc = ubs::check_shape("ubs_gatv2_fwd", F_s, F_d, heads, D, negative_slope)) return rc;
    UBS_REQUIRE(scores == nullptr || (((uintptr_t)scores % 16) == 0 && st_scores >= 0), "ubs_gatv2_fwd: scores must be 16-byte aligned");
    UBS_REQUIRE(n_seg >= 1 && n_dst_seg >= 0 && ld_out >= heads * D, "ubs_gatv2_fwd: bad segment description");
    UBS_REQUIRE(F_s != 4 || st_xsrc % 4 == 0, "ubs_gatv2_fwd: segment stride breaks 16-byte row alignment");
    UBS_REQUIRE(F_s != 2 || st_xsrc % 2 == 0, "ubs_gatv2_fwd: segment stride breaks 8-byte row alignment");
    UBS_REQUIRE(n_dst >= 0 && n_dst < (1ll << 31) && n_edges >= 0 && n_edges < (1ll << 31), "ubs_gatv2_fwd: sizes out of range");
    UBS_REQUIRE((smax == nullptr) == (ssum == nullptr), "ubs_gatv2_fwd: smax and ssum must both be given or both NULL");
    UBS_REQUIRE(!(flags & UBS_GAT_RESIDUAL) || W_res != nullptr, "ubs_gatv2_fwd: residual flag without W_res");
    UBS_REQUIRE(((uintptr_t)x_src % (F_s == 4 ? 16 : F_s == 2 ? 8 : 4)) == 0, "ubs_gatv2_fwd: x_src misaligned");
    if (n_dst == 0) return 0;
    ubs::GatArgs a{};
    a.x_src = x_src; a.x_dst = x_dst; a.indptr = indptr; a.src_idx = src_idx;
    a.W_src = W_src; a.b_src = b_src; a.W_dst = W_dst; a.b_dst = b_dst; a.attn = attn; a.W_res = W_res; a.b_res = b_res;
    a.out = out; a.smax = smax; a.ssum = ssum;
    a.n_dst = (int)n_dst; a.F_d = F_d; a.D = D; a.slope = negative_slope; a.flags = flags;
    a.n_dst_seg = (int)(n_dst_seg > 0 ? n_dst_seg : 1); a.st_xsrc = st_xsrc; a.st_xdst = st_xdst; a.st_ip = st_ip;
    a.st_sidx = st_sidx; a.ld_out = (int)ld_out; a.ld_gout = (int)ld_out;
    a.score = scores; a.st_score = st_scores;
    int rc = 0;
    UBS_DISPATCH_FS_HEADS(ubs::launch_fwd, a, n_edges, (cudaStream_t)stream)
    return rc;
}

extern "C" UBS_API int64_t ubs_gatv2_bwd_workspace(int64_t n_dst, int F_s, int F_d, int heads, int D) {
    return (int64_t)ubs::bwd_grid(n_dst) * ubs::gat_param_count(heads * D, F_s, F_d);
}

extern "C" UBS_API int ubs_gatv2_bwd(const float* x_src, const float* x_dst, const int32_t* indptr, const int32_t* src_idx,
                             const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                             const float* attn, const float* W_res, const float* b_res,
                             const float* out, const float* grad_out, const float* smax, const float* ssum,
                             float* grad_params, float* grad_x_src, float* grad_x_dst, float* workspace,
                             int64_t n_dst, int64_t n_edges, int64_t n_src, int F_s, int F_d, int heads, int D,
                             float negative_slope, int flags, void* stream) {
    (void)n_src;
    return ubs_gatv2_seg_bwd(x_src, x_dst, indptr, src_idx, W_src, b_src, W_dst, b_dst, attn, W_res, b_res, out,
                             grad_out, smax, ssum, grad_params, grad_x_src, grad_x_dst, workspace, 1, n_dst, n_edges,
                             0, 0, 0, 0, heads * D, heads * D, F_s, F_d, heads, D, negative_slope, flags, stream);
}

extern "C" UBS_API int ubs_gatv2_seg_bwd(const float* x_src, const float* x_dst, const int32_t* indptr,
                                 const int32_t* src_idx, const float* W_src, const float* b_src, const float* W_dst,
                                 const float* b_dst, const float* attn, const float* W_res, const float* b_res,
                                 const float* out, const float* grad_out, const float* smax, const float* ssum,
                                 float* grad_params, float* grad_x_src, float* grad_x_dst, float* workspace,
                                 int64_t n_seg, int64_t n_dst_seg, int64_t n_edges, int64_t st_xsrc, int64_t st_xdst,
                                 int64_t st_ip, int64_t st_sidx, int64_t ld_out, int64_t ld_gout, int F_s, int F_d,
                                 int heads, int D, float negative_slope, int flags, void* stream) {
    return ubs_gatv2_seg_bwd_scores(x_src, x_dst, indptr, src_idx, W_src, b_src, W_dst, b_dst, attn, W_res, b_res, out, grad_out,
                                    smax, ssum, nullptr, 0, grad_params, grad_x_src, grad_x_dst, workspace, n_seg, n_dst_seg,
                                    n_edges, st_xsrc, st_xdst, st_ip, st_sidx, ld_out, ld_gout, F_s, F_d, heads, D,
                                    negative_slope, flags, stream);
}

extern "C" UBS_API int ubs_gatv2_seg_bwd_scores(const float* x_src, const float* x_dst, const int32_t* indptr,
                                 const int32_t* src_idx, const float* W_src, const float* b_src, const float* W_dst,
                                 const float* b_dst, const float* attn, const float* W_res, const float* b_res,
                                 const float* out, const float* grad_out, const float* smax, const float* ssum,
                                 const float* scores, int64_t st_scores,
                                 float* grad_params, float* grad_x_src, float* grad_x_dst, float* workspace,
                                 int64_t n_seg, int64_t n_dst_seg, int64_t n_edges, int64_t st_xsrc, int64_t st_xdst,
                                 int64_t st_ip, int64_t st_sidx, int64_t ld_out, int64_t ld_gout, int F_s, int F_d,
                                 int heads, int D, float negative_slope, int flags, void* stream) {
    const int64_t n_dst = n_seg * n_dst_seg;
    if (int rc = ubs::check_shape("ubs_gatv2_bwd", F_s, F_d, heads, D, negative_slope)) return rc;
    UBS_REQUIRE(scores == nullptr || ((uintptr_t)scores % 16) == 0, "ubs_gatv2_bwd: scores must be 16-byte aligned");
    UBS_REQUIRE(n_seg >= 1 && n_dst_seg >= 0 && ld_out >= heads * D && ld_gout >= heads * D, "ubs_gatv2_bwd: bad segment description");
    UBS_REQUIRE(F_s != 4 || st_xsrc % 4 == 0, "ubs_gatv2_bwd: segment stride breaks 16-byte row alignment");
    UBS_REQUIRE(F_s != 2 || st_xsrc % 2 == 0, "ubs_gatv2_bwd: segment stride breaks 8-byte row alignment");
    UBS_REQUIRE(n_dst >= 0 && n_dst < (1ll << 31) && n_edges >= 0 && n_edges < (1ll << 31), "ubs_gatv2_bwd: sizes out of range");
    UBS_REQUIRE(smax && ssum && out && grad_out && grad_params && workspace, "ubs_gatv2_bwd: NULL argument");
    UBS_REQUIRE(((uintptr_t)x_src % (F_s == 4 ? 16 : F_s == 2 ? 8 : 4)) == 0, "ubs_gatv2_bwd: x_src misaligned");
    const int H = heads * D;
    UBS_REQUIRE(32 % heads == 0 && (H / heads) % (H / 32) == 0, "ubs_gatv2_bwd: head layout unsupported");
    const int P = ubs::gat_param_count(H, F_s, F_d);
    cudaStream_t st = (cudaStream_t)stream;
    if (n_dst == 0) {
        cudaMemsetAsync(grad_params, 0, sizeof(float) * P, st);
        return 0;
    }
    ubs::GatArgs a{};
    a.x_src = x_src; a.x_dst = x_dst; a.indptr = indptr; a.src_idx = src_idx;
    a.W_src = W_src; a.b_src = b_src; a.W_dst = W_dst; a.b_dst = b_dst; a.attn = attn; a.W_res = W_res; a.b_res = b_res;
    a.grad_out = grad_out; a.out_in = out; a.smax_in = smax; a.ssum_in = ssum;
    a.partial = workspace; a.grad_x_src = grad_x_src; a.grad_x_dst = grad_x_dst;
    a.n_dst = (int)n_dst; a.F_d = F_d; a.D = D; a.slope = negative_slope; a.flags = flags;
    a.n_dst_seg = (int)(n_dst_seg > 0 ? n_dst_seg : 1); a.st_xsrc = st_xsrc; a.st_xdst = st_xdst; a.st_ip = st_ip;
    a.st_sidx = st_sidx; a.ld_out = (int)ld_out; a.ld_gout = (int)ld_gout;
    a.score_in = grad_x_src == nullptr ? scores : nullptr; a.st_score = st_scores;
    int rc = 0;
    UBS_DISPATCH_FS_HEADS(ubs::launch_bwd, a, H, st)
    if (rc) return rc;
    ubs::reduce_partials_kernel<<<(P + 31) / 32, 32 * ubs::RED_SLICES, 0, st>>>(workspace, ubs::bwd_grid(n_dst), P, grad_params);
    return ubs::check_launch("ubs_gatv2_bwd(reduce)");
}
