// General-CSR GATv2 attention + aggregation on PRE-PROJECTED features (wide inputs: the synthetic sweep, F_in = H >= 32).
//
// dglnn.GATv2Conv (reference call sites algos/madrqn/agents/gnn_agents.py:92-97; DGL 0.9.0 forward)
// after the three dense projections el = fc_src(h_src), er = fc_dst(h_dst), res = res_fc(h_dst) (library GEMMs):
//     e = leaky_relu(el[u] + er[v]) ; s = <attn_k, e_k> ; alpha = edge_softmax_by_dst(s) ;
//     out[v] = act( sum_e alpha * el[u] + res[v] )
// i.e. DGL's gSDDMM(u_add_v) + leaky_relu + reduce + 5 edge_softmax kernels + gSpMM(u_mul_e, sum) (7 passes over an
// E x H edge tensor) as ONE gather pass: a warp owns a destination, lanes own CPL = H/32 contiguous channels, the
// source rows el[u] are gathered with coalesced 16-byte loads (a full row per warp request), the per-head score is a
// segmented shuffle reduction over the lanes of the head, softmax is online, nothing of size E x H is ever written.
// This is the HBM/L2-bound kernel of the path: algorithmic bytes = 4 (E H [gather] + E + N+1 + 3 N H) forward.
// The backward re-gathers el, recomputes the scores from the saved (max, sum) statistics and scatters grad_el with
// 16-byte vector atomics (red.global.add.v4.f32, sm_90+); star layouts (one edge per source) store instead.
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {
namespace aggr {

struct Args {
    const float* el; const float* er; const float* res; const int* indptr; const int* src_idx; const float* attn;
    float* out; float* smax; float* ssum;
    const float* grad_out; const float* out_in; const float* smax_in; const float* ssum_in;
    float* grad_el; float* grad_er; float* grad_res; float* partial;
    int n_dst; float slope; int flags;
};

template <int CPL>
__device__ __forceinline__ void load_chunk(const float* __restrict__ p, float (&v)[CPL]) {
    if constexpr (CPL % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CPL / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else if constexpr (CPL == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) v[i] = __ldg(p + i);
    }
}
template <int CPL>
__device__ __forceinline__ void store_chunk(float* p, const float (&v)[CPL]) {
    if constexpr (CPL % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CPL / 4; ++i)
            reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if constexpr (CPL == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) p[i] = v[i];
    }
}
template <int CPL>
__device__ __forceinline__ void atomic_add_chunk(float* p, const float (&v)[CPL]) {
    if constexpr (CPL % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CPL / 4; ++i)
            atomicAdd(reinterpret_cast<float4*>(p) + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
    } else if constexpr (CPL == 2) {
        atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) atomicAdd(p + i, v[i]);
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <int CPL, int HEADS>
__global__ void __launch_bounds__(256) gat_aggr_fwd_kernel(const Args a) {
    constexpr int H = 32 * CPL, LPH = 32 / HEADS;
    const int lane = threadIdx.x % 32;
    const int head = lane / LPH;
    const int c0 = lane * CPL;
    const bool relu = a.flags & UBS_GAT_RELU;
    const float slope = a.slope;
    float at[CPL];
    load_chunk<CPL>(a.attn + c0, at);
    const int total_warps = gridDim.x * 8;
    for (int v = blockIdx.x * 8 + threadIdx.x / 32; v < a.n_dst; v += total_warps) {
        const int beg = __ldg(a.indptr + v), end = __ldg(a.indptr + v + 1);
        float er[CPL], acc[CPL];
        load_chunk<CPL>(a.er + (size_t)v * H + c0, er);
#pragma unroll
        for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
        float m = -CUDART_INF_F, l = 0.f;
        for (int e0 = beg; e0 < end; e0 += 32) {
            const int cnt = min(32, end - e0);
            int u_lane = e0 + lane;                                   // star layout: source id == CSR slot
            if (a.src_idx != nullptr && lane < cnt) u_lane = __ldg(a.src_idx + e0 + lane);
            float nxt[CPL];
            load_chunk<CPL>(a.el + (size_t)__shfl_sync(0xffffffffu, u_lane, 0) * H + c0, nxt);
            for (int i = 0; i < cnt; ++i) {
                float x[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) x[j] = nxt[j];
                if (i + 1 < cnt)                                        // prefetch the next source row
                    load_chunk<CPL>(a.el + (size_t)__shfl_sync(0xffffffffu, u_lane, i + 1) * H + c0, nxt);
                float sp = 0.f;
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const float z = x[j] + er[j];
                    sp = fmaf(at[j], fmaxf(z, slope * z), sp);
                }
                const float s = group_sum<LPH>(sp);
                const float mn = fmaxf(m, s);
                const float sc = __expf(m - mn), p = __expf(s - mn);
                l = fmaf(l, sc, p);
#pragma unroll
                for (int j = 0; j < CPL; ++j) acc[j] = fmaf(acc[j], sc, p * x[j]);
                m = mn;
            }
        }
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        float o[CPL];
        if (a.res != nullptr) load_chunk<CPL>(a.res + (size_t)v * H + c0, o);
        else {
#pragma unroll
            for (int j = 0; j < CPL; ++j) o[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            o[j] = fmaf(acc[j], inv, o[j]);
            if (relu) o[j] = fmaxf(o[j], 0.f);
        }
        store_chunk<CPL>(a.out + (size_t)v * H + c0, o);
        if (a.smax != nullptr && lane % LPH == 0) {
            a.smax[(size_t)v * HEADS + head] = l > 0.f ? m : 0.f;
            a.ssum[(size_t)v * HEADS + head] = l;
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
template <int CPL, int HEADS>
__global__ void __launch_bounds__(256) gat_aggr_bwd_kernel(const Args a) {
    constexpr int H = 32 * CPL, LPH = 32 / HEADS;
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int head = lane / LPH;
    const int c0 = lane * CPL;
    const bool relu = a.flags & UBS_GAT_RELU;
    const float slope = a.slope;
    __shared__ float red[8][H];
    float at[CPL], g_at[CPL];
    load_chunk<CPL>(a.attn + c0, at);
#pragma unroll
    for (int j = 0; j < CPL; ++j) g_at[j] = 0.f;
    const int total_warps = gridDim.x * 8;
    for (int v = blockIdx.x * 8 + warp; v < a.n_dst; v += total_warps) {
        const int beg = __ldg(a.indptr + v), end = __ldg(a.indptr + v + 1);
        float gp[CPL], oo[CPL], er[CPL], ger[CPL];
        load_chunk<CPL>(a.grad_out + (size_t)v * H + c0, gp);
        load_chunk<CPL>(a.out_in + (size_t)v * H + c0, oo);
        load_chunk<CPL>(a.er + (size_t)v * H + c0, er);
        float dotp = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            if (relu && !(oo[j] > 0.f)) gp[j] = 0.f;
            ger[j] = 0.f;
        }
        if (a.grad_res != nullptr) store_chunk<CPL>(a.grad_res + (size_t)v * H + c0, gp);
        if (end > beg) {
            // <g'_k, ft_k> with ft = out - res wherever g' != 0
            float rs[CPL];
            if (a.res != nullptr) load_chunk<CPL>(a.res + (size_t)v * H + c0, rs);
            else {
#pragma unroll
                for (int j = 0; j < CPL; ++j) rs[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) dotp = fmaf(gp[j], oo[j] - rs[j], dotp);
            dotp = group_sum<LPH>(dotp);
            const float Mx = __ldg(a.smax_in + (size_t)v * HEADS + head);
            const float invL = 1.0f / __ldg(a.ssum_in + (size_t)v * HEADS + head);
            for (int e0 = beg; e0 < end; e0 += 32) {
                const int cnt = min(32, end - e0);
                int u_lane = e0 + lane;
                if (a.src_idx != nullptr && lane < cnt) u_lane = __ldg(a.src_idx + e0 + lane);
                float nxt[CPL];
                load_chunk<CPL>(a.el + (size_t)__shfl_sync(0xffffffffu, u_lane, 0) * H + c0, nxt);
                for (int i = 0; i < cnt; ++i) {
                    const int u = __shfl_sync(0xffffffffu, u_lane, i);
                    float x[CPL];
#pragma unroll
                    for (int j = 0; j < CPL; ++j) x[j] = nxt[j];
                    if (i + 1 < cnt)
                        load_chunk<CPL>(a.el + (size_t)__shfl_sync(0xffffffffu, u_lane, i + 1) * H + c0, nxt);
                    float z[CPL], y[CPL], sp = 0.f, dap = 0.f;
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        z[j] = x[j] + er[j];
                        y[j] = fmaxf(z[j], slope * z[j]);
                        sp = fmaf(at[j], y[j], sp);
                        dap = fmaf(gp[j], x[j], dap);
                    }
                    const float s = group_sum<LPH>(sp);
                    const float da = group_sum<LPH>(dap);
                    const float alpha = __expf(s - Mx) * invL;
                    const float ds = alpha * (da - dotp);
                    float gel[CPL];
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        g_at[j] = fmaf(ds, y[j], g_at[j]);
                        const float dz = ds * at[j] * (z[j] > 0.f ? 1.0f : slope);
                        ger[j] += dz;
                        gel[j] = fmaf(alpha, gp[j], dz);
                    }
                    float* dst = a.grad_el + (size_t)u * H + c0;
                    if (a.src_idx != nullptr) atomic_add_chunk<CPL>(dst, gel);
                    else store_chunk<CPL>(dst, gel);
                }
            }
        }
        store_chunk<CPL>(a.grad_er + (size_t)v * H + c0, ger);
    }
    // grad_attn: fixed-order CTA reduction, one partial row per CTA
#pragma unroll
    for (int j = 0; j < CPL; ++j) red[warp][c0 + j] = g_at[j];
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += 256) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][c];
        a.partial[(size_t)blockIdx.x * H + c] = t;
    }
}

__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ partial, int nparts, int P,
                                                          float* __restrict__ out) {
    __shared__ float sm[8][32];
    const int col = blockIdx.x * 32 + threadIdx.x % 32, slice = threadIdx.x / 32;
    float acc = 0.f;
    if (col < P)
        for (int p = slice; p < nparts; p += 8) acc += partial[(size_t)p * P + col];
    sm[slice][threadIdx.x % 32] = acc;
    __syncthreads();
    if (slice == 0 && col < P) {
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) t += sm[s][threadIdx.x];
        out[col] = t;
    }
}

static int grid_for(int64_t n_dst) {
    const int64_t need = (n_dst + 7) / 8, cap = (int64_t)kNumSMs * 8;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}
static int check(const char* fn, int heads, int D, float slope) {
    const int H = heads * D;
    if (!(H == 32 || H == 64 || H == 128 || H == 256)) { set_error("%s: heads*D must be 32, 64, 128 or 256 (got %d)", fn, H); return 2; }
    if (!(heads == 1 || heads == 2 || heads == 4 || heads == 8) || D % (H / 32) != 0) { set_error("%s: head layout unsupported (heads=%d, D=%d)", fn, heads, D); return 2; }
    if (!(slope >= 0.f && slope <= 1.f)) { set_error("%s: negative_slope must be in [0,1]", fn); return 2; }
    return 0;
}

}  // namespace aggr
}  // namespace ubs

#define UBS_AGGR_DISPATCH(KERNEL, grid, st, a)                                                       \
    switch ((H / 32) * 16 + heads) {                                                                 \
        case 1 * 16 + 1: ubs::aggr::KERNEL<1, 1><<<grid, 256, 0, st>>>(a); break;                    \
        case 1 * 16 + 2: ubs::aggr::KERNEL<1, 2><<<grid, 256, 0, st>>>(a); break;                    \
        case 1 * 16 + 4: ubs::aggr::KERNEL<1, 4><<<grid, 256, 0, st>>>(a); break;                    \
        case 1 * 16 + 8: ubs::aggr::KERNEL<1, 8><<<grid, 256, 0, st>>>(a); break;                    \
        case 2 * 16 + 1: ubs::aggr::KERNEL<2, 1><<<grid, 256, 0, st>>>(a); break;                    \
        case 2 * 16 + 2: ubs::aggr::KERNEL<2, 2><<<grid, 256, 0, st>>>(a); break;                    \
        case 2 * 16 + 4: ubs::aggr::KERNEL<2, 4><<<grid, 256, 0, st>>>(a); break;                    \
        case 2 * 16 + 8: ubs::aggr::KERNEL<2, 8><<<grid, 256, 0, st>>>(a); break;                    \
        case 4 * 16 + 1: ubs::aggr::KERNEL<4, 1><<<grid, 256, 0, st>>>(a); break;                    \
        case 4 * 16 + 2: ubs::aggr::KERNEL<4, 2><<<grid, 256, 0, st>>>(a); break;                    \
        case 4 * 16 + 4: ubs::aggr::KERNEL<4, 4><<<grid, 256, 0, st>>>(a); break;                    \
        case 4 * 16 + 8: ubs::aggr::KERNEL<4, 8><<<grid, 256, 0, st>>>(a); break;                    \
        case 8 * 16 + 1: ubs::aggr::KERNEL<8, 1><<<grid, 256, 0, st>>>(a); break;                    \
        case 8 * 16 + 2: ubs::aggr::KERNEL<8, 2><<<grid, 256, 0, st>>>(a); break;                    \
        case 8 * 16 + 4: ubs::aggr::KERNEL<8, 4><<<grid, 256, 0, st>>>(a); break;                    \
        case 8 * 16 + 8: ubs::aggr::KERNEL<8, 8><<<grid, 256, 0, st>>>(a); break;                    \
        default: ubs::set_error("unsupported (H, heads)"); return 2;                                  \
    }

extern "C" UBS_API int ubs_gat_aggr_fwd(const float* el, const float* er, const float* res, const int32_t* indptr,
                                        const int32_t* src_idx, const float* attn, float* out, float* smax, float* ssum,
                                        int64_t n_dst, int64_t n_edges, int heads, int D, float negative_slope, int flags,
                                        void* stream) {
    (void)n_edges;
    if (int rc = ubs::aggr::check("ubs_gat_aggr_fwd", heads, D, negative_slope)) return rc;
    UBS_REQUIRE(el && er && indptr && attn && out, "ubs_gat_aggr_fwd: NULL argument");
    UBS_REQUIRE((smax == nullptr) == (ssum == nullptr), "ubs_gat_aggr_fwd: smax and ssum go together");
    UBS_REQUIRE(n_dst >= 0 && n_dst < (1ll << 31), "ubs_gat_aggr_fwd: n_dst out of range");
    if (n_dst == 0) return 0;
    const int H = heads * D;
    ubs::aggr::Args a{};
    a.el = el; a.er = er; a.res = res; a.indptr = indptr; a.src_idx = src_idx; a.attn = attn;
    a.out = out; a.smax = smax; a.ssum = ssum; a.n_dst = (int)n_dst; a.slope = negative_slope; a.flags = flags;
    const int grid = ubs::aggr::grid_for(n_dst);
    cudaStream_t st = (cudaStream_t)stream;
    UBS_AGGR_DISPATCH(gat_aggr_fwd_kernel, grid, st, a)
    return ubs::check_launch("ubs_gat_aggr_fwd");
}

extern "C" UBS_API int64_t ubs_gat_aggr_bwd_workspace(int64_t n_dst, int heads, int D) {
    return (int64_t)ubs::aggr::grid_for(n_dst) * heads * D;
}

extern "C" UBS_API int ubs_gat_aggr_bwd(const float* el, const float* er, const float* res, const int32_t* indptr,
                                        const int32_t* src_idx, const float* attn, const float* out,
                                        const float* grad_out, const float* smax, const float* ssum, float* grad_el,
                                        float* grad_er, float* grad_res, float* grad_attn, float* workspace,
                                        int64_t n_dst, int64_t n_edges, int heads, int D, float negative_slope,
                                        int flags, void* stream) {
    (void)n_edges;
    if (int rc = ubs::aggr::check("ubs_gat_aggr_bwd", heads, D, negative_slope)) return rc;
    UBS_REQUIRE(el && er && indptr && attn && out && grad_out && smax && ssum && grad_el && grad_er && grad_attn && workspace,
                "ubs_gat_aggr_bwd: NULL argument");
    const int H = heads * D;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_dst == 0) { cudaMemsetAsync(grad_attn, 0, sizeof(float) * H, st); return 0; }
    ubs::aggr::Args a{};
    a.el = el; a.er = er; a.res = res; a.indptr = indptr; a.src_idx = src_idx; a.attn = attn;
    a.grad_out = grad_out; a.out_in = out; a.smax_in = smax; a.ssum_in = ssum;
    a.grad_el = grad_el; a.grad_er = grad_er; a.grad_res = grad_res; a.partial = workspace;
    a.n_dst = (int)n_dst; a.slope = negative_slope; a.flags = flags;
    const int grid = ubs::aggr::grid_for(n_dst);
    UBS_AGGR_DISPATCH(gat_aggr_bwd_kernel, grid, st, a)
    if (int rc = ubs::check_launch("ubs_gat_aggr_bwd")) return rc;
    ubs::aggr::reduce_rows_kernel<<<(H + 31) / 32, 256, 0, st>>>(workspace, grid, H, grad_attn);
    return ubs::check_launch("ubs_gat_aggr_bwd(reduce)");
}
