// Act-step kernel with TMA-staged weights (inference, H <= 64 class configs).
//
// Same math as agent_step_fwd_kernel (reference GnnAgent.forward after the relations + learner.act's epsilon-greedy,
// algos/madrqn/agents/gnn_agents.py:51-56, algos/madrqn/learner.py:75-78), different data movement.  One act step
// needs ALL 198 KB of weights once per CTA (16 agent rows each): streaming them through registers exposes the L2
// latency at the start of each of the five dependent layers.  Here one elected thread issues 1-D bulk async copies
// (cp.async.bulk -> the TMA engine, completion on an mbarrier) of the NEXT layer's packed weight block into one of two
// shared-memory buffers while the current layer's GEMM runs out of the other, so every GEMM reads its weights from
// shared memory and the copy engine, not the warps, waits for L2:
//     layer i   :  wait(mbar[i & 1])  ->  GEMM from buf[i & 1]
//     meanwhile :  bulk copy of layer i+1 into buf[(i+1) & 1]   (free since layer i-1 finished)
// 16 warps per CTA (512 threads), thread tile 4x4, split-K for the narrow layers.
//
// Fused observation relations (ubs_agent_act_rel_fwd): the two GATv2 star relations of GraphObservationEncoder
// (gnn_agents.py:92-97,103-104 -> dglnn.GATv2Conv) run INSIDE this kernel instead of as two launches in front of it.
// A CTA owns 16 consecutive agent rows; in the star layout their source rows are ONE contiguous range of the
// observation packet (rows [indptr[row0], indptr[row0 + 16])), so the neighbour lists of the tile are staged with one
// bulk async copy (TMA) per relation into the shared-memory buffer of the second weight layer — idle until the
// aggregator GEMM starts — next to the per-relation weight tables that ubs_gatv2_rel_pack precomputes once per
// parameter update.  One warp per destination row: lanes own edges (online softmax per lane, merged by shuffles),
// the result lands in the feature-major input tile of the aggregator GEMM: no xin round trip through HBM, no
// fork / join of relation streams, one launch per vector-step.
#include "agent_step.cuh"

namespace ubs {
namespace act {
#ifdef UBS_ACT_TRACE
__device__ long long ubs_act_trace[16];
#define ATR(slot) do { if (blockIdx.x == 5 && threadIdx.x == 0) ubs_act_trace[slot] = clock64(); } while (0)
#else
#define ATR(slot) do { } while (0)
#endif

constexpr int NT = 512;
constexpr int RP = 24;     // row stride of the feature-major activation tiles in THIS kernel: conflict-free mma A fragments

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ACT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ACT_DONE;\n"
        "bra ACT_WAIT;\n"
        "ACT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// one thread: arm the barrier with the byte count, then hand the copy to the TMA engine
__device__ __forceinline__ void bulk_load(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------- fused relations
// Tables of one relation (ubs_gatv2_rel_pack): wA[H] = W_src row (zero padded to 4), wR[H] = {W_res[0], W_res[1],
// b_res, b_src}, wD[H] = {W_dst[0], W_dst[1], b_src + b_dst, (1-s)/2 attn}, hP[heads][8] = (1+s)/2 sum_d attn *
// {W_src[0..3] | W_dst[0..1], b_src + b_dst, 0}: the same constants gatv2_fwd_kernel derives in its prologue.
__host__ __device__ inline int rel_table_floats(int H, int heads) { return 12 * (H + heads) + 8 * heads; }   // one pad float4 per head

struct RelSmem { int tabs, xs[2], cbuf, ips, xd, total; };     // float offsets inside the staging region
__host__ __device__ inline RelSmem rel_smem(const StepDims& d, const RelIn& r) {
    RelSmem s;
    int o = 0;
    auto take = [&](int n) { int p = o; o += (n + 3) & ~3; return p; };
    s.tabs = take(2 * rel_table_floats(d.H, r.heads));
    s.xs[0] = take(R * r.cap[0] * r.FS[0] + 4);
    s.xs[1] = take(R * r.cap[1] * r.FS[1] + 4);
    s.cbuf = take((NT / 32) * 2 * (d.H + r.heads));
    s.ips = take(2 * (R + 1));
    s.xd = take(2 * R);
    s.total = o;
    return s;
}

template <int FS>
__device__ __forceinline__ void lds_row(const float* p, float (&x)[FS]) {
    if constexpr (FS == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else if constexpr (FS == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        x[0] = t.x; x[1] = t.y;
    } else {
#pragma unroll
        for (int f = 0; f < FS; ++f) x[f] = p[f];
    }
}

// One warp, one destination: GATv2 over the source rows xs[beg .. end) (shared memory), result -> out[ch * RP].
// Same arithmetic as gatv2_fwd_kernel.  Heads are independent (own softmax, own output channels), so the warp is cut
// into HEADS groups of LPH = 32 / HEADS lanes: a group owns one head and walks the edges two per lane per pass
// (80 edges = 5 full passes of an 8-lane group; no idle lanes at the BASELINE degree).  The tables are padded by one
// float4 per head so that the groups' weight reads fall into different banks.
#ifndef UBS_ACT_EPL
#define UBS_ACT_EPL 5
#endif
template <int FS, int HEADS>
__device__ __noinline__ void gat_rel_row(const float* xs, int beg, int end, const float* tab, int D, float xv0, float xv1,
                                         int flags, float2* cbuf, float* out) {
    // EPL edges per lane per pass: the weight / destination operands of a channel (one LDS.128 + one LDS.64) are
    // reused for all of them; 5 makes the BASELINE degree (80 edges over an 8-lane head group) exactly two passes
    constexpr int LPH = 32 / HEADS, EPL = UBS_ACT_EPL;
    const int H = HEADS * D, Hp = H + HEADS, Dp = D + 1, lane = threadIdx.x & 31;
    const int k = lane / LPH, li = lane % LPH;
    const float4* wA = reinterpret_cast<const float4*>(tab);
    const float4* wR = wA + Hp;
    const float4* wD = wR + Hp;
    const float* hP = reinterpret_cast<const float*>(wD + Hp);
    const bool relu = flags & UBS_GAT_RELU, has_res = flags & UBS_GAT_RESIDUAL;
    for (int ch = lane; ch < H; ch += 32) {
        const int p = ch + ch / D;
        const float4 w = wD[p];
        cbuf[p] = make_float2(fmaf(w.y, xv1, fmaf(w.x, xv0, w.z)), w.w);
    }
    __syncwarp();
    const float4* wk = wA + k * Dp;
    const float2* ck = cbuf + k * Dp;
    const float4 pk = *reinterpret_cast<const float4*>(hP + k * 8);
    const float lin = fmaf(hP[k * 8 + 5], xv1, fmaf(hP[k * 8 + 4], xv0, hP[k * 8 + 6]));
    float m = -CUDART_INF_F, l = 0.f, acc[FS];
#pragma unroll
    for (int f = 0; f < FS; ++f) acc[f] = 0.f;
    for (int e = beg + li; e < end; e += EPL * LPH) {
        float x[EPL][FS];
        bool ok[EPL];
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            const int ej = e + j * LPH;
            ok[j] = ej < end;
            if (ok[j]) lds_row<FS>(xs + ej * FS, x[j]);
            else {
#pragma unroll
                for (int f = 0; f < FS; ++f) x[j][f] = 0.f;
            }
        }
        float sc2[EPL];
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            sc2[j] = fmaf(pk.x, x[j][0], lin);
            if constexpr (FS > 1) sc2[j] = fmaf(pk.y, x[j][1], sc2[j]);
            if constexpr (FS > 2) sc2[j] = fmaf(pk.z, x[j][2], sc2[j]);
            if constexpr (FS > 3) sc2[j] = fmaf(pk.w, x[j][3], sc2[j]);
        }
#pragma unroll 8
        for (int dd = 0; dd < D; ++dd) {
            const float4 w = wk[dd];
            const float2 c = ck[dd];
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                float z = c.x;
                z = fmaf(w.x, x[j][0], z);
                if constexpr (FS > 1) z = fmaf(w.y, x[j][1], z);
                if constexpr (FS > 2) z = fmaf(w.z, x[j][2], z);
                if constexpr (FS > 3) z = fmaf(w.w, x[j][3], z);
                sc2[j] = fmaf(c.y, fabsf(z), sc2[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            if (ok[j]) {
                const float mn = fmaxf(m, sc2[j]);
                const float sc = __expf(m - mn);
                const float p = __expf(sc2[j] - mn);
                l = fmaf(l, sc, p);
#pragma unroll
                for (int f = 0; f < FS; ++f) acc[f] = fmaf(acc[f], sc, p * x[j][f]);
                m = mn;
            }
        }
    }
    {
        const float Mx = group_max<LPH>(m);
        const float sc = (m == -CUDART_INF_F) ? 0.f : __expf(m - Mx);
        l = group_sum<LPH>(l * sc);
#pragma unroll
        for (int f = 0; f < FS; ++f) acc[f] = group_sum<LPH>(acc[f] * sc);
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        for (int dd = li; dd < D; dd += LPH) {
            const float4 w = wk[dd];
            const float4 r = wR[k * Dp + dd];
            float o = 0.f;
            if (l > 0.f) {
                float t = w.x * acc[0];
                if constexpr (FS > 1) t = fmaf(w.y, acc[1], t);
                if constexpr (FS > 2) t = fmaf(w.z, acc[2], t);
                if constexpr (FS > 3) t = fmaf(w.w, acc[3], t);
                o = fmaf(t, inv, r.w);
            }
            if (has_res) o += fmaf(r.y, xv1, fmaf(r.x, xv0, r.z));
            if (relu) o = fmaxf(o, 0.f);
            out[(k * D + dd) * RP] = o;
        }
    }
    __syncwarp();
}

template <int FS>
__device__ __forceinline__ void gat_rel_heads(int heads, const float* xs, int beg, int end, const float* tab, int D,
                                              float xv0, float xv1, int flags, float2* cbuf, float* out) {
    switch (heads) {
        case 1: gat_rel_row<FS, 1>(xs, beg, end, tab, D, xv0, xv1, flags, cbuf, out); break;
        case 2: gat_rel_row<FS, 2>(xs, beg, end, tab, D, xv0, xv1, flags, cbuf, out); break;
        case 4: gat_rel_row<FS, 4>(xs, beg, end, tab, D, xv0, xv1, flags, cbuf, out); break;
        default: gat_rel_row<FS, 8>(xs, beg, end, tab, D, xv0, xv1, flags, cbuf, out); break;
    }
}

// ---------------------------------------------------------------------------------------------- tensor-core layers
// Every dense layer of the step is a 16-row product: exactly the M of one mma.sync.m16n8k8 tile.  fp32 accuracy through
// the 3xTF32 split (hi = top 11 significand bits, lo = the rest; hi.hi + hi.lo + lo.hi, fp32 accumulate), done on the
// fly in registers.  Weights arrive in FRAGMENT ORDER (PackLayout f_*): the B operand of an MMA is one 8-byte shared
// load per lane, conflict free; activations stay feature-major with a row stride of 24 floats, which makes the
// A-fragment loads conflict free as well ((k0 + c) * 24 + g hits 32 different banks).  Column tile j goes to warp
// j mod 16; a warp keeps the accumulators of its (<= MAXT) tiles in registers and alternates two accumulator sets
// between even and odd k-steps so that no two MMAs in flight depend on each other.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi)) & 0xffffe000u;
}
__device__ __forceinline__ void mma8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// out[n * RP + r] = act(bias[n] + sum_k A[k * RP + r] W[k][n]),  n < 8 * ntiles.  All threads call; ends with __syncthreads().
template <int MAXT>
__device__ __forceinline__ void layer_mma(const float* Wf, const float* bias, const float* A, int Kd, float* out, int ntiles,
                                          bool relu) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    float acc[MAXT][2][3][4];
#pragma unroll
    for (int t = 0; t < MAXT; ++t)
#pragma unroll
        for (int s_ = 0; s_ < 2; ++s_)
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[t][s_][m][q] = 0.f;
    if (warp < ntiles) {
        const float* ap = A + c * RP + g;
        const float2* wp = reinterpret_cast<const float2*>(Wf) + warp * 32 + lane;
        const int nks = Kd >> 3;
        for (int ks = 0; ks < nks; ks += 2) {
#pragma unroll
            for (int s_ = 0; s_ < 2; ++s_) {
                if (ks + s_ < nks) {
                    const float* a0 = ap + (ks + s_) * 8 * RP;
                    uint32_t ah[4], al[4];
                    split_tf32(a0[0], ah[0], al[0]);
                    split_tf32(a0[8], ah[1], al[1]);
                    split_tf32(a0[4 * RP], ah[2], al[2]);
                    split_tf32(a0[4 * RP + 8], ah[3], al[3]);
#pragma unroll
                    for (int t = 0; t < MAXT; ++t) {
                        if (warp + 16 * t < ntiles) {
                            const float2 w = wp[((ks + s_) * ntiles + 16 * t) * 32];
                            uint32_t bh0, bl0, bh1, bl1;
                            split_tf32(w.x, bh0, bl0);
                            split_tf32(w.y, bh1, bl1);
                            mma8(acc[t][s_][1], al, bh0, bh1);
                            mma8(acc[t][s_][2], ah, bl0, bl1);
                            mma8(acc[t][s_][0], ah, bh0, bh1);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
            if (warp + 16 * t < ntiles) {
                const int n0 = 8 * (warp + 16 * t) + 2 * c;
                const float b0 = bias[n0], b1 = bias[n0 + 1];
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    v[q] = (acc[t][0][0][q] + acc[t][1][0][q]) + ((acc[t][0][1][q] + acc[t][1][1][q]) + (acc[t][0][2][q] + acc[t][1][2][q])) +
                           ((q & 1) ? b1 : b0);
                if (relu) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
                }
                out[n0 * RP + g] = v[0];
                out[(n0 + 1) * RP + g] = v[1];
                out[n0 * RP + g + 8] = v[2];
                out[(n0 + 1) * RP + g + 8] = v[3];
            }
        }
    }
    __syncthreads();
}

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

struct Layer { int off, n; };       // packed offset / float count (weights followed by their bias)

struct Plan {
    Layer layer[5]; int L;
    int capA, capB;                 // floats of the two weight buffers
    int act_rows;                   // activation tile rows (x RP floats)
    int scratch;                    // floats of split-K scratch (what is left of the 227 KB, at most NT*16)
};

__host__ __device__ inline Plan make_plan(const StepDims& d) {
    const PackLayout P = make_layout(d);
    Plan p{};
    int L = 0;
    auto pad4 = [](int n) { return (n + 3) & ~3; };
    // fragment-order weight blocks, each followed by its bias
    if (d.aggr()) p.layer[L++] = Layer{P.f_aggr, pad4(d.Fin * d.H) + pad4(d.H)};
    if (d.tarmac()) p.layer[L++] = Layer{P.f_vsq, pad4(2 * d.H * d.Vp()) + pad4(d.Vp())};
    p.layer[L++] = Layer{P.f_ih, pad4(d.Iih() * 3 * d.H) + pad4(3 * d.H)};
    p.layer[L++] = Layer{P.f_hh, pad4(d.H * 3 * d.H) + pad4(3 * d.H)};
    p.layer[L++] = Layer{P.f_out, pad4(d.H * d.Ap8()) + pad4(d.Ap8())};
    p.L = L;
    for (int i = 0; i < L; ++i) {
        int& cap = (i & 1) ? p.capB : p.capA;
        if (p.layer[i].n > cap) cap = p.layer[i].n;
    }
    // activation rows: [c | x | hp] contiguous, vsq, gi, gh (aliases xin), hn, q, alpha
    const int H = d.H, H3 = 3 * H;
    const int r_xin_gh = (d.aggr() ? (d.Fin > H3 ? d.Fin : H3) : H3);
    p.act_rows = (d.tarmac() ? d.M : 0) + H + H + (d.tarmac() ? d.Vp() : 0) + H3 + r_xin_gh + H + d.Ap8() + (d.tarmac() ? d.U : 0);
    p.scratch = 0;
    return p;
}

__host__ __device__ inline size_t smem_bytes(const Plan& p) {
    return ((size_t)p.capA + p.capB + (size_t)p.act_rows * RP + p.scratch) * sizeof(float) + 64;
}

__global__ void __launch_bounds__(NT, 1) agent_act_kernel(const StepArgs a) {
    extern __shared__ __align__(16) float sm[];
    const StepDims d = a.d;
    const Plan plan = make_plan(d);
    const int H = d.H, H3 = 3 * H, M = d.M, K = d.K, U = d.U, Vp = d.Vp(), A = d.A, Ap = d.Ap();
    const bool tm = d.tarmac(), ag = d.aggr();
    int o = 0;
    auto take = [&](int n) { float* p = sm + o; o += n; return p; };
    float* wbuf[2];
    wbuf[0] = take(plan.capA);
    wbuf[1] = take(plan.capB);
    float* sC = take(tm ? M * RP : 0);                 // [c | x | hp] contiguous: [c|x] feeds W_ih, [x|h] the comm projections
    float* sX = take(H * RP);
    float* sHp = take(H * RP);
    float* sVSQ = take(tm ? Vp * RP : 0);
    float* sGI = take(H3 * RP);
    float* sGH = take((ag ? (d.Fin > H3 ? d.Fin : H3) : H3) * RP);
    float* sXin = sGH;                                 // xin is dead once the aggregator GEMM has run
    float* sHn = take(H * RP);
    float* sQ = take(d.Ap8() * RP);
    float* sAl = take(tm ? U * RP : 0);
    float* scratch = take(plan.scratch);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + o);

    const bool fused_rel = a.rel.relpack != nullptr;
    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase[2] = {0u, 0u};
    auto issue = [&](int i) {                          // thread 0 only: start the copy of layer i into its buffer
        if (threadIdx.x == 0)
            bulk_load(wbuf[i & 1], a.packed + plan.layer[i].off, (uint32_t)plan.layer[i].n * 4u, bars + (i & 1));
    };
    auto wait = [&](int i) {
        mbar_wait(bars + (i & 1), phase[i & 1]);
        phase[i & 1] ^= 1u;
    };

    const int rpt = d.rows_per_tile();
    const int64_t row0 = (int64_t)blockIdx.x * rpt;
    const int n_valid = (int)min((int64_t)rpt, a.N - row0);
    const float scale = tm ? 1.0f / (float)K : 0.f;
    issue(0);
    // Programmatic dependent launch (UBS_ACT_PDL; both instructions are no-ops in a normal launch): let the next grid of
    // the stream start as soon as SMs free up, and do not touch anything the previous kernel may have written — the
    // hidden state, the packet — before that kernel has completed and flushed.  Above this line: barrier init and the
    // first weight layer's bulk copy (constant during a rollout).
    pdl_trigger();
    pdl_wait();
    load_tile(a.h0, row0, n_valid, H, H, sHp);

    for (int t = 0; t < a.T; ++t) {
    ATR(0);
        int li = 0;                                    // index of the layer about to run
        if (fused_rel) {
            // ---- the two observation relations, staged into the (still idle) buffer of weight layer 1 ----------------
            const RelIn& rl = a.rel;
            const RelSmem rs = rel_smem(d, rl);
            float* stage = wbuf[1];
            int* ips = reinterpret_cast<int*>(stage + rs.ips);
            float* xd = stage + rs.xd;
            if (threadIdx.x < 2 * (R + 1)) {
                const int rel = threadIdx.x / (R + 1), i = threadIdx.x - rel * (R + 1);
                ips[threadIdx.x] = __ldg(rl.indptr[rel] + row0 + min(i, n_valid));
            } else if (threadIdx.x >= 64 && threadIdx.x < 64 + 2 * R) {
                const int i = threadIdx.x - 64, r = i >> 1, f = i & 1;
                xd[i] = (r < n_valid && f < rl.F_d) ? __ldg(rl.x_dst + (row0 + r) * rl.F_d + f) : 0.f;
            }
            __syncthreads();
            int shift[2];
#pragma unroll
            for (int rel = 0; rel < 2; ++rel) shift[rel] = (ips[rel * (R + 1)] * rl.FS[rel]) & 3;
            if (threadIdx.x == 0) {
                // neighbour lists of the tile: rows [ip[row0], ip[row0 + n_valid]) are contiguous in the packet (star layout)
                uint32_t bytes[2];
                const float* src[2];
                const uint32_t tab_bytes = 2u * (uint32_t)rel_table_floats(H, rl.heads) * 4u;
                uint32_t total = tab_bytes;
#pragma unroll
                for (int rel = 0; rel < 2; ++rel) {
                    const int64_t fbeg = (int64_t)ips[rel * (R + 1)] * rl.FS[rel];
                    const int64_t fend = (int64_t)ips[rel * (R + 1) + n_valid] * rl.FS[rel];
                    const int64_t abeg = fbeg & ~(int64_t)3, aend = (fend + 3) & ~(int64_t)3;
                    src[rel] = rl.x_src[rel] + abeg;
                    bytes[rel] = fend > fbeg ? (uint32_t)(aend - abeg) * 4u : 0u;
                    total += bytes[rel];
                }
                mbar_expect(bars + 2, total);
                bulk_copy(stage + rs.tabs, rl.relpack, tab_bytes, bars + 2);
#pragma unroll
                for (int rel = 0; rel < 2; ++rel)
                    if (bytes[rel]) bulk_copy(stage + rs.xs[rel], src[rel], bytes[rel], bars + 2);
            }
            mbar_wait(bars + 2, 0);
            {
                const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;      // warp = destination row of the tile
                float2* cb = reinterpret_cast<float2*>(stage + rs.cbuf) + r * (H + rl.heads);
                const int D = H / rl.heads;
#pragma unroll
                for (int rel = 0; rel < 2; ++rel) {
                    float* out = sXin + rel * H * RP + r;
                    if (r < n_valid) {
                        const int b0 = ips[rel * (R + 1)];
                        const int beg = ips[rel * (R + 1) + r] - b0, end = ips[rel * (R + 1) + r + 1] - b0;
                        const float* xs = stage + rs.xs[rel] + shift[rel];
                        const float* tab = stage + rs.tabs + rel * rel_table_floats(H, rl.heads);
                        switch (rl.FS[rel]) {
                            case 1: gat_rel_heads<1>(rl.heads, xs, beg, end, tab, D, xd[2 * r], xd[2 * r + 1], rl.gat_flags, cb, out); break;
                            case 2: gat_rel_heads<2>(rl.heads, xs, beg, end, tab, D, xd[2 * r], xd[2 * r + 1], rl.gat_flags, cb, out); break;
                            case 3: gat_rel_heads<3>(rl.heads, xs, beg, end, tab, D, xd[2 * r], xd[2 * r + 1], rl.gat_flags, cb, out); break;
                            default: gat_rel_heads<4>(rl.heads, xs, beg, end, tab, D, xd[2 * r], xd[2 * r + 1], rl.gat_flags, cb, out); break;
                        }
                    } else {
                        for (int ch = lane; ch < H; ch += 32) out[ch * RP] = 0.f;
                    }
                }
            }
            __syncthreads();
            // the staging region is about to be overwritten by a bulk copy (async proxy) after generic-proxy accesses
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        } else {
            load_tile(a.xin + t * a.st_xin, row0, n_valid, d.Fin, d.Fin, ag ? sXin : sX);
        }
        __syncthreads();
        ATR(1);
        if (ag) {
            issue(li + 1);
            wait(li);
            layer_mma<1>(wbuf[li & 1], wbuf[li & 1] + ((d.Fin * H + 3) & ~3), sXin, d.Fin, sX, H / 8, true);
            ++li;
        ATR(2);
        }
        if (tm) {
            issue(li + 1);
            wait(li);
            layer_mma<1>(wbuf[li & 1], wbuf[li & 1] + ((2 * H * Vp + 3) & ~3), sX, 2 * H, sVSQ, Vp / 8, false);
            ATR(3);
            ++li;
            const uint32_t* mk = a.mask + t * a.st_mask;
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
            ATR(4);
        }
        // gi = W_ih [x ‖ c] + b_ih (smem order [c | x]),  gh = W_hh h + b_hh
        issue(li + 1);
        wait(li);
        layer_mma<2>(wbuf[li & 1], wbuf[li & 1] + ((d.Iih() * H3 + 3) & ~3), tm ? sC : sX, d.Iih(), sGI, H3 / 8, false);
        ATR(5);
        ++li;
        issue(li + 1);
        wait(li);
        layer_mma<2>(wbuf[li & 1], wbuf[li & 1] + ((H * H3 + 3) & ~3), sHp, H, sGH, H3 / 8, false);
        ATR(6);
        ++li;
        for (int p = threadIdx.x; p < H * R; p += NT) {
            const int ch = p / R, r = p - ch * R;
            const float rr = sigmoidf_(sGI[ch * RP + r] + sGH[ch * RP + r]);
            const float zz = sigmoidf_(sGI[(H + ch) * RP + r] + sGH[(H + ch) * RP + r]);
            const float nn = tanhf(fmaf(rr, sGH[(2 * H + ch) * RP + r], sGI[(2 * H + ch) * RP + r]));
            sHn[ch * RP + r] = fmaf(zz, sHp[ch * RP + r] - nn, nn);
        }
        __syncthreads();
        ATR(7);
        wait(li);
        layer_mma<1>(wbuf[li & 1], wbuf[li & 1] + ((H * d.Ap8() + 3) & ~3), sHn, H, sQ, d.Ap8() / 8, false);
        ATR(8);
        if (t + 1 < a.T) issue(0);                     // every layer of this step has finished: layer 0's buffer is free
        store_tile(a.h_out + t * a.st_h, row0, n_valid, H, H, sHn);
        store_tile(a.q + t * a.st_q, row0, n_valid, A, A, sQ);
        if (a.actions != nullptr && threadIdx.x < n_valid) {
            const int r = threadIdx.x;
            int best = 0;
            float bv = sQ[r];
            for (int c = 1; c < A; ++c) { const float v = sQ[c * RP + r]; if (v > bv) { bv = v; best = c; } }
            const int64_t ai = t * a.st_act + row0 + r;
            int64_t act_ = best;
            if (a.eg_u != nullptr && __ldg(a.eg_u + ai) <= __ldg(a.eg_eps)) act_ = __ldg(a.eg_a + ai);
            a.actions[ai] = act_;
        }
        for (int p = threadIdx.x; p < H * RP; p += NT) sHp[p] = sHn[p];
        __syncthreads();
        ATR(9);
    }
}

}  // namespace act

// Launches the TMA-staged act kernel when the configuration fits (inference, weights of two consecutive layers +
// activations <= 227 KB of shared memory); *handled = false tells the caller to use the register-streaming kernel.
bool agent_act_fits(const StepDims& d) {
    static const bool enabled = [] { const char* e = getenv("UBS_ACT_TMA"); return !(e && e[0] == '0'); }();
    if (!enabled) return false;
    if (!d.mma_ok() || 3 * d.H / 8 > 32 || d.H / 8 > 16 || (d.tarmac() && d.Vp() / 8 > 16)) return false;   // <= 2 column tiles per warp
    const act::Plan plan = act::make_plan(d);
    if (act::smem_bytes(plan) > 227 * 1024) return false;
    for (int i = 0; i < plan.L; ++i)
        if ((size_t)plan.layer[i].n * 4 >= (1u << 20)) return false;     // mbarrier tx-count range
    return true;
}

// The fused-relation prologue stages tables + neighbour lists in the buffer of weight layer 1.
static bool agent_act_rel_fits(const StepDims& d, const RelIn& r) {
    if (!d.aggr() || d.Fin != 2 * d.H || !agent_act_fits(d)) return false;
    if (r.heads != 1 && r.heads != 2 && r.heads != 4 && r.heads != 8) return false;
    if (d.H % r.heads || r.F_d < 1 || r.F_d > 2) return false;
    for (int i = 0; i < 2; ++i)
        if (r.FS[i] < 1 || r.FS[i] > 4 || r.cap[i] < 0) return false;
    const act::Plan plan = act::make_plan(d);
    return act::rel_smem(d, r).total <= plan.capB;
}

int launch_agent_act(const StepArgs& a, cudaStream_t st, bool* handled) {
    *handled = false;
    if (a.sv_gate != nullptr) return 0;                // training saves: streaming / resident-weight kernels
    if ((reinterpret_cast<uintptr_t>(a.packed) & 15u) != 0) return 0;   // bulk copies need 16-byte aligned sources
    if (!agent_act_fits(a.d)) return 0;
    const act::Plan plan = act::make_plan(a.d);
    const size_t smem = act::smem_bytes(plan);
    // every configuration that fits needs more than the 48 KB default: opt in to the device maximum once per
    // process (idempotent, so a racing second thread only repeats the same call)
    static const cudaError_t attr_rc = cudaFuncSetAttribute(act::agent_act_kernel,
                                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_rc != cudaSuccess) {
        set_error("ubs_agent_act_fwd: cannot opt in to 227 KB of shared memory: %s", cudaGetErrorString(attr_rc));
        return 1;
    }
    const int rpt = a.d.rows_per_tile();
    const unsigned grid = (unsigned)((a.N + rpt - 1) / rpt);
    // UBS_ACT_PDL: programmatic dependent launch — the grid may start while the previous kernel of the stream drains;
    // everything it reads before pdl_wait() (the packed weights) must not be written by that kernel
    const cudaError_t lrc = launch_pdl(act::agent_act_kernel, grid, (unsigned)act::NT, smem, st, a.pdl != 0, a);
    if (lrc != cudaSuccess) {
        set_error("ubs_agent_act_fwd: launch failed: %s", cudaGetErrorString(lrc));
        return 1;
    }
    *handled = true;
    return check_launch("ubs_agent_act_fwd(tma)");
}

#ifdef UBS_ACT_TRACE
}  // namespace ubs
extern "C" UBS_API int ubs_act_trace_read(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, ubs::act::ubs_act_trace, sizeof(long long) * 16);
}
namespace ubs {
#endif
// ---------------------------------------------------------------------------------------------- relation tables
struct RelPackArgs {
    const float *W_src, *b_src, *W_dst, *b_dst, *attn, *W_res, *b_res;
    int FS, FD, heads, D, flags; float slope; float* out;
};

__global__ void __launch_bounds__(256) gatv2_rel_pack_kernel(const RelPackArgs a) {
    const int H = a.heads * a.D, Hp = H + a.heads;          // padded: channel ch lives at index ch + ch / D
    float4* wA = reinterpret_cast<float4*>(a.out);
    float4* wR = wA + Hp;
    float4* wD = wR + Hp;
    float* hP = reinterpret_cast<float*>(wD + Hp);
    for (int i = threadIdx.x; i < 3 * Hp; i += 256) wA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int ch = threadIdx.x; ch < H; ch += 256) {
        const int p = ch + ch / a.D;
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        for (int f = 0; f < a.FS; ++f) w[f] = a.W_src[ch * a.FS + f];
        wA[p] = make_float4(w[0], w[1], w[2], w[3]);
        const float bs = a.b_src ? a.b_src[ch] : 0.f;
        const float bd = a.b_dst ? a.b_dst[ch] : 0.f;
        float r0 = 0.f, r1 = 0.f, rb = 0.f;
        if (a.flags & UBS_GAT_RESIDUAL) {
            r0 = a.W_res[ch * a.FD];
            r1 = a.FD > 1 ? a.W_res[ch * a.FD + 1] : 0.f;
            rb = a.b_res ? a.b_res[ch] : 0.f;
        }
        wR[p] = make_float4(r0, r1, rb, bs);
        wD[p] = make_float4(a.W_dst[ch * a.FD], a.FD > 1 ? a.W_dst[ch * a.FD + 1] : 0.f, bs + bd,
                            0.5f * (1.0f - a.slope) * a.attn[ch]);
    }
    __syncthreads();
    if (threadIdx.x < a.heads * 8) {
        const int k = threadIdx.x / 8, j = threadIdx.x % 8;
        float acc = 0.f;
        for (int d0 = 0; d0 < a.D; ++d0) {
            const int ch = k * a.D + d0, p = ch + k;
            const float at = a.attn[ch];
            float w = 0.f;
            if (j < 4) w = j == 0 ? wA[p].x : j == 1 ? wA[p].y : j == 2 ? wA[p].z : wA[p].w;
            else if (j == 4) w = wD[p].x;
            else if (j == 5) w = wD[p].y;
            else if (j == 6) w = wD[p].z;
            acc = fmaf(at, w, acc);
        }
        hP[threadIdx.x] = 0.5f * (1.0f + a.slope) * acc;
    }
}

}  // namespace ubs

extern "C" UBS_API int64_t ubs_gatv2_rel_pack_size(int heads, int D) { return ubs::act::rel_table_floats(heads * D, heads); }

extern "C" UBS_API int ubs_gatv2_rel_pack(const float* W_src, const float* b_src, const float* W_dst, const float* b_dst,
                                          const float* attn, const float* W_res, const float* b_res, int F_s, int F_d,
                                          int heads, int D, float negative_slope, int flags, float* out, void* stream) {
    UBS_REQUIRE(W_src && W_dst && attn && out, "ubs_gatv2_rel_pack: NULL argument");
    UBS_REQUIRE(F_s >= 1 && F_s <= 4 && F_d >= 1 && F_d <= 2 && heads >= 1 && heads <= 8 && D >= 1,
                "ubs_gatv2_rel_pack: shape outside the fused relation kernels (F_s <= 4, F_d <= 2, heads <= 8)");
    UBS_REQUIRE(!(flags & UBS_GAT_RESIDUAL) || W_res, "ubs_gatv2_rel_pack: residual weights missing");
    UBS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "ubs_gatv2_rel_pack: out must be 16-byte aligned");
    ubs::RelPackArgs a{W_src, b_src, W_dst, b_dst, attn, W_res, b_res, F_s, F_d, heads, D, flags, negative_slope, out};
    ubs::gatv2_rel_pack_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_gatv2_rel_pack");
}

static ubs::RelIn mk_rel(const float* relpack, const float* x_gt, const int32_t* ip_seen, int F_gt, int cap_gt,
                         const float* x_ubs, const int32_t* ip_near, int F_ubs, int cap_ubs, const float* x_agent, int F_d,
                         int heads, int gat_flags) {
    ubs::RelIn r{};
    r.relpack = relpack;
    r.x_src[0] = x_gt; r.indptr[0] = ip_seen; r.FS[0] = F_gt; r.cap[0] = cap_gt;
    r.x_src[1] = x_ubs; r.indptr[1] = ip_near; r.FS[1] = F_ubs; r.cap[1] = cap_ubs;
    r.x_dst = x_agent; r.F_d = F_d; r.heads = heads; r.gat_flags = gat_flags;
    return r;
}

extern "C" UBS_API int ubs_agent_act_rel_supported(int H, int M, int K, int A, int U, int flags, int heads, int F_gt,
                                                   int cap_gt, int F_ubs, int cap_ubs, int F_d) {
    const ubs::StepDims d = ubs::mk_dims(H, M, K, A, U, 2 * H, flags);
    const ubs::RelIn r = mk_rel(nullptr, nullptr, nullptr, F_gt, cap_gt, nullptr, nullptr, F_ubs, cap_ubs, nullptr, F_d, heads, 0);
    return ubs::agent_act_rel_fits(d, r) ? 1 : 0;
}

extern "C" UBS_API int ubs_agent_act_rel_fwd(int H, int M, int K, int A, int U, int flags, const float* packed,
                                             const float* relpack, const float* x_gt, const int32_t* ip_seen, int F_gt,
                                             int cap_gt, const float* x_ubs, const int32_t* ip_near, int F_ubs, int cap_ubs,
                                             const float* x_agent, int F_d, int heads, int gat_flags, const float* h0,
                                             const uint32_t* mask, float* h_out, float* q, int64_t* actions,
                                             const float* eg_u, const int64_t* eg_a, const float* eg_eps, int64_t n_rows,
                                             void* stream) {
    ubs::StepArgs a{};
    a.d = ubs::mk_dims(H, M, K, A, U, 2 * H, flags);
    if (int rc = ubs::check_dims("ubs_agent_act_rel_fwd", a.d)) return rc;
    UBS_REQUIRE(packed && relpack && x_gt && ip_seen && x_ubs && ip_near && x_agent && h0 && h_out && q,
                "ubs_agent_act_rel_fwd: NULL argument");
    UBS_REQUIRE(!a.d.tarmac() || mask, "ubs_agent_act_rel_fwd: TarMAC needs the block mask");
    UBS_REQUIRE(!a.d.tarmac() || n_rows % a.d.U == 0, "ubs_agent_act_rel_fwd: n_rows must be a multiple of agents per env");
    UBS_REQUIRE(eg_u == nullptr || (eg_a && eg_eps && actions), "ubs_agent_act_rel_fwd: incomplete epsilon-greedy arguments");
    a.rel = mk_rel(relpack, x_gt, ip_seen, F_gt, cap_gt, x_ubs, ip_near, F_ubs, cap_ubs, x_agent, F_d, heads, gat_flags & ~UBS_ACT_PDL);
    a.pdl = (gat_flags & UBS_ACT_PDL) ? 1 : 0;
    UBS_REQUIRE(ubs::agent_act_rel_fits(a.d, a.rel), "ubs_agent_act_rel_fwd: configuration outside the fused kernel "
                "(ask ubs_agent_act_rel_supported first)");
    for (const void* p : {(const void*)packed, (const void*)relpack, (const void*)x_gt, (const void*)x_ubs})
        UBS_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15u) == 0, "ubs_agent_act_rel_fwd: bulk-copy sources must be 16-byte aligned");
    if (n_rows == 0) return 0;
    a.packed = packed; a.h0 = h0; a.mask = mask; a.st_mask = n_rows;
    a.h_out = h_out; a.st_h = n_rows * H; a.q = q; a.st_q = n_rows * A; a.actions = actions; a.st_act = n_rows;
    a.eg_u = eg_u; a.eg_a = eg_a; a.eg_eps = eg_eps; a.N = n_rows; a.T = 1;
    bool handled = false;
    const int rc = ubs::launch_agent_act(a, (cudaStream_t)stream, &handled);
    UBS_REQUIRE(handled || rc, "ubs_agent_act_rel_fwd: the fused kernel did not take the launch");
    return rc;
}
