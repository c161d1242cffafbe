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
#include "agent_step.cuh"

namespace ubs {
namespace act {

constexpr int NT = 512;

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

// out[j*RP + r] = act(bias[j] + sum_k W[k*ldw + j] * A[k*RP + r]);  W, bias, A in shared memory.  mode 0 store, 1 relu.
__device__ __noinline__ void gemm_s(const float* W, int ldw, const float* bias, const float* A, int Kd, float* out,
                                       int Nout, int mode, float* scratch, int scratch_cap) {
    const int ntc = Nout >> 2, tiles = ntc * 4;
    int ksplit = 1;                                    // split K while threads and partial-sum scratch allow
    while (ksplit < 8 && tiles * ksplit * 2 <= NT && Kd >= ksplit * 16 && (2 * ksplit - 1) * tiles * 16 <= scratch_cap) ksplit *= 2;
    const int kchunk = (((Kd + ksplit - 1) / ksplit) + 3) & ~3;
    for (int base = 0; base < tiles * ksplit; base += NT) {
        const int t = base + threadIdx.x;
        const bool active = t < tiles * ksplit;
        const int ks = active ? t / tiles : 0, tile = active ? t - ks * tiles : 0;
        const int ct = tile % ntc, rt = tile / ntc;
        float acc[4][4];
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
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = 4 * ct + c;
                const float b = bias[j];
                float4 v = make_float4(acc[0][c] + b, acc[1][c] + b, acc[2][c] + b, acc[3][c] + b);
                if (mode == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4*>(out + j * RP + 4 * rt) = v;
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
    if (d.aggr()) p.layer[L++] = Layer{P.t_aggr, pad4(d.Fin * d.H) + pad4(d.H)};
    if (d.tarmac()) p.layer[L++] = Layer{P.t_vsq, pad4(2 * d.H * d.Vp()) + pad4(d.Vp())};
    p.layer[L++] = Layer{P.t_ih, pad4(d.Iih() * 3 * d.H) + pad4(3 * d.H)};
    p.layer[L++] = Layer{P.t_hh, pad4(d.H * 3 * d.H) + pad4(3 * d.H)};
    p.layer[L++] = Layer{P.t_out, pad4(d.H * d.Ap()) + pad4(d.Ap())};
    p.L = L;
    for (int i = 0; i < L; ++i) {
        int& cap = (i & 1) ? p.capB : p.capA;
        if (p.layer[i].n > cap) cap = p.layer[i].n;
    }
    // activation rows: [c | x | hp] contiguous, vsq, gi, gh (aliases xin), hn, q, alpha
    const int H = d.H, H3 = 3 * H;
    const int r_xin_gh = (d.aggr() ? (d.Fin > H3 ? d.Fin : H3) : H3);
    p.act_rows = (d.tarmac() ? d.M : 0) + H + H + (d.tarmac() ? d.Vp() : 0) + H3 + r_xin_gh + H + d.Ap() + (d.tarmac() ? d.U : 0);
    const int budget = (227 * 1024 - 64) / 4 - p.capA - p.capB - p.act_rows * RP;
    p.scratch = budget < NT * 16 ? (budget < 0 ? 0 : budget & ~3) : NT * 16;
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
    float* sQ = take(Ap * RP);
    float* sAl = take(tm ? U * RP : 0);
    float* scratch = take(plan.scratch);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + o);

    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
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
    load_tile(a.h0, row0, n_valid, H, H, sHp);

    for (int t = 0; t < a.T; ++t) {
        int li = 0;                                    // index of the layer about to run
        load_tile(a.xin + t * a.st_xin, row0, n_valid, d.Fin, d.Fin, ag ? sXin : sX);
        __syncthreads();
        if (ag) {
            issue(li + 1);
            wait(li);
            gemm_s(wbuf[li & 1], H, wbuf[li & 1] + ((d.Fin * H + 3) & ~3), sXin, d.Fin, sX, H, 1, scratch, plan.scratch);
            ++li;
        }
        if (tm) {
            issue(li + 1);
            wait(li);
            gemm_s(wbuf[li & 1], Vp, wbuf[li & 1] + ((2 * H * Vp + 3) & ~3), sX, 2 * H, sVSQ, Vp, 0, scratch, plan.scratch);
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
        }
        // gi = W_ih [x ‖ c] + b_ih (smem order [c | x]),  gh = W_hh h + b_hh
        issue(li + 1);
        wait(li);
        gemm_s(wbuf[li & 1], H3, wbuf[li & 1] + ((d.Iih() * H3 + 3) & ~3), tm ? sC : sX, d.Iih(), sGI, H3, 0, scratch, plan.scratch);
        ++li;
        issue(li + 1);
        wait(li);
        gemm_s(wbuf[li & 1], H3, wbuf[li & 1] + ((H * H3 + 3) & ~3), sHp, H, sGH, H3, 0, scratch, plan.scratch);
        ++li;
        for (int p = threadIdx.x; p < H * R; p += NT) {
            const int ch = p / R, r = p - ch * R;
            const float rr = sigmoidf_(sGI[ch * RP + r] + sGH[ch * RP + r]);
            const float zz = sigmoidf_(sGI[(H + ch) * RP + r] + sGH[(H + ch) * RP + r]);
            const float nn = tanhf(fmaf(rr, sGH[(2 * H + ch) * RP + r], sGI[(2 * H + ch) * RP + r]));
            sHn[ch * RP + r] = fmaf(zz, sHp[ch * RP + r] - nn, nn);
        }
        __syncthreads();
        wait(li);
        gemm_s(wbuf[li & 1], Ap, wbuf[li & 1] + ((H * Ap + 3) & ~3), sHn, H, sQ, Ap, 0, scratch, plan.scratch);
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
    }
}

}  // namespace act

// Launches the TMA-staged act kernel when the configuration fits (inference, weights of two consecutive layers +
// activations <= 227 KB of shared memory); *handled = false tells the caller to use the register-streaming kernel.
bool agent_act_fits(const StepDims& d) {
    static const bool enabled = [] { const char* e = getenv("UBS_ACT_TMA"); return !(e && e[0] == '0'); }();
    if (!enabled) return false;
    const act::Plan plan = act::make_plan(d);
    if (act::smem_bytes(plan) > 227 * 1024 || plan.scratch < 3 * d.H * 16) return false;   // the 3H-wide GRU GEMMs need one K split
    for (int i = 0; i < plan.L; ++i)
        if ((size_t)plan.layer[i].n * 4 >= (1u << 20)) return false;     // mbarrier tx-count range
    return true;
}

int launch_agent_act(const StepArgs& a, cudaStream_t st, bool* handled) {
    *handled = false;
    if (a.sv_gate != nullptr) return 0;                // training saves: streaming / resident-weight kernels
    if ((reinterpret_cast<uintptr_t>(a.packed) & 15u) != 0) return 0;   // bulk copies need 16-byte aligned sources
    if (!agent_act_fits(a.d)) return 0;
    const act::Plan plan = act::make_plan(a.d);
    const size_t smem = act::smem_bytes(plan);
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(act::agent_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    const int rpt = a.d.rows_per_tile();
    const unsigned grid = (unsigned)((a.N + rpt - 1) / rpt);
    act::agent_act_kernel<<<grid, act::NT, smem, st>>>(a);
    *handled = true;
    return check_launch("ubs_agent_act_fwd(tma)");
}

}  // namespace ubs
