// Fused recurrent agent step (forward and backward), per-step or persistent over a whole sequence.
//
// Replaces, per timestep, everything reference GnnAgent.forward does after the two GATv2 relations
// (algos/madrqn/agents/gnn_agents.py:51-56): the aggregator Linear(2H,H)+ReLU (:99,106-107), TarMAC.forward
// (:248-271: f_val/f_sign/f_que on [x ‖ h.detach()], u_dot_v / key_size, edge_softmax, u_mul_e+sum, GRUCell on [x ‖ c])
// or the plain GRUCell when c is None (:29,55; drqn gnn_agents.py:20,28), and the Q head f_out (:43-46,56).
// DGL + PyTorch run ~25 kernels for this; here it is ONE kernel per call, and with n_steps > 1 ONE kernel for a
// whole BPTT sequence: env instances never exchange data (the talk graph is block-diagonal per env), so a CTA
// owns a tile of <=16 agent rows (whole envs) and walks all timesteps with the hidden state resident in shared
// memory — no grid-wide synchronisation between timesteps.
//
// Layout: activations live in shared memory feature-major ("K-major", [feature][row], row stride RP) so that every
// dense layer is the same register-tiled FP32 GEMM: thread tile = 4 output features x 4 rows, one coalesced 16-byte
// weight load (weights pre-transposed once per parameter update into a packed buffer, streamed from L2) and one
// 16-byte shared load per 16 FMAs; small layers are split over K to keep all 256 threads busy.
// FP32 FMA on purpose: 1e-5 parity rules out single-pass TF32, and with 16 rows per SM a 3xTF32 tcgen05 tile
// (M >= 64) would leave 3/4 of the SMs idle.
#include "agent_step.cuh"

namespace ubs {

constexpr int NT = 256;   // threads per CTA

struct PackArgs {
    StepDims d;
    const float *W_aggr, *b_aggr, *W_val, *b_val, *W_sign, *b_sign, *W_que, *b_que, *W_ih, *b_ih, *W_hh, *b_hh, *W_out, *b_out;
    float* packed;
};

// One thread per packed element; gathers from the PyTorch-layout parameters.
__global__ void __launch_bounds__(256) agent_pack_kernel(const PackArgs a) {
    const StepDims d = a.d;
    const PackLayout L = make_layout(d);
    const int H = d.H, H3 = 3 * H, M = d.M, K = d.K, V = d.V(), Vp = d.Vp(), A = d.A, Ap = d.Ap(), I = d.Iih();
    for (int i = blockIdx.x * 256 + threadIdx.x; i < L.total; i += gridDim.x * 256) {
        float v = 0.f;
        auto vsq_w = [&](int row, int col) -> float {          // concatenated [W_val; W_sign; W_que] (V x 2H)
            if (row < M) return a.W_val[row * 2 * H + col];
            if (row < M + K) return a.W_sign[(row - M) * 2 * H + col];
            if (row < V) return a.W_que[(row - M - K) * 2 * H + col];
            return 0.f;
        };
        auto vsq_b = [&](int row) -> float {
            if (row < M) return a.b_val[row];
            if (row < M + K) return a.b_sign[row - M];
            if (row < V) return a.b_que[row - M - K];
            return 0.f;
        };
        // forward GRU input order in smem is [c | x] (so that [x | h] is contiguous for the comm projections)
        auto ih_col = [&](int k) -> int { return d.tarmac() ? (k < M ? H + k : k - M) : k; };
        if (d.aggr() && i >= L.t_aggr && i < L.t_aggr + d.Fin * H) { int j = i - L.t_aggr; v = a.W_aggr[(j % H) * d.Fin + j / H]; }
        else if (d.aggr() && i >= L.b_aggr && i < L.b_aggr + H) v = a.b_aggr[i - L.b_aggr];
        else if (d.tarmac() && i >= L.t_vsq && i < L.t_vsq + 2 * H * Vp) { int j = i - L.t_vsq; v = vsq_w(j % Vp, j / Vp); }
        else if (d.tarmac() && i >= L.b_vsq && i < L.b_vsq + Vp) v = vsq_b(i - L.b_vsq);
        else if (i >= L.t_ih && i < L.t_ih + I * H3) { int j = i - L.t_ih; v = a.W_ih[(j % H3) * I + ih_col(j / H3)]; }
        else if (i >= L.b_ih && i < L.b_ih + H3) v = a.b_ih[i - L.b_ih];
        else if (i >= L.t_hh && i < L.t_hh + H * H3) { int j = i - L.t_hh; v = a.W_hh[(j % H3) * H + j / H3]; }
        else if (i >= L.b_hh && i < L.b_hh + H3) v = a.b_hh[i - L.b_hh];
        else if (i >= L.t_out && i < L.t_out + H * Ap) { int j = i - L.t_out; int c = j % Ap; v = c < A ? a.W_out[c * H + j / Ap] : 0.f; }
        else if (i >= L.b_out && i < L.b_out + Ap) { int c = i - L.b_out; v = c < A ? a.b_out[c] : 0.f; }
        else if (d.aggr() && i >= L.o_aggr && i < L.o_aggr + H * d.Fin) v = a.W_aggr[i - L.o_aggr];
        else if (d.tarmac() && i >= L.o_vsq && i < L.o_vsq + Vp * 2 * H) { int j = i - L.o_vsq; v = vsq_w(j / (2 * H), j % (2 * H)); }
        else if (i >= L.o_ih && i < L.o_ih + H3 * I) v = a.W_ih[i - L.o_ih];
        else if (i >= L.o_hh && i < L.o_hh + H3 * H) v = a.W_hh[i - L.o_hh];
        else if (i >= L.o_out && i < L.o_out + Ap * H) { int j = i - L.o_out; v = (j / H) < A ? a.W_out[j] : 0.f; }
        else if (i >= L.f_aggr && d.mma_ok()) {
            // fragment-order copies: element e of lane `ln` of tile j of k-step ks = W[k = 8ks + ln%4 + 4e][n = 8j + ln/4]
            auto frag = [&](int j_, int N, int& k, int& n) {
                const int e = j_ & 1, ln = (j_ >> 1) & 31, tile = (j_ >> 6) % (N / 8), ks = (j_ >> 6) / (N / 8);
                k = 8 * ks + (ln & 3) + 4 * e;
                n = 8 * tile + (ln >> 2);
            };
            int k, n;
            const int A8 = d.Ap8();
            if (d.aggr() && i < L.f_aggr + d.Fin * H) { frag(i - L.f_aggr, H, k, n); v = a.W_aggr[n * d.Fin + k]; }
            else if (d.aggr() && i >= L.fb_aggr && i < L.fb_aggr + H) v = a.b_aggr[i - L.fb_aggr];
            else if (d.tarmac() && i >= L.f_vsq && i < L.f_vsq + 2 * H * Vp) { frag(i - L.f_vsq, Vp, k, n); v = vsq_w(n, k); }
            else if (d.tarmac() && i >= L.fb_vsq && i < L.fb_vsq + Vp) v = vsq_b(i - L.fb_vsq);
            else if (i >= L.f_ih && i < L.f_ih + I * H3) { frag(i - L.f_ih, H3, k, n); v = a.W_ih[n * I + ih_col(k)]; }
            else if (i >= L.fb_ih && i < L.fb_ih + H3) v = a.b_ih[i - L.fb_ih];
            else if (i >= L.f_hh && i < L.f_hh + H * H3) { frag(i - L.f_hh, H3, k, n); v = a.W_hh[n * H + k]; }
            else if (i >= L.fb_hh && i < L.fb_hh + H3) v = a.b_hh[i - L.fb_hh];
            else if (i >= L.f_out && i < L.f_out + H * A8) { frag(i - L.f_out, A8, k, n); v = n < A ? a.W_out[n * H + k] : 0.f; }
            else if (i >= L.fb_out && i < L.fb_out + A8) { int c = i - L.fb_out; v = c < A ? a.b_out[c] : 0.f; }
        }
        a.packed[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------- tile GEMM
// out[j*RP + r] (op)= bias[j] + sum_{k<Kd} W[k*ldw + j] * A[k*RP + r]      j < Nout (multiple of 4), r < 16
// W is K-major in global memory (L2 resident), A / out are feature-major shared tiles.  mode: 0 store, 1 store+relu,
// 2 accumulate into out.  All threads of the CTA must call; ends with __syncthreads().
// __noinline__: the kernels call it 5-6 times; one copy keeps the hot loop inside the instruction cache.
__device__ __noinline__ void tile_gemm(const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                                       const float* A, int Kd, float* out, int Nout, int mode, float* scratch) {
    const int ntc = Nout >> 2, tiles = ntc * 4;
    int ksplit = 1;
    while (ksplit < 8 && tiles * ksplit * 2 <= NT && Kd >= ksplit * 16) ksplit *= 2;
    const int kchunk = (((Kd + ksplit - 1) / ksplit) + 3) & ~3;
    for (int base = 0; base < tiles * ksplit; base += NT) {
        const int t = base + threadIdx.x;
        const bool active = t < tiles * ksplit;
        const int ks = active ? t / tiles : 0, tile = active ? t - ks * tiles : 0;
        const int ct = tile % ntc, rt = tile / ntc;
        float acc[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[c][r] = 0.f;
        if (active) {
            // Weights stream from L2 (~250+ cycle latency).  Two register buffers of KB rows are filled and consumed
            // in ping-pong (no register copies): KB 16-byte loads per thread are in flight while the other buffer
            // feeds 16*KB FMAs.
            constexpr int KB = 8;
            const int k0 = ks * kchunk, k1 = min(Kd, k0 + kchunk);
            const float* wp = W + 4 * ct;
            const float* ap = A + 4 * rt;
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            auto load_block = [&](float4 (&w)[KB], int kb) {
#pragma unroll
                for (int i = 0; i < KB; ++i)
                    w[i] = (kb + i < k1) ? __ldg(reinterpret_cast<const float4*>(wp + (size_t)(kb + i) * ldw)) : zero4;
            };
            auto fma_block = [&](const float4 (&w)[KB], int kb) {
                float4 xa[KB];                                  // all shared loads of the block first
#pragma unroll
                for (int i = 0; i < KB; ++i)
                    xa[i] = *reinterpret_cast<const float4*>(ap + min(kb + i, k1 - 1) * RP);   // rows past k1: zero weights
#pragma unroll
                for (int i = 0; i < KB; ++i) {
                    const float4 ww = w[i];
                    const float4 x = xa[i];
                    acc[0][0] = fmaf(ww.x, x.x, acc[0][0]); acc[0][1] = fmaf(ww.x, x.y, acc[0][1]);
                    acc[0][2] = fmaf(ww.x, x.z, acc[0][2]); acc[0][3] = fmaf(ww.x, x.w, acc[0][3]);
                    acc[1][0] = fmaf(ww.y, x.x, acc[1][0]); acc[1][1] = fmaf(ww.y, x.y, acc[1][1]);
                    acc[1][2] = fmaf(ww.y, x.z, acc[1][2]); acc[1][3] = fmaf(ww.y, x.w, acc[1][3]);
                    acc[2][0] = fmaf(ww.z, x.x, acc[2][0]); acc[2][1] = fmaf(ww.z, x.y, acc[2][1]);
                    acc[2][2] = fmaf(ww.z, x.z, acc[2][2]); acc[2][3] = fmaf(ww.z, x.w, acc[2][3]);
                    acc[3][0] = fmaf(ww.w, x.x, acc[3][0]); acc[3][1] = fmaf(ww.w, x.y, acc[3][1]);
                    acc[3][2] = fmaf(ww.w, x.z, acc[3][2]); acc[3][3] = fmaf(ww.w, x.w, acc[3][3]);
                }
            };
            float4 wa[KB], wb[KB];
            load_block(wa, k0);
            for (int kb = k0; kb < k1; kb += 2 * KB) {
                load_block(wb, kb + KB);
                fma_block(wa, kb);
                load_block(wa, kb + 2 * KB);
                if (kb + KB < k1) fma_block(wb, kb + KB);
            }
        }
        if (ksplit > 1) {                                   // tiles*ksplit <= NT here: single pass of the base loop
            if (active && ks > 0) {
                float4* sp = reinterpret_cast<float4*>(scratch + ((ks - 1) * tiles + tile) * 16);
#pragma unroll
                for (int c = 0; c < 4; ++c) sp[c] = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
            }
            __syncthreads();
            if (active && ks == 0) {
                for (int s = 1; s < ksplit; ++s) {
                    const float4* sp = reinterpret_cast<const float4*>(scratch + ((s - 1) * tiles + tile) * 16);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 v = sp[c];
                        acc[c][0] += v.x; acc[c][1] += v.y; acc[c][2] += v.z; acc[c][3] += v.w;
                    }
                }
            }
        }
        if (active && ks == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = 4 * ct + c;
                const float b = bias ? __ldg(bias + j) : 0.f;
                float4* op = reinterpret_cast<float4*>(out + j * RP + 4 * rt);
                float4 v = make_float4(acc[c][0] + b, acc[c][1] + b, acc[c][2] + b, acc[c][3] + b);
                if (mode == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                if (mode == 2) { const float4 o = *op; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                *op = v;
            }
        }
    }
    __syncthreads();
}

// global (rows x F, row-major, leading dim ld) -> smem feature-major tile; rows beyond n_valid are zero.
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

struct SmemPlan {
    int xin, c, x, hp, vsq, gi, gh, hn, q, alpha, scratch, total;     // offsets in floats
};
__host__ __device__ inline SmemPlan make_smem(const StepDims& d, bool backward) {
    SmemPlan p;
    int o = 0;
    auto take = [&](int rows) { int r = o; o += rows * RP; return r; };
    p.xin = take(d.Fin);
    p.c = take(d.tarmac() ? d.M : 0);          // c, x, hp contiguous: [c|x] feeds W_ih, [x|h] feeds the comm projections
    p.x = take(d.H);
    p.hp = take(d.H);
    p.vsq = take(d.tarmac() ? d.Vp() : 0);
    p.gi = take(3 * d.H);
    p.gh = take(3 * d.H);
    p.hn = take(d.H);
    p.q = take(d.Ap());
    p.alpha = take(d.tarmac() ? d.U : 0);      // alpha[i*RP + r]: weight of block-local source i for destination row r
    if (backward) o += 0;
    p.scratch = o; o += NT * 16;
    p.total = o;
    return p;
}

// ---------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(NT, 1) agent_step_fwd_kernel(const StepArgs a) {
    extern __shared__ __align__(16) float sm[];
    const StepDims d = a.d;
    const PackLayout L = make_layout(d);
    const SmemPlan P = make_smem(d, false);
    const int H = d.H, H3 = 3 * H, M = d.M, K = d.K, U = d.U, Vp = d.Vp(), A = d.A, Ap = d.Ap();
    const int rpt = d.rows_per_tile();
    const int64_t row0 = (int64_t)blockIdx.x * rpt;
    const int n_valid = (int)min((int64_t)rpt, a.N - row0);
    const float* pk = a.packed;
    float *sXin = sm + P.xin, *sC = sm + P.c, *sX = sm + P.x, *sHp = sm + P.hp, *sVSQ = sm + P.vsq, *sGI = sm + P.gi,
          *sGH = sm + P.gh, *sHn = sm + P.hn, *sQ = sm + P.q, *sAl = sm + P.alpha, *scratch = sm + P.scratch;
    const bool training = a.sv_gate != nullptr;
    const float scale = d.tarmac() ? 1.0f / (float)K : 0.f;

    load_tile(a.h0, row0, n_valid, H, H, sHp);
    for (int t = 0; t < a.T; ++t) {
        load_tile(a.xin + t * a.st_xin, row0, n_valid, d.Fin, d.Fin, d.aggr() ? sXin : sX);
        __syncthreads();
        if (d.aggr()) tile_gemm(pk + L.t_aggr, H, pk + L.b_aggr, sXin, d.Fin, sX, H, 1, scratch);
        if (d.tarmac()) {
            // [v | s | q] = W_vsq [x ‖ h]   (h enters detached: the backward never routes grad_vsq into h)
            tile_gemm(pk + L.t_vsq, Vp, pk + L.b_vsq, sX, 2 * H, sVSQ, Vp, 0, scratch);
            // scores e[i -> r] = <s_i, q_r> / key_size for sources i of r's env block
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
        // GRU projections: gi = W_ih [x ‖ c] + b_ih (smem order [c | x]), gh = W_hh h + b_hh
        tile_gemm(pk + L.t_ih, H3, pk + L.b_ih, d.tarmac() ? sC : sX, d.Iih(), sGI, H3, 0, scratch);
        tile_gemm(pk + L.t_hh, H3, pk + L.b_hh, sHp, H, sGH, H3, 0, scratch);
        for (int p = threadIdx.x; p < H * R; p += NT) {
            const int ch = p / R, r = p - ch * R;
            const float rr = sigmoidf_(sGI[ch * RP + r] + sGH[ch * RP + r]);
            const float zz = sigmoidf_(sGI[(H + ch) * RP + r] + sGH[(H + ch) * RP + r]);
            const float ghn = sGH[(2 * H + ch) * RP + r];
            const float nn = tanhf(fmaf(rr, ghn, sGI[(2 * H + ch) * RP + r]));
            sHn[ch * RP + r] = fmaf(zz, sHp[ch * RP + r] - nn, nn);
            if (training) {                    // reuse the gate tiles as [r | z | n | ghn] for the save below
                sGI[ch * RP + r] = rr; sGI[(H + ch) * RP + r] = zz; sGI[(2 * H + ch) * RP + r] = nn;
                sGH[ch * RP + r] = ghn;
            }
        }
        __syncthreads();
        tile_gemm(pk + L.t_out, Ap, pk + L.b_out, sHn, H, sQ, Ap, 0, scratch);
        // ---- write-back
        store_tile(a.h_out + t * a.st_h, row0, n_valid, H, H, sHn);
        store_tile(a.q + t * a.st_q, row0, n_valid, A, A, sQ);
        if (a.actions != nullptr && threadIdx.x < n_valid) {
            const int r = threadIdx.x;
            int best = 0;
            float bv = sQ[r];
            for (int c = 1; c < A; ++c) { const float v = sQ[c * RP + r]; if (v > bv) { bv = v; best = c; } }
            const int64_t ai = t * a.st_act + row0 + r;
            int64_t act = best;
            if (a.eg_u != nullptr && __ldg(a.eg_u + ai) <= __ldg(a.eg_eps)) act = __ldg(a.eg_a + ai);
            a.actions[ai] = act;
        }
        if (training) {
            const int64_t n = a.N;
            const int I = d.Iih();
            float* xc = a.sv_xc + (size_t)t * n * I;
            store_tile(xc, row0, n_valid, H, I, sX);                                  // [x | c] in PyTorch order
            if (d.tarmac()) {
                store_tile(xc + H, row0, n_valid, M, I, sC);
                store_tile(a.sv_vsq + (size_t)t * n * Vp, row0, n_valid, Vp, Vp, sVSQ);
                store_tile(a.sv_alpha + (size_t)t * n * U, row0, n_valid, U, U, sAl);
            }
            float* gt = a.sv_gate + (size_t)t * n * 4 * H;
            store_tile(gt, row0, n_valid, H3, 4 * H, sGI);
            store_tile(gt + H3, row0, n_valid, H, 4 * H, sGH);
        }
        __syncthreads();
        // next step: h_prev <- h_new
        for (int p = threadIdx.x; p < H * RP; p += NT) sHp[p] = sHn[p];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- backward
// Reverse-time walk of one row tile.  Per step t (T-1 .. 0), with dh = grad wrt h_t:
//   dh += dq_t W_out ; gates' -> dgi, dgh, dh_direct = dh * z ; dh_{t-1} = dh_direct + dgh W_hh ;
//   [dx | dc] = dgi W_ih ; attention' (dc -> dvsq) ; dx += dvsq W_vsq[:, :H] (h is detached there) ;
//   dpre = dx * 1[x > 0] ; d_xin = dpre W_aggr.
// dgi / dgh / dvsq / dpre are stashed; the parameter gradients are batched GEMMs over the whole sequence afterwards.
__global__ void __launch_bounds__(NT, 1) agent_step_bwd_kernel(const StepArgs a) {
    extern __shared__ __align__(16) float sm[];
    const StepDims d = a.d;
    const PackLayout L = make_layout(d);
    const int H = d.H, H3 = 3 * H, M = d.M, K = d.K, U = d.U, Vp = d.Vp(), A = d.A, Ap = d.Ap(), I = d.Iih();
    const int rpt = d.rows_per_tile();
    const int64_t row0 = (int64_t)blockIdx.x * rpt;
    const int n_valid = (int)min((int64_t)rpt, a.N - row0);
    const int64_t n = a.N;
    const float* pk = a.packed;
    // smem plan (feature-major tiles)
    int o = 0;
    auto take = [&](int rows) { float* p = sm + o; o += rows * RP; return p; };
    float* sDH = take(H);                         // grad wrt h_t (carry)
    float* sDQ = take(Ap);
    float* sDGI = take(H3);
    float* sDGH = take(H3);
    float* sDXC = take(I);                        // [dx | dc]   (PyTorch order)
    float* sVSQ = take(d.tarmac() ? Vp : 0);
    float* sDVSQ = take(d.tarmac() ? Vp : 0);
    float* sAl = take(d.tarmac() ? U : 0);
    float* sDS = take(d.tarmac() ? U : 0);
    float* sDPRE = take(d.aggr() ? H : 0);
    float* sDXIN = take(d.aggr() ? d.Fin : 0);
    float* scratch = sm + o;
    const float scale = d.tarmac() ? 1.0f / (float)K : 0.f;

    if (a.dh_last != nullptr) load_tile(a.dh_last, row0, n_valid, H, H, sDH);
    else for (int p = threadIdx.x; p < H * RP; p += NT) sDH[p] = 0.f;
    for (int t = a.T - 1; t >= 0; --t) {
        // dq_t (zero padded to Ap rows)
        for (int i = threadIdx.x; i < R * Ap; i += NT) {
            const int r = i / Ap, c = i - r * Ap;
            sDQ[c * RP + r] = (r < n_valid && c < A) ? __ldg(a.dq + ((size_t)t * n + row0 + r) * A + c) : 0.f;
        }
        __syncthreads();
        tile_gemm(pk + L.o_out, H, nullptr, sDQ, Ap, sDH, H, 2, scratch);
        // GRU gates backward (saved r, z, n, ghn; h_{t-1} from h_out[t-1] or h0)
        const float* gt = a.sv_gate + (size_t)t * n * 4 * H;
        const float* hprev = t > 0 ? a.h_out + (t - 1) * a.st_h : a.h0;
        for (int r = threadIdx.x >> 5; r < R; r += NT / 32)
        for (int ch = threadIdx.x & 31; ch < H; ch += 32) {
            float dr = 0.f, dz = 0.f, dn = 0.f, dnr = 0.f, dir = 0.f;
            if (r < n_valid) {
                const float* g = gt + (row0 + r) * 4 * H;
                const float rr = __ldg(g + ch), zz = __ldg(g + H + ch), nn = __ldg(g + 2 * H + ch), ghn = __ldg(g + H3 + ch);
                const float gv = sDH[ch * RP + r];
                dn = gv * (1.0f - zz) * (1.0f - nn * nn);
                dz = gv * (__ldg(hprev + (row0 + r) * H + ch) - nn) * zz * (1.0f - zz);
                dr = dn * ghn * rr * (1.0f - rr);
                dnr = dn * rr;
                dir = gv * zz;
            }
            sDGI[ch * RP + r] = dr; sDGI[(H + ch) * RP + r] = dz; sDGI[(2 * H + ch) * RP + r] = dn;
            sDGH[ch * RP + r] = dr; sDGH[(H + ch) * RP + r] = dz; sDGH[(2 * H + ch) * RP + r] = dnr;
            sDH[ch * RP + r] = dir;                      // becomes the carry after adding dgh W_hh
        }
        __syncthreads();
        store_tile(a.st_dgi + (size_t)t * n * H3, row0, n_valid, H3, H3, sDGI);
        store_tile(a.st_dgh + (size_t)t * n * H3, row0, n_valid, H3, H3, sDGH);
        tile_gemm(pk + L.o_hh, H, nullptr, sDGH, H3, sDH, H, 2, scratch);              // carry: grad wrt h_{t-1}
        tile_gemm(pk + L.o_ih, I, nullptr, sDGI, H3, sDXC, I, 0, scratch);             // [dx | dc]
        if (d.tarmac()) {
            float* sDC = sDXC + H * RP;
            load_tile(a.sv_vsq + (size_t)t * n * Vp, row0, n_valid, Vp, Vp, sVSQ);
            load_tile(a.sv_alpha + (size_t)t * n * U, row0, n_valid, U, U, sAl);
            __syncthreads();
            // d_alpha[i -> r] = <dc_r, v_i>
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
            // dvsq: rows [0,M) dv_u = sum_j alpha[u -> j] dc_j ; [M,M+K) ds_u = scale sum_j ds[u -> j] q_j ;
            //       [M+K,V) dq_r = scale sum_i ds[i -> r] s_i ; padding rows zero
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
            store_tile(a.st_dvsq + (size_t)t * n * Vp, row0, n_valid, Vp, Vp, sDVSQ);
            tile_gemm(pk + L.o_vsq, 2 * H, nullptr, sDVSQ, Vp, sDXC, H, 2, scratch);   // dx += dvsq W_vsq[:, :H]
        }
        if (d.aggr()) {
            const float* xc = a.sv_xc + (size_t)t * n * I;
            for (int r = threadIdx.x >> 5; r < R; r += NT / 32)
            for (int ch = threadIdx.x & 31; ch < H; ch += 32) {
                const bool on = r < n_valid && __ldg(xc + (row0 + r) * I + ch) > 0.f;
                sDPRE[ch * RP + r] = on ? sDXC[ch * RP + r] : 0.f;
            }
            __syncthreads();
            store_tile(a.st_dpre + (size_t)t * n * H, row0, n_valid, H, H, sDPRE);
            tile_gemm(pk + L.o_aggr, d.Fin, nullptr, sDPRE, H, sDXIN, d.Fin, 0, scratch);
            store_tile(a.d_xin + (size_t)t * n * d.Fin, row0, n_valid, d.Fin, d.Fin, sDXIN);
        } else {
            store_tile(a.d_xin + (size_t)t * n * d.Fin, row0, n_valid, H, d.Fin, sDXC);
        }
        __syncthreads();
    }
    if (a.d_h0 != nullptr) store_tile(a.d_h0, row0, n_valid, H, H, sDH);
}

static int bwd_smem_floats(const StepDims& d) {
    int rows = d.H + d.Ap() + 6 * d.H + d.Iih();
    if (d.tarmac()) rows += 2 * d.Vp() + 2 * d.U;
    if (d.aggr()) rows += d.H + d.Fin;
    return rows * RP + NT * 16;
}

int check_dims(const char* fn, const StepDims& d) {
    if (d.H < 16 || d.H % 4 || d.H > 256) { set_error("%s: hidden size %d unsupported (multiple of 4, 16..256)", fn, d.H); return 2; }
    if (d.A < 1 || d.A > 64) { set_error("%s: n_actions %d unsupported", fn, d.A); return 2; }
    if (d.tarmac() && (d.U < 1 || d.U > R)) { set_error("%s: agents per env must be 1..%d for the fused step (got %d)", fn, R, d.U); return 2; }
    if (d.tarmac() && (d.M < 4 || d.M % 4 || d.K < 1)) { set_error("%s: msg_size must be a multiple of 4", fn); return 2; }
    if (d.aggr() ? (d.Fin % 4 != 0 || d.Fin < 4) : (d.Fin != d.H)) { set_error("%s: bad input width %d", fn, d.Fin); return 2; }
    return 0;
}

StepDims mk_dims(int H, int M, int K, int A, int U, int Fin, int flags) {
    StepDims d;
    d.H = H; d.M = (flags & UBS_STEP_TARMAC) ? M : 0; d.K = (flags & UBS_STEP_TARMAC) ? K : 0; d.A = A;
    d.U = (flags & UBS_STEP_TARMAC) ? U : 1; d.Fin = Fin; d.flags = flags;
    return d;
}

}  // namespace ubs

extern "C" UBS_API int64_t ubs_agent_pack_size(int H, int M, int K, int A, int U, int Fin, int flags) {
    return ubs::make_layout(ubs::mk_dims(H, M, K, A, U, Fin, flags)).total;
}

extern "C" UBS_API int ubs_agent_pack(int H, int M, int K, int A, int U, int Fin, int flags,
                                      const float* W_aggr, const float* b_aggr, const float* W_val, const float* b_val,
                                      const float* W_sign, const float* b_sign, const float* W_que, const float* b_que,
                                      const float* W_ih, const float* b_ih, const float* W_hh, const float* b_hh,
                                      const float* W_out, const float* b_out, float* packed, void* stream) {
    ubs::PackArgs a{};
    a.d = ubs::mk_dims(H, M, K, A, U, Fin, flags);
    if (int rc = ubs::check_dims("ubs_agent_pack", a.d)) return rc;
    UBS_REQUIRE(W_ih && b_ih && W_hh && b_hh && W_out && b_out && packed, "ubs_agent_pack: NULL parameter");
    UBS_REQUIRE(!a.d.aggr() || (W_aggr && b_aggr), "ubs_agent_pack: aggregator parameters missing");
    UBS_REQUIRE(!a.d.tarmac() || (W_val && b_val && W_sign && b_sign && W_que && b_que), "ubs_agent_pack: TarMAC parameters missing");
    a.W_aggr = W_aggr; a.b_aggr = b_aggr; a.W_val = W_val; a.b_val = b_val; a.W_sign = W_sign; a.b_sign = b_sign;
    a.W_que = W_que; a.b_que = b_que; a.W_ih = W_ih; a.b_ih = b_ih; a.W_hh = W_hh; a.b_hh = b_hh;
    a.W_out = W_out; a.b_out = b_out; a.packed = packed;
    const int total = ubs::make_layout(a.d).total;
    ubs::agent_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_agent_pack");
}

extern "C" UBS_API int ubs_agent_act_uses_tma(int H, int M, int K, int A, int U, int Fin, int flags) {
    return ubs::agent_act_fits(ubs::mk_dims(H, M, K, A, U, Fin, flags)) ? 1 : 0;
}

extern "C" UBS_API int ubs_agent_seq_fwd(int H, int M, int K, int A, int U, int Fin, int flags, const float* packed,
                                         const float* xin, const float* h0, const uint32_t* mask,
                                         float* h_out, float* q, int64_t* actions,
                                         float* sv_xc, float* sv_vsq, float* sv_alpha, float* sv_gate,
                                         int64_t n_rows, int n_steps, void* stream) {
    return ubs_agent_act_fwd(H, M, K, A, U, Fin, flags, packed, xin, h0, mask, h_out, q, actions, nullptr, nullptr,
                             nullptr, sv_xc, sv_vsq, sv_alpha, sv_gate, n_rows, n_steps, stream);
}

extern "C" UBS_API int ubs_agent_act_fwd(int H, int M, int K, int A, int U, int Fin, int flags, const float* packed,
                                         const float* xin, const float* h0, const uint32_t* mask,
                                         float* h_out, float* q, int64_t* actions,
                                         const float* eg_u, const int64_t* eg_a, const float* eg_eps,
                                         float* sv_xc, float* sv_vsq, float* sv_alpha, float* sv_gate,
                                         int64_t n_rows, int n_steps, void* stream) {
    ubs::StepArgs a{};
    a.d = ubs::mk_dims(H, M, K, A, U, Fin, flags);
    if (int rc = ubs::check_dims("ubs_agent_seq_fwd", a.d)) return rc;
    UBS_REQUIRE(packed && xin && h0 && h_out && q, "ubs_agent_seq_fwd: NULL argument");
    UBS_REQUIRE(!a.d.tarmac() || mask, "ubs_agent_seq_fwd: TarMAC needs the block mask");
    UBS_REQUIRE(!a.d.tarmac() || n_rows % a.d.U == 0, "ubs_agent_seq_fwd: n_rows must be a multiple of agents per env");
    const bool training = sv_gate != nullptr;
    UBS_REQUIRE(!training || (sv_xc && (!a.d.tarmac() || (sv_vsq && sv_alpha))), "ubs_agent_seq_fwd: incomplete save buffers");
    if (n_rows == 0 || n_steps == 0) return 0;
    a.packed = packed; a.xin = xin; a.st_xin = n_rows * Fin; a.h0 = h0; a.mask = mask; a.st_mask = n_rows;
    a.h_out = h_out; a.st_h = n_rows * H; a.q = q; a.st_q = n_rows * A; a.actions = actions; a.st_act = n_rows;
    a.eg_u = eg_u; a.eg_a = eg_a; a.eg_eps = eg_eps;
    UBS_REQUIRE(eg_u == nullptr || (eg_a && eg_eps && actions), "ubs_agent_act_fwd: incomplete epsilon-greedy arguments");
    a.sv_xc = sv_xc; a.sv_vsq = sv_vsq; a.sv_alpha = sv_alpha; a.sv_gate = sv_gate; a.N = n_rows; a.T = n_steps;
    // inference: weights staged through shared memory by the copy engine when two layers + activations fit
    // (UBS_ACT_TMA=0 forces the register-streaming kernel, for A/B measurements)
    if (!training) {
        bool handled = false;
        const int rc = ubs::launch_agent_act(a, (cudaStream_t)stream, &handled);
        if (handled) return rc;
    }
    const int rpt = a.d.rows_per_tile();
    const unsigned grid = (unsigned)((n_rows + rpt - 1) / rpt);
    const size_t smem = (size_t)ubs::make_smem(a.d, false).total * sizeof(float);
    UBS_REQUIRE(smem <= 227 * 1024, "ubs_agent_seq_fwd: tile does not fit shared memory");
    UBS_OPT_IN_SMEM(ubs::agent_step_fwd_kernel, "ubs_agent_seq_fwd");
    ubs::agent_step_fwd_kernel<<<grid, ubs::NT, smem, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_agent_seq_fwd");
}

extern "C" UBS_API int ubs_agent_seq_bwd(int H, int M, int K, int A, int U, int Fin, int flags, const float* packed,
                                         const float* h0, const float* h_out, const float* sv_xc, const float* sv_vsq,
                                         const float* sv_alpha, const float* sv_gate, const float* dq,
                                         const float* dh_last, float* d_xin, float* d_h0, float* st_dgi, float* st_dgh,
                                         float* st_dvsq, float* st_dpre, int64_t n_rows, int n_steps, void* stream) {
    ubs::StepArgs a{};
    a.d = ubs::mk_dims(H, M, K, A, U, Fin, flags);
    if (int rc = ubs::check_dims("ubs_agent_seq_bwd", a.d)) return rc;
    UBS_REQUIRE(packed && h0 && h_out && sv_xc && sv_gate && dq && d_xin && st_dgi && st_dgh, "ubs_agent_seq_bwd: NULL argument");
    UBS_REQUIRE(!a.d.tarmac() || (sv_vsq && sv_alpha && st_dvsq), "ubs_agent_seq_bwd: TarMAC buffers missing");
    UBS_REQUIRE(!a.d.aggr() || st_dpre, "ubs_agent_seq_bwd: aggregator stash missing");
    if (n_rows == 0 || n_steps == 0) return 0;
    a.packed = packed; a.h0 = h0; a.h_out = const_cast<float*>(h_out); a.st_h = n_rows * H;
    a.sv_xc = const_cast<float*>(sv_xc); a.sv_vsq = const_cast<float*>(sv_vsq);
    a.sv_alpha = const_cast<float*>(sv_alpha); a.sv_gate = const_cast<float*>(sv_gate);
    a.dq = dq; a.dh_last = dh_last; a.d_xin = d_xin; a.d_h0 = d_h0;
    a.st_dgi = st_dgi; a.st_dgh = st_dgh; a.st_dvsq = st_dvsq; a.st_dpre = st_dpre; a.N = n_rows; a.T = n_steps;
    const int rpt = a.d.rows_per_tile();
    const unsigned grid = (unsigned)((n_rows + rpt - 1) / rpt);
    const size_t smem = (size_t)ubs::bwd_smem_floats(a.d) * sizeof(float);
    UBS_REQUIRE(smem <= 227 * 1024, "ubs_agent_seq_bwd: tile does not fit shared memory");
    UBS_OPT_IN_SMEM(ubs::agent_step_bwd_kernel, "ubs_agent_seq_bwd");
    ubs::agent_step_bwd_kernel<<<grid, ubs::NT, smem, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_agent_seq_bwd");
}
