// Shared declarations of the fused agent-step kernels (agent_step.cu, agent_act.cu).
#pragma once
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {

constexpr int R = 16;     // agent rows per CTA tile
constexpr int RP = 20;    // padded row stride of feature-major smem tiles (16-byte aligned)

struct StepDims {
    int H, M, K, A, U, Fin, flags;
    __host__ __device__ int V() const { return M + 2 * K; }                 // [v | s | q]
    __host__ __device__ int Vp() const { return (V() + 3) & ~3; }
    __host__ __device__ int Ap() const { return (A + 3) & ~3; }
    __host__ __device__ bool aggr() const { return flags & UBS_STEP_AGGR; }
    __host__ __device__ bool tarmac() const { return flags & UBS_STEP_TARMAC; }
    __host__ __device__ int Iih() const { return tarmac() ? H + M : H; }    // GRU input width
    __host__ __device__ int rows_per_tile() const { return tarmac() ? (R / U) * U : R; }
    __host__ __device__ int Ap8() const { return (A + 7) & ~7; }
    // every contraction / output width a multiple of 8: the layers can run as mma.sync m16n8k8 tiles (agent_act.cu)
    __host__ __device__ bool mma_ok() const {
        return H % 8 == 0 && Fin % 8 == 0 && (!tarmac() || (M % 8 == 0 && Vp() % 8 == 0));
    }
};

// Packed parameter buffer (floats).  "t_*" = transposed (K-major) copies for the forward GEMMs, "o_*" = original
// row-major (out, in) copies which are K-major for the backward products.  Every offset is a multiple of 4.
struct PackLayout {
    int t_aggr, b_aggr, t_vsq, b_vsq, t_ih, b_ih, t_hh, b_hh, t_out, b_out;
    int o_aggr, o_vsq, o_ih, o_hh, o_out;
    // "f_*": the K-major matrices again in mma FRAGMENT ORDER ([k-step][8-column tile][lane][2] = W[8ks + lane%4 (+4)][8j +
    // lane/4]: one conflict-free 8-byte shared load per B operand), each followed by a copy of its bias ("fb_*"); the Q head
    // is padded to a multiple of 8 columns.  Present only when d.mma_ok().
    int f_aggr, fb_aggr, f_vsq, fb_vsq, f_ih, fb_ih, f_hh, fb_hh, f_out, fb_out;
    int total;
};

__host__ __device__ inline PackLayout make_layout(const StepDims& d) {
    PackLayout L;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
    const int H = d.H, H3 = 3 * d.H;
    L.t_aggr = take(d.aggr() ? d.Fin * H : 0);       L.b_aggr = take(d.aggr() ? H : 0);
    L.t_vsq = take(d.tarmac() ? 2 * H * d.Vp() : 0);  L.b_vsq = take(d.tarmac() ? d.Vp() : 0);
    L.t_ih = take(d.Iih() * H3);                      L.b_ih = take(H3);
    L.t_hh = take(H * H3);                            L.b_hh = take(H3);
    L.t_out = take(H * d.Ap());                       L.b_out = take(d.Ap());
    L.o_aggr = take(d.aggr() ? H * d.Fin : 0);
    L.o_vsq = take(d.tarmac() ? d.Vp() * 2 * H : 0);
    L.o_ih = take(H3 * d.Iih());
    L.o_hh = take(H3 * H);
    L.o_out = take(d.Ap() * H);
    const bool mm = d.mma_ok();
    L.f_aggr = take(mm && d.aggr() ? d.Fin * H : 0);          L.fb_aggr = take(mm && d.aggr() ? H : 0);
    L.f_vsq = take(mm && d.tarmac() ? 2 * H * d.Vp() : 0);    L.fb_vsq = take(mm && d.tarmac() ? d.Vp() : 0);
    L.f_ih = take(mm ? d.Iih() * H3 : 0);                     L.fb_ih = take(mm ? H3 : 0);
    L.f_hh = take(mm ? H * H3 : 0);                           L.fb_hh = take(mm ? H3 : 0);
    L.f_out = take(mm ? H * d.Ap8() : 0);                     L.fb_out = take(mm ? d.Ap8() : 0);
    L.total = o;
    return L;
}

// Observation relations folded into the act step (agent_act.cu): the two GATv2 star relations of
// GraphObservationEncoder (gnn_agents.py:92-97,103-104) read straight from an observation packet.
struct RelIn {
    const float* relpack;        // per-relation tables of ubs_gatv2_rel_pack, relation 0 then relation 1 (NULL: not fused)
    const float* x_src[2];       // compacted source rows (star layout: row id == CSR slot)
    const int* indptr[2];        // (n_rows + 1) cumulative in-degrees
    int FS[2];                   // source feature widths (1..4)
    int cap[2];                  // max in-degree of a destination (G, U-1): bounds the shared-memory staging area
    const float* x_dst; int F_d; // (n_rows, F_d) destination (agent) features, F_d <= 2
    int heads; int gat_flags;    // UBS_GAT_* flags
};

struct StepArgs {
    StepDims d;
    RelIn rel;
    const float* packed;
    // forward, sequence-strided: tensor[t] = base + t * stride (strides in floats / elements)
    const float* xin;  int64_t st_xin;      // (T, N, Fin)
    const float* h0;                        // (N, H) hidden state entering step 0
    const uint32_t* mask; int64_t st_mask;  // (T, N)
    float* h_out; int64_t st_h;             // (T, N, H)   hidden state after each step
    float* q; int64_t st_q;                 // (T, N, A)
    int64_t* actions; int64_t st_act;       // (T, N) argmax, nullable
    const float* eg_u; const int64_t* eg_a; const float* eg_eps;   // epsilon-greedy: action = u <= *eps ? a : argmax
    // saved for backward (nullable => inference)
    float* sv_xc;   // (T, N, Iih)  [x | c]   (PyTorch GRU input order)
    float* sv_vsq;  // (T, N, Vp)
    float* sv_alpha;// (T, N, U)
    float* sv_gate; // (T, N, 4H)   r | z | n | (W_hn h + b_hn)
    int64_t N; int T;
    // backward
    const float* dq;       // (T, N, A)
    const float* dh_last;  // (N, H) gradient flowing into the last hidden state, nullable
    float* d_xin;          // (T, N, Fin)
    float* d_h0;           // (N, H), nullable
    float* st_dgi;         // (T, N, 3H)  stash for the batched weight-gradient GEMMs
    float* st_dgh;         // (T, N, 3H)
    float* st_dvsq;        // (T, N, Vp)
    float* st_dpre;        // (T, N, H)   grad of the aggregator pre-activation
    // launch option (host side): programmatic dependent launch of the act kernel (UBS_ACT_PDL)
    int pdl;
};

int launch_agent_act(const StepArgs& a, cudaStream_t st, bool* handled);   // agent_act.cu
bool agent_act_fits(const StepDims& d);
StepDims mk_dims(int H, int M, int K, int A, int U, int Fin, int flags);   // agent_step.cu
int check_dims(const char* fn, const StepDims& d);

}  // namespace ubs
