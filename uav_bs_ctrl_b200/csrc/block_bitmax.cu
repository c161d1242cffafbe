// DiscreteComm message passing over block-diagonal comm graphs (reference algos/madrqn/agents/gnn_agents.py:151-193):
//   per EDGE u->v and bit k:  y = softmax((f_enc([x_u ‖ h_u])[k, 0:2] + gumbel[e, k, 0:2]) / tau)   (F.gumbel_softmax, :173)
//                             m[e, k, :] = one_hot(argmax y) - y.detach() + y                      (hard=True: exactly 0 / 1)
//   per destination v:        c_v = max over in-edges of m  (element-wise OR, :176-179), zeros without in-edges
// DGL runs this as an edge UDF (E x 2M logits gathered per edge) + a degree-bucketed mailbox max.  Here the encoder
// output stays per NODE (N x 2M), the noise is read per edge in the reference's edge-id order, and the max is a U-term
// scan selected by the destination's bit mask — no edge list, no mailbox, no atomics.
//
// Edge ids.  The reference builds the comm graph src-major (env_wrappers.py:141-144) and dgl.batch concatenates the
// envs, so edge (i -> j) of env b has id  eoff[b] + sum_{i' < i} outdeg(i') + #{j' < j : i -> j'}; the out-neighbour
// sets come from the destinations' masks (bit i of mask[j'] <=> edge i -> j').
//
// Gradient of the max: torch.max(dim) gives ONE winner per (destination, feature) — the first maximal mailbox entry,
// i.e. the lowest source index (mailbox rows are in edge-id order); messages are exactly 0 / 1, so ties are the rule.
// The forward records the winner's local source index; the backward (thread per (source, bit), deterministic) sends
// grad_c through the straight-through term y of the winning edges only and through the 2-class softmax.
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {

struct BitMaxArgs {
    const float* logits; int64_t ld_logits;   // (N, 2M) f_enc output per node
    const float* expo;                        // (E, M, 2) Exponential(1) draws in edge-id order (gumbel = -log)
    const uint32_t* mask;                     // (N) in-neighbour bit masks
    const int64_t* eoff;                      // (N / U) first edge id of every env
    float* out; int64_t ld_out;               // (N, 2M)
    uint8_t* winner;                          // (N, 2M) local source index of the winning edge, 255 = none
    const float* grad_out; int64_t ld_go;
    float* grad_logits; int64_t ld_gl;
    int64_t n; int U, M; float inv_tau;
};

// bit set of the destinations that source `i` of the env starting at row b0 sends to
__device__ __forceinline__ uint32_t out_bits(const uint32_t* __restrict__ mask, int64_t b0, int U, int i) {
    uint32_t ob = 0;
    for (int j = 0; j < U; ++j) ob |= ((__ldg(mask + b0 + j) >> i) & 1u) << j;
    return ob;
}

// straight-through message of edge e, bit k: class probabilities y0, y1 and the hard choice
__device__ __forceinline__ int edge_bits(const BitMaxArgs& a, int64_t u, int64_t e, int k, float& y0, float& y1) {
    const float2 l = __ldg(reinterpret_cast<const float2*>(a.logits + u * a.ld_logits) + k);
    const float2 x = __ldg(reinterpret_cast<const float2*>(a.expo + (e * a.M + k) * 2));
    const float z0 = (l.x - logf(x.x)) * a.inv_tau, z1 = (l.y - logf(x.y)) * a.inv_tau;
    const float mx = fmaxf(z0, z1);
    const float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
    const float s = e0 + e1;
    y0 = e0 / s; y1 = e1 / s;
    return y1 > y0 ? 1 : 0;                    // torch's argmax: first index on ties
}

__global__ void __launch_bounds__(256) block_bitmax_fwd_kernel(const BitMaxArgs a) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= a.n * a.M) return;
    const int64_t v = idx / a.M;
    const int k = (int)(idx - v * a.M), U = a.U;
    const int64_t env = v / U, b0 = env * U;
    const int j = (int)(v - b0);
    const uint32_t mk = __ldg(a.mask + v);
    int win0 = 255, win1 = 255, first = 255;
    int64_t ebase = __ldg(a.eoff + env);
    for (int i = 0; i < U; ++i) {
        const uint32_t ob = out_bits(a.mask, b0, U, i);
        if ((mk >> i) & 1u) {
            const int64_t e = ebase + __popc(ob & ((1u << j) - 1u));
            float y0, y1;
            const int c = edge_bits(a, b0 + i, e, k, y0, y1);
            if (first == 255) first = i;
            if (c == 0 && win0 == 255) win0 = i;
            if (c == 1 && win1 == 255) win1 = i;
        }
        ebase += __popc(ob);
    }
    // no edge chose the class: the max over the mailbox is 0 and its first row (lowest source) is the winner
    float* o = a.out + v * a.ld_out + 2 * k;
    o[0] = win0 != 255 ? 1.f : 0.f;
    o[1] = win1 != 255 ? 1.f : 0.f;
    if (a.winner != nullptr) {
        a.winner[v * 2 * a.M + 2 * k] = (uint8_t)(win0 != 255 ? win0 : first);
        a.winner[v * 2 * a.M + 2 * k + 1] = (uint8_t)(win1 != 255 ? win1 : first);
    }
}

__global__ void __launch_bounds__(256) block_bitmax_bwd_kernel(const BitMaxArgs a) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= a.n * a.M) return;
    const int64_t u = idx / a.M;
    const int k = (int)(idx - u * a.M), U = a.U;
    const int64_t env = u / U, b0 = env * U;
    const int i = (int)(u - b0);
    int64_t ebase = __ldg(a.eoff + env);
    for (int ii = 0; ii < i; ++ii) ebase += __popc(out_bits(a.mask, b0, U, ii));
    const uint32_t ob = out_bits(a.mask, b0, U, i);
    float g0 = 0.f, g1 = 0.f;
    for (int j = 0; j < U; ++j) {
        if (!((ob >> j) & 1u)) continue;
        const int64_t v = b0 + j;
        const uint8_t* w = a.winner + v * 2 * a.M + 2 * k;
        const bool t0 = w[0] == i, t1 = w[1] == i;
        if (!(t0 || t1)) continue;
        const int64_t e = ebase + __popc(ob & ((1u << j) - 1u));
        float y0, y1;
        edge_bits(a, u, e, k, y0, y1);
        const float2 go = __ldg(reinterpret_cast<const float2*>(a.grad_out + v * a.ld_go) + k);
        const float d0 = t0 ? go.x : 0.f, d1 = t1 ? go.y : 0.f;          // gradient reaching y of this edge
        const float dot = y0 * d0 + y1 * d1;                            // softmax backward, then the 1 / tau of the logits
        g0 += a.inv_tau * y0 * (d0 - dot);
        g1 += a.inv_tau * y1 * (d1 - dot);
    }
    *reinterpret_cast<float2*>(a.grad_logits + u * a.ld_gl + 2 * k) = make_float2(g0, g1);
}

}  // namespace ubs

static int bitmax_check(const char* fn, const void* p0, const void* p1, const void* p2, const void* p3, int64_t n, int block, int M,
                        float tau, int64_t ld) {
    UBS_REQUIRE(p0 && p1 && p2 && p3 && n >= 0, "%s: NULL argument", fn);
    UBS_REQUIRE(block >= 1 && block <= 32 && n % block == 0, "%s: block must be in [1, 32] and divide n", fn);
    UBS_REQUIRE(M >= 1 && ld >= 2 * M && ld % 2 == 0 && tau > 0.f && n * (int64_t)M < (1ll << 40), "%s: bad sizes", fn);
    return 0;
}

extern "C" UBS_API int ubs_block_bitmax_fwd(const float* logits, int64_t ld_logits, const float* expo, const uint32_t* mask,
                                            const int64_t* env_edge_offset, float* out, int64_t ld_out, uint8_t* winner,
                                            int64_t n, int block, int msg_size, float tau, void* stream) {
    if (int rc = bitmax_check("ubs_block_bitmax_fwd", logits, expo, mask, env_edge_offset, n, block, msg_size, tau, ld_logits)) return rc;
    UBS_REQUIRE(out && ld_out >= 2 * msg_size, "ubs_block_bitmax_fwd: bad output");
    UBS_REQUIRE(((uintptr_t)logits % 8) == 0 && ((uintptr_t)expo % 8) == 0, "ubs_block_bitmax_fwd: logits / noise must be 8-byte aligned");
    if (n == 0) return 0;
    ubs::BitMaxArgs a{};
    a.logits = logits; a.ld_logits = ld_logits; a.expo = expo; a.mask = mask; a.eoff = env_edge_offset; a.out = out;
    a.ld_out = ld_out; a.winner = winner; a.n = n; a.U = block; a.M = msg_size; a.inv_tau = 1.0f / tau;
    ubs::block_bitmax_fwd_kernel<<<(unsigned)((n * msg_size + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_block_bitmax_fwd");
}

extern "C" UBS_API int ubs_block_bitmax_bwd(const float* logits, int64_t ld_logits, const float* expo, const uint32_t* mask,
                                            const int64_t* env_edge_offset, const uint8_t* winner, const float* grad_out,
                                            int64_t ld_go, float* grad_logits, int64_t ld_gl, int64_t n, int block,
                                            int msg_size, float tau, void* stream) {
    if (int rc = bitmax_check("ubs_block_bitmax_bwd", logits, expo, mask, env_edge_offset, n, block, msg_size, tau, ld_logits)) return rc;
    UBS_REQUIRE(winner && grad_out && grad_logits && ld_go >= 2 * msg_size && ld_gl >= 2 * msg_size && ld_go % 2 == 0 && ld_gl % 2 == 0,
                "ubs_block_bitmax_bwd: bad gradient buffers");
    if (n == 0) return 0;
    ubs::BitMaxArgs a{};
    a.logits = logits; a.ld_logits = ld_logits; a.expo = expo; a.mask = mask; a.eoff = env_edge_offset;
    a.winner = const_cast<uint8_t*>(winner); a.grad_out = grad_out; a.ld_go = ld_go; a.grad_logits = grad_logits; a.ld_gl = ld_gl;
    a.n = n; a.U = block; a.M = msg_size; a.inv_tau = 1.0f / tau;
    ubs::block_bitmax_bwd_kernel<<<(unsigned)((n * msg_size + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_block_bitmax_bwd");
}
