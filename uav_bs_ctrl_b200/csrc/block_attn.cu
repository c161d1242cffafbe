// TarMAC targeted attention over block-diagonal comm graphs.
//
// Replaces the edge ops of reference TarMAC.forward (algos/madrqn/agents/gnn_agents.py:261-266):
//   apply_edges(u_dot_v('s','q','e')); e / key_size; edge_softmax; update_all(u_mul_e('v','a','m'), sum('m','c'))
// (DGL: gSDDMM(dot) + 5 softmax kernels + gSpMM) with one kernel forward and two backward.
//
// The `talk` relation of a batch of envs is block-diagonal with blocks of U agents
// (algos/madrqn/utils/env_wrappers.py:139-154 + algos/common.py:40-47), so the in-edge set of a destination is
// a U-bit mask and attention rows are dense length-U vectors: no index arrays, no atomics, deterministic.
//   forward : one warp per destination; lane i scores source i of the block; softmax by shuffles;
//             c_v = sum_i alpha_i val_i with lanes over the message channels (coalesced row reads).
//   backward: pass 1 (warp per destination) d_alpha -> d_score (written dense), grad_q;
//             pass 2 (warp per source) gathers its column of alpha / d_score: grad_val, grad_s.
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {

struct AttnArgs {
    const float* s; const float* q; const float* val; const uint32_t* mask;
    int64_t ld_s, ld_q, ld_v;
    float* c; float* alpha;
    const float* grad_c; float* grad_s; float* grad_q; float* grad_val; float* ds;
    int64_t ld_gs, ld_gq, ld_gv;
    int n, U, K, M; float scale;
};

__global__ void __launch_bounds__(256) block_attn_fwd_kernel(const AttnArgs a) {
    const int lane = threadIdx.x % 32;
    const int v = blockIdx.x * 8 + threadIdx.x / 32;
    if (v >= a.n) return;
    const int b0 = (v / a.U) * a.U;
    const uint32_t mk = __ldg(a.mask + v);
    const bool valid = lane < a.U && ((mk >> lane) & 1u);
    float score = -CUDART_INF_F;
    if (valid) {
        const float* sp = a.s + (size_t)(b0 + lane) * a.ld_s;
        const float* qp = a.q + (size_t)v * a.ld_q;
        float acc = 0.f;
        for (int k = 0; k < a.K; ++k) acc = fmaf(__ldg(sp + k), __ldg(qp + k), acc);
        score = acc * a.scale;
    }
    const float mx = warp_max(score);
    const float p = valid ? expf(score - mx) : 0.f;
    const float den = warp_sum(p);
    const float al = den > 0.f ? p / den : 0.f;
    if (lane < a.U) a.alpha[(size_t)v * a.U + lane] = al;
    for (int mb = 0; mb < a.M; mb += 32) {                      // warp-uniform trip count: shuffles inside
        const int m0 = mb + lane;
        float acc = 0.f;
        for (int i = 0; i < a.U; ++i) {
            const float ai = __shfl_sync(0xffffffffu, al, i);
            if (ai != 0.f && m0 < a.M) acc = fmaf(ai, __ldg(a.val + (size_t)(b0 + i) * a.ld_v + m0), acc);
        }
        if (m0 < a.M) a.c[(size_t)v * a.M + m0] = acc;
    }
}

// pass 1: per destination v
__global__ void __launch_bounds__(256) block_attn_bwd_dst_kernel(const AttnArgs a) {
    const int lane = threadIdx.x % 32;
    const int v = blockIdx.x * 8 + threadIdx.x / 32;
    if (v >= a.n) return;
    const int b0 = (v / a.U) * a.U;
    const float al = lane < a.U ? __ldg(a.alpha + (size_t)v * a.U + lane) : 0.f;
    // d_alpha_i = <grad_c[v], val_i>
    float da = 0.f;
    for (int i = 0; i < a.U; ++i) {
        float part = 0.f;
        for (int m0 = lane; m0 < a.M; m0 += 32)
            part = fmaf(__ldg(a.grad_c + (size_t)v * a.M + m0), __ldg(a.val + (size_t)(b0 + i) * a.ld_v + m0), part);
        part = warp_sum(part);
        if (lane == i) da = part;
    }
    const float sum_ad = warp_sum(al * da);
    const float ds = al * (da - sum_ad);                        // d score (before the 1/key_size scale)
    if (lane < a.U) a.ds[(size_t)v * a.U + lane] = ds;
    // grad_q[v] = scale * sum_i ds_i s_i
    for (int kb = 0; kb < a.K; kb += 32) {
        const int k = kb + lane;
        float acc = 0.f;
        for (int i = 0; i < a.U; ++i) {
            const float di = __shfl_sync(0xffffffffu, ds, i);
            if (k < a.K) acc = fmaf(di, __ldg(a.s + (size_t)(b0 + i) * a.ld_s + k), acc);
        }
        if (k < a.K) a.grad_q[(size_t)v * a.ld_gq + k] = acc * a.scale;
    }
}

// pass 2: per source u (block-local index i): column i of alpha / ds over the destinations of the block
__global__ void __launch_bounds__(256) block_attn_bwd_src_kernel(const AttnArgs a) {
    const int lane = threadIdx.x % 32;
    const int u = blockIdx.x * 8 + threadIdx.x / 32;
    if (u >= a.n) return;
    const int b0 = (u / a.U) * a.U, i = u - b0;
    float al = 0.f, ds = 0.f;
    if (lane < a.U) {
        al = __ldg(a.alpha + (size_t)(b0 + lane) * a.U + i);
        ds = a.ds[(size_t)(b0 + lane) * a.U + i];
    }
    for (int mb = 0; mb < a.M; mb += 32) {
        const int m0 = mb + lane;
        float acc = 0.f;
        for (int j = 0; j < a.U; ++j) {
            const float aj = __shfl_sync(0xffffffffu, al, j);
            if (m0 < a.M) acc = fmaf(aj, __ldg(a.grad_c + (size_t)(b0 + j) * a.M + m0), acc);
        }
        if (m0 < a.M) a.grad_val[(size_t)u * a.ld_gv + m0] = acc;
    }
    for (int kb = 0; kb < a.K; kb += 32) {
        const int k = kb + lane;
        float acc = 0.f;
        for (int j = 0; j < a.U; ++j) {
            const float dj = __shfl_sync(0xffffffffu, ds, j);
            if (k < a.K) acc = fmaf(dj, __ldg(a.q + (size_t)(b0 + j) * a.ld_q + k), acc);
        }
        if (k < a.K) a.grad_s[(size_t)u * a.ld_gs + k] = acc * a.scale;
    }
}

static int check_attn(const char* fn, int64_t n, int block, int K, int M) {
    if (block < 1 || block > 32) { set_error("%s: block must be in 1..32 (got %d)", fn, block); return 2; }
    if (n % block != 0) { set_error("%s: n_nodes (%lld) is not a multiple of block (%d)", fn, (long long)n, block); return 2; }
    if (K < 1 || M < 1 || n >= (1ll << 31)) { set_error("%s: bad sizes", fn); return 2; }
    return 0;
}

}  // namespace ubs

extern "C" UBS_API int ubs_block_attn_fwd(const float* s, int64_t ld_s, const float* q, int64_t ld_q, const float* val,
                                  int64_t ld_v, const uint32_t* mask, float* c, float* alpha, int64_t n_nodes,
                                  int block, int key_size, int msg_size, float scale, void* stream) {
    if (int rc = ubs::check_attn("ubs_block_attn_fwd", n_nodes, block, key_size, msg_size)) return rc;
    if (n_nodes == 0) return 0;
    ubs::AttnArgs a{};
    a.s = s; a.q = q; a.val = val; a.mask = mask; a.ld_s = ld_s; a.ld_q = ld_q; a.ld_v = ld_v;
    a.c = c; a.alpha = alpha; a.n = (int)n_nodes; a.U = block; a.K = key_size; a.M = msg_size; a.scale = scale;
    ubs::block_attn_fwd_kernel<<<(unsigned)((n_nodes + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_block_attn_fwd");
}

extern "C" UBS_API int ubs_block_attn_bwd(const float* s, int64_t ld_s, const float* q, int64_t ld_q, const float* val,
                                  int64_t ld_v, const uint32_t* mask, const float* alpha, const float* grad_c,
                                  float* grad_s, int64_t ld_gs, float* grad_q, int64_t ld_gq, float* grad_val,
                                  int64_t ld_gv, float* ds_work, int64_t n_nodes, int block, int key_size,
                                  int msg_size, float scale, void* stream) {
    if (int rc = ubs::check_attn("ubs_block_attn_bwd", n_nodes, block, key_size, msg_size)) return rc;
    if (n_nodes == 0) return 0;
    ubs::AttnArgs a{};
    a.s = s; a.q = q; a.val = val; a.mask = mask; a.ld_s = ld_s; a.ld_q = ld_q; a.ld_v = ld_v;
    a.alpha = const_cast<float*>(alpha); a.grad_c = grad_c;
    a.grad_s = grad_s; a.grad_q = grad_q; a.grad_val = grad_val; a.ds = ds_work;
    a.ld_gs = ld_gs; a.ld_gq = ld_gq; a.ld_gv = ld_gv;
    a.n = (int)n_nodes; a.U = block; a.K = key_size; a.M = msg_size; a.scale = scale;
    const unsigned grid = (unsigned)((n_nodes + 7) / 8);
    ubs::block_attn_bwd_dst_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    if (int rc = ubs::check_launch("ubs_block_attn_bwd(dst)")) return rc;
    ubs::block_attn_bwd_src_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_block_attn_bwd(src)");
}
