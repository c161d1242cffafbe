// Mean aggregation over block-diagonal comm graphs: the reduce step of the reference's BaseComm / CommNet protocols
// (algos/madrqn/agents/gnn_agents.py:125-133,144 and :205-214,224: update_all(udf_msg, udf_reduce) with
// `nodes.mailbox['m'].mean(1)`; DGL runs it as a degree-bucketed gather + mean, zeros for nodes without messages).
//
// Messages of these protocols depend on the SOURCE node only, and the `talk` relation of a batch of envs is
// block-diagonal with blocks of U agents (env_wrappers.py:139-154 + algos/common.py:40-47), so
//   c_v = (1 / deg_v) * sum_{u in mask_v} msg[u]          (0 when deg_v = 0)
// is a U-term sum selected by the destination's bit mask: no edge list, no atomics, deterministic in both directions.
//   forward : thread per (destination, feature), sources in index order, coalesced over features
//   backward: thread per (source, feature) gathers grad_out[v] / deg_v over the destinations v that list it
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {

__global__ void __launch_bounds__(256) block_mean_fwd_kernel(const float* __restrict__ msg, int64_t ld_msg,
                                                             const uint32_t* __restrict__ mask, float* __restrict__ out,
                                                             int64_t ld_out, int64_t n, int U, int F) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n * F) return;
    const int64_t v = i / F;
    const int f = (int)(i - v * F);
    const int64_t b0 = (v / U) * U;
    const uint32_t mk = __ldg(mask + v);
    float acc = 0.f;
    for (int u = 0; u < U; ++u)
        if ((mk >> u) & 1u) acc += __ldg(msg + (b0 + u) * ld_msg + f);
    const int deg = __popc(mk);
    out[v * ld_out + f] = deg > 0 ? acc / (float)deg : 0.f;
}

__global__ void __launch_bounds__(256) block_mean_bwd_kernel(const float* __restrict__ grad_out, int64_t ld_go,
                                                             const uint32_t* __restrict__ mask, float* __restrict__ grad_msg,
                                                             int64_t ld_gm, int64_t n, int U, int F) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n * F) return;
    const int64_t u = i / F;
    const int f = (int)(i - u * F);
    const int64_t b0 = (u / U) * U;
    const int ul = (int)(u - b0);
    float acc = 0.f;
    for (int v = 0; v < U; ++v) {
        const uint32_t mk = __ldg(mask + b0 + v);
        if ((mk >> ul) & 1u) acc += __ldg(grad_out + (b0 + v) * ld_go + f) / (float)__popc(mk);
    }
    grad_msg[u * ld_gm + f] = acc;
}

}  // namespace ubs

extern "C" UBS_API int ubs_block_mean_fwd(const float* msg, int64_t ld_msg, const uint32_t* mask, float* out, int64_t ld_out,
                                          int64_t n, int block, int F, void* stream) {
    UBS_REQUIRE(msg && mask && out && n >= 0, "ubs_block_mean_fwd: NULL argument");
    UBS_REQUIRE(block >= 1 && block <= 32 && n % block == 0, "ubs_block_mean_fwd: block must be in [1, 32] and divide n");
    UBS_REQUIRE(F >= 1 && ld_msg >= F && ld_out >= F && n * (int64_t)F < (1ll << 40), "ubs_block_mean_fwd: bad sizes");
    if (n == 0) return 0;
    const int64_t blocks = (n * F + 255) / 256;
    ubs::block_mean_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(msg, ld_msg, mask, out, ld_out, n, block, F);
    return ubs::check_launch("ubs_block_mean_fwd");
}

extern "C" UBS_API int ubs_block_mean_bwd(const float* grad_out, int64_t ld_go, const uint32_t* mask, float* grad_msg,
                                          int64_t ld_gm, int64_t n, int block, int F, void* stream) {
    UBS_REQUIRE(grad_out && mask && grad_msg && n >= 0, "ubs_block_mean_bwd: NULL argument");
    UBS_REQUIRE(block >= 1 && block <= 32 && n % block == 0, "ubs_block_mean_bwd: block must be in [1, 32] and divide n");
    UBS_REQUIRE(F >= 1 && ld_go >= F && ld_gm >= F, "ubs_block_mean_bwd: bad sizes");
    if (n == 0) return 0;
    const int64_t blocks = (n * F + 255) / 256;
    ubs::block_mean_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, ld_go, mask, grad_msg, ld_gm, n, block, F);
    return ubs::check_launch("ubs_block_mean_bwd");
}
