// nn.GRUCell gate math (reference call sites algos/madrqn/agents/gnn_agents.py:29,55,246,270;
// algos/drqn/agents/gnn_agents.py:20,28), gate order (r, z, n), SURVEY.md Appendix A.3:
//   r = sigma(gi_r + gh_r), z = sigma(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1 - z) n + z h
// given the two projections gi = W_ih x + b_ih and gh = W_hh h + b_hh.  Pure streaming kernels (HBM/L2 bound):
// one thread per (row, channel), consecutive threads on consecutive channels => coalesced 128-byte rows.
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {

__global__ void __launch_bounds__(256) gru_gates_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                           const float* __restrict__ h, float* __restrict__ out,
                                                           int64_t n, int H) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n * H) return;
    const int64_t row = idx / H;
    const int c = (int)(idx - row * H);
    const float* gir = gi + row * 3 * H;
    const float* ghr = gh + row * 3 * H;
    const float r = sigmoidf_(__ldg(gir + c) + __ldg(ghr + c));
    const float z = sigmoidf_(__ldg(gir + H + c) + __ldg(ghr + H + c));
    const float nn = tanhf(fmaf(r, __ldg(ghr + 2 * H + c), __ldg(gir + 2 * H + c)));
    out[idx] = fmaf(z, __ldg(h + idx) - nn, nn);            // (1-z) n + z h
}

__global__ void __launch_bounds__(256) gru_gates_bwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                           const float* __restrict__ h, const float* __restrict__ go,
                                                           float* __restrict__ ggi, float* __restrict__ ggh,
                                                           float* __restrict__ gh_direct, int64_t n, int H) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= n * H) return;
    const int64_t row = idx / H;
    const int c = (int)(idx - row * H);
    const float* gir = gi + row * 3 * H;
    const float* ghr = gh + row * 3 * H;
    const float r = sigmoidf_(__ldg(gir + c) + __ldg(ghr + c));
    const float z = sigmoidf_(__ldg(gir + H + c) + __ldg(ghr + H + c));
    const float ghn = __ldg(ghr + 2 * H + c);
    const float nn = tanhf(fmaf(r, ghn, __ldg(gir + 2 * H + c)));
    const float g = __ldg(go + idx);
    const float dn = g * (1.0f - z) * (1.0f - nn * nn);      // d pre-activation of n
    const float dz = g * (__ldg(h + idx) - nn) * z * (1.0f - z);
    const float dr = dn * ghn * r * (1.0f - r);
    float* a = ggi + row * 3 * H;
    float* b = ggh + row * 3 * H;
    a[c] = dr; a[H + c] = dz; a[2 * H + c] = dn;
    b[c] = dr; b[H + c] = dz; b[2 * H + c] = dn * r;
    gh_direct[idx] = g * z;
}

}  // namespace ubs

extern "C" UBS_API int ubs_gru_gates_fwd(const float* gi, const float* gh, const float* h, float* h_out, int64_t n, int H,
                                 void* stream) {
    UBS_REQUIRE(n >= 0 && H > 0, "ubs_gru_gates_fwd: bad sizes");
    if (n == 0) return 0;
    const int64_t total = n * H;
    ubs::gru_gates_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gi, gh, h, h_out, n, H);
    return ubs::check_launch("ubs_gru_gates_fwd");
}

extern "C" UBS_API int ubs_gru_gates_bwd(const float* gi, const float* gh, const float* h, const float* grad_out,
                                 float* grad_gi, float* grad_gh, float* grad_h_direct, int64_t n, int H, void* stream) {
    UBS_REQUIRE(n >= 0 && H > 0, "ubs_gru_gates_bwd: bad sizes");
    if (n == 0) return 0;
    const int64_t total = n * H;
    ubs::gru_gates_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        gi, gh, h, grad_out, grad_gi, grad_gh, grad_h_direct, n, H);
    return ubs::check_launch("ubs_gru_gates_bwd");
}
