// Column sums over the T*N rows of a BPTT window — the bias gradients of the recurrent block (db_ih | db_vsq | db_hh =
// column sums of the [dgi | dvsq | dgh] stash) and of the aggregator (db_aggr) — optionally fused with the ReLU
// backward of the aggregator (dpre = dx * (x > 0)), which would otherwise be three more passes over the same rows.
//
// HBM-bound streaming reduction: 4 CTAs per SM, each owns a contiguous slab of rows; a thread keeps one float4 column
// group and walks the slab's rows (stride = rows per pass), four independent loads in flight; the CTA's row groups are
// added in a fixed order through shared memory and ONE partial row per CTA goes to the workspace; a second small
// kernel adds the partials in a fixed order.  Deterministic: no atomics, the result does not depend on timing.
#include "common.cuh"
#include "../../include/ubs_gnn.h"

namespace ubs {
namespace colsum {

constexpr int NT = 256, MAX_C4 = 256;          // up to 1024 columns

struct Args {
    const float* X; const float* Y; float* dX; float* ws;
    long long ldx, ldy, ldd, R, rows_per_cta;
    int C4, rpp;                                // float4 column groups; rows per pass = NT / C4
};

template <bool RELU_BWD>
__global__ void __launch_bounds__(NT, 4) colsum_kernel(const Args a) {
    __shared__ float4 part[NT];
    const int cg = threadIdx.x % a.C4, rg = threadIdx.x / a.C4;
    const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
    const long long r1 = r0 + a.rows_per_cta < a.R ? r0 + a.rows_per_cta : a.R;
    float4 acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rg < a.rpp) {
        for (long long r = r0 + rg; r < r1; r += 4LL * a.rpp) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ru = r + (long long)u * a.rpp;
                if (ru < r1) {
                    float4 x = __ldcs(reinterpret_cast<const float4*>(a.X + ru * a.ldx) + cg);
                    if (RELU_BWD) {
                        const float4 y = __ldcs(reinterpret_cast<const float4*>(a.Y + ru * a.ldy) + cg);
                        x.x = y.x > 0.f ? x.x : 0.f; x.y = y.y > 0.f ? x.y : 0.f;
                        x.z = y.z > 0.f ? x.z : 0.f; x.w = y.w > 0.f ? x.w : 0.f;
                        *(reinterpret_cast<float4*>(a.dX + ru * a.ldd) + cg) = x;
                    }
                    acc[u].x += x.x; acc[u].y += x.y; acc[u].z += x.z; acc[u].w += x.w;
                }
            }
        }
    }
    float4 s;
    s.x = (acc[0].x + acc[1].x) + (acc[2].x + acc[3].x);
    s.y = (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y);
    s.z = (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z);
    s.w = (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w);
    part[threadIdx.x] = s;
    __syncthreads();
    if (rg == 0) {
        for (int g = 1; g < a.rpp; ++g) {                       // fixed order over the CTA's row groups
            const float4 p = part[g * a.C4 + cg];
            s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
        }
        reinterpret_cast<float4*>(a.ws + (size_t)blockIdx.x * a.C4 * 4)[cg] = s;
    }
}

// out[c] = sum_p ws[p, c]: a block owns 32 columns, its 32 warps take the partial rows p = w, w + 32, ... (8 independent
// loads in flight per thread: a few dependent rounds for hundreds of partials), combined in warp order
constexpr int FIN_WARPS = 32;
__global__ void __launch_bounds__(32 * FIN_WARPS) colsum_final_kernel(const float* __restrict__ ws, int nparts, int C, float* __restrict__ out) {
    __shared__ float sm[FIN_WARPS][32];
    const int lane = threadIdx.x % 32, w = threadIdx.x / 32, col = blockIdx.x * 32 + lane;
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    if (col < C) {
        for (int p = w; p < nparts; p += 8 * FIN_WARPS) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (p + u * FIN_WARPS < nparts) acc[u] += ws[(size_t)(p + u * FIN_WARPS) * C + col];
        }
    }
    sm[w][lane] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    __syncthreads();
    if (w == 0 && col < C) {
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < FIN_WARPS; ++s) t += sm[s][lane];
        out[col] = t;
    }
}

static int zero(float* out, int C, cudaStream_t st, const char* what) {
    const cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * C, st);
    if (e != cudaSuccess) { set_error("%s: cudaMemsetAsync failed: %s", what, cudaGetErrorString(e)); return 1; }
    return 0;
}

static int grid_for(long long R, int rpp, long long* rows_per_cta) {
    long long ctas = (long long)kNumSMs * 4;
    long long per = (R + ctas - 1) / ctas;
    per = (per + rpp - 1) / rpp * rpp;                          // whole passes
    if (per < rpp) per = rpp;
    *rows_per_cta = per;
    return (int)((R + per - 1) / per);
}

static int launch(const float* X, long long ldx, const float* Y, long long ldy, float* dX, long long ldd, long long R, int C,
                  float* out, float* ws, cudaStream_t st, const char* what) {
    Args a{};
    a.X = X; a.Y = Y; a.dX = dX; a.ws = ws; a.ldx = ldx; a.ldy = ldy; a.ldd = ldd; a.R = R;
    a.C4 = C / 4; a.rpp = NT / a.C4;
    const int grid = grid_for(R, a.rpp, &a.rows_per_cta);
    if (Y) colsum_kernel<true><<<grid, NT, 0, st>>>(a);
    else colsum_kernel<false><<<grid, NT, 0, st>>>(a);
    if (int rc = check_launch(what)) return rc;
    colsum_final_kernel<<<(C + 31) / 32, 32 * FIN_WARPS, 0, st>>>(ws, grid, C, out);
    return check_launch(what);
}

}  // namespace colsum
}  // namespace ubs

extern "C" UBS_API int64_t ubs_colsum_workspace(int C) { return (int64_t)ubs::kNumSMs * 4 * C; }

#define UBS_COLSUM_CHECK(name)                                                                                                 \
    UBS_REQUIRE(X && out && workspace && R >= 0, name ": NULL argument");                                                      \
    UBS_REQUIRE(C >= 4 && C % 4 == 0 && C <= 4 * ubs::colsum::MAX_C4, name ": C must be a multiple of 4, <= 1024 (got %d)", C); \
    UBS_REQUIRE(ldx % 4 == 0 && ldx >= C && ((uintptr_t)X % 16) == 0, name ": rows must be 16-byte aligned")

// out[c] = sum_r X[r * ldx + c]                                    (torch: X.sum(0), learner.py bias gradients via autograd)
extern "C" UBS_API int ubs_colsum(const float* X, int64_t ldx, int64_t R, int C, float* out, float* workspace, void* stream) {
    UBS_COLSUM_CHECK("ubs_colsum");
    if (R == 0) return ubs::colsum::zero(out, C, (cudaStream_t)stream, "ubs_colsum");
    return ubs::colsum::launch(X, ldx, nullptr, 0, nullptr, 0, R, C, out, workspace, (cudaStream_t)stream, "ubs_colsum");
}

// dX = X * (Y > 0), out[c] = sum_r dX[r, c]: ReLU backward of the aggregator layer + its bias gradient (dX may alias X)
extern "C" UBS_API int ubs_relu_bwd_colsum(const float* X, int64_t ldx, const float* Y, int64_t ldy, float* dX, int64_t ldd,
                                           int64_t R, int C, float* out, float* workspace, void* stream) {
    UBS_COLSUM_CHECK("ubs_relu_bwd_colsum");
    UBS_REQUIRE(Y && dX && ldy % 4 == 0 && ldd % 4 == 0 && ldy >= C && ldd >= C && ((uintptr_t)Y % 16) == 0 && ((uintptr_t)dX % 16) == 0,
                "ubs_relu_bwd_colsum: rows must be 16-byte aligned");
    if (R == 0) return ubs::colsum::zero(out, C, (cudaStream_t)stream, "ubs_relu_bwd_colsum");
    return ubs::colsum::launch(X, ldx, Y, ldy, dX, ldd, R, C, out, workspace, (cudaStream_t)stream, "ubs_relu_bwd_colsum");
}
