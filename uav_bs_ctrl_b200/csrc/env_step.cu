// Device-resident vectorised MultiUbsCoverageEnv + on-device observation-graph builder (include/ubs_env.h).
//
// Kernel 1 (env_step_kernel): one CTA per env instance runs env_core.h's env_run — move, distances, channel gains,
// greedy RB scheduling, rates, running averages, fairness, reward, visibility — with the env's working set in shared
// memory, and leaves the env's visible GT / UBS rows compacted in (agent, slot) order in a staging area.
// Kernel 2 (env_pack_kernel): one CTA per env turns the per-agent degrees of ALL envs into the packet's global
// `indptr` (exclusive prefix over the B*U agent rows) and copies the staged rows to their place in `x_gt` / `x_ubs`:
// the node / edge order of the reference's dgl.batch(per-agent graphs) + dgl.merge + dgl.batch(envs)
// (env_wrappers.py:65-89,137; algos/common.py:40-47) by pointer arithmetic, no graph objects.
//
// Compiled with -fmad=false (see Makefile): the env arithmetic must round like numpy's separate multiply / add.
#include <stdlib.h>

#include "common.cuh"
#include "env_core.h"

namespace ubs_env {

__device__ long long g_phase_clock[32];   // UBS_ENV_PROFILE=1: clock64() of CTA 0 / thread 0 after every phase barrier

struct DevCtx {
    int tid, nthr;
    bool prof;
    __device__ void sync() const { __syncthreads(); }
    __device__ void mark(int k) const {
        if (prof && tid == 0 && k < 32) g_phase_clock[k] = clock64();
    }
};

struct StepArgs {
    ubs_env_cfg cfg;
    ubs_env_state st;
    ubs_env_packet pk;
    const int64_t* actions;
    int32_t* scratch;
    int64_t B;
    int is_reset, profile;
};

// Both kernels trigger their dependents at once and read nothing before pdl_wait(): the act kernel that follows the pack
// kernel inside a rollout graph is launched programmatically (UBS_ACT_PDL) and stages its first weight layer meanwhile.
// (Launching these two programmatically as well was measured: no gain — 18.9 us per env step either way.)
__global__ void __launch_bounds__(256) env_step_kernel(const __grid_constant__ StepArgs a) {
    ubs::pdl_trigger();
    ubs::pdl_wait();
    extern __shared__ double smem_d[];
    const ubs_env_cfg& c = a.cfg;
    Work w;
    w.carve(smem_d, c.n_ubs, c.n_gts, c.n_rbs);
    const Scratch sc(c, a.B);
    const int64_t b = blockIdx.x;
    EnvOut o;
    o.x_agent = reinterpret_cast<float*>(a.pk.packet + a.pk.off_x_agent);
    o.mask = a.pk.packet + a.pk.off_mask;
    o.rew = reinterpret_cast<float*>(a.pk.packet + a.pk.off_rew);
    o.done = reinterpret_cast<float*>(a.pk.packet + a.pk.off_done);
    o.bad = reinterpret_cast<float*>(a.pk.packet + a.pk.off_bad);
    o.stage_gt = reinterpret_cast<float*>(a.scratch + sc.off_stage_gt + b * sc.gt_stride);
    o.stage_ubs = reinterpret_cast<float*>(a.scratch + sc.off_stage_ubs + b * sc.ubs_stride);
    o.off_seen = a.scratch + sc.off_off_seen;
    o.off_near = a.scratch + sc.off_off_near;
    o.tot = a.scratch + sc.off_tot;
    o.state = a.pk.off_state >= 0 ? reinterpret_cast<float*>(a.pk.packet + a.pk.off_state) : nullptr;
    o.flat = a.pk.off_flat >= 0 ? reinterpret_cast<float*>(a.pk.packet + a.pk.off_flat) : nullptr;
    o.ld_flat = a.pk.ld_flat;
    DevCtx ctx{(int)threadIdx.x, (int)blockDim.x, a.profile != 0 && blockIdx.x == 0};
    ctx.mark(0);
    env_run(c, a.st, b, a.actions, a.is_reset != 0, w, o, ctx);
}

__global__ void __launch_bounds__(128) env_pack_kernel(const __grid_constant__ StepArgs a) {
    ubs::pdl_trigger();
    ubs::pdl_wait();
    __shared__ int red[2][4];
    const ubs_env_cfg& c = a.cfg;
    const Scratch sc(c, a.B);
    const int U = c.n_ubs, Fg = c.fair_service ? 4 : 3;
    const int64_t b = blockIdx.x, N = a.B * U;
    const int32_t* tot = a.scratch + sc.off_tot;
    // rows of all envs before this one: one pass over the per-env totals
    int s1 = 0, s2 = 0;
    for (int64_t i = threadIdx.x; i < b; i += 128) { s1 += tot[2 * i]; s2 += tot[2 * i + 1]; }
    s1 = __reduce_add_sync(0xffffffffu, s1);
    s2 = __reduce_add_sync(0xffffffffu, s2);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    const int my_seen = tot[2 * b], my_near = tot[2 * b + 1];
    int o1 = 0, o2 = 0;
    if (threadIdx.x < U) {
        o1 = a.scratch[sc.off_off_seen + b * U + threadIdx.x];
        o2 = a.scratch[sc.off_off_near + b * U + threadIdx.x];
    }
    __syncthreads();
    const int base_seen = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    const int base_near = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    int32_t* ip_seen = a.pk.packet + a.pk.off_ip_seen;
    int32_t* ip_near = a.pk.packet + a.pk.off_ip_near;
    if (threadIdx.x < U) {
        ip_seen[b * U + threadIdx.x] = base_seen + o1;
        ip_near[b * U + threadIdx.x] = base_near + o2;
    }
    if (b == a.B - 1 && threadIdx.x == 0) {
        ip_seen[N] = base_seen + my_seen;
        ip_near[N] = base_near + my_near;
    }
    const int32_t* sg = a.scratch + sc.off_stage_gt + b * sc.gt_stride;
    int32_t* xg = a.pk.packet + a.pk.off_x_gt + (int64_t)base_seen * Fg;
    for (int i = threadIdx.x; i < my_seen * Fg; i += 128) xg[i] = sg[i];
    const int32_t* su = a.scratch + sc.off_stage_ubs + b * sc.ubs_stride;
    int32_t* xu = a.pk.packet + a.pk.off_x_ubs + (int64_t)base_near * 2;
    for (int i = threadIdx.x; i < my_near * 2; i += 128) xu[i] = su[i];
}

static int check_cfg(const ubs_env_cfg* c, const char* fn) {
    using ubs::set_error;
    if (!c) { set_error("%s: NULL config", fn); return 2; }
    if (c->n_ubs < 1 || c->n_ubs > UBS_ENV_MAX_UBS) { set_error("%s: n_ubs must be in [1, %d] (got %d)", fn, UBS_ENV_MAX_UBS, c->n_ubs); return 2; }
    if (c->n_gts < 1 || c->n_gts > 512) { set_error("%s: n_gts must be in [1, 512] (got %d)", fn, c->n_gts); return 2; }
    if (c->n_rbs < 1 || c->n_rbs > 64) { set_error("%s: n_rbs must be in [1, 64] (got %d)", fn, c->n_rbs); return 2; }
    if (c->n_actions < 1 || c->n_actions > UBS_ENV_MAX_ACTIONS) { set_error("%s: n_actions must be in [1, %d] (got %d)", fn, UBS_ENV_MAX_ACTIONS, c->n_actions); return 2; }
    if (!(c->range_pos > 0) || !(c->max_rate > 0)) { set_error("%s: range_pos / max_rate must be positive", fn); return 2; }
    return 0;
}

static int launch(const char* fn, const ubs_env_cfg* cfg, const ubs_env_state* st, const int64_t* actions,
                  const ubs_env_packet* pk, int32_t* scratch, int64_t B, bool is_reset, void* stream) {
    if (int rc = check_cfg(cfg, fn)) return rc;
    UBS_REQUIRE(st && pk && scratch, "%s: NULL argument", fn);
    UBS_REQUIRE(st->pos_ubs && st->pos_gts && st->avg_rate && st->rate && st->prior && st->t && st->info && st->sched,
                "%s: NULL state pointer", fn);
    UBS_REQUIRE(pk->packet != nullptr, "%s: NULL packet", fn);
    UBS_REQUIRE(pk->off_flat < 0 || pk->ld_flat >= 2 + cfg->n_gts * (cfg->fair_service ? 5 : 4) + (cfg->n_ubs - 1) * 3,
                "%s: ld_flat smaller than the flattened observation", fn);
    UBS_REQUIRE(is_reset || actions != nullptr, "%s: NULL actions", fn);
    UBS_REQUIRE(B >= 0 && B * cfg->n_ubs < (1ll << 31) / (cfg->n_gts > 0 ? cfg->n_gts : 1), "%s: batch out of range", fn);
    if (B == 0) return 0;
    StepArgs a;
    a.cfg = *cfg; a.st = *st; a.pk = *pk; a.actions = actions; a.scratch = scratch; a.B = B; a.is_reset = is_reset ? 1 : 0;
    static const bool env_profile = getenv("UBS_ENV_PROFILE") != nullptr;      // read once per process
    a.profile = env_profile;
    const size_t smem = Work::bytes(cfg->n_ubs, cfg->n_gts, cfg->n_rbs);
    UBS_REQUIRE(smem <= 227 * 1024, "%s: env working set (%zu B) exceeds shared memory", fn, smem);
    if (smem > 48 * 1024) {                      // opt in to the device maximum once per process (thread-safe static)
        static const cudaError_t attr_rc = cudaFuncSetAttribute(env_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        UBS_REQUIRE(attr_rc == cudaSuccess, "%s: cannot opt in to 227 KB of shared memory", fn);
    }
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = ubs::launch_pdl(env_step_kernel, (unsigned)B, 256u, smem, s, false, a);
    if (e != cudaSuccess) { ubs::set_error("%s: launch failed: %s", fn, cudaGetErrorString(e)); return 1; }
    if (int rc = ubs::check_launch(fn)) return rc;
    e = ubs::launch_pdl(env_pack_kernel, (unsigned)B, 128u, 0, s, false, a);
    if (e != cudaSuccess) { ubs::set_error("%s: launch failed: %s", fn, cudaGetErrorString(e)); return 1; }
    return ubs::check_launch(fn);
}

}  // namespace ubs_env

// Diagnostic (tools/env_profile.py): cycle stamps of the last profiled ubs_env_step launch (UBS_ENV_PROFILE=1).
extern "C" UBS_ENV_API int ubs_env_phase_clocks(int64_t* out32) {
    long long h[32];
    if (cudaMemcpyFromSymbol(h, ubs_env::g_phase_clock, sizeof(h)) != cudaSuccess) return 1;
    for (int i = 0; i < 32; ++i) out32[i] = h[i];
    return 0;
}

extern "C" UBS_ENV_API int64_t ubs_env_scratch_words(const ubs_env_cfg* cfg, int64_t B) {
    if (!cfg || B < 0) return -1;
    return ubs_env::Scratch(*cfg, B).words;
}

extern "C" UBS_ENV_API int ubs_env_reset(const ubs_env_cfg* cfg, const ubs_env_state* st, const ubs_env_packet* pk,
                                         int32_t* scratch, int64_t B, void* stream) {
    return ubs_env::launch("ubs_env_reset", cfg, st, nullptr, pk, scratch, B, true, stream);
}

extern "C" UBS_ENV_API int ubs_env_step(const ubs_env_cfg* cfg, const ubs_env_state* st, const int64_t* actions,
                                        const ubs_env_packet* pk, int32_t* scratch, int64_t B, void* stream) {
    return ubs_env::launch("ubs_env_step", cfg, st, actions, pk, scratch, B, false, stream);
}
