// One step of ONE MultiUbsCoverageEnv instance, written once for the device (a CTA per env, threads stride over the
// parallel phases) and for a serial host build (oracle/env_host.cpp: test infrastructure only).
//
// Follows envs/mubs_cov/mubs_cov.py statement by statement: step :104-127, _transmit_data :129-211, get_obs_agent
// :216-242, _get_reward :296-312, _get_terminate :314-316; channel model envs/common.py:45-55; Jain index :19-25;
// observation-graph layout algos/madrqn/utils/env_wrappers.py:69-89,139-154.  Comments of the form `np: ...` state
// the dtype / summation order numpy (>= 2, NEP 50) uses for the statement being restated — the arithmetic below
// reproduces it operation by operation (compile with FMA contraction OFF).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/ubs_env.h"

#if defined(__CUDACC__)
#define UBS_HD __host__ __device__
#else
#define UBS_HD
#endif

namespace ubs_env {

// numpy's pairwise summation of a contiguous float32 vector (numpy/_core/src/umath/loops_utils.h.src): what
// ndarray.sum() / np.mean() of a 1-D float32 array computes.  `at(i)` yields element i; the recursion of the original
// (blocks of <= 128 elements, halves rounded to multiples of 8) is unrolled at compile time (2 levels: n <= 512 = the n_gts limit of the env).
template <class F>
UBS_HD inline float np_sum_block(const F& at, int lo, int n) {
    if (n < 8) {
        float res = -0.0f;
        for (int i = 0; i < n; ++i) res += at(lo + i);
        return res;
    }
    float r0 = at(lo), r1 = at(lo + 1), r2 = at(lo + 2), r3 = at(lo + 3), r4 = at(lo + 4), r5 = at(lo + 5),
          r6 = at(lo + 6), r7 = at(lo + 7);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 += at(lo + i); r1 += at(lo + i + 1); r2 += at(lo + i + 2); r3 += at(lo + i + 3);
        r4 += at(lo + i + 4); r5 += at(lo + i + 5); r6 += at(lo + i + 6); r7 += at(lo + i + 7);
    }
    float res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += at(lo + i);
    return res;
}
template <int DEPTH, class F>
UBS_HD inline float np_sum_rec(const F& at, int lo, int n) {
    if constexpr (DEPTH == 0) {
        return np_sum_block(at, lo, n);
    } else {
        if (n <= 128) return np_sum_block(at, lo, n);
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum_rec<DEPTH - 1>(at, lo, n2) + np_sum_rec<DEPTH - 1>(at, lo + n2, n - n2);
    }
}
template <class F>
UBS_HD inline float np_sum_fn(const F& at, int n) { return np_sum_rec<2>(at, 0, n); }
#if defined(__CUDACC__)
__noinline__          // one copy of the unrolled blocks for the three call sites
#endif
UBS_HD inline float np_sum_f32(const float* a, int n) {
    return np_sum_fn([a](int i) { return a[i]; }, n);
}

// Carves the per-env workspace (shared memory on the device) out of one 8-byte aligned buffer.
struct Work {
    double *pos_u, *gain;
    float *pos_g, *d_u2g, *d_u2u, *pitf, *rate, *avg, *tmp, *tmp2, *f_rate, *f_avg, *scal;
    int *prior, *occ, *nsch, *sch_ubs, *sch_rb, *slot, *nslot, *deg, *ncov;
    unsigned char* order;

    UBS_HD static size_t bytes(int U, int G, int R) {
        size_t n = 8 * ((size_t)U * 2 + (size_t)U * G);
        n += 4 * ((size_t)G * 2 + (size_t)U * G + (size_t)U * U + (size_t)U * G + 6 * (size_t)G + 8);
        n += 4 * ((size_t)G + (size_t)U * R + U + 3 * (size_t)G + (size_t)U * G + (size_t)U * U + 5 * (size_t)U + 4);
        n += (size_t)G * U;
        return (n + 15) & ~(size_t)15;
    }
    UBS_HD void carve(void* base, int U, int G, int R) {
        double* d = (double*)base;
        pos_u = d; d += U * 2;
        gain = d; d += (size_t)U * G;
        float* f = (float*)d;
        pos_g = f; f += G * 2;
        d_u2g = f; f += (size_t)U * G;
        d_u2u = f; f += U * U;
        pitf = f; f += (size_t)U * G;
        rate = f; f += G;
        avg = f; f += G;
        tmp = f; f += G;
        tmp2 = f; f += G;
        f_rate = f; f += G;
        f_avg = f; f += G;
        scal = f; f += 8;
        int* i = (int*)f;
        prior = i; i += G;
        occ = i; i += U * R;
        nsch = i; i += U;
        sch_ubs = i; i += G;
        sch_rb = i; i += G;
        slot = i; i += (size_t)U * G;
        nslot = i; i += U * U;
        deg = i; i += 5 * U + 4;          // deg_seen[U] | deg_near[U] | off_seen[U] | off_near[U] | not-idle[U]
        ncov = i; i += G;
        order = (unsigned char*)i;
    }
};

struct SerialCtx {
    int tid = 0, nthr = 1;
    void sync() const {}
    void mark(int) const {}
};

// Where the results of env b go.
struct EnvOut {
    float* x_agent;     // packet section, (N, 2)
    int32_t* mask;      // packet section, (N)
    float* rew;         // (N)
    float* done;        // (B)
    float* bad;         // (B)
    float* stage_gt;    // this env's staging rows (U*G, F_gt): compacted in (agent, slot) order
    float* stage_ubs;   // (U*(U-1), 2)
    int32_t* off_seen;  // (N) all envs: env-local exclusive prefix of the seen degrees
    int32_t* off_near;  // (N)
    int32_t* tot;       // (B, 2): seen / near rows of each env
    float* state;       // (B, 2U + Fs*G) global state rows, or nullptr
    float* flat;        // (N, ld_flat) flattened local observations [agent | gt rows with flag | ubs rows with flag], or nullptr
    int64_t ld_flat;
};


// ---- device-only helpers: warp-cooperative versions of the two serial phases (same arithmetic, same order) --------
#if defined(__CUDA_ARCH__)
template <int UMAX>
__device__ __forceinline__ float itf_sum(float p, unsigned occb) {
    float v = 0.f;
#pragma unroll
    for (int i2 = 0; i2 < UMAX; ++i2) {
        const float pi = __shfl_sync(0xffffffffu, p, i2);
        v += ((occb >> i2) & 1u) ? pi : 0.f;
    }
    return v;
}

// Greedy RB scheduling by warp 0.  Lane rb owns resource block rb (bit i of `occb` = UBS i transmits on it); the
// interference sum of a block is accumulated over the UBSs in index order exactly like the serial loop.
__device__ inline void sched_warp(const ubs_env_cfg& c, const Work& w, int lane) {
    const int U = c.n_ubs, G = c.n_gts, R = c.n_rbs;
    int* list = w.slot;                                   // scratch (the compaction slots are computed later)
    int cnt = 0;
    for (int k0 = 0; k0 < G; k0 += 32) {                  // covered GTs in priority order
        const int k = k0 + lane;
        const int m = k < G ? w.prior[k] : -1;
        const bool f = m >= 0 && w.ncov[m] > 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (f) list[cnt + __popc(bal & ((1u << lane) - 1u))] = m;
        cnt += __popc(bal);
    }
    __syncwarp();
    unsigned occb = 0;                                    // lane rb: UBSs transmitting on RB rb
    unsigned long long nsch = 0;                          // uniform: 4 bits per UBS (U <= 16, R <= 15)
    int m_next = cnt > 0 ? list[0] : 0;
    for (int q = 0; q < cnt; ++q) {
        const int m = m_next;
        if (q + 1 < cnt) m_next = list[q + 1];            // off the critical path
        const float p = lane < U ? w.pitf[lane * G + m] : 0.f;
        const int nc = w.ncov[m];
        int i = -1;
        for (int j = 0; j < nc; ++j) {                    // nearest covering UBS with a free RB
            const int cand = w.order[(size_t)m * U + j];
            if ((int)((nsch >> (4 * cand)) & 15ull) < R) { i = cand; break; }
        }
        if (i < 0) continue;                              // warp-uniform
        // interference on RB `lane` if m were served there: the UBSs already on that RB, added in index order
        // (adding +0.f for the others / for i2 >= U leaves the non-negative sum unchanged bit for bit)
        const float v = U <= 8 ? itf_sum<8>(p, occb) : itf_sum<16>(p, occb);
        const bool idle = lane < R && !((occb >> i) & 1u);
        // first minimum among the idle RBs: v >= +0, so the uint order of the bit patterns is the float order
        const unsigned key = idle ? __float_as_uint(v) : 0xffffffffu;
        const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
        const int best = __ffs(__ballot_sync(0xffffffffu, key == kmin)) - 1;
        if (lane == best) occb |= 1u << i;
        nsch += 1ull << (4 * i);
        if (lane == 0) {
            w.occ[i * R + best] = m;
            w.sch_ubs[m] = i;
            w.sch_rb[m] = best;
        }
    }
    __syncwarp();
}
#endif

template <class Ctx>
UBS_HD inline void env_run(const ubs_env_cfg& c, const ubs_env_state& st, int64_t b, const int64_t* actions,
                           bool is_reset, const Work& w, const EnvOut& o, Ctx& ctx) {
    const int U = c.n_ubs, G = c.n_gts, R = c.n_rbs;
    const int tid = ctx.tid, nthr = ctx.nthr;
    const int Fg = c.fair_service ? 4 : 3;
    double* info = st.info + b * UBS_ENV_INFO;

    // running episode statistics: fetched now by thread 0 so that their latency hides behind the whole step
    double inf0 = 0.0, inf1 = 0.0, inf2 = 0.0, inf3 = 0.0;
    if (tid == 0 && !is_reset) { inf0 = info[0]; inf1 = info[1]; inf2 = info[2]; inf3 = info[3]; }

    // ---- load state; step(): self.t += 1; pos_ubs = clip(pos_ubs + avail_moves[actions], 0, range_pos)  (:105-109)
    const int t = is_reset ? 0 : st.t[b] + 1;
    for (int i = tid; i < U * 2; i += nthr) {
        double p = st.pos_ubs[b * U * 2 + i];                              // np: float64
        if (!is_reset) {
            long long a = actions[b * U + i / 2];
            if (a < 0) a = 0;
            if (a >= c.n_actions) a = c.n_actions - 1;
            p = p + c.moves[a][i & 1];
            p = p < 0.0 ? 0.0 : p;                                         // np.clip = minimum(maximum(x, lo), hi)
            p = p > c.range_pos ? c.range_pos : p;
        }
        w.pos_u[i] = p;
    }
    for (int i = tid; i < G * 2; i += nthr) w.pos_g[i] = st.pos_gts[b * G * 2 + i];
    for (int m = tid; m < G; m += nthr) {
        w.avg[m] = is_reset ? 0.0f : st.avg_rate[b * G + m];
        w.prior[m] = st.prior[b * G + m];
        w.sch_ubs[m] = -1;
        w.sch_rb[m] = -1;
    }
    for (int i = tid; i < U * R; i += nthr) w.occ[i] = -1;
    for (int i = tid; i < U; i += nthr) { w.nsch[i] = 0; w.deg[4 * U + i] = 0; }
    ctx.sync();
    ctx.mark(1);

    // ---- _transmit_data step 1: distances (:133-139).  np: norm of a float64 difference stored into a float32 array
    const float r_cov = (float)c.r_cov, r_sns = (float)c.r_sns, r_comm = (float)c.r_comm;
    for (int p = tid; p < U * G; p += nthr) {
        const int i = p / G, m = p - i * G;
        const double dx = (double)w.pos_g[m * 2] - w.pos_u[i * 2], dy = (double)w.pos_g[m * 2 + 1] - w.pos_u[i * 2 + 1];
        const float dl = (float)sqrt(dx * dx + dy * dy);
        w.d_u2g[p] = dl;
        // channel gain (common.py:45-55).  np: p_los stays float32 (python scalars are weak); np.square(h_ubs) is a
        // float64 scalar, so d, fspl, pl and the gain are float64.  Only needed where the link can carry signal or
        // interference (d <= r_cov): p_itf and the served links are masked by it (:164,176,185).
        float pitf = 0.f;
        double g = 0.0;
        if (dl <= r_cov) {
            const float e = expf((float)(-c.chan_b) * (atanf((float)c.h_ubs / (dl + 1e-5f)) - (float)c.chan_a));
            const float p_los = 1.0f / (1.0f + (float)c.chan_a * e);
            const double d3 = sqrt((double)(dl * dl) + c.h_ubs * c.h_ubs);
            const double q = c.c_fspl * d3 / 3e8;
            const double fspl = q * q;
            const double pl = (double)p_los * fspl * c.k_los + (double)(1.0f - p_los) * fspl * c.k_nlos;
            g = 1.0 / pl;
            pitf = (float)(c.p_tx * g);       // p_itf[i, :, rb] = p_tx * g[i] * mask_itf[i]  (:176): float64 -> float32
        }
        w.gain[p] = g;
        w.pitf[p] = pitf;
    }
    for (int p = tid; p < U * U; p += nthr) {
        const int i = p / U, j = p - i * U;
        const double dx = w.pos_u[j * 2] - w.pos_u[i * 2], dy = w.pos_u[j * 2 + 1] - w.pos_u[i * 2 + 1];
        w.d_u2u[p] = (float)sqrt(dx * dx + dy * dy);
    }
    ctx.sync();
    ctx.mark(2);
    // nearest_ubs = np.argsort(d_u2g[:, m]) (:167) restricted to the UBSs that cover m (the loop over nearest_ubs can
    // only pick those, :169): insertion sort = numpy's small-array path, ties by index
    for (int m = tid; m < G; m += nthr) {
        unsigned char* ord = w.order + (size_t)m * U;
        int n = 0;
        for (int i = 0; i < U; ++i) {
            const float key = w.d_u2g[i * G + m];
            if (!(key <= r_cov)) continue;
            int j = n++;
            while (j > 0 && w.d_u2g[ord[j - 1] * G + m] > key) { ord[j] = ord[j - 1]; --j; }
            ord[j] = (unsigned char)i;
        }
        w.ncov[m] = n;
    }
    ctx.sync();
    ctx.mark(3);

    // ---- step 2: greedy RB scheduling in priority order (:166-178) — inherently sequential over GTs
#if defined(__CUDA_ARCH__)
    if (R <= 15 && U <= 16) {
        if (tid < 32) sched_warp(c, w, tid);
    } else
#endif
    if (tid == 0) {
        for (int k = 0; k < G; ++k) {
            const int m = w.prior[k];
            const unsigned char* ord = w.order + (size_t)m * U;
            for (int jj = 0; jj < w.ncov[m]; ++jj) {
                const int i = ord[jj];
                if (w.nsch[i] < R) {
                    int best = -1;
                    float bestv = 0.f;
                    for (int rb = 0; rb < R; ++rb) {
                        if (w.occ[i * R + rb] >= 0) continue;    // itf_per_chan[occupied] = nan
                        float v = 0.f;                            // np: p_itf[:, m, :].sum(0) adds the rows in order
                        for (int i2 = 0; i2 < U; ++i2) v += (w.occ[i2 * R + rb] >= 0) ? w.pitf[i2 * G + m] : 0.f;
                        if (best < 0 || v < bestv) { best = rb; bestv = v; }   // nanargmin: first minimum
                    }
                    w.occ[i * R + best] = m;
                    w.nsch[i] += 1;
                    w.sch_ubs[m] = i;
                    w.sch_rb[m] = best;
                    break;
                }
            }
        }
    }
    ctx.sync();
    ctx.mark(4);

    // ---- rates (:181-187) and running averages (:193, np: float32 throughout; self.t is a weak python int)
    for (int m = tid; m < G; m += nthr) {
        float r = 0.f;
        const int i = w.sch_ubs[m];
        if (i >= 0) {
            const int rb = w.sch_rb[m];
            // p_itf[:, m, rb] with p_itf[i, m, rb] = 0 for the served GT (:177); np: (U,1) float32 copy .sum() -> pairwise
            const float itf = np_sum_fn([&](int i2) {
                const int oc = w.occ[i2 * R + rb];
                return (oc >= 0 && oc != m) ? w.pitf[i2 * G + m] : 0.f;
            }, U);
            const double sinr = (c.p_tx * w.gain[i * G + m]) / ((double)itf + c.bw * c.n0);
            r = (float)(c.bw * log2(1.0 + sinr) * 1e-6);                       // np: float64, stored into float32
            if (r != 0.f) w.deg[4 * U + i] = 1;                                // rate_per_ubs[i] != 0 (:188): UBS i is not idle
        }
        w.rate[m] = r;
        const float av = (w.avg[m] * (float)t + r) / (float)(t + 1);
        w.avg[m] = av;
        const float x = av < 1e-6f ? 1e-6f : av;                               // np.clip(x, 1e-6, inf)
        w.tmp[m] = x;
        w.tmp2[m] = x * x;
        // observation features that depend on the GT only (:237-240).  np: float32 / float64 -> float64
        w.f_rate[m] = (float)((double)r / c.max_rate);
        w.f_avg[m] = (float)((double)av / c.max_rate * (double)G / (double)(U * R));
    }
    ctx.sync();
    ctx.mark(5);
    // three pairwise sums (numpy order), one per warp on the device
    {
        // (the first warps are busy ranking the GTs below)
        const int w0 = nthr >= 256 ? 96 : 0, w1 = nthr >= 256 ? 128 : (nthr >= 96 ? 32 : 0),
                  w2 = nthr >= 256 ? 160 : (nthr >= 96 ? 64 : 0);
        if (tid == w0) w.scal[5] = np_sum_f32(w.tmp, G);
        if (tid == w1) w.scal[6] = np_sum_f32(w.tmp2, G);
        if (tid == w2) w.scal[7] = np_sum_f32(w.rate, G);
    }
    // prior_gts = argsort(avg_rate_per_gt) (:199), ties by index: rank by counting
    for (int m = tid; m < G; m += nthr) {
        const float a = w.avg[m];
        int rank = 0;
        for (int m2 = 0; m2 < G; ++m2) {
            const float a2 = w.avg[m2];
            rank += (a2 < a || (a2 == a && m2 < m)) ? 1 : 0;
        }
        st.prior[b * G + rank] = m;
    }
    // per-agent compaction slots of visible GTs / UBSs (env_wrappers.py:71: rows with flag == 1, in order)
#if defined(__CUDA_ARCH__)
    {
        const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
        for (int i = wid; i < U; i += nw) {
            int cnt = 0;
            for (int m0 = 0; m0 < G; m0 += 32) {
                const int m = m0 + lane;
                const bool f = m < G && w.d_u2g[i * G + m] <= r_sns;
                const unsigned bal = __ballot_sync(0xffffffffu, f);
                if (m < G) w.slot[i * G + m] = f ? cnt + __popc(bal & ((1u << lane) - 1u)) : -1;
                cnt += __popc(bal);
            }
            int cn = 0;
            for (int j0 = 0; j0 < U; j0 += 32) {
                const int j = j0 + lane;
                const bool f = j < U && j != i && w.d_u2u[i * U + j] <= r_comm;
                const unsigned bal = __ballot_sync(0xffffffffu, f);
                if (j < U) w.nslot[i * U + j] = f ? cn + __popc(bal & ((1u << lane) - 1u)) : -1;
                cn += __popc(bal);
            }
            if (lane == 0) { w.deg[i] = cnt; w.deg[U + i] = cn; }
        }
    }
#else
    for (int i = 0; i < U; ++i) {
        int cnt = 0;
        for (int m = 0; m < G; ++m) w.slot[i * G + m] = (w.d_u2g[i * G + m] <= r_sns) ? cnt++ : -1;
        w.deg[i] = cnt;
        cnt = 0;
        for (int j = 0; j < U; ++j) w.nslot[i * U + j] = (j != i && w.d_u2u[i * U + j] <= r_comm) ? cnt++ : -1;
        w.deg[U + i] = cnt;
    }
#endif
    ctx.sync();
    ctx.mark(6);
    // ---- reward (:296-312), terminate (:314-316), collisions (:142-143)
    const bool done = (t == c.episode_limit) && !is_reset;
    // fairness / utility (:195-198, np: float32), recomputed by every thread that needs them
    const float s1 = w.scal[5], s2 = w.scal[6], rsum = w.scal[7];
    const float fair = (s1 * s1) / ((float)G * s2);                            // compute_jain_fairness_index
    const float mean = rsum / (float)G;                                        // np.mean of float32
    const float gu = fair * mean;                                              // global_util
    for (int i = tid; i < U; i += nthr) {
        bool coll = false;
        for (int j = 0; j < U; ++j) coll = coll || (j != i && (double)w.d_u2u[i * U + j] < c.safe_dist);
        // np: rew_scale * float32 stays float32, / max_rate (np.float64) promotes to float64
        const float base = c.fair_service ? gu : mean;
        double lr = (double)((float)c.rew_scale * base) / c.max_rate;
        lr = lr * (w.deg[4 * U + i] ? 1.0 : 0.0);                              // idle UBSs receive no reward (:305-306)
        if (c.avoid_collision) lr = (coll ? 0.0 : 1.0) * lr - (coll ? 1.0 : 0.0) * c.penalty;
        o.rew[b * U + i] = is_reset ? 0.f : (float)lr;
        w.tmp[i] = is_reset ? 0.f : (float)lr;
        w.nsch[i] = coll ? 1 : 0;                                              // reuse: collision flag
        // own features (:223) + talk mask: bit i of mask[dst j] <=> d_u2u[i, j] <= r_comm  (env_wrappers.py:141-144)
        o.x_agent[(b * U + i) * 2] = (float)(w.pos_u[i * 2] / c.range_pos);
        o.x_agent[(b * U + i) * 2 + 1] = (float)(w.pos_u[i * 2 + 1] / c.range_pos);
        uint32_t mk = 0;
        for (int s = 0; s < U && s < 32; ++s) mk |= (w.d_u2u[s * U + i] <= r_comm) ? (1u << s) : 0u;
        o.mask[b * U + i] = (int32_t)mk;
    }
    ctx.sync();
    ctx.mark(8);
    if (tid == 0) {
        int a1 = 0, a2 = 0, ncoll = 0;
        double rmean = 0.0;
        for (int i = 0; i < U; ++i) {
            w.deg[2 * U + i] = a1;
            w.deg[3 * U + i] = a2;
            o.off_seen[b * U + i] = a1;                                         // env-local exclusive prefix
            o.off_near[b * U + i] = a2;
            a1 += w.deg[i];
            a2 += w.deg[U + i];
            ncoll += w.nsch[i];
            rmean += (double)w.tmp[i];
        }
        o.tot[2 * b] = a1;
        o.tot[2 * b + 1] = a2;
        o.done[b] = done ? 1.f : 0.f;
        o.bad[b] = done ? 1.f : 0.f;                                            // BadMask: t == episode_limit (:122)
        st.t[b] = t;
        info[0] = inf0 + rmean / U;                                             // ep_ret
        info[1] = (double)((float)inf1 + rsum * (float)c.dt / 1e3f);            // total_throughput (np: float32)
        info[2] = inf2 + ncoll / 2.0;                                           // n_colls
        info[3] = (double)(((float)inf3 * (float)t + gu) / (float)(t + 1));     // avg_global_util (np: float32)
        info[4] = fair;
        info[5] = gu;
        info[6] = t;
        info[7] = 0.0;
    }
    ctx.sync();
    ctx.mark(9);

    // ---- persistent state + observation rows (:225-240), compacted per env into the staging area
    for (int i = tid; i < U * 2; i += nthr) st.pos_ubs[b * U * 2 + i] = w.pos_u[i];
    for (int m = tid; m < G; m += nthr) {
        st.avg_rate[b * G + m] = w.avg[m];
        st.rate[b * G + m] = w.rate[m];
        st.sched[(b * G + m) * 2] = w.sch_ubs[m];
        st.sched[(b * G + m) * 2 + 1] = w.sch_rb[m];
    }
    // get_state() (:244-262): [pos_ubs / range_pos | per GT: pos / range_pos, rate / max_rate, avg-rate feature]
    if (o.state != nullptr) {
        const int Fs = c.fair_service ? 4 : 3;
        float* srow = o.state + (size_t)b * (2 * U + Fs * G);
        for (int i = tid; i < U * 2; i += nthr) srow[i] = (float)(w.pos_u[i] / c.range_pos);
        for (int m = tid; m < G; m += nthr) {
            float* g4 = srow + 2 * U + m * Fs;
            // np: float32 positions (DenseHotSpot, Debug) are divided in float32, float64 positions in float64
            g4[0] = c.gts_f64 ? (float)((double)w.pos_g[m * 2] / c.range_pos) : w.pos_g[m * 2] / (float)c.range_pos;
            g4[1] = c.gts_f64 ? (float)((double)w.pos_g[m * 2 + 1] / c.range_pos) : w.pos_g[m * 2 + 1] / (float)c.range_pos;
            g4[2] = w.f_rate[m];
            if (c.fair_service) g4[3] = w.f_avg[m];
        }
    }
    const double ngt = fmin(c.range_pos, c.r_sns), nub = fmin(c.range_pos, c.r_comm);
    // flattened observation rows keep every GT / UBS slot (zeros when not visible) with its flag (:225-240)
    const int flat_gt0 = 2, flat_ubs0 = 2 + G * (1 + Fg), flat_dim = flat_ubs0 + (U - 1) * 3;
    for (int p = tid; p < U * G; p += nthr) {
        const int s = w.slot[p];
        const int i = p / G, m = p - i * G;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
        if (s >= 0) {
            r0 = (float)(((double)w.pos_g[m * 2] - w.pos_u[i * 2]) / ngt);       // np: float32 - float64 -> float64
            r1 = (float)(((double)w.pos_g[m * 2 + 1] - w.pos_u[i * 2 + 1]) / ngt);
            r2 = w.f_rate[m];
            r3 = w.f_avg[m];
            float* row = o.stage_gt + (size_t)(w.deg[2 * U + i] + s) * Fg;
            row[0] = r0; row[1] = r1; row[2] = r2;
            if (c.fair_service) row[3] = r3;
        }
        if (o.flat != nullptr) {
            float* fr = o.flat + (size_t)(b * U + i) * o.ld_flat + flat_gt0 + m * (1 + Fg);
            fr[0] = s >= 0 ? 1.f : 0.f;
            fr[1] = r0; fr[2] = r1; fr[3] = r2;
            if (c.fair_service) fr[4] = r3;
        }
    }
    for (int p = tid; p < U * U; p += nthr) {
        const int s = w.nslot[p];
        const int i = p / U, j = p - i * U;
        float r0 = 0.f, r1 = 0.f;
        if (s >= 0) {
            r0 = (float)((w.pos_u[j * 2] - w.pos_u[i * 2]) / nub);
            r1 = (float)((w.pos_u[j * 2 + 1] - w.pos_u[i * 2 + 1]) / nub);
            float* row = o.stage_ubs + (size_t)(w.deg[3 * U + i] + s) * 2;
            row[0] = r0; row[1] = r1;
        }
        if (o.flat != nullptr && j != i) {
            float* fr = o.flat + (size_t)(b * U + i) * o.ld_flat + flat_ubs0 + (j < i ? j : j - 1) * 3;
            fr[0] = s >= 0 ? 1.f : 0.f;
            fr[1] = r0; fr[2] = r1;
        }
    }
    if (o.flat != nullptr) {
        for (int p = tid; p < U * 2; p += nthr)                                   // own features (:223)
            o.flat[(size_t)(b * U + p / 2) * o.ld_flat + (p & 1)] = (float)(w.pos_u[p] / c.range_pos);
        const int pad = (int)o.ld_flat - flat_dim;
        for (int p = tid; p < U * pad; p += nthr) o.flat[(size_t)(b * U + p / pad) * o.ld_flat + flat_dim + p % pad] = 0.f;
    }
}

// Scratch layout (words) shared by the step and pack kernels and the host harness.
struct Scratch {
    UBS_HD static int64_t al(int64_t x) { return (x + 3) & ~(int64_t)3; }
    int64_t off_off_seen, off_off_near, off_tot, off_stage_gt, off_stage_ubs, words, gt_stride, ubs_stride;
    UBS_HD Scratch(const ubs_env_cfg& c, int64_t B) {
        const int64_t N = B * c.n_ubs;
        const int Fg = c.fair_service ? 4 : 3;
        gt_stride = al((int64_t)c.n_ubs * c.n_gts * Fg);
        ubs_stride = al((int64_t)c.n_ubs * (c.n_ubs > 1 ? c.n_ubs - 1 : 1) * 2);
        off_off_seen = 0;
        off_off_near = al(N);
        off_tot = off_off_near + al(N);
        off_stage_gt = off_tot + al(2 * B);
        off_stage_ubs = off_stage_gt + B * gt_stride;
        words = off_stage_ubs + B * ubs_stride;
    }
};

}  // namespace ubs_env
