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
// ndarray.sum() / np.mean() of a 1-D float32 array computes.
UBS_HD inline float np_sum_f32(const float* a, int n) {
    if (n < 8) {
        float res = -0.0f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_sum_f32(a, n2) + np_sum_f32(a + n2, n - n2);
}

// Carves the per-env workspace (shared memory on the device) out of one 8-byte aligned buffer.
struct Work {
    double *pos_u, *gain;
    float *pos_g, *d_u2g, *d_u2u, *pitf, *rate, *avg, *tmp, *tmp2, *scal;
    int *prior, *occ, *nsch, *sch_ubs, *sch_rb, *slot, *nslot, *deg;
    unsigned char* order;

    UBS_HD static size_t bytes(int U, int G, int R) {
        size_t n = 8 * ((size_t)U * 2 + (size_t)U * G);
        n += 4 * ((size_t)G * 2 + (size_t)U * G + (size_t)U * U + (size_t)U * G + 4 * (size_t)G + 8);
        n += 4 * ((size_t)G + (size_t)U * R + U + 2 * (size_t)G + (size_t)U * G + (size_t)U * U + 4 * (size_t)U + 4);
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
        scal = f; f += 8;
        int* i = (int*)f;
        prior = i; i += G;
        occ = i; i += U * R;
        nsch = i; i += U;
        sch_ubs = i; i += G;
        sch_rb = i; i += G;
        slot = i; i += (size_t)U * G;
        nslot = i; i += U * U;
        deg = i; i += 4 * U + 4;          // deg_seen[U] | deg_near[U] | off_seen[U] | off_near[U] | totals
        order = (unsigned char*)i;
    }
};

struct SerialCtx {
    int tid = 0, nthr = 1;
    void sync() const {}
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
    int32_t* deg_seen;  // (N) all envs
    int32_t* deg_near;  // (N)
};

// scal[] slots
enum { S_FAIR = 0, S_GU = 1, S_MEAN = 2, S_AGU = 3, S_TPUT = 4 };

template <class Ctx>
UBS_HD inline void env_run(const ubs_env_cfg& c, const ubs_env_state& st, int64_t b, const int64_t* actions,
                           bool is_reset, const Work& w, const EnvOut& o, Ctx& ctx) {
    const int U = c.n_ubs, G = c.n_gts, R = c.n_rbs;
    const int tid = ctx.tid, nthr = ctx.nthr;
    const int Fg = c.fair_service ? 4 : 3;
    double* info = st.info + b * UBS_ENV_INFO;

    // ---- load state; step(): self.t += 1; pos_ubs = clip(pos_ubs + avail_moves[actions], 0, range_pos)  (:105-109)
    int t = is_reset ? 0 : st.t[b] + 1;
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
    for (int i = tid; i < U; i += nthr) w.nsch[i] = 0;
    ctx.sync();

    // ---- _transmit_data step 1: distances (:133-139).  np: norm of a float64 difference stored into a float32 array
    const float r_cov = (float)c.r_cov, r_sns = (float)c.r_sns, r_comm = (float)c.r_comm;
    for (int p = tid; p < U * G; p += nthr) {
        const int i = p / G, m = p - i * G;
        const double dx = (double)w.pos_g[m * 2] - w.pos_u[i * 2], dy = (double)w.pos_g[m * 2 + 1] - w.pos_u[i * 2 + 1];
        const float dl = (float)sqrt(dx * dx + dy * dy);
        w.d_u2g[p] = dl;
        // channel gain (common.py:45-55).  np: p_los stays float32 (python scalars are weak); np.square(h_ubs) is a
        // float64 scalar, so d, fspl, pl and the gain are float64.
        const float e = expf((float)(-c.chan_b) * (atanf((float)c.h_ubs / (dl + 1e-5f)) - (float)c.chan_a));
        const float p_los = 1.0f / (1.0f + (float)c.chan_a * e);
        const double d3 = sqrt((double)(dl * dl) + c.h_ubs * c.h_ubs);
        const double q = c.c_fspl * d3 / 3e8;
        const double fspl = q * q;
        const double pl = (double)p_los * fspl * c.k_los + (double)(1.0f - p_los) * fspl * c.k_nlos;
        const double g = 1.0 / pl;
        w.gain[p] = g;
        // p_itf[i, :, rb] = p_tx * g[i] * mask_itf[i]  (:176), float64 product stored into the float32 p_itf
        w.pitf[p] = (dl <= r_cov) ? (float)(c.p_tx * g) : 0.0f;
    }
    for (int p = tid; p < U * U; p += nthr) {
        const int i = p / U, j = p - i * U;
        const double dx = w.pos_u[j * 2] - w.pos_u[i * 2], dy = w.pos_u[j * 2 + 1] - w.pos_u[i * 2 + 1];
        w.d_u2u[p] = (float)sqrt(dx * dx + dy * dy);
    }
    ctx.sync();
    // nearest_ubs = np.argsort(d_u2g[:, m])  (:167): insertion sort (numpy's small-array path), ties by index
    for (int m = tid; m < G; m += nthr) {
        unsigned char* ord = w.order + (size_t)m * U;
        for (int i = 0; i < U; ++i) {
            const float key = w.d_u2g[i * G + m];
            int j = i;
            while (j > 0 && w.d_u2g[ord[j - 1] * G + m] > key) { ord[j] = ord[j - 1]; --j; }
            ord[j] = (unsigned char)i;
        }
    }
    ctx.sync();

    // ---- step 2: greedy RB scheduling in priority order (:166-178) — inherently sequential over GTs
    if (tid == 0) {
        for (int k = 0; k < G; ++k) {
            const int m = w.prior[k];
            const unsigned char* ord = w.order + (size_t)m * U;
            for (int jj = 0; jj < U; ++jj) {
                const int i = ord[jj];
                const float dim = w.d_u2g[i * G + m];
                if (!(dim <= r_cov)) break;                       // sorted by distance: nobody further covers m either
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

    // ---- rates (:181-187)
    for (int m = tid; m < G; m += nthr) {
        float r = 0.f;
        const int i = w.sch_ubs[m];
        if (i >= 0) {
            const int rb = w.sch_rb[m];
            float a[UBS_ENV_MAX_UBS];
            for (int i2 = 0; i2 < U; ++i2) {
                const int oc = w.occ[i2 * R + rb];
                a[i2] = (oc >= 0 && oc != m) ? w.pitf[i2 * G + m] : 0.f;       // p_itf[i, m, rb] = 0 for the served GT (:177)
            }
            const float itf = np_sum_f32(a, U);                                // np: (U,1) float32 copy .sum() -> pairwise
            const double sinr = (c.p_tx * w.gain[i * G + m]) / ((double)itf + c.bw * c.n0);
            r = (float)(c.bw * log2(1.0 + sinr) * 1e-6);                       // np: float64, stored into float32
        }
        w.rate[m] = r;
    }
    ctx.sync();

    // ---- step 3 (:193-199): running averages (np: float32 throughout; self.t is a weak python int)
    for (int m = tid; m < G; m += nthr) {
        const float a = (w.avg[m] * (float)t + w.rate[m]) / (float)(t + 1);
        w.avg[m] = a;
        const float x = a < 1e-6f ? 1e-6f : a;                                 // np.clip(x, 1e-6, inf)
        w.tmp[m] = x;
        w.tmp2[m] = x * x;
    }
    ctx.sync();
    if (tid == 0) {
        const float s1 = np_sum_f32(w.tmp, G), s2 = np_sum_f32(w.tmp2, G);
        const float fair = (s1 * s1) / ((float)G * s2);                        // compute_jain_fairness_index
        const float rsum = np_sum_f32(w.rate, G);
        const float mean = rsum / (float)G;                                    // np.mean of float32
        const float gu = fair * mean;                                          // global_util
        const float agu0 = is_reset ? 0.f : (float)info[3];
        const float agu = (agu0 * (float)t + gu) / (float)(t + 1);
        w.scal[S_FAIR] = fair;
        w.scal[S_GU] = gu;
        w.scal[S_MEAN] = mean;
        w.scal[S_AGU] = agu;
        w.scal[S_TPUT] = rsum * (float)c.dt / 1e3f;
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
    ctx.sync();

    // ---- reward (:296-312), terminate (:314-316), collisions (:142-143)
    const bool done = (t == c.episode_limit) && !is_reset;
    if (tid < U || nthr == 1) {
        for (int i = (nthr == 1 ? 0 : tid); i < U; i += (nthr == 1 ? 1 : U)) {
            bool coll = false;
            for (int j = 0; j < U; ++j) coll = coll || (j != i && (double)w.d_u2u[i * U + j] < c.safe_dist);
            double rate_ubs = 0.0;                                             // rate_per_ubs (:188); only `== 0` is used
            for (int m = 0; m < G; ++m) rate_ubs += (w.sch_ubs[m] == i) ? (double)w.rate[m] : 0.0;
            // np: rew_scale * float32 stays float32, / max_rate (np.float64) promotes to float64
            const float base = c.fair_service ? w.scal[S_GU] : w.scal[S_MEAN];
            double lr = (double)((float)c.rew_scale * base) / c.max_rate;
            lr = lr * (rate_ubs == 0.0 ? 0.0 : 1.0);
            if (c.avoid_collision) lr = (coll ? 0.0 : 1.0) * lr - (coll ? 1.0 : 0.0) * c.penalty;
            o.rew[b * U + i] = is_reset ? 0.f : (float)lr;
            w.tmp[i] = is_reset ? 0.f : (float)lr;
            w.nsch[i] = coll ? 1 : 0;                                          // reuse: collision flag
            // own features (:223) + talk mask: bit i of mask[dst j] <=> d_u2u[i, j] <= r_comm  (env_wrappers.py:141-144)
            o.x_agent[(b * U + i) * 2] = (float)(w.pos_u[i * 2] / c.range_pos);
            o.x_agent[(b * U + i) * 2 + 1] = (float)(w.pos_u[i * 2 + 1] / c.range_pos);
            uint32_t mk = 0;
            for (int s = 0; s < U && s < 32; ++s) mk |= (w.d_u2u[s * U + i] <= r_comm) ? (1u << s) : 0u;
            o.mask[b * U + i] = (int32_t)mk;
            // per-agent compaction slots of visible GTs / UBSs (env_wrappers.py:71: rows with flag == 1, in order)
            int cnt = 0;
            for (int m = 0; m < G; ++m) w.slot[i * G + m] = (w.d_u2g[i * G + m] <= r_sns) ? cnt++ : -1;
            w.deg[i] = cnt;
            cnt = 0;
            for (int j = 0; j < U; ++j) w.nslot[i * U + j] = (j != i && w.d_u2u[i * U + j] <= r_comm) ? cnt++ : -1;
            w.deg[U + i] = cnt;
        }
    }
    ctx.sync();
    if (tid == 0) {
        int a1 = 0, a2 = 0, ncoll = 0;
        double rmean = 0.0;
        for (int i = 0; i < U; ++i) {
            w.deg[2 * U + i] = a1;
            w.deg[3 * U + i] = a2;
            a1 += w.deg[i];
            a2 += w.deg[U + i];
            o.deg_seen[b * U + i] = w.deg[i];
            o.deg_near[b * U + i] = w.deg[U + i];
            ncoll += w.nsch[i];
            rmean += (double)w.tmp[i];
        }
        o.done[b] = done ? 1.f : 0.f;
        o.bad[b] = done ? 1.f : 0.f;                                            // BadMask: t == episode_limit (:122)
        st.t[b] = t;
        info[0] = (is_reset ? 0.0 : info[0]) + rmean / U;                       // ep_ret
        info[1] = (double)((is_reset ? 0.f : (float)info[1]) + w.scal[S_TPUT]); // total_throughput (np: float32)
        info[2] = (is_reset ? 0.0 : info[2]) + ncoll / 2.0;                     // n_colls
        info[3] = w.scal[S_AGU];
        info[4] = w.scal[S_FAIR];
        info[5] = w.scal[S_GU];
        info[6] = t;
        info[7] = 0.0;
    }
    ctx.sync();

    // ---- persistent state + observation rows (:225-240), compacted per env into the staging area
    for (int i = tid; i < U * 2; i += nthr) st.pos_ubs[b * U * 2 + i] = w.pos_u[i];
    for (int m = tid; m < G; m += nthr) {
        st.avg_rate[b * G + m] = w.avg[m];
        st.rate[b * G + m] = w.rate[m];
        st.sched[(b * G + m) * 2] = w.sch_ubs[m];
        st.sched[(b * G + m) * 2 + 1] = w.sch_rb[m];
    }
    const double ngt = fmin(c.range_pos, c.r_sns), nub = fmin(c.range_pos, c.r_comm);
    const double fair_scale_num = (double)G, fair_scale_den = (double)(U * R);
    for (int p = tid; p < U * G; p += nthr) {
        const int s = w.slot[p];
        if (s < 0) continue;
        const int i = p / G, m = p - i * G;
        float* row = o.stage_gt + (size_t)(w.deg[2 * U + i] + s) * Fg;
        row[0] = (float)(((double)w.pos_g[m * 2] - w.pos_u[i * 2]) / ngt);       // np: float32 - float64 -> float64
        row[1] = (float)(((double)w.pos_g[m * 2 + 1] - w.pos_u[i * 2 + 1]) / ngt);
        row[2] = (float)((double)w.rate[m] / c.max_rate);
        if (c.fair_service) row[3] = (float)((double)w.avg[m] / c.max_rate * fair_scale_num / fair_scale_den);
    }
    for (int p = tid; p < U * U; p += nthr) {
        const int s = w.nslot[p];
        if (s < 0) continue;
        const int i = p / U, j = p - i * U;
        float* row = o.stage_ubs + (size_t)(w.deg[3 * U + i] + s) * 2;
        row[0] = (float)((w.pos_u[j * 2] - w.pos_u[i * 2]) / nub);
        row[1] = (float)((w.pos_u[j * 2 + 1] - w.pos_u[i * 2 + 1]) / nub);
    }
}

// Scratch layout (words) shared by the step and pack kernels and the host harness.
struct Scratch {
    UBS_HD static int64_t al(int64_t x) { return (x + 3) & ~(int64_t)3; }
    int64_t off_deg_seen, off_deg_near, off_stage_gt, off_stage_ubs, words, gt_stride, ubs_stride;
    UBS_HD Scratch(const ubs_env_cfg& c, int64_t B) {
        const int64_t N = B * c.n_ubs;
        const int Fg = c.fair_service ? 4 : 3;
        gt_stride = al((int64_t)c.n_ubs * c.n_gts * Fg);
        ubs_stride = al((int64_t)c.n_ubs * (c.n_ubs > 1 ? c.n_ubs - 1 : 1) * 2);
        off_deg_seen = 0;
        off_deg_near = al(N);
        off_stage_gt = off_deg_near + al(N);
        off_stage_ubs = off_stage_gt + B * gt_stride;
        words = off_stage_ubs + B * ubs_stride;
    }
};

}  // namespace ubs_env
