/* ubs_env.h — C ABI of the device-resident, vectorised MultiUbsCoverageEnv + on-device observation-graph builder.
 *
 * SURVEY.md §8(f) rows 1 and 2.  One call advances B independent env instances by one step and writes the NEXT
 * observation of all of them straight into an observation packet of a sequence arena (uav_bs_ctrl_b200/arena.py),
 * i.e. it replaces, for B envs at once and without leaving the device,
 *   - MultiUbsCoverageEnv.step / reset / _transmit_data / get_obs / _get_reward / _get_terminate
 *                                          (envs/mubs_cov/mubs_cov.py:85-127,129-211,213-242,296-316)
 *   - AirToGroundChannel.estimate_chan_gain, compute_jain_fairness_index   (envs/common.py:19-25,45-55)
 *   - GraphObservation.build_obs_graph / local_observation, MultiUbsCoverageWrapper.build_comm_graph + dgl.merge
 *                                          (algos/madrqn/utils/env_wrappers.py:65-89,122-154)
 *   - common.cat -> dgl.batch over the B env instances                       (algos/common.py:40-47)
 * Arithmetic follows the reference statement by statement, including which intermediates numpy keeps in float32 and
 * which in float64 (NEP 50 promotion, numpy >= 2; see tests/golden/make_env_golden.py) and numpy's pairwise
 * summation order; argsort ties are broken by index (a valid np.argsort; numpy's own tie order is unspecified).
 * Initial layouts (maps.py set_positions) come either from the caller (RNG-matched host sampler) or from
 * ubs_env_sample_layouts (device, Philox).
 *
 * Conventions: as ubs_gnn.h — device pointers owned by the caller, asynchronous launches on `stream`, 0 = success,
 * message through ubs_last_error().  The config struct is passed by host pointer and copied into the launch.
 */
#ifndef UBS_ENV_H_
#define UBS_ENV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define UBS_ENV_API __attribute__((visibility("default")))
#else
#define UBS_ENV_API
#endif

#define UBS_ENV_MAX_ACTIONS 33
#define UBS_ENV_MAX_UBS 32
#define UBS_ENV_INFO 8 /* doubles per env: ep_ret, total_throughput, n_colls, avg_global_util, fair_idx, global_util, t, - */

/* Static parameters of a map + env (maps.py:7-34, mubs_cov.py:13-66); every derived constant is computed by the host
 * with the reference's own expression so that it is bit-identical (uav_bs_ctrl_b200/envs.py). */
typedef struct {
    int32_t n_ubs, n_gts, n_rbs, n_actions, episode_limit, fair_service, avoid_collision;
    int32_t gts_f64; /* the map returns float64 GT positions (HotSpot, Map, DenseHotSpotV2): get_state divides in float64 */
    double range_pos, r_cov, r_sns, r_comm; /* metres; r_sns / r_comm may be +inf */
    double dt, rew_scale;                   /* maps.py: dt, reward_scale_rate */
    double h_ubs, p_tx, n0, bw;             /* mubs_cov.py:14-17 */
    double c_fspl;                          /* 4 * pi * fc                    (common.py:52) */
    double chan_a, chan_b, k_los, k_nlos;   /* a, b, 10**(eta_los/20), 10**(eta_nlos/20) (common.py:31-54) */
    double max_rate;                        /* mubs_cov.py:36-39 */
    double safe_dist, penalty;              /* mubs_cov.py:20-21 */
    double moves[UBS_ENV_MAX_ACTIONS][2];   /* avail_moves (mubs_cov.py:61-65) */
} ubs_env_cfg;

/* Per-env state, all device pointers, B envs back to back. */
typedef struct {
    double* pos_ubs;   /* (B, U, 2)  float64 like the reference (maps.py returns float64 UBS positions) */
    float* pos_gts;    /* (B, G, 2)  float32 (maps.py:107) */
    float* avg_rate;   /* (B, G)     avg_rate_per_gt */
    float* rate;       /* (B, G)     rate_per_gt of the last step */
    int32_t* prior;    /* (B, G)     prior_gts: GT ids, highest priority first */
    int32_t* t;        /* (B)        timer */
    double* info;      /* (B, UBS_ENV_INFO) */
    int32_t* sched;    /* (B, G, 2)  (serving UBS, RB) of each GT or (-1,-1); diagnostic / tests */
} ubs_env_state;

/* Where one observation packet lives (word = 4 bytes; offsets from `packet`; uav_bs_ctrl_b200/arena.py PacketLayout). */
typedef struct {
    int32_t* packet;
    int64_t off_x_gt, off_x_ubs, off_x_agent, off_ip_seen, off_ip_near, off_mask, off_rew, off_done, off_bad;
    int64_t off_state; /* (B, 2U + (3 + fair_service) G) global state of get_state() (mubs_cov.py:244-262) for QMIX; -1: not stored */
    int64_t off_flat;  /* (B*U, ld_flat) flattened local observations for the MLP encoder (FlattenedObservation,
                          env_wrappers.py:41-54: gym flattens the Dict in key order agent | gt | ubs, flags included);
                          -1: not stored */
    int64_t ld_flat;   /* row pitch in floats, >= 2 + G (2 + 3 + fair_service) ... see envs.flat_obs_dim(); the pad is zero */
} ubs_env_packet;

/* Parameters of the map's reset distribution (maps.py set_positions) for the on-device layout sampler. */
typedef struct {
    int32_t kind;        /* 0: Map (uniform grid points), 1: HotSpot (maps.py:64-77), 2: DenseHotSpot (maps.py:97-113) */
    int32_t n_ubs, n_gts, n_grps, gts_per_grp; /* n_grps / gts_per_grp: DenseHotSpot only */
    int32_t ubs_cells;   /* UBS grid points per axis: range_pos (kind 0) or range_pos // min_dist */
    int32_t range_spot;  /* hotspot side in cells: smallest s with s*s >= n_gts (HotSpot) / n_grps (DenseHotSpot) */
    int32_t spot_cells;  /* hotspot origins per axis: range_pos // min_dist // range_spot */
    double min_dist, range_pos, r_cov;
} ubs_env_layout_cfg;

/* reset(), first half, on the device: samples pos_ubs / pos_gts / prior of all B instances from the map's reset
 * distribution (Map / HotSpot / DenseHotSpot.set_positions + np.random.permutation(n_gts), mubs_cov.py:96-98) with a
 * counter-based Philox stream addressed by (seed, instance, episode): same distribution as the reference's host RNG,
 * not the same draws (the RNG-matched host sampler stays available through the caller: envs.sample_layouts).
 * Follow with ubs_env_reset.                                                                                        */
UBS_ENV_API int ubs_env_sample_layouts(const ubs_env_layout_cfg* layout, const ubs_env_state* st, uint64_t seed,
                                       uint32_t episode, int64_t B, void* stream);

/* Words of device scratch ubs_env_step / ubs_env_reset need for B envs (staging of the per-env compacted rows). */
UBS_ENV_API int64_t ubs_env_scratch_words(const ubs_env_cfg* cfg, int64_t B);

/* reset(): t = 0, running averages cleared, positions / initial priorities as currently stored in `st`
 * (the caller has just written pos_ubs, pos_gts, prior), one _transmit_data, observation -> packet with
 * reward = 0, done = bad = 0.                                                                                       */
UBS_ENV_API int ubs_env_reset(const ubs_env_cfg* cfg, const ubs_env_state* st, const ubs_env_packet* pk, int32_t* scratch,
                              int64_t B, void* stream);

/* step(actions): actions (B*U) int64 (what the fused act kernel writes).  Writes the next observation, the reward of
 * this transition, done and bad_mask (t == episode_limit) into the packet.                                          */
UBS_ENV_API int ubs_env_step(const ubs_env_cfg* cfg, const ubs_env_state* st, const int64_t* actions,
                             const ubs_env_packet* pk, int32_t* scratch, int64_t B, void* stream);

/* Diagnostic: with UBS_ENV_PROFILE=1 in the environment, CTA 0 stamps clock64() after every phase barrier of the step
 * kernel; this copies the 32 stamps of the last launch to the host (synchronises the device).                        */
UBS_ENV_API int ubs_env_phase_clocks(int64_t* out32);

#ifdef __cplusplus
}
#endif
#endif /* UBS_ENV_H_ */
