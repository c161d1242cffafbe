// On-device episode reset: initial UBS / GT layouts and GT priorities for B env instances in one launch.
//
// Replaces, for B instances at once and without the host, what the reference draws at every `reset()`:
//   Map.set_positions / HotSpot.set_positions / DenseHotSpot.set_positions      (envs/mubs_cov/maps.py:30-34,64-77,97-113)
//   select_from_cube  (random.sample of the grid points of a cube)                (envs/common.py:13-16)
//   self.prior_gts = np.random.permutation(self.n_gts)                            (envs/mubs_cov/mubs_cov.py:98)
// The reference consumes python's `random` and numpy's legacy global RNG; reproducing those streams bit for bit on the
// device would serialise B Mersenne twisters, so this sampler draws from a counter-based Philox4x32-10 stream instead:
// the DISTRIBUTION is the reference's (tests compare it with the RNG-matched host sampler `envs.sample_layouts`), the
// individual draws are not.  Every draw is addressed by (seed, env instance, episode, purpose, index): no RNG state, the
// same (seed, episode) always gives the same layouts, on any grid size.
//
// One CTA per env instance.  "n distinct cells of a small cube" (hotspot cells, the row shuffle, the priority
// permutation) = rank of independent random keys, computed by counting in parallel (n <= 1024); the UBS cells come
// from a cube that may be large (range_pos^2) and are few (<= 32): sequential rejection by one thread.
#include "common.cuh"
#include "../../include/ubs_env.h"

namespace ubs_env {

constexpr int NT = 256, MAX_KEYS = 1024;

struct Philox { uint32_t k0, k1; };

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
}
// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1) -> 4 x 32 random bits
__device__ __forceinline__ void philox4(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct Rng {
    uint32_t k0, k1, env, episode;
    // 32 random bits of draw `i` of stream `purpose`
    __device__ uint32_t u32(uint32_t purpose, uint32_t i) const {
        uint32_t c[4] = {i, purpose, env, episode};
        philox4(c, k0, k1);
        return c[0];
    }
    // uniform integer in [0, n)
    __device__ uint32_t below(uint32_t purpose, uint32_t i, uint32_t n) const { return (uint32_t)(((uint64_t)u32(purpose, i) * n) >> 32); }
    // uniform double in [0, 1) with 53 random bits (numpy's random_sample construction)
    __device__ double uniform(uint32_t purpose, uint32_t i) const {
        uint32_t c[4] = {i, purpose, env, episode};
        philox4(c, k0, k1);
        return ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) / 9007199254740992.0;
    }
};

struct LayoutArgs {
    ubs_env_layout_cfg L;
    ubs_env_state st;
    uint64_t seed; uint32_t episode; int B;
};

// rank[i] = position of key i in the sorted order of keys[0..n) (ties by index): a uniformly random permutation
__device__ void rank_keys(const uint32_t* keys, int n, int* rank) {
    for (int i = threadIdx.x; i < n; i += NT) {
        const uint32_t k = keys[i];
        int r = 0;
        for (int j = 0; j < n; ++j) {
            const uint32_t kj = keys[j];
            r += (kj < k) || (kj == k && j < i);
        }
        rank[i] = r;
    }
}

__global__ void __launch_bounds__(NT) env_layout_kernel(const LayoutArgs a) {
    __shared__ uint32_t keys[MAX_KEYS];
    __shared__ int rank[MAX_KEYS];
    __shared__ int cell_of[MAX_KEYS];          // k-th sampled hotspot cell
    __shared__ float gx[MAX_KEYS], gy[MAX_KEYS];
    __shared__ int spot[2];
    const ubs_env_layout_cfg& L = a.L;
    const int b = blockIdx.x, U = L.n_ubs, G = L.n_gts;
    const Rng rng{(uint32_t)a.seed, (uint32_t)(a.seed >> 32), (uint32_t)b, a.episode};
    double* pos_ubs = a.st.pos_ubs + (size_t)b * U * 2;
    float* pos_gts = a.st.pos_gts + (size_t)b * G * 2;
    int32_t* prior = a.st.prior + (size_t)b * G;
    const double ubs_pitch = L.kind == 0 ? 1.0 : L.min_dist;

    // ---- UBS cells: n_ubs distinct grid points (select_from_cube(n_ubs, 0, cells, 2)); hotspot origin
    if (threadIdx.x == 0) {
        const uint32_t C = (uint32_t)L.ubs_cells * (uint32_t)L.ubs_cells;
        uint32_t chosen[UBS_ENV_MAX_UBS];
        uint32_t draw = 0;
        for (int i = 0; i < U; ++i) {
            uint32_t cell;
            bool dup;
            do {
                cell = rng.below(1, draw++, C);
                dup = false;
                for (int j = 0; j < i; ++j) dup |= chosen[j] == cell;
            } while (dup);
            chosen[i] = cell;
            pos_ubs[2 * i] = ubs_pitch * (double)(cell / L.ubs_cells);
            pos_ubs[2 * i + 1] = ubs_pitch * (double)(cell % L.ubs_cells);
        }
        if (L.kind != 0) {
            spot[0] = (int)rng.below(2, 0, (uint32_t)L.spot_cells);
            spot[1] = (int)rng.below(2, 1, (uint32_t)L.spot_cells);
        }
    }
    // ---- hotspot cells: n distinct cells of the range_spot x range_spot cube, in random order
    const int S = L.kind == 0 ? 0 : L.range_spot * L.range_spot;
    const int n_cells = L.kind == 1 ? G : L.n_grps;
    for (int i = threadIdx.x; i < S; i += NT) keys[i] = rng.u32(3, (uint32_t)i);
    __syncthreads();
    if (S > 0) rank_keys(keys, S, rank);
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += NT)
        if (rank[i] < n_cells) cell_of[rank[i]] = i;
    __syncthreads();

    // ---- GT positions (before the shuffle)
    for (int g = threadIdx.x; g < G; g += NT) {
        double x, y;
        if (L.kind == 0) {                                   // Map: uniform grid points (duplicates resolved below)
            x = (double)rng.below(4, 2 * g, (uint32_t)L.ubs_cells);
            y = (double)rng.below(4, 2 * g + 1, (uint32_t)L.ubs_cells);
        } else {
            const int k = L.kind == 1 ? g : g / L.gts_per_grp;
            const int cell = cell_of[k];
            const double ox = L.min_dist * L.range_spot * spot[0], oy = L.min_dist * L.range_spot * spot[1];
            x = ox + L.min_dist * (double)(cell / L.range_spot);
            y = oy + L.min_dist * (double)(cell % L.range_spot);
            if (L.kind == 2) {                               // r_cov * (rand(gts_per_grp, 2) - 0.5) around the group centre
                x += L.r_cov * (rng.uniform(5, 2 * g) - 0.5);
                y += L.r_cov * (rng.uniform(5, 2 * g + 1) - 0.5);
            }
        }
        gx[g] = fminf(fmaxf((float)x, 0.f), (float)L.range_pos);      // np.clip(pos_gts, 0, range_pos) on the float32 array
        gy[g] = fminf(fmaxf((float)y, 0.f), (float)L.range_pos);
    }
    __syncthreads();
    if (L.kind == 0 && threadIdx.x == 0) {                    // distinct grid points: redraw duplicates (rare: G << cells^2)
        uint32_t draw = 0;
        for (int g = 1; g < G; ++g) {
            bool dup = true;
            while (dup) {
                dup = false;
                for (int j = 0; j < g; ++j) dup |= gx[j] == gx[g] && gy[j] == gy[g];
                if (dup) {
                    gx[g] = (float)rng.below(6, draw++, (uint32_t)L.ubs_cells);
                    gy[g] = (float)rng.below(6, draw++, (uint32_t)L.ubs_cells);
                }
            }
        }
    }
    // ---- np.random.shuffle(pos_gts) (HotSpot / DenseHotSpot): a uniformly random row permutation
    for (int i = threadIdx.x; i < G; i += NT) keys[i] = rng.u32(7, (uint32_t)i);
    __syncthreads();
    rank_keys(keys, G, rank);
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += NT) {
        const int dst = L.kind == 0 ? g : rank[g];
        pos_gts[2 * dst] = gx[g];
        pos_gts[2 * dst + 1] = gy[g];
    }
    __syncthreads();
    // ---- prior_gts = np.random.permutation(n_gts)
    for (int i = threadIdx.x; i < G; i += NT) keys[i] = rng.u32(8, (uint32_t)i);
    __syncthreads();
    rank_keys(keys, G, rank);
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += NT) prior[rank[g]] = g;
}

}  // namespace ubs_env

extern "C" UBS_ENV_API int ubs_env_sample_layouts(const ubs_env_layout_cfg* L, const ubs_env_state* st, uint64_t seed,
                                                  uint32_t episode, int64_t B, void* stream) {
    UBS_REQUIRE(L && st && st->pos_ubs && st->pos_gts && st->prior, "ubs_env_sample_layouts: NULL argument");
    UBS_REQUIRE(L->kind >= 0 && L->kind <= 2, "ubs_env_sample_layouts: map kind %d has no device sampler (0 grid, 1 HotSpot, 2 DenseHotSpot)", L->kind);
    UBS_REQUIRE(L->n_ubs >= 1 && L->n_ubs <= UBS_ENV_MAX_UBS && L->n_gts >= 1 && L->n_gts <= ubs_env::MAX_KEYS,
                "ubs_env_sample_layouts: n_ubs <= %d and n_gts <= %d", UBS_ENV_MAX_UBS, ubs_env::MAX_KEYS);
    UBS_REQUIRE(L->ubs_cells >= 1 && (int64_t)L->ubs_cells * L->ubs_cells >= L->n_ubs, "ubs_env_sample_layouts: UBS grid too small");
    if (L->kind != 0) {
        const int S = L->range_spot * L->range_spot;
        const int n = L->kind == 1 ? L->n_gts : L->n_grps;
        UBS_REQUIRE(L->range_spot >= 1 && S <= ubs_env::MAX_KEYS && S >= n && L->spot_cells >= 1,
                    "ubs_env_sample_layouts: hotspot of %d cells cannot hold %d points (or exceeds %d cells)", S, n, ubs_env::MAX_KEYS);
        UBS_REQUIRE(L->kind != 2 || (L->gts_per_grp >= 1 && L->n_grps * L->gts_per_grp == L->n_gts),
                    "ubs_env_sample_layouts: n_gts != n_grps * gts_per_grp");
    }
    if (B == 0) return 0;
    ubs_env::LayoutArgs a{*L, *st, seed, episode, (int)B};
    ubs_env::env_layout_kernel<<<(unsigned)B, ubs_env::NT, 0, (cudaStream_t)stream>>>(a);
    return ubs::check_launch("ubs_env_sample_layouts");
}
