// ta_her.cuh -- hindsight relabelling plan for a [T][n] rollout.
//
// Restates Buffer_gridworld.her_func (soa/env_buffer.py:101-143) per finished episode:
//   :108     states_mask, indices, counts = np.unique(buffer_her['p'][:,4,0:2], return_index=True, axis=0)
//            -> first-occurrence record index of every distinct agent position, in sorted (y,x) order
//   :115     episode_idxs = np.random.choice(indices, size=min(4, len(indices)), replace=False)
//   :118-127 for each chosen index > 0: the episode prefix 0..index is copied with g = the position
//            reached at `index`, r[index] = 0.9, d[index] = 1
// Buffer_gridworld.pre_her_func (soa/env_buffer.py:145-210) is the same selection over the 9-frame records of
// train_ppo_predictor.py, which exist from the episode's 5th step on: record j there is step j + 4 here, so it
// is this kernel with first_rec = 4 (candidates = steps >= 4, `index > 0` = step > 4); the four shifted pad
// records it appends are the 5-frame records of the prefix's last four steps (see her.py).
// The copies are not materialised: the kernel emits, for every record and each of the 4 relabel
// slots, whether the record belongs to that slot's prefix and with which goal ("plan"); the host
// turns the non-empty entries into (source record, goal, reward) triples.
//
// One warp per env walks its episode segments (a segment starts at t = 0 or after a done record
// and is relabelled only if it ENDS inside the window); lanes hold the segment's positions (two
// per lane, segments are at most 50 records long because max_steps = 50).
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int HER_WARPS = 4;
constexpr int HER_MAXLEN = 64;
constexpr uint32_t HER_NONE = 0xFFFFu;

__global__ void __launch_bounds__(32 * HER_WARPS)
her_plan_kernel(const float *__restrict__ p, const uint8_t *__restrict__ done, int T, long long n, int first_rec, uint32_t seed_lo,
                uint32_t seed_hi, unsigned long long env_id0, const uint8_t *__restrict__ chosen_in, uint8_t *uniq_out,
                uint8_t *m_out, uint16_t *plan) {
    __shared__ uint8_t s_sorted[HER_WARPS][HER_MAXLEN];
    __shared__ uint8_t s_perm[HER_WARPS][HER_MAXLEN];
    __shared__ uint8_t s_chosen[HER_WARPS][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long env = (long long)blockIdx.x * HER_WARPS + warp;
    if (env >= n) return;
    int t0 = 0;
    for (int tb = 0; tb < T; tb += 32) {
        const int t = tb + lane;
        uint32_t ends = __ballot_sync(0xFFFFFFFFu, t < T && done[(long long)t * n + env] != 0);
        while (ends) {
            const int t1 = tb + __ffs(ends) - 1;
            ends &= ends - 1;
            const int L = t1 - t0 + 1;
            if (L <= HER_MAXLEN) {
                // positions of the segment's records: i = lane and i = lane + 32
                uint32_t key[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = lane + 32 * h;
                    key[h] = HER_NONE;
                    if (i < L) {
                        const float *q = p + ((long long)(t0 + i) * n + env) * 10 + 8;  // p[t][env][4][0..1] = (y, x)
                        key[h] = ((uint32_t)(int)q[0] << 5) | (uint32_t)(int)q[1];
                    }
                }
                // first occurrence of each distinct position
                // (records before first_rec are no candidates: pre_her_func only sees records from the 5th step on)
                bool first[2] = {lane >= first_rec && lane < L, lane + 32 >= first_rec && lane + 32 < L};
                for (int j = first_rec; j < L; j++) {
                    const uint32_t kj = j < 32 ? __shfl_sync(0xFFFFFFFFu, key[0], j) : __shfl_sync(0xFFFFFFFFu, key[1], j - 32);
                    if (j < lane && kj == key[0]) first[0] = false;
                    if (j < lane + 32 && kj == key[1]) first[1] = false;
                }
                // np.unique order: rank among the distinct positions, sorted by (y, x)
                int rank[2] = {0, 0};
                const uint32_t f0 = __ballot_sync(0xFFFFFFFFu, first[0]), f1 = __ballot_sync(0xFFFFFFFFu, first[1]);
                for (int j = 0; j < L; j++) {
                    const uint32_t kj = j < 32 ? __shfl_sync(0xFFFFFFFFu, key[0], j) : __shfl_sync(0xFFFFFFFFu, key[1], j - 32);
                    const bool fj = j < 32 ? (f0 >> j) & 1u : (f1 >> (j - 32)) & 1u;
                    if (fj) {
                        rank[0] += kj < key[0];
                        rank[1] += kj < key[1];
                    }
                }
                const int m = __popc(f0) + __popc(f1);
                if (first[0]) s_sorted[warp][rank[0]] = (uint8_t)lane;
                if (first[1]) s_sorted[warp][rank[1]] = (uint8_t)(lane + 32);
                __syncwarp();
                const long long rec = (long long)t1 * n + env;
                if (m_out && lane == 0) m_out[rec] = (uint8_t)m;
                if (uniq_out) {
                    for (int i = lane; i < HER_MAXLEN; i += 32) uniq_out[rec * HER_MAXLEN + i] = i < m ? s_sorted[warp][i] : 0xFF;
                }
                const int k = m < 4 ? m : 4;
                if (chosen_in) {
                    if (lane < 4) s_chosen[warp][lane] = lane < k ? chosen_in[rec * 4 + lane] : 0xFF;
                } else {
                    // production draw: k distinct entries by a partial Fisher-Yates over one Philox block
                    for (int i = lane; i < HER_MAXLEN; i += 32) s_perm[warp][i] = (uint8_t)i;
                    __syncwarp();
                    if (lane == 0) {
                        const unsigned long long gid = env_id0 + (unsigned long long)env;
                        uint32_t w[4];
                        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t1, 0x48455221u, seed_lo, seed_hi, w);
                        for (int c = 0; c < 4; c++) {
                            if (c < k) {
                                const int j = c + (int)__umulhi(w[c], (uint32_t)(m - c));
                                const uint8_t a = s_perm[warp][c], b = s_perm[warp][j];
                                s_perm[warp][c] = b;
                                s_perm[warp][j] = a;
                                s_chosen[warp][c] = s_sorted[warp][b];
                            } else {
                                s_chosen[warp][c] = 0xFF;
                            }
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int idx = s_chosen[warp][c];
                    if (idx == 0xFF || idx <= first_rec || idx >= L) continue;  // :121 `if index > 0`
                    const uint32_t goal = idx < 32 ? __shfl_sync(0xFFFFFFFFu, key[0], idx) : __shfl_sync(0xFFFFFFFFu, key[1], idx - 32);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int i = lane + 32 * h;
                        if (i <= idx) plan[((long long)(t0 + i) * n + env) * 4 + c] = (uint16_t)(goal | (i == idx ? 0x8000u : 0u));
                    }
                }
                __syncwarp();
            }
            t0 = t1 + 1;
        }
    }
}

}  // namespace ta
