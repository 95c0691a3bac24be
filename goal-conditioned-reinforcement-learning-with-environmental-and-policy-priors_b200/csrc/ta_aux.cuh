// ta_aux.cuh -- reset / observe / state export-import / featurise kernels.
//
//   reset_*            MiniGridEnv.reset             gym_minigrid/minigrid.py:947-980
//                      Twoarmy_v4._gen_grid          gym_minigrid/envs/twoarmy_v4.py:38-80
//                      Twoarmy_v4.__init__ (hard)    gym_minigrid/envs/twoarmy_v4.py:9-36
//   observe_kernel     MiniGridEnv.gen_obs           gym_minigrid/minigrid.py:1443-1496
//   state_matrix_*     Env_transact.matrix_env       soa/env_buffer.py:300-318
//                      Env_transact.data_env         soa/env_buffer.py:320-334
//   stack_roll_kernel  frame-stack roll              soa/train_ppo.py:116-121,
//                      np.tile at episode start      soa/env_buffer.py:420-423
#pragma once
#include "ta_common.cuh"

namespace ta {

// One thread per (env, record word): rebuild the packed grid of masked envs from the template.
__global__ void reset_grid_kernel(uint32_t *grid, const uint32_t *tmpl, const uint8_t *mask, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * REC_WORDS) return;
    const long long e = i / REC_WORDS;
    if (mask && !mask[e]) return;
    grid[i] = tmpl[(int)(i - e * REC_WORDS)];
}

// One thread per env: agent, step_count, ball objects; hard also re-runs __init__'s flags.
__global__ void reset_scalar_kernel(uint4 *sc0, uint4 *sc1, const uint8_t *mask, int hard, long long n) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (mask && !mask[e]) return;
    uint4 s0 = sc0[e], s1 = sc1[e];
    uint32_t fl = (s0.x >> 16) & 0xFFu, risk = s0.x >> 24, err = (s1.w >> 16) & 0xFFu;
    if (hard) {
        fl = FLAGS_INIT;
        risk = 0;
        err = 0;
        s0.z = 0;  // step_move
    }
    s0.x = 3u | (15u << 8) | (fl << 16) | (risk << 24);
    s0.y = 0;  // step_count
    s1.x = MID_INIT;
    s1.y = ALL_NONE3;
    s1.z = ALL_NONE3;
    s1.w = NOPOS | (err << 16);
    sc0[e] = s0;
    sc1[e] = s1;
}

// One thread per (env, view cell): gen_obs of the current state.
__global__ void observe_kernel(const uint32_t *grid, const uint4 *sc0, uint8_t *obs, int V, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int VV = V * V;
    if (i >= n * VV) return;
    const long long e = i / VV;
    const int k = (int)(i - e * VV), vi = k / V, vj = k - vi * V;
    const uint32_t m = sc0[e].x;
    const int x = (int)(m & 0xFFu) - V / 2 + vi, y = (int)((m >> 8) & 0xFFu) - (V - 1) + vj;
    uint32_t code = inb(x, y) ? cell_get(grid + e * REC_WORDS, x, y) : C_WALL;
    if (vi == V / 2 && vj == V - 1) code = C_EMPTY;
    uint8_t *o = obs + i * 3;
    o[0] = (uint8_t)(0x08060201u >> (8 * code));
    o[1] = (uint8_t)(0x01040500u >> (8 * code));
    o[2] = 0;
}

// gen_obs for ANY agent_dir and either visibility mode (minigrid.py:1443-1496): get_view_exts (:1262-1293) ->
// Grid.slice (:641-660, off-grid = wall) -> rotate_left x (dir + 1) (:627-639) -> Grid.process_vis (:795-832) when
// see_through_walls is False -> agent cell := carrying (None) -> Grid.encode(vis_mask) (:749-772).
// Dead code for the registered Twoarmy envs (agent_dir == 3 and see_through_walls=True are hard-coded,
// twoarmy_v4.py:34,68), so this is the plain form: ONE THREAD PER ENV keeps the rotated V x V view as one 2-bit code
// word pair per row and the visibility flood as one bit mask per row, exactly the reference's sweep order
// (rows bottom-up; in a row left-to-right, then right-to-left; only walls block: Wall.see_behind, :422).
__global__ void observe_general_kernel(const uint32_t *grid, const uint4 *sc0, const uint8_t *dirs, int dir_all,
                                       const uint8_t *stws, int stw_all, uint8_t *obs, int V, long long n) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t m = sc0[e].x;
    const int ax = (int)(m & 0xFFu), ay = (int)((m >> 8) & 0xFFu);
    const int dir = (dirs ? dirs[e] : dir_all) & 3;
    const bool see_through = (stws ? stws[e] : stw_all) != 0;
    const int hs = V / 2;
    int topX, topY;  // get_view_exts
    if (dir == 0) { topX = ax; topY = ay - hs; }
    else if (dir == 1) { topX = ax - hs; topY = ay; }
    else if (dir == 2) { topX = ax - V + 1; topY = ay - hs; }
    else { topX = ax - hs; topY = ay - V + 1; }
    const uint32_t *rec = grid + e * REC_WORDS;
    unsigned long long code[GS];   // row r of the rotated view: 2 bits per column
    uint32_t wall[GS], vis[GS];
    for (int r = 0; r < V; r++) {
        unsigned long long cw = 0;
        uint32_t ww = 0;
        for (int c = 0; c < V; c++) {
            // one rotate_left maps old (col i, row j) to new (col j, row V-1-i); undo it dir+1 times
            int ci = c, rj = r;
            for (int k = 0; k <= dir; k++) {
                const int oi = V - 1 - rj, oj = ci;
                ci = oi;
                rj = oj;
            }
            const int x = topX + ci, y = topY + rj;
            const uint32_t cd = inb(x, y) ? cell_get(rec, x, y) : C_WALL;
            cw |= (unsigned long long)cd << (2 * c);
            ww |= (cd == C_WALL ? 1u : 0u) << c;
        }
        code[r] = cw;
        wall[r] = ww;
        vis[r] = see_through ? 0xFFFFFFFFu : 0u;
    }
    if (!see_through) {  // Grid.process_vis
        vis[V - 1] = 1u << hs;
        for (int j = V - 1; j >= 0; j--) {
            for (int i = 0; i < V - 1; i++) {
                if (!((vis[j] >> i) & 1u) || ((wall[j] >> i) & 1u)) continue;
                vis[j] |= 1u << (i + 1);
                if (j > 0) vis[j - 1] |= (1u << (i + 1)) | (1u << i);
            }
            for (int i = V - 1; i >= 1; i--) {
                if (!((vis[j] >> i) & 1u) || ((wall[j] >> i) & 1u)) continue;
                vis[j] |= 1u << (i - 1);
                if (j > 0) vis[j - 1] |= (1u << (i - 1)) | (1u << i);
            }
        }
    }
    code[V - 1] &= ~(3ull << (2 * hs));  // the agent's own cell holds `carrying` = None
    uint8_t *o = obs + e * (3ll * V * V);
    for (int i = 0; i < V; i++)       // image[i][j][c]: column first
        for (int j = 0; j < V; j++) {
            const uint32_t cd = (uint32_t)(code[j] >> (2 * i)) & 3u;
            const bool seen = (vis[j] >> i) & 1u;
            o[0] = seen ? (uint8_t)(0x08060201u >> (8 * cd)) : 0;
            o[1] = seen ? (uint8_t)(0x01040500u >> (8 * cd)) : 0;
            o[2] = 0;
            o += 3;
        }
}

// Export / import: ta_env_state records (328 B, see include/twoarmy_b200.h).
struct EnvStateRec {
    uint8_t grid[289];
    uint8_t agent_x, agent_y, flags, risk_count, error;
    uint8_t balls[10][2];
    uint8_t pad_[2];
    int32_t step_count, step_move;
    uint32_t t;
};
static_assert(sizeof(EnvStateRec) == 328, "ta_env_state layout");

__device__ __forceinline__ void unpack_ball(uint32_t p, uint8_t *o) {
    if (p == NOPOS) {
        o[0] = 0xFF;
        o[1] = 0xFF;
    } else {
        o[0] = (uint8_t)pos_x(p);
        o[1] = (uint8_t)pos_y(p);
    }
}
__device__ __forceinline__ uint32_t pack_ball(const uint8_t *o) {
    return (o[0] == 0xFF) ? NOPOS : pack_pos(o[0] & 31, o[1] & 31);
}

__global__ void export_kernel(const uint32_t *grid, const uint4 *sc0, const uint4 *sc1, EnvStateRec *out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * NCELL) return;
    const long long e = i / NCELL;
    const int c = (int)(i - e * NCELL);  // reference order c = y*17 + x
    const int y = c / GS, x = c - y * GS;
    out[e].grid[c] = (uint8_t)cell_get(grid + e * REC_WORDS, x, y);
    if (c == 0) {
        const uint4 s0 = sc0[e], s1 = sc1[e];
        EnvStateRec &r = out[e];
        r.agent_x = (uint8_t)(s0.x & 0xFFu);
        r.agent_y = (uint8_t)((s0.x >> 8) & 0xFFu);
        r.flags = (uint8_t)((s0.x >> 16) & 0xFFu);
        r.risk_count = (uint8_t)(s0.x >> 24);
        r.error = (uint8_t)((s1.w >> 16) & 0xFFu);
        for (int k = 0; k < 3; k++) unpack_ball(ball_get(s1.x, k), r.balls[k]);
        for (int k = 0; k < 3; k++) unpack_ball(ball_get(s1.y, k), r.balls[3 + k]);
        for (int k = 0; k < 3; k++) unpack_ball(ball_get(s1.z, k), r.balls[6 + k]);
        unpack_ball(ball_get(s1.w, 0), r.balls[9]);
        r.pad_[0] = r.pad_[1] = 0;
        r.step_count = (int32_t)s0.y;
        r.step_move = (int32_t)s0.z;
        r.t = s0.w;
    }
}

// One thread per (env, record word): gathers its 16 cells from the reference-order grid.
__global__ void import_kernel(uint32_t *grid, uint4 *sc0, uint4 *sc1, const EnvStateRec *in, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * REC_WORDS) return;
    const long long e = i / REC_WORDS;
    const int w = (int)(i - e * REC_WORDS);
    uint32_t word = 0;
    for (int b = 0; b < 16; b++) {
        const int c = w * 16 + b;  // column-major cell index x*17 + y
        if (c < NCELL) {
            const int x = c / GS, y = c - x * GS;
            word |= (uint32_t)(in[e].grid[y * GS + x] & 3u) << (2 * b);
        } else {
            word |= C_WALL << (2 * b);  // padding cells always hold the wall code (ta_step.cuh)
        }
    }
    grid[i] = word;
    if (w == 0) {
        const EnvStateRec &r = in[e];
        uint4 s0, s1;
        s0.x = (uint32_t)r.agent_x | ((uint32_t)r.agent_y << 8) | ((uint32_t)r.flags << 16) | ((uint32_t)r.risk_count << 24);
        s0.y = (uint32_t)r.step_count;
        s0.z = (uint32_t)r.step_move;
        s0.w = r.t;
        s1.x = pack_ball(r.balls[0]) | (pack_ball(r.balls[1]) << 10) | (pack_ball(r.balls[2]) << 20);
        s1.y = pack_ball(r.balls[3]) | (pack_ball(r.balls[4]) << 10) | (pack_ball(r.balls[5]) << 20);
        s1.z = pack_ball(r.balls[6]) | (pack_ball(r.balls[7]) << 10) | (pack_ball(r.balls[8]) << 20);
        s1.w = pack_ball(r.balls[9]) | ((uint32_t)r.error << 16);
        sc0[e] = s0;
        sc1[e] = s1;
    }
}

// get_full_render (minigrid.py:1514-1563 -> Grid.render :712-747): every cell is one of 16 cached tile
// images (cell code | agent here << 2 | highlighted << 3) of ts x ts x 3 bytes; the frame is a blit.
// One thread per output byte (a frame row is 51*ts bytes, rarely a multiple of 4).  env_ids selects the
// envs to draw (NULL = the first m).  Highlighted = inside the agent's view window, if `highlight`.
__global__ void __launch_bounds__(256) render_kernel(const uint32_t *grid, const uint4 *sc0, const uint8_t *atlas,
                                                    const long long *env_ids, long long m, int ts, int highlight, int V,
                                                    uint8_t *out) {
    const int side = GS * ts;
    const long long per_env = (long long)side * side * 3, total = m * per_env;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long k = i / per_env;
        const int r = (int)(i - k * per_env);
        const int Y = r / (side * 3), q = r - Y * side * 3, X = q / 3, c = q - 3 * X;
        const int y = Y / ts, x = X / ts, ry = Y - y * ts, rx = X - x * ts;
        const long long e = env_ids ? env_ids[k] : k;
        const uint32_t a = sc0[e].x;
        const int ax = (int)(a & 0xFFu), ay = (int)((a >> 8) & 0xFFu);
        uint32_t idx = cell_get(grid + e * REC_WORDS, x, y);
        if (x == ax && y == ay) idx |= 4u;
        if (highlight && x >= ax - V / 2 && x <= ax + V / 2 && y >= ay - (V - 1) && y <= ay) idx |= 8u;
        out[i] = atlas[((idx * ts + ry) * ts + rx) * 3 + c];
    }
}

// matrix_env (env_buffer.py:300-318): None 0.9, wall -0.9, ball -0.5, goal 0.9, then the
// agent's cell 0.3; index y*17 + x.  The float64 LUT is cast to float32 exactly as
// train_ppo.py:122 does on store.  Compact code: 0 -> 0.9, 1 -> -0.9, 2 -> -0.5, 4 -> 0.3.
__device__ __forceinline__ uint32_t matrix_code(uint32_t cell, bool agent) {
    return agent ? 4u : (cell == C_GOAL ? 0u : cell);
}
__device__ __forceinline__ float matrix_value(uint32_t mc) {
    return mc == 0u ? 0.9f : (mc == 1u ? -0.9f : (mc == 2u ? -0.5f : 0.3f));
}

// One CTA of 320 threads handles 4 envs: the packed column-major records are staged through
// shared memory so that both the HBM read and the row-major write are contiguous.
constexpr int SM_ENVS = 4;
__global__ void __launch_bounds__(320) state_matrix_kernel(const uint32_t *grid, const uint4 *sc0, uint8_t *codes,
                                                          float *matrix, float *place, long long n) {
    __shared__ uint32_t sg[SM_ENVS * REC_WORDS];
    __shared__ uint32_t sa[SM_ENVS];
    const long long e0 = (long long)blockIdx.x * SM_ENVS;
    const int cnt = (int)((n - e0) < SM_ENVS ? (n - e0) : SM_ENVS);
    for (int i = threadIdx.x; i < cnt * REC_WORDS; i += blockDim.x) sg[i] = grid[e0 * REC_WORDS + i];
    if (threadIdx.x < cnt) {
        const uint32_t m = sc0[e0 + threadIdx.x].x;
        sa[threadIdx.x] = m & 0xFFFFu;
        if (place) {
            place[(e0 + threadIdx.x) * 2 + 0] = (float)((m >> 8) & 0xFFu);  // (row, col) = (y, x)
            place[(e0 + threadIdx.x) * 2 + 1] = (float)(m & 0xFFu);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * NCELL; i += blockDim.x) {
        const int e = i / NCELL, c = i - e * NCELL, y = c / GS, x = c - y * GS;
        const uint32_t a = sa[e];
        const uint32_t mc = matrix_code(cell_get(sg + e * REC_WORDS, x, y), x == (int)(a & 0xFFu) && y == (int)(a >> 8));
        if (codes) codes[e0 * NCELL + i] = (uint8_t)mc;
        if (matrix) matrix[e0 * NCELL + i] = matrix_value(mc);
    }
}

// Frame-stack roll fused with matrix_env: s [n][5][289], p [n][5][2].  ST = float (the LUT
// applied) or uint8_t (the compact codes; the LUT is applied by the network's loader).
template <typename ST>
__device__ __forceinline__ ST stack_value(uint32_t mc);
template <>
__device__ __forceinline__ float stack_value<float>(uint32_t mc) { return matrix_value(mc); }
template <>
__device__ __forceinline__ uint8_t stack_value<uint8_t>(uint32_t mc) { return (uint8_t)mc; }

template <typename ST>
__global__ void __launch_bounds__(320) stack_roll_kernel(const uint32_t *grid, const uint4 *sc0, ST *s, float *p,
                                                        const uint8_t *init_mask, int init, long long n) {
    __shared__ uint32_t sg[SM_ENVS * REC_WORDS];
    __shared__ uint32_t sa[SM_ENVS];
    __shared__ uint8_t sinit[SM_ENVS];
    const long long e0 = (long long)blockIdx.x * SM_ENVS;
    const int cnt = (int)((n - e0) < SM_ENVS ? (n - e0) : SM_ENVS);
    for (int i = threadIdx.x; i < cnt * REC_WORDS; i += blockDim.x) sg[i] = grid[e0 * REC_WORDS + i];
    if (threadIdx.x < cnt) {
        sa[threadIdx.x] = sc0[e0 + threadIdx.x].x & 0xFFFFu;
        // init call: masked envs are tiled, the others are left untouched (2 = skip)
        sinit[threadIdx.x] = (uint8_t)(init ? ((!init_mask || init_mask[e0 + threadIdx.x]) ? 1 : 2) : 0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * NCELL; i += blockDim.x) {
        const int e = i / NCELL, c = i - e * NCELL, y = c / GS, x = c - y * GS;
        const uint32_t a = sa[e];
        const ST v = stack_value<ST>(matrix_code(cell_get(sg + e * REC_WORDS, x, y), x == (int)(a & 0xFFu) && y == (int)(a >> 8)));
        ST *row = s + (e0 + e) * 5 * NCELL + c;
        if (sinit[e] == 2) continue;
        if (sinit[e]) {
#pragma unroll
            for (int f = 0; f < 5; f++) row[f * NCELL] = v;
        } else {
            const ST f1 = row[1 * NCELL], f2 = row[2 * NCELL], f3 = row[3 * NCELL], f4 = row[4 * NCELL];
            row[0] = f1; row[1 * NCELL] = f2; row[2 * NCELL] = f3; row[3 * NCELL] = f4; row[4 * NCELL] = v;
        }
    }
    if (p && threadIdx.x < cnt * 2) {
        const int e = threadIdx.x >> 1, comp = threadIdx.x & 1;
        const uint32_t a = sa[e];
        const float v = comp == 0 ? (float)(a >> 8) : (float)(a & 0xFFu);
        float *row = p + (e0 + e) * 10 + comp;
        if (sinit[e] == 2) {
        } else if (sinit[e]) {
#pragma unroll
            for (int f = 0; f < 5; f++) row[f * 2] = v;
        } else {
            const float f1 = row[2], f2 = row[4], f3 = row[6], f4 = row[8];
            row[0] = f1; row[2] = f2; row[4] = f3; row[6] = f4; row[8] = v;
        }
    }
}

}  // namespace ta
