// ta_step.cuh -- the fused Twoarmy step + gen_obs kernel (sm_100a).
//
// Reference functions restated here (paths relative to the reference root):
//   Twoarmy_v4.step            gym_minigrid/envs/twoarmy_v4.py:82-322
//   Twoarmy_v6.step            gym_minigrid/envs/twoarmy_v6.py:83-325
//   MiniGridEnv.step           gym_minigrid/minigrid.py:1333-1441
//   gen_obs / gen_obs_grid     gym_minigrid/minigrid.py:1443-1496 (+ get_view_exts :1262-1293,
//                              Grid.slice :641-660, Grid.encode :749-772)
//   MiniGridEnv.reset          gym_minigrid/minigrid.py:947-980 (autoreset tail)
//
// Mapping: one warp owns a tile of 32 envs.  The tile's grids (9248 B) arrive in shared
// memory with one TMA bulk copy; lane e then runs env e's transition on its staged grid
// (phase A: balls, patrols, agent move), the whole warp builds the 32 observations
// cooperatively into a 16-env staging buffer that leaves with one TMA bulk store per half
// tile (phase B), and lane e finishes the step (phase C: wall blocks, patrol spawn, reward,
// episode end) before the grids go back with one more bulk store.  The obs is built between
// A and C because the reference builds it inside MiniGridEnv.step, i.e. before the wall
// blocks / patrol balls of the same step appear (ordering fact a).
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int STEP_WARPS = 9;
constexpr int G_BYTES = TILE * NCELL;  // 9248
constexpr int G_PAD_BEFORE = 144;      // the V=17 window read may start up to 134 B before the tile
constexpr int G_PAD_AFTER = 128;       // ... and end up to 121 B after it
constexpr int TAB_TOP = 0;             // [17 ay][17 j0][16]  wall mask for rows above the grid
constexpr int TAB_COL = 17 * 17 * 16;  // [4][17 j0][16]      wall mask for columns off the grid
constexpr int TAB_TEMPLATE = TAB_COL + 4 * 17 * 16;  // 289 B initial grid, column-major
constexpr int TAB_BYTES = TAB_TEMPLATE + 304;        // 6016
constexpr int STAGE_MAX = HALF * 3 * GS * GS;        // 13872
constexpr int WARP_SMEM = 16 + 128 + G_PAD_BEFORE + G_BYTES + G_PAD_AFTER + STAGE_MAX;  // 23536
constexpr int STEP_SMEM = TAB_BYTES + 16 + STEP_WARPS * WARP_SMEM;

struct StepArgs {
    uint8_t *grid;
    uint4 *sc0;
    uint4 *sc1;
    const uint8_t *tables;
    const void *actions;
    const uint8_t *draws;
    uint8_t *obs;
    float *reward;
    uint8_t *term;
    uint8_t *trunc;
    uint8_t *consumed;
    long long n;
    int ntiles;
    int version;
    int flags;
    int action_dtype;
    uint32_t seed_lo, seed_hi;
    unsigned long long env_id0;
};

struct DrawSrc {
    bool replay;
    uint32_t rec_lo, rec_hi;  // replay record, slot s = byte s
    uint32_t w[4];
    bool have;
    uint32_t k0, k1, c0, c1, t;
    uint32_t consumed;
};

__device__ __forceinline__ int draw(DrawSrc &d, int slot, int lo) {
    d.consumed |= 1u << slot;
    if (d.replay) return (int)(((slot < 4 ? d.rec_lo : d.rec_hi) >> (8 * (slot & 3))) & 0xFFu);
    if (!d.have) {
        philox4x32_10(d.c0, d.c1, d.t, 0u, d.k0, d.k1, d.w);
        d.have = true;
    }
    if (slot == 0) return lo + (int)__umulhi(d.w[0], 10u);
    if (slot <= 3) return lo + (int)((d.w[1] >> (2 * (slot - 1))) & 3u);
    return lo + (int)((d.w[2] >> (slot - 5)) & 1u);
}

__device__ __forceinline__ void put_cell(uint8_t *G, int x, int y, uint32_t code) {
    if (inb(x, y)) G[x * GS + y] = (uint8_t)code;
}

// clear every old cell, then put each ball at old+(dx,dy); a put that leaves the grid is the
// swallowed AssertionError of twoarmy_v4.py:102-111 / :126-129: the ball keeps its cur_pos
template <int NB>
__device__ __forceinline__ void move_group(uint8_t *G, uint32_t (&p)[NB], int dx, int dy, bool fixed_y8) {
    int ox[NB], oy[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) {
        ox[k] = pos_x(p[k]);
        oy[k] = pos_y(p[k]);
        G[ox[k] * GS + oy[k]] = (uint8_t)C_EMPTY;
    }
#pragma unroll
    for (int k = 0; k < NB; k++) {
        int nx = ox[k] + dx, ny = fixed_y8 ? 8 : oy[k] + dy;
        if (inb(nx, ny)) {
            G[nx * GS + ny] = (uint8_t)C_BALL;
            p[k] = pack_pos(nx, ny);
        }
    }
}

// ---- observation builders -----------------------------------------------------------------
// Output layout per env: image[i][j][c], i = view column, j = view row, c in (type,color,state)
// (Grid.encode, minigrid.py:749-772).  With agent_dir == 3 and see_through_walls the four
// rotate_left calls are the identity: view cell (i,j) is grid cell (ax - V/2 + i, ay - V+1 + j),
// off-grid cells are walls, and the agent's own cell (V/2, V-1) is empty.
constexpr uint32_t TYPE_LUT = 0x08060201u;   // empty 1, wall 2, ball 6, goal 8
constexpr uint32_t COLOR_LUT = 0x01040500u;  // -, grey 5, yellow 4, green 1

template <int V>
__device__ __forceinline__ void fill_stage_generic(uint8_t *stage, const uint8_t *g, const uint32_t *meta, int half,
                                                   int lane) {
    constexpr int VV = V * V;
    for (int c = lane; c < HALF * VV; c += 32) {
        int e = c / VV, k = c - e * VV, i = k / V, j = k - i * V;
        uint32_t m = meta[half * HALF + e];
        int x = (int)(m & 0xFFu) - V / 2 + i, y = (int)((m >> 8) & 0xFFu) - (V - 1) + j;
        uint32_t code = inb(x, y) ? g[(half * HALF + e) * NCELL + x * GS + y] : C_WALL;
        if (i == V / 2 && j == V - 1) code = C_EMPTY;
        stage[3 * c + 0] = (uint8_t)(TYPE_LUT >> (8 * code));
        stage[3 * c + 1] = (uint8_t)(COLOR_LUT >> (8 * code));
        stage[3 * c + 2] = 0;
    }
}

// 16 consecutive view cells k0..k0+15 of env e (V = 17) as 16 code bytes.  Because the view
// is as wide as the grid and the staged grid is column-major, they are the 16 bytes at
// g[e][k0 + (ax-8)*17 + (ay-16)], except (a) rows above the grid and (b) columns off the grid,
// which become walls through two 16-byte masks looked up by (ay, k0 mod 17) and by
// (column-off-grid bits, k0 mod 17), and (c) the agent's own cell k = 152.
__device__ __forceinline__ uint4 codes16_v17(const uint8_t *gpad, const uint8_t *tab, uint32_t m, int e, int k0) {
    const int ax = (int)(m & 0xFFu), ay = (int)((m >> 8) & 0xFFu);
    // i0 = floor(k0 / 17), valid for k0 in [-16, 288]
    const int kk = k0 + 17;
    const int i0 = kk / 17 - 1, j0 = kk - (i0 + 1) * 17;
    const int addr = G_PAD_BEFORE + e * NCELL + k0 + (ax - 8) * GS + (ay - 16);
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(gpad) + (addr >> 2);
    const uint32_t sh = (uint32_t)(addr & 3) * 8u;
    uint32_t v0 = wp[0], v1 = wp[1], v2 = wp[2], v3 = wp[3], v4 = wp[4];
    uint4 c;
    c.x = __funnelshift_r(v0, v1, sh);
    c.y = __funnelshift_r(v1, v2, sh);
    c.z = __funnelshift_r(v2, v3, sh);
    c.w = __funnelshift_r(v3, v4, sh);
    const int x0 = ax - 8 + i0;
    const int colbits = ((unsigned)x0 > 16u ? 1 : 0) | ((unsigned)(x0 + 1) > 16u ? 2 : 0);
    const uint4 mt = *reinterpret_cast<const uint4 *>(tab + TAB_TOP + (ay * 17 + j0) * 16);
    const uint4 mc = *reinterpret_cast<const uint4 *>(tab + TAB_COL + (colbits * 17 + j0) * 16);
    uint32_t mx = mt.x | mc.x, my = mt.y | mc.y, mz = mt.z | mc.z, mw = mt.w | mc.w;
    // (c & ~m) | (WALL & m)
    c.x = (c.x & ~mx) | (0x01010101u & mx);
    c.y = (c.y & ~my) | (0x01010101u & my);
    c.z = (c.z & ~mz) | (0x01010101u & mz);
    c.w = (c.w & ~mw) | (0x01010101u & mw);
    // agent cell: view (8,16) -> k = 152
    const int d = 152 - k0;
    if ((unsigned)d < 16u) {
        const uint32_t clr = ~(0xFFu << (8 * (d & 3)));
        if ((d >> 2) == 0) c.x &= clr;
        else if ((d >> 2) == 1) c.y &= clr;
        else if ((d >> 2) == 2) c.z &= clr;
        else c.w &= clr;
    }
    return c;
}

// 4 code bytes -> 12 obs bytes (type,color,0 per cell) as 3 words
__device__ __forceinline__ void expand4(uint32_t w, uint32_t &o0, uint32_t &o1, uint32_t &o2) {
    const uint32_t sel = __byte_perm(w | (w >> 4), 0u, 0x4420u);  // c0 | c1<<4 | c2<<8 | c3<<12
    const uint32_t t4 = __byte_perm(TYPE_LUT, 0u, sel);
    const uint32_t c4 = __byte_perm(COLOR_LUT, 0u, sel);
    o0 = __byte_perm(t4, c4, 0x1040u) & 0xFF00FFFFu;  // t0 c0 0 t1
    o1 = __byte_perm(t4, c4, 0x6205u) & 0xFFFF00FFu;  // c1 0 t2 c2
    o2 = __byte_perm(t4, c4, 0x0730u) & 0x00FFFF00u;  // 0 t3 c3 0
}

__device__ __forceinline__ void expand16_store(uint8_t *stage, int run, uint4 c) {
    uint4 a, b, d;
    expand4(c.x, a.x, a.y, a.z);
    expand4(c.y, a.w, b.x, b.y);
    expand4(c.z, b.z, b.w, d.x);
    expand4(c.w, d.y, d.z, d.w);
    uint4 *dst = reinterpret_cast<uint4 *>(stage + run * 48);
    dst[0] = a;
    dst[1] = b;
    dst[2] = d;
}

// One half tile (16 envs x 289 cells = 289 runs of 16 cells -> 13872 B) for V = 17.
__device__ __forceinline__ void fill_stage_v17(uint8_t *stage, const uint8_t *gpad, const uint8_t *tab,
                                               const uint32_t *meta, int half, int lane) {
    constexpr int RUNS = HALF * NCELL / 16;  // 289
    for (int run = lane; run < RUNS; run += 32) {
        const int q0 = run * 16;
        const int e = q0 / NCELL, k0 = q0 - e * NCELL;
        const int ge = half * HALF + e;
        expand16_store(stage, run, codes16_v17(gpad, tab, meta[ge], ge, k0));
    }
    __syncwarp();
    // runs that straddle two envs: run 18c holds the last c cells of env c-1 and the first
    // 16-c cells of env c (c = 1..15); redo them with both halves merged
    if (lane >= 1 && lane < HALF) {
        const int c = lane, run = 18 * c;
        const int ea = half * HALF + c - 1, eb = ea + 1;
        uint4 ca = codes16_v17(gpad, tab, meta[ea], ea, NCELL - c);
        uint4 cb = codes16_v17(gpad, tab, meta[eb], eb, -c);
        // bytes [0,c) from ca, bytes [c,16) from cb
        uint32_t mk[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            int nb = c - 4 * w;  // how many low bytes of this word come from ca
            mk[w] = nb <= 0 ? 0u : (nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u));
        }
        uint4 cm;
        cm.x = (ca.x & mk[0]) | (cb.x & ~mk[0]);
        cm.y = (ca.y & mk[1]) | (cb.y & ~mk[1]);
        cm.z = (ca.z & mk[2]) | (cb.z & ~mk[2]);
        cm.w = (ca.w & mk[3]) | (cb.w & ~mk[3]);
        expand16_store(stage, run, cm);
    }
}

__device__ __forceinline__ int load_action(const void *actions, int dtype, long long i) {
    if (dtype == 0) return reinterpret_cast<const int *>(actions)[i];
    if (dtype == 1) return (int)reinterpret_cast<const uint8_t *>(actions)[i];
    long long a = reinterpret_cast<const long long *>(actions)[i];
    return a > 1000 ? 1000 : (a < -1000 ? -1000 : (int)a);
}

template <int V, bool FAST>
__global__ void __launch_bounds__(STEP_WARPS * 32, 1) step_obs_kernel(const StepArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *tab = smem;
    uint64_t *tab_bar = reinterpret_cast<uint64_t *>(smem + TAB_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *wbase = smem + TAB_BYTES + 16 + warp * WARP_SMEM;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase);
    uint32_t *meta = reinterpret_cast<uint32_t *>(wbase + 16);
    uint8_t *gpad = wbase + 16 + 128;
    uint8_t *g = gpad + G_PAD_BEFORE;
    uint8_t *stage = gpad + G_PAD_BEFORE + G_BYTES + G_PAD_AFTER;
    constexpr int OBS_ENV = 3 * V * V;
    constexpr int STAGE_BYTES = HALF * OBS_ENV;

    if (threadIdx.x == 0) mbar_init(tab_bar, 1);
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(tab_bar, TAB_BYTES);
        bulk_g2s(tab, a.tables, TAB_BYTES, tab_bar);
    }
    uint32_t phase = 0;
    bool tab_ready = false;
    const bool v4 = a.version == 4;

    for (int it = warp;; it += STEP_WARPS) {
        const long long tile = (long long)blockIdx.x + (long long)gridDim.x * it;
        if (tile >= a.ntiles) break;
        // previous tile's bulk stores must have finished reading g / stage
        if (lane == 0) {
            bulk_wait_read0();
            mbar_expect_tx(bar, G_BYTES);
            bulk_g2s(g, a.grid + tile * G_BYTES, G_BYTES, bar);
        }
        const long long env = tile * TILE + lane;
        const bool live = env < a.n;
        uint4 s0 = a.sc0[env], s1 = a.sc1[env];
        int act = live ? load_action(a.actions, a.action_dtype, env) : 6;
        DrawSrc d;
        d.replay = a.draws != nullptr;
        d.rec_lo = d.rec_hi = 0xFFFFFFFFu;
        if (d.replay && live) {
            uint2 r = reinterpret_cast<const uint2 *>(a.draws)[env];
            d.rec_lo = r.x;
            d.rec_hi = r.y;
        }
        d.have = false;
        d.consumed = 0;
        d.k0 = a.seed_lo;
        d.k1 = a.seed_hi;
        {
            unsigned long long gid = a.env_id0 + (unsigned long long)env;
            d.c0 = (uint32_t)gid;
            d.c1 = (uint32_t)(gid >> 32);
        }

        int ax = (int)(s0.x & 0xFFu), ay = (int)((s0.x >> 8) & 0xFFu);
        uint32_t fl = (s0.x >> 16) & 0xFFu;
        int risk = (int)(s0.x >> 24);
        int step_count = (int)s0.y, step_move = (int)s0.z;
        uint32_t tcount = s0.w;
        uint32_t mid[3] = {ball_get(s1.x, 0), ball_get(s1.x, 1), ball_get(s1.x, 2)};
        uint32_t o1[3] = {ball_get(s1.y, 0), ball_get(s1.y, 1), ball_get(s1.y, 2)};
        uint32_t o2[4] = {ball_get(s1.z, 0), ball_get(s1.z, 1), ball_get(s1.z, 2), ball_get(s1.w, 0)};
        uint32_t err = (s1.w >> 16) & 0xFFu;

        // ---- pre-checks: situations where the reference raises (documented divergence:
        // the env is left untouched and an error bit is set) ---------------------------------
        if (act >= 7) act = 0;  // twoarmy_v4.py:84-85
        int adx = 0, ady = 0;
        if (act == 0) adx = -1;
        else if (act == 1) adx = 1;
        else if (act == 2) ady = -1;
        else if (act == 3) ady = 1;
        bool skip = false;
        if (!(act == 0 || act == 1 || act == 2 || act == 3 || act == 6)) {
            err |= ERR_BAD_ACTION;
            skip = true;
        } else if ((fl & F_PATROL) && (o1[0] == NOPOS || o2[0] == NOPOS)) {
            err |= ERR_NONE_POS;
            skip = true;
        } else if (!inb(ax + adx, ay + ady)) {
            err |= ERR_OOB_MOVE;
            skip = true;
        }

        mbar_wait(bar, phase);
        phase ^= 1u;
        uint8_t *G = g + lane * NCELL;
        bool term = false, trunc = false;
        d.t = tcount;

        // ---- phase A: everything up to and including the agent move -------------------------
        if (!skip) {
            tcount += 1u;
            step_move += 1;
            const int m6 = step_move % 6, m4 = step_move & 3;
            {  // mid-row balls, twoarmy_v4.py:95-111
                const int dx = (m6 == 1 || m6 == 0) ? 1 : ((m6 == 2 || m6 == 3) ? -1 : 0);
                move_group<3>(G, mid, dx, 0, true);
            }
            if (v4) {
                if (fl & F_UPD_L) {  // twoarmy_v4.py:115-144
                    fl &= ~F_UPD_H;
                    if (m4 == 2 || m6 == 3 || m6 == 0 || draw(d, 0, 0) == 6) {
                        if (fl & F_PATROL) {
                            if (fl & F_UP1) {
                                move_group<3>(G, o1, 0, -1, false);
                                if (pos_y(o1[0]) == 3) fl &= ~F_UP1;
                            } else {
                                move_group<3>(G, o1, 0, 1, false);
                                if (pos_y(o1[2]) == 7) fl |= F_UP1;
                            }
                        }
                    }
                }
                if (fl & F_UPD_H) {  // twoarmy_v4.py:147-176
                    fl &= ~F_UPD_L;
                    if (m6 == 0 || m6 == 2 || m6 == 3 || m6 == 5 || m6 == 4 || draw(d, 0, 0) == 6) {
                        if (fl & F_PATROL) {
                            if (fl & F_RIGHT2) {
                                move_group<4>(G, o2, 1, 0, false);
                                if (pos_x(o2[3]) == 11) fl &= ~F_RIGHT2;
                            } else {
                                move_group<4>(G, o2, -1, 0, false);
                                if (pos_x(o2[0]) == 5) fl |= F_RIGHT2;
                            }
                        }
                    }
                }
            }
            // MiniGridEnv.step, minigrid.py:1333-1441
            step_count += 1;
            const int tx = ax + adx, ty = ay + ady;
            const uint32_t c = G[tx * GS + ty];
            if (c == C_EMPTY || c == C_GOAL) {
                ax = tx;
                ay = ty;
            }
            if (c == C_GOAL) term = true;
            if (step_count >= 50) trunc = true;
        }
        meta[lane] = (uint32_t)ax | ((uint32_t)ay << 8);
        __syncwarp();

        // ---- phase B: observations of all 32 envs, 16 at a time ----------------------------
        if (!tab_ready) {  // mask tables (V=17 path) and the reset template
            mbar_wait(tab_bar, 0);
            tab_ready = true;
        }
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            if (half == 1) {
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
            }
            if (FAST) fill_stage_v17(stage, gpad, tab, meta, half, lane);
            else fill_stage_generic<V>(stage, g, meta, half, lane);
            fence_proxy_async();
            __syncwarp();
            const long long e0 = tile * TILE + half * HALF;
            long long nvalid = a.n - e0;
            nvalid = nvalid < 0 ? 0 : (nvalid > HALF ? HALF : nvalid);
            uint8_t *dst = a.obs + e0 * OBS_ENV;
            if (nvalid == HALF) {
                if (lane == 0) {
                    bulk_s2g(dst, stage, STAGE_BYTES);
                    bulk_commit();
                }
            } else {  // ragged last tile: plain stores
                for (int i = lane; i < (int)nvalid * OBS_ENV; i += 32) dst[i] = stage[i];
                __syncwarp();
            }
        }

        // ---- phase C: rest of Twoarmy.step ---------------------------------------------------
        int reward = R_STEP;  // twoarmy_v4.py:180
        bool need_reset = false;
        if (!skip) {
            if (!(fl & F_PONE) && (ax > 3 || ay < 14)) {  // twoarmy_v4.py:181-195, twoarmy_v6.py:182-198
                int i = v4 ? draw(d, 1, 9) : 11;
                put_cell(G, 4, i, C_WALL); put_cell(G, 5, i, C_WALL);
                put_cell(G, 4, i + 1, C_WALL); put_cell(G, 5, i + 1, C_WALL);
                i = v4 ? draw(d, 2, 6) : 8;
                put_cell(G, i, 11, C_WALL); put_cell(G, i, 12, C_WALL);
                put_cell(G, i + 1, 11, C_WALL); put_cell(G, i + 1, 12, C_WALL);
                fl |= F_PONE;
            }
            if (v4 && !(fl & F_PATROL) && ay <= 8) {  // twoarmy_v4.py:212-225
                const int i = draw(d, 3, 6);
                const int bx[4] = {i, i + 1, i, i + 1}, by[4] = {4, 4, 5, 5};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (inb(bx[k], by[k])) {
                        G[bx[k] * GS + by[k]] = (uint8_t)C_BALL;
                        o2[k] = pack_pos(bx[k], by[k]);
                    }
                d.consumed |= 1u << 4;  // :221 choice(range(4,5)) == 4, no generator words
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    G[12 * GS + 4 + k] = (uint8_t)C_BALL;
                    o1[k] = pack_pos(12, 4 + k);
                }
                fl |= F_PATROL;
            }
            const uint32_t ap = pack_pos(ax, ay);
            // twoarmy_v4.py:228-240
            if (ap == mid[0] || ap == mid[1] || ap == mid[2]) {
                reward = R_HIT;
                trunc = true;
            }
            if (ay == pos_y(mid[0]) + 1 && (ax == pos_x(mid[0]) || ax == pos_x(mid[1]) || ax == pos_x(mid[2])))
                reward = R_RISK;
            if (fl & F_PATROL) {  // twoarmy_v4.py:242-280
                if (ay == pos_y(o2[2]) + 1 && (ax == pos_x(o2[2]) || ax == pos_x(o2[3]))) reward = R_RISK;
                if (ax == pos_x(o2[0]) - 1 && (ay == pos_y(o2[0]) || ay == pos_y(o2[2]))) reward = R_RISK;
                if (ax == pos_x(o2[1]) + 1 && (ay == pos_y(o2[1]) || ay == pos_y(o2[3]))) reward = R_RISK;
                if (ax == pos_x(o1[0]) - 1 && (ay == pos_y(o1[0]) || ay == pos_y(o1[1]) || ay == pos_y(o1[2])))
                    reward = R_RISK;
                if (ap == o1[0] || ap == o1[1] || ap == o1[2] || ap == o2[0] || ap == o2[1] || ap == o2[2] ||
                    ap == o2[3]) {
                    reward = R_HIT;
                    trunc = true;
                }
            }
            if ((fl & F_FIRST) && ay == 7) {  // twoarmy_v4.py:282-285
                reward = R_ROOM2;
                fl &= ~F_FIRST;
            }
            if (reward == R_RISK) {  // twoarmy_v4.py:287-291
                risk += 1;
                if (risk > 5) trunc = true;
            }
            if (term || trunc) {  // twoarmy_v4.py:293-315
                if (term) reward = R_GOAL;
                step_move = 0;
                fl &= ~(F_PONE | F_PATROL);
                fl |= F_FIRST;
                risk = 0;
                if (draw(d, 5, 0) == 1) fl = (fl & ~F_UP1) | F_RIGHT2;
                else fl = (fl | F_UP1) & ~F_RIGHT2;
                if (draw(d, 6, 0) == 1) fl = (fl & ~F_UPD_H) | F_UPD_L;
                else fl = (fl | F_UPD_H) & ~F_UPD_L;
                need_reset = (a.flags & 1) != 0;
            }
        }
        if (live) {
            a.reward[env] = reward_value(reward);
            a.term[env] = term ? 1 : 0;
            a.trunc[env] = trunc ? 1 : 0;
            if (a.consumed) a.consumed[env] = (uint8_t)d.consumed;
        }
        // autoreset: MiniGridEnv.reset (minigrid.py:947-980) -- grid, balls, agent, step_count
        if (need_reset) {
            ax = 3; ay = 15; step_count = 0;
            mid[0] = ball_get(MID_INIT, 0); mid[1] = ball_get(MID_INIT, 1); mid[2] = ball_get(MID_INIT, 2);
            o1[0] = o1[1] = o1[2] = NOPOS;
            o2[0] = o2[1] = o2[2] = o2[3] = NOPOS;
        }
        __syncwarp();
        uint32_t rmask = __ballot_sync(0xFFFFFFFFu, need_reset);
        while (rmask) {
            const int e = __ffs(rmask) - 1;
            rmask &= rmask - 1;
            for (int i = lane; i < NCELL; i += 32) g[e * NCELL + i] = tab[TAB_TEMPLATE + i];
        }
        s0.x = (uint32_t)ax | ((uint32_t)ay << 8) | (fl << 16) | ((uint32_t)risk << 24);
        s0.y = (uint32_t)step_count;
        s0.z = (uint32_t)step_move;
        s0.w = tcount;
        s1.x = mid[0] | (mid[1] << 10) | (mid[2] << 20);
        s1.y = o1[0] | (o1[1] << 10) | (o1[2] << 20);
        s1.z = o2[0] | (o2[1] << 10) | (o2[2] << 20);
        s1.w = o2[3] | (err << 16);
        a.sc0[env] = s0;
        a.sc1[env] = s1;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(a.grid + tile * G_BYTES, g, G_BYTES);
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait_all0();
}

}  // namespace ta
