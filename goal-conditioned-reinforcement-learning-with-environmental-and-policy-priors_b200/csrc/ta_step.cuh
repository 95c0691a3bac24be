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
// Mapping: a CTA of 2 warps owns a tile of 32 envs; 14 CTAs are resident per SM, so the 2048
// tiles of a 65536-env batch are all on chip at once.  The tile's packed grids (2560 B) arrive
// in shared memory with one TMA bulk copy and the 32 envs' scalar state lives in warp 0's
// registers for the whole launch.  For each of the launch's T steps (T = 1 for ta_step, the
// rollout length for ta_rollout) warp 0 runs the per-env transition, lane e = env e (phase A:
// balls, patrols, agent move); both warps then build the 32 observations cooperatively and
// stream them to HBM with 16-byte stores (phase B); warp 0 finishes the step (phase C: wall
// blocks, patrol spawn, reward, episode end, autoreset).  After the last step the state goes
// back (grids with one TMA bulk store).  The obs is built between A and C because the
// reference builds it inside MiniGridEnv.step, i.e. before the wall blocks / patrol balls of
// the same step appear (ordering fact a).
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int STEP_THREADS = 64;
constexpr int STEP_WARPS = STEP_THREADS / 32;
constexpr int STEP_CTAS_PER_SM = 14;  // 14 x 148 = 2072 resident tiles >= the 2048 tiles of 65536 envs
constexpr int G_BYTES = TILE * REC_BYTES;  // 2560
constexpr int G_PAD_BEFORE = 48;           // bytes: the V=17 window may start 134 cells early
constexpr int G_PAD_AFTER = 48;            // ... and end 134 cells late
constexpr int STAGE_BYTES = 6144;          // staging for the per-cell (generic view size) obs path
// shared memory map (bytes)
constexpr int SM_TAB = 0;
constexpr int SM_BARS = TAB_SMEM_BYTES;                       // 2 mbarriers
constexpr int SM_META = SM_BARS + 16;                    // 32 x u32: agent x | y<<8
constexpr int SM_GPAD = SM_META + 128;                   // padded grid tile
constexpr int SM_STAGE = SM_GPAD + G_PAD_BEFORE + G_BYTES + G_PAD_AFTER;
constexpr int STEP_SMEM = SM_STAGE + STAGE_BYTES;        // 14560
static_assert(SM_GPAD % 16 == 0 && SM_STAGE % 16 == 0 && (SM_GPAD + G_PAD_BEFORE) % 16 == 0, "alignment");

struct StepArgs {
    uint32_t *grid;
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
    int T;  // env steps per launch; outputs / actions / draws are [T][n]
    int version;
    int flags;
    int action_dtype;
    uint32_t seed_lo, seed_hi;
    unsigned long long env_id0;
};

// The step's draws, one byte per call-site slot: either the replay
// record the caller supplied (verification mode) or the values of ONE Philox block computed up
// front for every lane (production mode; cheaper than branching into it at each site).
struct DrawSrc {
    uint32_t rec_lo, rec_hi;  // slot s = byte s
    uint32_t consumed;
};

__device__ __forceinline__ void philox_record(DrawSrc &d, uint32_t c0, uint32_t c1, uint32_t t, uint32_t k0, uint32_t k1) {
    uint32_t w[4];
    philox4x32_10(c0, c1, t, 0u, k0, k1, w);
    d.rec_lo = __umulhi(w[0], 10u) | ((9u + (w[1] & 3u)) << 8) | ((6u + ((w[1] >> 2) & 3u)) << 16) |
               ((6u + ((w[1] >> 4) & 3u)) << 24);
    d.rec_hi = (4u) | ((w[2] & 1u) << 8) | (((w[2] >> 1) & 1u) << 16);
}

__device__ __forceinline__ int draw(DrawSrc &d, int slot) {
    d.consumed |= 1u << slot;
    return (int)(((slot < 4 ? d.rec_lo : d.rec_hi) >> (8 * (slot & 3))) & 0xFFu);
}

__device__ __forceinline__ void put_cell(uint32_t *G, int x, int y, uint32_t code) {
    if (inb(x, y)) cell_set(G, x, y, code);
}

// clear every old cell, then put each ball at old+(dx,dy); a put that leaves the grid is the
// swallowed AssertionError of twoarmy_v4.py:102-111 / :126-129: the ball keeps its cur_pos
template <int NB>
__device__ __forceinline__ void move_group(uint32_t *G, uint32_t (&p)[NB], int dx, int dy, bool fixed_y8) {
    int ox[NB], oy[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) {
        ox[k] = pos_x(p[k]);
        oy[k] = pos_y(p[k]);
        cell_set(G, ox[k], oy[k], C_EMPTY);
    }
#pragma unroll
    for (int k = 0; k < NB; k++) {
        int nx = ox[k] + dx, ny = fixed_y8 ? 8 : oy[k] + dy;
        if (inb(nx, ny)) {
            cell_set(G, nx, ny, C_BALL);
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

// Per-env record the scalar warp leaves for the observation builders:
//   bits 0..15  cellbase = pad + e*320 + (ax-8)*17 + (ay-16): where view cell k=0 of env e sits
//               in the padded shared-memory tile, in cells
//   bits 16..20 agent y      bits 24..28 agent x
__device__ __forceinline__ uint32_t make_meta(int e, int ax, int ay) {
    return (uint32_t)(G_PAD_BEFORE * 4 + e * REC_CELLS + (ax - 8) * GS + (ay - 16)) | ((uint32_t)ay << 16) |
           ((uint32_t)ax << 24);
}

// 16 consecutive view cells k0..k0+15 of one env (V = 17) as one word of 2-bit codes.  Because
// the view is as wide as the grid and the record is column-major, they are the 32 bits at
// cell offset cellbase + k0 -- one funnel shift -- except (a) rows above the grid and (b)
// columns off the grid, which become walls through two masks looked up by (ay, j0) and by
// (column-off-grid bits, j0), and (c) the agent's own cell k = 152.  i0p1 = floor(k0/17)+1,
// j0 = k0 mod 17; valid for k0 in [-16, 288].
__device__ __forceinline__ uint32_t codes16_v17(const uint32_t *gpadw, const uint8_t *tab, uint32_t m, int k0, int i0p1,
                                                int j0) {
    const int cell = (int)(m & 0xFFFFu) + k0;
    const int ay = (int)((m >> 16) & 31u), ax = (int)(m >> 24);
    const int wi = cell >> 4;
    const uint32_t win = __funnelshift_r(gpadw[wi], gpadw[wi + 1], (uint32_t)(cell & 15) * 2u);
    int s = ax + i0p1 - 1;  // grid column of view column i0, plus 8; columns < 0 or > 16 are walls
    s = s > 31 ? 31 : (s < 0 ? 0 : s);
    const uint32_t colbits = ((0xFE0000FFu >> s) & 1u) | (((0xFF00007Fu >> s) & 1u) << 1);
    const uint32_t mk = reinterpret_cast<const uint32_t *>(tab + TAB_TOP)[ay * 17 + j0] |
                        reinterpret_cast<const uint32_t *>(tab + TAB_COL)[colbits * 17 + j0];
    uint32_t c = (win & ~mk) | (0x55555555u & mk);
    const int d = 152 - k0;  // agent cell: view (8,16)
    if ((unsigned)d < 16u) c &= ~(3u << (2 * d));
    return c;
}

// 16 packed cells -> 48 obs bytes through the 256-entry (4 cells -> 12 bytes) table, stored as
// three 16-byte vectors at dst (16-byte aligned)
__device__ __forceinline__ void expand16_store(uint32_t c, const uint8_t *tab, uint4 *dst) {
    const uint4 *lut = reinterpret_cast<const uint4 *>(tab + TAB_LUT);
    const uint4 a = lut[c & 0xFFu], b = lut[(c >> 8) & 0xFFu], cc = lut[(c >> 16) & 0xFFu], d = lut[c >> 24];
    dst[0] = make_uint4(a.x, a.y, a.z, b.x);
    dst[1] = make_uint4(b.y, b.z, cc.x, cc.y);
    dst[2] = make_uint4(cc.z, d.x, d.y, d.z);
}

// V = 17, whole tile: the obs block (32 x 867 B = 27744 B) is 578 runs of 16 cells / 48 bytes.
// Every lane builds one run per iteration; the 30 runs that straddle two envs are skipped in
// the main loop and rebuilt by warp 3 (which has one block fewer) from both envs.
// XPOSE: the warp's 32 x 48 B go through a shared-memory transposer so that each global store
// instruction writes 512 contiguous bytes; otherwise every lane stores its own 48 bytes.
template <bool XPOSE>
__device__ __forceinline__ void emit_run(uint32_t c, bool valid, const uint8_t *tab, uint4 *d4, int run, uint4 *xp,
                                         int lane) {
    if (XPOSE) {
        if (valid) expand16_store(c, tab, xp + 3 * lane);
        __syncwarp();
        const int u0 = (run - lane) * 3;  // first uint4 of this warp-iteration's 1536-byte block
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int src_lane = (32 * k + lane) / 3;  // the lane whose run produced this uint4
            const bool ok = __shfl_sync(0xFFFFFFFFu, (int)valid, src_lane) != 0;
            if (ok) d4[u0 + 32 * k + lane] = xp[32 * k + lane];
        }
        __syncwarp();
    } else {
        if (valid) expand16_store(c, tab, d4 + 3 * run);
    }
}

template <bool XPOSE>
__device__ __forceinline__ void obs_tile_v17(uint8_t *dst, const uint32_t *gpadw, const uint8_t *tab,
                                             const uint32_t *runtab, const uint32_t *meta, uint8_t *xpose, int warp,
                                             int lane) {
    constexpr int BLOCKS = 19;  // ceil(578 / 32)
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    uint4 *xp = reinterpret_cast<uint4 *>(xpose);
#pragma unroll 2
    for (int bi = warp; bi < BLOCKS; bi += STEP_WARPS) {
        const int run = bi * 32 + lane;
        const uint32_t rt = __ldg(runtab + run);  // coalesced, L1-resident after the first tile
        const int e = (int)(rt & 31u), k0 = (int)((rt >> 5) & 511u), j0 = (int)((rt >> 14) & 31u),
                  i0p1 = (int)((rt >> 19) & 31u);
        const uint32_t c = codes16_v17(gpadw, tab, meta[e], k0, i0p1, j0);
        emit_run<XPOSE>(c, (rt >> 24) == 0u, tab, d4, run, xp, lane);
    }
    if (warp == STEP_WARPS - 1) {
        // boundary between env cidx-1 and env cidx; env 16 starts exactly on a run boundary
        const int cidx = lane;
        const bool valid = lane >= 1 && lane != 16;
        const int run = (NCELL * cidx) >> 4;
        const int k0a = valid ? 16 * run - NCELL * (cidx - 1) : 274;  // [274,288]: column 16 of env cidx-1
        const int k0b = k0a - NCELL;                                   // [-15,-1]: "column -1" of env cidx
        const uint32_t ca = codes16_v17(gpadw, tab, meta[valid ? cidx - 1 : 0], k0a, 17, k0a - 272);
        const uint32_t cb = codes16_v17(gpadw, tab, meta[valid ? cidx : 1], k0b, 0, k0b + 17);
        const uint32_t keep = (1u << (2 * (NCELL - k0a))) - 1u;
        if (valid) expand16_store((ca & keep) | (cb & ~keep), tab, d4 + 3 * run);
    }
}

// Any other odd V <= 17: CE envs at a time are expanded per cell into the shared staging area
// and copied out with coalesced 4-byte stores.
template <int V>
__device__ __forceinline__ void obs_tile_generic(uint8_t *dst, long long limit_bytes, const uint32_t *gw,
                                                 const uint32_t *meta, uint8_t *stage, int tid) {
    constexpr int VV = V * V, OBS = 3 * VV;
    constexpr int CE = (32 * OBS <= STAGE_BYTES) ? 32 : (16 * OBS <= STAGE_BYTES) ? 16 : (8 * OBS <= STAGE_BYTES) ? 8 : 4;
    static_assert(CE * OBS <= STAGE_BYTES, "staging too small");
    for (int e0 = 0; e0 < TILE; e0 += CE) {
        for (int c = tid; c < CE * VV; c += STEP_THREADS) {
            const int e = c / VV, k = c - e * VV, i = k / V, j = k - i * V;
            const uint32_t m = meta[e0 + e];
            const int x = (int)(m >> 24) - V / 2 + i, y = (int)((m >> 16) & 31u) - (V - 1) + j;
            uint32_t code = inb(x, y) ? cell_get(gw + (e0 + e) * REC_WORDS, x, y) : C_WALL;
            if (i == V / 2 && j == V - 1) code = C_EMPTY;
            stage[3 * c + 0] = (uint8_t)(TYPE_LUT >> (8 * code));
            stage[3 * c + 1] = (uint8_t)(COLOR_LUT >> (8 * code));
            stage[3 * c + 2] = 0;
        }
        __syncthreads();
        const long long base = (long long)e0 * OBS;
        for (int w = tid; w < CE * OBS / 4; w += STEP_THREADS) {
            const long long b0 = base + 4ll * w;
            if (b0 + 4 <= limit_bytes) *reinterpret_cast<uint32_t *>(dst + b0) = reinterpret_cast<const uint32_t *>(stage)[w];
            else if (b0 < limit_bytes)
                for (int b = 0; b < (int)(limit_bytes - b0); b++) dst[b0 + b] = stage[4 * w + b];
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int load_action(const void *actions, int dtype, long long i) {
    if (dtype == 0) return reinterpret_cast<const int *>(actions)[i];
    if (dtype == 1) return (int)reinterpret_cast<const uint8_t *>(actions)[i];
    long long a = reinterpret_cast<const long long *>(actions)[i];
    return a > 1000 ? 1000 : (a < -1000 ? -1000 : (int)a);
}

template <int V, bool FAST>
__global__ void __launch_bounds__(STEP_THREADS, STEP_CTAS_PER_SM) step_obs_kernel(const StepArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *tab = smem + SM_TAB;
    uint64_t *tab_bar = reinterpret_cast<uint64_t *>(smem + SM_BARS);
    uint64_t *bar = tab_bar + 1;
    uint32_t *meta = reinterpret_cast<uint32_t *>(smem + SM_META);
    uint8_t *gpad = smem + SM_GPAD;
    uint32_t *gw = reinterpret_cast<uint32_t *>(gpad + G_PAD_BEFORE);
    uint8_t *stage = smem + SM_STAGE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int OBS_ENV = 3 * V * V;

    if (tid == 0) {
        mbar_init(tab_bar, 1);
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(tab_bar, TAB_SMEM_BYTES);
        bulk_g2s(tab, a.tables, TAB_SMEM_BYTES, tab_bar);
    }
    __syncthreads();
    uint32_t phase = 0;
    const bool v4 = a.version == 4;
    bool first = true;
    const uint32_t *runtab = reinterpret_cast<const uint32_t *>(a.tables + TAB_RUN);

    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        // ---- tile prologue (warp 0): state of 32 envs -> registers, packed grids -> smem -------
        const long long env = tile * TILE + lane;
        const bool live = env < a.n;
        int ax = 0, ay = 0, risk = 0, step_count = 0, step_move = 0;
        uint32_t fl = 0, tcount = 0, err = 0;
        uint32_t mid[3] = {0, 0, 0}, o1[3] = {0, 0, 0}, o2[4] = {0, 0, 0, 0};
        uint32_t *G = gw + lane * REC_WORDS;
        long long nvalid = a.n - tile * TILE;
        nvalid = nvalid > TILE ? TILE : nvalid;
        if (warp == 0) {
            if (lane == 0) {
                bulk_wait_read0();  // the previous tile's grid store has finished reading smem
                mbar_expect_tx(bar, G_BYTES);
                bulk_g2s(gw, a.grid + tile * (G_BYTES / 4), G_BYTES, bar);
            }
            const uint4 s0 = a.sc0[env], s1 = a.sc1[env];
            ax = (int)(s0.x & 0xFFu);
            ay = (int)((s0.x >> 8) & 0xFFu);
            fl = (s0.x >> 16) & 0xFFu;
            risk = (int)(s0.x >> 24);
            step_count = (int)s0.y;
            step_move = (int)s0.z;
            tcount = s0.w;
            mid[0] = ball_get(s1.x, 0); mid[1] = ball_get(s1.x, 1); mid[2] = ball_get(s1.x, 2);
            o1[0] = ball_get(s1.y, 0); o1[1] = ball_get(s1.y, 1); o1[2] = ball_get(s1.y, 2);
            o2[0] = ball_get(s1.z, 0); o2[1] = ball_get(s1.z, 1); o2[2] = ball_get(s1.z, 2);
            o2[3] = ball_get(s1.w, 0);
            err = (s1.w >> 16) & 0xFFu;
        }

        // ---- T env steps on the resident tile -------------------------------------------------
        for (int t = 0; t < a.T; t++) {
            const long long out_idx = (long long)t * a.n + env;  // [T][n] outputs
            bool skip = false, term = false, trunc = false;
            DrawSrc d;
            d.rec_lo = d.rec_hi = 0xFFFFFFFFu;
            d.consumed = 0;
            // ---- warp 0, phase A: everything up to and including the agent move -------------
            if (warp == 0) {
                int act = live ? load_action(a.actions, a.action_dtype, out_idx) : 6;
                if (a.draws != nullptr) {
                    if (live) {
                        const uint2 r = reinterpret_cast<const uint2 *>(a.draws)[out_idx];
                        d.rec_lo = r.x;
                        d.rec_hi = r.y;
                    }
                } else {
                    const unsigned long long gid = a.env_id0 + (unsigned long long)env;
                    philox_record(d, (uint32_t)gid, (uint32_t)(gid >> 32), tcount, a.seed_lo, a.seed_hi);
                }
                // situations where the reference raises (documented divergence: the env is left
                // untouched and an error bit is set)
                if (act >= 7) act = 0;  // twoarmy_v4.py:84-85
                int adx = 0, ady = 0;
                if (act == 0) adx = -1;
                else if (act == 1) adx = 1;
                else if (act == 2) ady = -1;
                else if (act == 3) ady = 1;
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
                if (t == 0) {
                    mbar_wait(bar, phase);
                    phase ^= 1u;
                }
                if (!skip) {
                    tcount += 1u;
                    step_move += 1;
                    const int m6 = step_move % 6, m4 = step_move & 3;
                    {  // mid-row balls, twoarmy_v4.py:95-111
                        const int dx = (m6 == 1 || m6 == 0) ? 1 : ((m6 == 2 || m6 == 3) ? -1 : 0);
                        move_group<3>(G, mid, dx, 0, true);
                    }
                    if (v4) {
                        bool mv1 = false, mv2 = false;
                        if (fl & F_UPD_L) {  // twoarmy_v4.py:115-144 (the draw is short-circuited)
                            fl &= ~F_UPD_H;
                            if (m4 == 2 || m6 == 3 || m6 == 0 || draw(d, 0) == 6) mv1 = (fl & F_PATROL) != 0;
                        }
                        if (fl & F_UPD_H) {  // twoarmy_v4.py:147-176
                            fl &= ~F_UPD_L;
                            if (m6 != 1 || draw(d, 0) == 6) mv2 = (fl & F_PATROL) != 0;
                        }
                        if (mv1) {
                            const bool up = (fl & F_UP1) != 0;
                            move_group<3>(G, o1, 0, up ? -1 : 1, false);
                            if (up) {
                                if (pos_y(o1[0]) == 3) fl &= ~F_UP1;
                            } else if (pos_y(o1[2]) == 7) {
                                fl |= F_UP1;
                            }
                        }
                        if (mv2) {
                            const bool right = (fl & F_RIGHT2) != 0;
                            move_group<4>(G, o2, right ? 1 : -1, 0, false);
                            if (right) {
                                if (pos_x(o2[3]) == 11) fl &= ~F_RIGHT2;
                            } else if (pos_x(o2[0]) == 5) {
                                fl |= F_RIGHT2;
                            }
                        }
                    }
                    // MiniGridEnv.step, minigrid.py:1333-1441
                    step_count += 1;
                    const int tx = ax + adx, ty = ay + ady;
                    const uint32_t c = cell_get(G, tx, ty);
                    if (c == C_EMPTY || c == C_GOAL) {
                        ax = tx;
                        ay = ty;
                    }
                    if (c == C_GOAL) term = true;
                    if (step_count >= 50) trunc = true;
                }
                meta[lane] = make_meta(lane, ax, ay);
            }
            if (first) {  // LUT, masks and the reset template
                mbar_wait(tab_bar, 0);
                first = false;
            }
            __syncthreads();

            // ---- phase B: observations of all 32 envs, every warp ----------------------------
            {
                uint8_t *dst = a.obs + ((long long)t * a.n + tile * TILE) * OBS_ENV;
                if (FAST && nvalid == TILE && (((long long)t * a.n * OBS_ENV) & 15) == 0) {
                    if (a.flags & 2)
                        obs_tile_v17<true>(dst, reinterpret_cast<const uint32_t *>(gpad), tab, runtab, meta,
                                           stage + warp * 1536, warp, lane);
                    else
                        obs_tile_v17<false>(dst, reinterpret_cast<const uint32_t *>(gpad), tab, runtab, meta, stage, warp,
                                            lane);
                } else {  // other view sizes, and the ragged last tile of a batch
                    obs_tile_generic<V>(dst, nvalid * OBS_ENV, gw, meta, stage, tid);
                }
            }
            __syncthreads();

            // ---- warp 0, phase C: rest of Twoarmy.step ----------------------------------------
            if (warp == 0) {
                int reward = R_STEP;  // twoarmy_v4.py:180
                bool need_reset = false;
                if (!skip) {
                    if (!(fl & F_PONE) && (ax > 3 || ay < 14)) {  // twoarmy_v4.py:181-195, twoarmy_v6.py:182-198
                        int i = v4 ? draw(d, 1) : 11;
                        put_cell(G, 4, i, C_WALL); put_cell(G, 5, i, C_WALL);
                        put_cell(G, 4, i + 1, C_WALL); put_cell(G, 5, i + 1, C_WALL);
                        i = v4 ? draw(d, 2) : 8;
                        put_cell(G, i, 11, C_WALL); put_cell(G, i, 12, C_WALL);
                        put_cell(G, i + 1, 11, C_WALL); put_cell(G, i + 1, 12, C_WALL);
                        fl |= F_PONE;
                    }
                    if (v4 && !(fl & F_PATROL) && ay <= 8) {  // twoarmy_v4.py:212-225
                        const int i = draw(d, 3);
                        const int bx[4] = {i, i + 1, i, i + 1}, by[4] = {4, 4, 5, 5};
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (inb(bx[k], by[k])) {
                                cell_set(G, bx[k], by[k], C_BALL);
                                o2[k] = pack_pos(bx[k], by[k]);
                            }
                        d.consumed |= 1u << 4;  // :221 choice(range(4,5)) == 4, no generator words
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            cell_set(G, 12, 4 + k, C_BALL);
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
                        if (draw(d, 5) == 1) fl = (fl & ~F_UP1) | F_RIGHT2;
                        else fl = (fl | F_UP1) & ~F_RIGHT2;
                        if (draw(d, 6) == 1) fl = (fl & ~F_UPD_H) | F_UPD_L;
                        else fl = (fl | F_UPD_H) & ~F_UPD_L;
                        need_reset = (a.flags & 1) != 0;
                    }
                }
                if (live) {
                    a.reward[out_idx] = reward_value(reward);
                    a.term[out_idx] = term ? 1 : 0;
                    a.trunc[out_idx] = trunc ? 1 : 0;
                    if (a.consumed) a.consumed[out_idx] = (uint8_t)d.consumed;
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
                    if (lane < REC_WORDS)
                        gw[e * REC_WORDS + lane] = reinterpret_cast<const uint32_t *>(tab + TAB_TEMPLATE)[lane];
                }
                __syncwarp();
            }
        }

        // ---- tile epilogue (warp 0): registers -> state arrays, packed grids -> HBM ------------
        if (warp == 0) {
            uint4 s0, s1;
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
                bulk_s2g(a.grid + tile * (G_BYTES / 4), gw, G_BYTES);
                bulk_commit();
            }
        }
    }
    if (tid == 0) bulk_wait_all0();
}

}  // namespace ta
