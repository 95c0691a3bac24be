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
// Mapping: ONE WARP owns a tile of 32 envs and is its own CTA, so tiles never wait for each
// other (no __syncthreads anywhere); up to 16-19 such CTAs are resident per SM, which puts all
// 2048 tiles of a 65536-env batch on chip in one wave.  Per tile:
//   load    the tile's 2560 B of packed grids as five perfectly coalesced 16-byte loads per lane, scattered
//           into guard-padded shared-memory slots; the scalar state (2 x 16 B per env) goes straight
//           into the lanes' registers.  (Per-lane TMA bulk copies were tried first: a bulk copy with
//           lane-dependent addresses is issued by a 32-trip serial loop, ~290 instructions each way,
//           and the mbarrier wait spins; plain vector loads cost 15 instructions and stall silently.)
//   phase A lane e = env e: balls, patrols, agent move            (twoarmy_v4.py:82-179)
//   obs     the tile's 32 observations are one contiguous byte block in HBM.  It is produced as
//           "runs" of 16 view cells = one 32-bit word of 2-bit codes = 48 output bytes; lane l
//           builds run it*32+l, expands it with byte permutes (no table), writes its 48 bytes to
//           a shared-memory ring slot, and every 32 runs (1536 B) leave with one TMA bulk store
//   phase C lane e = env e: wall blocks, patrol spawn, reward, episode end, autoreset
//                                                                  (twoarmy_v4.py:180-322)
//   store   scalars from registers, the packed grids with five coalesced 16-byte stores per lane
// With T > 1 (ta_rollout) A/obs/C repeat T times on the resident tile.  The obs is built
// between A and C because the reference builds it inside MiniGridEnv.step, i.e. before the wall
// blocks / patrol balls of the same step appear (ordering fact a).
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int STEP_MAX_WARPS = 8;  // independent warps (tiles) per CTA; they share nothing but the launch
// shared-memory grid tile: env e's 20-word record sits at word GUARD0_WORDS + e*SLOT_WORDS; every
// other word is a guard full of wall codes, so that view columns left / right of the grid (and
// the record's own padding cells 289..319, kept at "wall" in HBM) read as walls without a mask.
// The furthest a view reaches is 134 cells before and 150 cells after a record.
constexpr int SLOT_WORDS = 36;
constexpr int GUARD0_WORDS = 12;
constexpr int GRID_S_WORDS = GUARD0_WORDS + TILE * SLOT_WORDS;  // 1164
constexpr uint32_t WALLS16 = 0x55555555u;                        // 16 wall codes
constexpr int RUN_BYTES = 48;                                    // 16 cells x (type,color,state)
constexpr int ITER_BYTES = 32 * RUN_BYTES;                       // one warp iteration = 1536 B
constexpr int CHUNK_ITERS = 2;                                   // warp iterations per bulk store
constexpr int CHUNK_BYTES = CHUNK_ITERS * ITER_BYTES;            // 3072 B
constexpr int RING = 2;                                          // chunk buffers per warp

constexpr int round16(int x) { return (x + 15) / 16 * 16; }

template <int V>
struct ObsCfg {
    static constexpr int VV = V * V;
    static constexpr int OBS = 3 * VV;   // bytes per env
    static constexpr int L = 2 * VV;     // bits of one env's packed view
    static constexpr int RUNS = 2 * VV;  // 16-cell runs in a 32-env tile (32 * VV / 16)
    static constexpr int ITERS = (RUNS + 31) / 32;
    static constexpr int CHUNKS = (ITERS + CHUNK_ITERS - 1) / CHUNK_ITERS;
    static constexpr int P = (L + 31) / 32;                              // words per env view (V < 17)
    static constexpr int EV_STRIDE = (V == 17) ? 0 : ((P + 1) | 1);      // odd: conflict-free pass 1
    static constexpr int EV_ROWS = 34;                                   // env 32, 33: readable padding
    // shared memory map (bytes)
    static constexpr int SM_META = 16;                                   // 32 x uint4 (V = 17)
    static constexpr int SM_HEAD = SM_META + 512;                        // 33 words (V = 17)
    static constexpr int SM_GRID = SM_HEAD + 144;
    static constexpr int SM_EV = SM_GRID + GRID_S_WORDS * 4;
    static constexpr int SM_RING = SM_EV + round16(EV_ROWS * EV_STRIDE * 4);
    static constexpr int SMEM = (SM_RING + RING * CHUNK_BYTES + 127) / 128 * 128;  // per warp
    static_assert((RING & (RING - 1)) == 0, "RING must be a power of two");
    static_assert(SM_GRID % 16 == 0 && (SM_GRID + GUARD0_WORDS * 4) % 16 == 0 && SM_RING % 16 == 0, "alignment");
};

struct StepArgs {
    uint32_t *grid;
    uint4 *sc0;
    uint4 *sc1;
    const uint32_t *tmpl;  // [20] the _gen_grid record, [20..23] TYPE_LUT, COLOR_LUT twice
    const void *actions;
    const uint8_t *draws;
    uint8_t *obs;
    float *reward;
    uint8_t *term;
    uint8_t *trunc;
    uint8_t *consumed;
    uint8_t *status;  // nullable [T][npad]: reward index | terminated << 3 | truncated << 4 (the host call's one-byte result)
    long long n;
    int ntiles;
    int T;  // env steps per launch; outputs / actions / draws are [T][n]
    int version;
    int flags;  // bit 0 autoreset, bit 1 never use bulk stores for the obs (test hook),
                // bits 2,3 timing experiments only: skip the obs pass / the scalar phases,
                // bit 4 observe only: gen_obs() of the current state, nothing else read or written
                // bit 5 packed observations: `obs` receives the tile's cell stream as 2-bit codes (uint32 [ntiles][2*V*V],
                //       word r = the 16 cells of run r) instead of the 3 bytes per cell of Grid.encode -- the transfer
                //       form of ta_step_host, expanded to the same bytes by the host side of that call
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

// 8 two-bit codes (low 16 bits of x) -> 8 nibbles
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    return x;
}

// 4 cells (4 selector nibbles in the low 16 bits of sel) -> 12 obs bytes.  Two permutes look the
// type / colour bytes up in the register-resident tables, three more interleave them with the
// zero state byte (selector nibble 8 = sign-replicate a byte < 0x80 = 0x00).
// (__byte_perm masks the selector to 3 bits per nibble, so the PTX instruction is used directly.)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ void expand4(uint32_t sel, uint32_t tl, uint32_t cl, uint32_t &w0, uint32_t &w1, uint32_t &w2) {
    const uint32_t t = prmt(tl, 0u, sel), c = prmt(cl, 0u, sel);
    w0 = prmt(t, c, 0x1840u);  // t0 c0 0  t1
    w1 = prmt(t, c, 0x6285u);  // c1 0  t2 c2
    w2 = prmt(t, c, 0x8738u);  // 0  t3 c3 0
}

// 16 packed cells -> 48 obs bytes, stored as three 16-byte vectors at dst (shared memory)
// (tl, cl: TYPE_LUT / COLOR_LUT held in registers by the caller)
__device__ __forceinline__ void expand16_store(uint32_t c, uint32_t tl, uint32_t cl, uint4 *dst) {
    const uint32_t sa = spread16(c & 0xFFFFu), sb = spread16(c >> 16);
    uint32_t w[12];
    expand4(sa, tl, cl, w[0], w[1], w[2]);
    expand4(sa >> 16, tl, cl, w[3], w[4], w[5]);
    expand4(sb, tl, cl, w[6], w[7], w[8]);
    expand4(sb >> 16, tl, cl, w[9], w[10], w[11]);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
}

// V = 17.  The view is as wide as the grid and the record is column-major, so the 16 view cells
// k0..k0+15 of an env are the 32 bits at cell offset cellbase + k0 of the guard-padded tile
// (one funnel shift), except that rows above the grid (they alias the previous column's tail)
// become walls -- a 34-bit periodic mask kept per env as two words.  The agent's own cell (view
// cell 152) must read as empty: the kernel clears it in the shared-memory grid for the duration
// of the obs pass.  meta = (cellbase, mask lo, mask hi, -).
__device__ __forceinline__ uint4 make_meta17(int e, int ax, int ay) {
    const int h = 16 - ay;  // rows above the grid
    const unsigned long long rm = (1ull << (2 * (h < 0 ? 0 : h))) - 1ull;
    const unsigned long long r2 = rm | (rm << 34);
    return make_uint4((uint32_t)((GUARD0_WORDS + e * SLOT_WORDS) * 16 + (ax - 8) * GS + (ay - 16)), (uint32_t)r2,
                      (uint32_t)(r2 >> 32), 0u);
}
__device__ __forceinline__ uint32_t fetch17(const uint32_t *gs, uint4 m, int k0) {
    const int j0 = k0 - GS * ((k0 * 241) >> 12);  // k0 mod 17, exact for k0 < 320
    const int cell = (int)m.x + k0, wi = cell >> 4;
    const uint32_t win = __funnelshift_r(gs[wi], gs[wi + 1], (uint32_t)(cell & 15) * 2u);
    const uint32_t mk = __funnelshift_rc(m.y, m.z, (uint32_t)(2 * j0));
    return (win & ~mk) | (WALLS16 & mk);
}

// V < 17, pass 1: lane e packs env e's V x V view (2 bits per cell, column after column) into
// ev[e][0..P).  Everything about the bit positions is known at compile time.
template <int V>
__device__ __forceinline__ void build_envview(const uint32_t *gs, uint32_t *ev, int lane, int ax, int ay) {
    using C = ObsCfg<V>;
    uint32_t out[C::P + 1];
#pragma unroll
    for (int w = 0; w <= C::P; w++) out[w] = 0u;
    int h = V - 1 - ay;  // rows above the grid
    h = h < 0 ? 0 : h;
    const uint32_t mk = (1u << (2 * h)) - 1u;  // h <= 14
    const int base = (GUARD0_WORDS + lane * SLOT_WORDS) * 16 + (ax - V / 2) * GS + (ay - (V - 1));
#pragma unroll
    for (int i = 0; i < V; i++) {
        const int cell = base + GS * i, wi = cell >> 4;
        uint32_t bits = __funnelshift_r(gs[wi], gs[wi + 1], (uint32_t)(cell & 15) * 2u) & ((1u << (2 * V)) - 1u);
        bits = (bits & ~mk) | (WALLS16 & mk & ((1u << (2 * V)) - 1u));
        if (i == V / 2) bits &= ~(3u << (2 * (V - 1)));  // the agent's own cell
        const int pos = i * 2 * V, w = pos >> 5, s = pos & 31;
        out[w] |= bits << s;
        if (s + 2 * V > 32) out[w + 1] |= bits >> (32 - s);
    }
#pragma unroll
    for (int w = 0; w < C::P; w++) ev[lane * C::EV_STRIDE + w] = out[w];
}

// The 16 cells of run r of the tile's cell stream (env e = 16r / VV, possibly continuing into
// env e+1, and for V = 3 into e+2) as one word of 2-bit codes.
template <int V>
__device__ __forceinline__ uint32_t run_codes(int r, const uint32_t *gs, const uint4 *meta, const uint32_t *head,
                                              const uint32_t *ev) {
    using C = ObsCfg<V>;
    if constexpr (V == 17) {
        const int q0 = 16 * r, e = q0 / NCELL, k0 = q0 - NCELL * e;
        uint32_t c = fetch17(gs, meta[e], k0);
        const int nb = 2 * (NCELL - k0);  // bits env e still has from k0 on
        if (nb < 32) c = (c & ((1u << nb) - 1u)) | (head[e + 1] << nb);
        return c;
    } else {
        constexpr int S = C::EV_STRIDE;
        const int p = 32 * r, e = p / C::L, off = p - C::L * e, wi = off >> 5;
        uint32_t c = __funnelshift_r(ev[e * S + wi], ev[e * S + wi + 1], (uint32_t)(off & 31));
        const int nb = C::L - off;
        if (nb < 32) {
            c = (c & ((1u << nb) - 1u)) | (ev[(e + 1) * S] << nb);
            if (V == 3 && nb + C::L < 32) c |= ev[(e + 2) * S] << (nb + C::L);
        }
        return c;
    }
}

// slow path (obs slice not 16-byte aligned, ragged last tile, or the test hook): nbytes of a
// staging buffer -> global memory with per-lane stores
__device__ __forceinline__ void copy_out(uint8_t *dst, const uint8_t *src, int nbytes, int lane) {
    if ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
        const int nw = nbytes > 0 ? nbytes >> 2 : 0;
        for (int w = lane; w < nw; w += 32) reinterpret_cast<uint32_t *>(dst)[w] = reinterpret_cast<const uint32_t *>(src)[w];
        for (int b = 4 * nw + lane; b < nbytes; b += 32) dst[b] = src[b];
    } else {
        for (int b = lane; b < nbytes; b += 32) dst[b] = src[b];
    }
}

__device__ __forceinline__ int load_action(const void *actions, int dtype, long long i) {
    if (dtype == 0) return reinterpret_cast<const int *>(actions)[i];
    if (dtype == 1) return (int)reinterpret_cast<const uint8_t *>(actions)[i];
    long long a = reinterpret_cast<const long long *>(actions)[i];
    return a > 1000 ? 1000 : (a < -1000 ? -1000 : (int)a);
}

template <int V>
__global__ void __launch_bounds__(32 * STEP_MAX_WARPS) step_obs_kernel(const StepArgs a) {
    using C = ObsCfg<V>;
    extern __shared__ __align__(128) uint8_t smem_all[];
    const int warps_per_cta = (int)(blockDim.x >> 5), warp = (int)(threadIdx.x >> 5);
    uint8_t *smem = smem_all + warp * C::SMEM;  // every warp has its own slice
    uint4 *meta = reinterpret_cast<uint4 *>(smem + C::SM_META);
    uint32_t *head = reinterpret_cast<uint32_t *>(smem + C::SM_HEAD);
    uint32_t *gs = reinterpret_cast<uint32_t *>(smem + C::SM_GRID);
    uint32_t *ev = reinterpret_cast<uint32_t *>(smem + C::SM_EV);
    uint8_t *ring = smem + C::SM_RING;
    const int lane = threadIdx.x & 31;
    uint32_t *G = gs + GUARD0_WORDS + lane * SLOT_WORDS;  // this lane's env record

    // Programmatic dependent launch (flags bits 6..8, set by the host when the launch carries the attribute).  The
    // kernel that follows in the stream may be scheduled once every CTA of this grid has executed launch_dependents
    // (or exited); it may then run its preamble while this grid drains, and blocks in griddep_wait() until this grid
    // has completed and flushed.  Nothing is WRITTEN to global memory before griddep_wait(), and the only global
    // reads before it are the constant template and -- flags bit 6, which the host sets only when the kernel launched
    // before this one did not write this handle's state -- the first tile's env state, so that its DRAM latency
    // overlaps the previous grid's tail instead of heading this one.
    //   trigger (flags bits 7,8): 0 at the top (the next grid's CTAs queue for SM resources at once), 1 right after this
    //   grid's own wait (so at most one grid runs ahead: the grid before this one has completed), 2 after the last tile's
    //   observation pass
    const int pdl_trigger = (a.flags >> 7) & 3;
    if (pdl_trigger == 0) griddep_launch_dependents();
    // one-time setup: the reset template word of this lane (guards: see below)
    const uint32_t tmpl_word = lane < REC_WORDS ? __ldg(a.tmpl + lane) : 0u;
    const long long tile_first = (long long)blockIdx.x * warps_per_cta + warp, tile_stride = (long long)gridDim.x * warps_per_cta;
    uint4 piece[5], s0, s1;
    bool preloaded = false;
    if ((a.flags & 64) && tile_first < a.ntiles) {
        const uint4 *gsrc = reinterpret_cast<const uint4 *>(a.grid + tile_first * (TILE * REC_WORDS));
#pragma unroll
        for (int k = 0; k < 5; k++) piece[k] = gsrc[lane + 32 * k];
        s0 = a.sc0[tile_first * TILE + lane];
        s1 = a.sc1[tile_first * TILE + lane];
        preloaded = true;
    }
    __syncwarp();
    bool first_tile = true;
    griddep_wait();
    if (pdl_trigger == 1) griddep_launch_dependents();
    // the two look-up words live in ordinary registers: loaded through a lane-dependent address
    // (tmpl[20..23] = TYPE, COLOR, TYPE, COLOR) so that ptxas cannot turn them back into immediates
    // / uniform registers, which it re-materialises in front of every permute
    const uint32_t type_lut = __ldg(a.tmpl + REC_WORDS + 2 * (lane & 1));
    const uint32_t color_lut = __ldg(a.tmpl + REC_WORDS + 1 + 2 * (lane & 1));
    uint32_t gi = 0;  // obs chunks emitted so far (ring slot = gi % RING)
    const bool v4 = a.version == 4;

    for (long long tile = tile_first; tile < a.ntiles; tile += tile_stride) {
        // ---- tile prologue: packed grids -> smem, state of 32 envs -> registers ------------------
        const long long env = tile * TILE + lane;
        const bool live = env < a.n;
        long long nvalid = a.n - tile * TILE;
        nvalid = nvalid > TILE ? TILE : nvalid;
        // the tile's 32 records are 160 contiguous 16-byte pieces: piece q = lane + 32k belongs to env q / 5
        if (!preloaded) {
            const uint4 *gsrc = reinterpret_cast<const uint4 *>(a.grid + tile * (TILE * REC_WORDS));
#pragma unroll
            for (int k = 0; k < 5; k++) piece[k] = gsrc[lane + 32 * k];
            s0 = a.sc0[env];
            s1 = a.sc1[env];
        }
        preloaded = false;
        if (first_tile) {  // guard words (never written again), while the loads are in flight
            first_tile = false;
            const uint4 w4 = make_uint4(WALLS16, WALLS16, WALLS16, WALLS16);
            uint4 *g4 = reinterpret_cast<uint4 *>(G + REC_WORDS);
            g4[0] = w4; g4[1] = w4; g4[2] = w4; g4[3] = w4;
            if (lane < GUARD0_WORDS / 4) reinterpret_cast<uint4 *>(gs)[lane] = w4;
            if constexpr (V == 17) {
                if (lane == 0) head[32] = 0u;
            } else {
                for (int w = lane; w < 2 * C::EV_STRIDE; w += 32) ev[32 * C::EV_STRIDE + w] = 0u;
            }
        }
        int ax = (int)(s0.x & 0xFFu), ay = (int)((s0.x >> 8) & 0xFFu), risk = (int)(s0.x >> 24);
        uint32_t fl = (s0.x >> 16) & 0xFFu;
        int step_count = (int)s0.y, step_move = (int)s0.z;
        uint32_t tcount = s0.w;
        uint32_t mid[3] = {ball_get(s1.x, 0), ball_get(s1.x, 1), ball_get(s1.x, 2)};
        uint32_t o1[3] = {ball_get(s1.y, 0), ball_get(s1.y, 1), ball_get(s1.y, 2)};
        uint32_t o2[4] = {ball_get(s1.z, 0), ball_get(s1.z, 1), ball_get(s1.z, 2), ball_get(s1.w, 0)};
        uint32_t err = (s1.w >> 16) & 0xFFu;

        // ---- T env steps on the resident tile -------------------------------------------------
        for (int t = 0; t < a.T; t++) {
            const long long out_idx = (long long)t * a.n + env;  // [T][n] outputs
            bool skip = false, term = false, trunc = false;
            DrawSrc d;
            d.rec_lo = d.rec_hi = 0xFFFFFFFFu;
            d.consumed = 0;
            // ---- phase A: everything up to and including the agent move ----------------------
            const bool observe_only = (a.flags & 16) != 0;
            int act = (live && !observe_only) ? load_action(a.actions, a.action_dtype, out_idx) : 6;
            if (observe_only) {
            } else if (a.draws != nullptr) {
                if (live) {
                    const uint2 r = reinterpret_cast<const uint2 *>(a.draws)[out_idx];
                    d.rec_lo = r.x;
                    d.rec_hi = r.y;
                }
            } else {
                const unsigned long long gid = a.env_id0 + (unsigned long long)env;
                philox_record(d, (uint32_t)gid, (uint32_t)(gid >> 32), tcount, a.seed_lo, a.seed_hi);
            }
            if (act >= 7) act = 0;  // twoarmy_v4.py:84-85
            int adx = 0, ady = 0;
            if (act == 0) adx = -1;
            else if (act == 1) adx = 1;
            else if (act == 2) ady = -1;
            else if (act == 3) ady = 1;
            // situations where the reference raises (documented divergence: the env is left
            // untouched and an error bit is set)
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
            if (t == 0) {  // the packed grids have landed: scatter the pieces into the padded slots
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int q = lane + 32 * k, e = (q * 205) >> 10, part = q - 5 * e;  // q / 5 exact for q < 1024
                    *reinterpret_cast<uint4 *>(gs + GUARD0_WORDS + e * SLOT_WORDS + 4 * part) = piece[k];
                }
                __syncwarp();
            }
            if (a.flags & (8 | 16)) skip = true;
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

            // ---- observations of all 32 envs (gen_obs) ------------------------------------------
            uint32_t agent_cell = C_EMPTY;
            if constexpr (V == 17) {
                if (inb(ax, ay)) {  // the agent's own cell reads as empty (minigrid.py:1472-1476)
                    agent_cell = cell_get(G, ax, ay);
                    if (agent_cell != C_EMPTY) cell_set(G, ax, ay, C_EMPTY);
                }
                const uint4 m = make_meta17(lane, ax, ay);
                meta[lane] = m;
                __syncwarp();  // every lane's grid edits are visible
                head[lane] = fetch17(gs, m, 0);
            } else {
                __syncwarp();
                build_envview<V>(gs, ev, lane, ax, ay);
            }
            __syncwarp();
            if (a.flags & 32) {
                uint32_t *cdst = reinterpret_cast<uint32_t *>(a.obs) + ((long long)t * a.ntiles + tile) * C::RUNS;
#pragma unroll 1
                for (int it = 0; it < C::ITERS; it++) {
                    const int r = it * 32 + lane;
                    if (r < C::RUNS) cdst[r] = run_codes<V>(r, gs, meta, head, ev);
                }
            } else if (!(a.flags & 4)) {
                uint8_t *dst = a.obs + ((long long)t * a.n + tile * TILE) * C::OBS;
                const bool fast = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && !(a.flags & 2) && nvalid == TILE;
                if (fast) {
                    // whole tile, aligned: CHUNK_ITERS x 1536 B are staged, then leave with one bulk store
#pragma unroll 1
                    for (int ch = 0; ch < C::CHUNKS; ch++) {
                        uint8_t *slot = ring + (gi & (RING - 1)) * CHUNK_BYTES;
                        gi++;
                        if (lane == 0) bulk_wait_read<RING - 1>();  // the slot's previous bulk store has read it
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < CHUNK_ITERS; k++) {
                            const int it = ch * CHUNK_ITERS + k;
                            if (it < C::ITERS) {
                                int r = it * 32 + lane;
                                r = r < C::RUNS ? r : C::RUNS - 1;
                                const uint32_t c = run_codes<V>(r, gs, meta, head, ev);
                                expand16_store(c, type_lut, color_lut,
                                               reinterpret_cast<uint4 *>(slot + k * ITER_BYTES + lane * RUN_BYTES));
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            int len = C::RUNS * RUN_BYTES - ch * CHUNK_BYTES;
                            len = len > CHUNK_BYTES ? CHUNK_BYTES : len;
                            bulk_s2g(dst + ch * CHUNK_BYTES, slot, (uint32_t)len);
                            bulk_commit();
                        }
                    }
                } else {
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    const int valid_bytes = (int)nvalid * C::OBS;
#pragma unroll 1
                    for (int it = 0; it < C::ITERS; it++) {
                        int r = it * 32 + lane;
                        r = r < C::RUNS ? r : C::RUNS - 1;
                        const uint32_t c = run_codes<V>(r, gs, meta, head, ev);
                        expand16_store(c, type_lut, color_lut, reinterpret_cast<uint4 *>(ring + lane * RUN_BYTES));
                        __syncwarp();
                        int len = (C::RUNS - it * 32) * RUN_BYTES;
                        len = len > ITER_BYTES ? ITER_BYTES : len;
                        const int rem = valid_bytes - it * ITER_BYTES;
                        copy_out(dst + it * ITER_BYTES, ring, len < rem ? len : rem, lane);
                        __syncwarp();
                    }
                }
            }
            if (pdl_trigger == 2 && t == a.T - 1 && tile + tile_stride >= a.ntiles) griddep_launch_dependents();
            if constexpr (V == 17) {
                __syncwarp();
                if (agent_cell != C_EMPTY) cell_set(G, ax, ay, agent_cell);
            }
            __syncwarp();  // the obs pass has read the grids; phase C edits them

            // ---- phase C: rest of Twoarmy.step -----------------------------------------------
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
            if (live && !observe_only) {
                if (a.status) {
                    a.status[(long long)t * a.ntiles * TILE + env] = (uint8_t)(reward | (term ? 8 : 0) | (trunc ? 16 : 0));
                } else {
                    a.reward[out_idx] = reward_value(reward);
                    a.term[out_idx] = term ? 1 : 0;
                    a.trunc[out_idx] = trunc ? 1 : 0;
                }
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
                if (lane < REC_WORDS) gs[GUARD0_WORDS + e * SLOT_WORDS + lane] = tmpl_word;
            }
            __syncwarp();
        }

        // ---- tile epilogue: registers -> state arrays, packed grids -> HBM -------------------------
        if (a.flags & 16) {  // observe only: the state is unchanged
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            continue;
        }
        uint4 o0, o1w;
        o0.x = (uint32_t)ax | ((uint32_t)ay << 8) | (fl << 16) | ((uint32_t)risk << 24);
        o0.y = (uint32_t)step_count;
        o0.z = (uint32_t)step_move;
        o0.w = tcount;
        o1w.x = mid[0] | (mid[1] << 10) | (mid[2] << 20);
        o1w.y = o1[0] | (o1[1] << 10) | (o1[2] << 20);
        o1w.z = o2[0] | (o2[1] << 10) | (o2[2] << 20);
        o1w.w = o2[3] | (err << 16);
        a.sc0[env] = o0;
        a.sc1[env] = o1w;
        __syncwarp();
        uint4 *gdst = reinterpret_cast<uint4 *>(a.grid + tile * (TILE * REC_WORDS));
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int q = lane + 32 * k, e = (q * 205) >> 10, part = q - 5 * e;
            gdst[q] = *reinterpret_cast<const uint4 *>(gs + GUARD0_WORDS + e * SLOT_WORDS + 4 * part);
        }
        if (lane == 0) bulk_wait_read<0>();  // the obs ring may be rewritten (next tile) or released (exit)
        __syncwarp();
    }
}

}  // namespace ta
