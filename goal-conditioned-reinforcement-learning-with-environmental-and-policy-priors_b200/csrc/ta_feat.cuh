// ta_feat.cuh -- featuriser kernels, vectorised (16-byte stores, HBM-bound).
//
//   frame_codes_tile_kernel   Env_transact.matrix_env / data_env   soa/env_buffer.py:300-334
//   stack_push_codes_tile_kernel (uint8 codes), stack_push_kernel<float> (float32 LUT values)
//                       Env_transact.reset's np.tile (env_buffer.py:420-423) + the frame-stack roll
//                       np.delete(x,0,0); np.append(x,[new],0) (soa/train_ppo.py:116-121) in ONE
//                       out-of-place pass: out = [base[1..4], current frame], base = the previous
//                       stack, or the tiled reset frame for envs whose episode has just restarted
//
// A CTA owns FEAT_ENVS = 16 consecutive envs, so its slice of every output array starts on a
// 16-byte boundary (16*289 and 16*1445 elements) and is produced as 16-byte chunks: one thread =
// one chunk = one vector store.  The packed column-major records of the 16 envs are staged in
// shared memory once (coalesced) and decoded per output element.
#pragma once
#include "ta_aux.cuh"

namespace ta {

constexpr int FEAT_ENVS = 16;
constexpr int FEAT_THREADS = 256;
constexpr int STACK_ELEMS = 5 * NCELL;   // 1445 per env
constexpr int SHIFT_ELEMS = 4 * NCELL;   // 1156: frames 0..3 of the output come from frames 1..4

// matrix_env code of row-major cell c = y*17 + x of one staged record
__device__ __forceinline__ uint32_t frame_code(const uint32_t *rec, uint32_t agent_xy, int c) {
    const int y = (c * 241) >> 12, x = c - GS * y;  // c / 17 exact for c < 320
    return matrix_code(cell_get(rec, x, y), (uint32_t)(x | (y << 8)) == agent_xy);
}
// the same for the state MiniGridEnv.reset leaves behind (_gen_grid, agent at (3,15))
__device__ __forceinline__ uint32_t reset_frame_code(int c) {
    const int y = (c * 241) >> 12, x = c - GS * y;
    return matrix_code(initial_cell(x, y), x == 3 && y == 15);
}

template <typename ST>
struct Chunk;
template <>
struct Chunk<uint8_t> {
    static constexpr int K = 16;
    uint32_t w[4] = {0, 0, 0, 0};
    __device__ __forceinline__ void set(int k, uint32_t code) { w[k >> 2] |= code << (8 * (k & 3)); }
    __device__ __forceinline__ void store(uint8_t *dst) const { *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]); }
    __device__ __forceinline__ uint8_t get(int k) const { return (uint8_t)(w[k >> 2] >> (8 * (k & 3))); }
    static __device__ __forceinline__ uint8_t conv(uint32_t code) { return (uint8_t)code; }
};
template <>
struct Chunk<float> {
    static constexpr int K = 4;
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    __device__ __forceinline__ void set(int k, uint32_t code) { f[k] = matrix_value(code); }
    __device__ __forceinline__ void store(float *dst) const { *reinterpret_cast<float4 *>(dst) = make_float4(f[0], f[1], f[2], f[3]); }
    __device__ __forceinline__ float get(int k) const { return f[k]; }
    static __device__ __forceinline__ float conv(uint32_t code) { return matrix_value(code); }
};

__device__ __forceinline__ void stage_records(const uint32_t *grid, const uint4 *sc0, long long e0, int cnt, uint32_t *sg,
                                              uint32_t *sa) {
    for (int i = threadIdx.x; i < cnt * REC_WORDS; i += FEAT_THREADS) sg[i] = grid[e0 * REC_WORDS + i];
    if (threadIdx.x < cnt) sa[threadIdx.x] = sc0[e0 + threadIdx.x].x & 0xFFFFu;  // x | y << 8
}

// s_prev, s_out: ST [n][5][289]; p_prev, p_out: float32 [n][5][2]; prev_done uint8 [n] (nullable);
// init_all != 0: every env starts from the tiled reset frame (s_prev / p_prev may be NULL then).
template <typename ST>
__global__ void __launch_bounds__(FEAT_THREADS) stack_push_kernel(const uint32_t *grid, const uint4 *sc0, const ST *s_prev,
                                                                 ST *s_out, const float *p_prev, float *p_out,
                                                                 const uint8_t *prev_done, int init_all, long long n) {
    __shared__ uint32_t sg[FEAT_ENVS * REC_WORDS];
    __shared__ uint32_t sa[FEAT_ENVS];
    __shared__ uint8_t sdone[FEAT_ENVS];
    __shared__ uint8_t sreset[NCELL + 3];
    constexpr int K = Chunk<ST>::K;
    const long long e0 = (long long)blockIdx.x * FEAT_ENVS;
    const int cnt = (int)((n - e0) < FEAT_ENVS ? (n - e0) : FEAT_ENVS);
    stage_records(grid, sc0, e0, cnt, sg, sa);
    if (threadIdx.x < cnt) sdone[threadIdx.x] = (uint8_t)(init_all || (prev_done && prev_done[e0 + threadIdx.x]));
    for (int c = threadIdx.x; c < NCELL; c += FEAT_THREADS) sreset[c] = (uint8_t)reset_frame_code(c);
    __syncthreads();
    const int total = cnt * STACK_ELEMS;
    const ST *in = s_prev + e0 * STACK_ELEMS;  // never dereferenced where sdone is set
    ST *out = s_out + e0 * STACK_ELEMS;
    for (int q0 = threadIdx.x * K; q0 < total; q0 += FEAT_THREADS * K) {
        Chunk<ST> ch;
        const int e = q0 / STACK_ELEMS, j = q0 - STACK_ELEMS * e;
        if (j + K <= SHIFT_ELEMS && !sdone[e]) {
            // whole chunk = 16 bytes of the previous stack, one frame (289 elements) further on
            if constexpr (K == 16) {
                // in + q0 + 288 is 16-byte aligned like out + q0; the 17th byte completes the shift by one
                const uint4 a = *reinterpret_cast<const uint4 *>(in + q0 + NCELL - 1);
                const uint32_t last = in[q0 + NCELL + 15];
                ch.w[0] = __funnelshift_r(a.x, a.y, 8);
                ch.w[1] = __funnelshift_r(a.y, a.z, 8);
                ch.w[2] = __funnelshift_r(a.z, a.w, 8);
                ch.w[3] = (a.w >> 8) | (last << 24);
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) ch.f[k] = in[q0 + NCELL + k];
            }
        } else {
            int ee = e, jj = j;
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (q0 + k < total) {
                    if (jj >= SHIFT_ELEMS) ch.set(k, frame_code(sg + ee * REC_WORDS, sa[ee], jj - SHIFT_ELEMS));
                    else if (sdone[ee]) ch.set(k, sreset[jj % NCELL]);
                    else if constexpr (K == 16) ch.w[k >> 2] |= (uint32_t)in[ee * STACK_ELEMS + jj + NCELL] << (8 * (k & 3));
                    else ch.f[k] = in[ee * STACK_ELEMS + jj + NCELL];
                }
                if (++jj == STACK_ELEMS) { jj = 0; ee++; }
            }
        }
        if (q0 + K <= total) ch.store(out + q0);
        else
            for (int k = 0; q0 + k < total; k++) out[q0 + k] = ch.get(k);
    }
    if (p_out && threadIdx.x < cnt * 10) {  // data_env: (y, x) rows; reset position (15, 3)
        const int e = threadIdx.x / 10, r = threadIdx.x - 10 * e, f = r >> 1, comp = r & 1;
        float v;
        if (f == 4) v = comp ? (float)(sa[e] & 0xFFu) : (float)(sa[e] >> 8);
        else if (sdone[e]) v = comp ? 3.0f : 15.0f;
        else v = p_prev[(e0 + e) * 10 + r + 2];
        p_out[(e0 + e) * 10 + r] = v;
    }
}

// ---- uint8 codes, tile form --------------------------------------------------------------------
// The CTA's whole output slice is assembled in shared memory and leaves with ONE TMA bulk store.
// The new frame is produced by "column threads": thread (env, x) reads the 34 bits of grid column x
// once and scatters its 17 codes (stride 17 bytes) -- about 5 instructions per cell instead of the
// ~25 a per-byte decode of the column-major record costs.
__device__ __forceinline__ void write_frame_columns(const uint32_t *sg, const uint32_t *sa, uint8_t *tile, int cnt,
                                                    int env_stride, int frame_off) {
    for (int i = threadIdx.x; i < cnt * GS; i += FEAT_THREADS) {
        const int e = i / GS, x = i - GS * e;
        const uint32_t *rec = sg + e * REC_WORDS;
        const int bit = 2 * GS * x, w = bit >> 5, sh = bit & 31;        // column x = cells x*17 .. x*17+16
        const uint32_t w1 = w + 1 < REC_WORDS ? rec[w + 1] : 0u, w2 = w + 2 < REC_WORDS ? rec[w + 2] : 0u;
        const uint32_t lo = __funnelshift_r(rec[w], w1, sh);            // rows 0..15
        const uint32_t hi = __funnelshift_r(w1, w2, sh);                // row 16 in its low 2 bits
        uint8_t *dst = tile + e * env_stride + frame_off + x;
#pragma unroll
        for (int y = 0; y < GS; y++) {
            const uint32_t cell = y < 16 ? (lo >> (2 * y)) & 3u : hi & 3u;
            dst[y * GS] = (uint8_t)((0x0210u >> (4 * cell)) & 0xFu);    // empty 0, wall 1, ball 2, goal 0 (0.9 like empty)
        }
        const uint32_t a = sa[e];
        if ((int)(a & 0xFFu) == x && (a >> 8) < (uint32_t)GS) dst[(a >> 8) * GS] = 4;  // the agent's cell (0.3)
    }
}

// nbytes of a shared-memory tile -> global (both 16-byte aligned): one bulk store + a byte tail
__device__ __forceinline__ void tile_out(uint8_t *dst, const uint8_t *tile, int nbytes) {
    fence_proxy_async();
    __syncthreads();
    const int nb16 = nbytes & ~15;
    if (threadIdx.x == 0 && nb16) {
        bulk_s2g(dst, tile, (uint32_t)nb16);
        bulk_commit();
        bulk_wait_read<0>();
    }
    for (int b = nb16 + threadIdx.x; b < nbytes; b += FEAT_THREADS) dst[b] = tile[b];
}

__global__ void __launch_bounds__(FEAT_THREADS) frame_codes_tile_kernel(const uint32_t *grid, const uint4 *sc0, uint8_t *codes,
                                                                       float *matrix, float *place, long long n) {
    __shared__ uint32_t sg[FEAT_ENVS * REC_WORDS];
    __shared__ uint32_t sa[FEAT_ENVS];
    __shared__ __align__(16) uint8_t tile[FEAT_ENVS * NCELL];
    const long long e0 = (long long)blockIdx.x * FEAT_ENVS;
    const int cnt = (int)((n - e0) < FEAT_ENVS ? (n - e0) : FEAT_ENVS);
    stage_records(grid, sc0, e0, cnt, sg, sa);
    __syncthreads();
    if (place && threadIdx.x < cnt * 2) {
        const int e = threadIdx.x >> 1;
        place[e0 * 2 + threadIdx.x] = (threadIdx.x & 1) ? (float)(sa[e] & 0xFFu) : (float)(sa[e] >> 8);
    }
    write_frame_columns(sg, sa, tile, cnt, NCELL, 0);
    __syncthreads();
    const int total = cnt * NCELL;
    if (matrix) {  // the LUT applied: 4 codes -> one 16-byte store
        float *out = matrix + e0 * NCELL;
        for (int q0 = threadIdx.x * 4; q0 < total; q0 += FEAT_THREADS * 4) {
            if (q0 + 4 <= total) {
                const uint32_t c4 = *reinterpret_cast<const uint32_t *>(tile + q0);
                *reinterpret_cast<float4 *>(out + q0) = make_float4(matrix_value(c4 & 0xFFu), matrix_value((c4 >> 8) & 0xFFu),
                                                                    matrix_value((c4 >> 16) & 0xFFu), matrix_value(c4 >> 24));
            } else {
                for (int k = 0; q0 + k < total; k++) out[q0 + k] = matrix_value(tile[q0 + k]);
            }
        }
    }
    if (codes) tile_out(codes + e0 * NCELL, tile, total);
}

// does the 16-byte chunk at byte q0 of a CTA's slice of [env][5][289] touch frames 0..3 of some env?  (Chunks that lie
// wholly inside a new frame -- j in [1156, 1429] -- are produced by the column threads alone and never loaded.)
__device__ __forceinline__ bool push_chunk_needed(int q0) {
    const int j = q0 - STACK_ELEMS * (q0 / STACK_ELEMS);
    return j < SHIFT_ELEMS || j + 16 > STACK_ELEMS;
}

// MINB = CTAs per SM the register allocation is bounded for (5: no bound hit, 48 registers; 6: 40; 8: 32 with a few spills)
template <int MINB>
__global__ void __launch_bounds__(FEAT_THREADS, MINB) stack_push_codes_tile_kernel(const uint32_t *grid, const uint4 *sc0,
                                                                            const uint8_t *s_prev, uint8_t *s_out,
                                                                            const float *p_prev, float *p_out,
                                                                            const uint8_t *prev_done, int init_flags, long long n) {
    __shared__ uint32_t sg[FEAT_ENVS * REC_WORDS];
    __shared__ uint32_t sa[FEAT_ENVS];
    __shared__ uint32_t sdone_bits;
    __shared__ __align__(16) uint8_t tile[FEAT_ENVS * STACK_ELEMS];
    const long long e0 = (long long)blockIdx.x * FEAT_ENVS;
    const int cnt = (int)((n - e0) < FEAT_ENVS ? (n - e0) : FEAT_ENVS);
    const int total = cnt * STACK_ELEMS;
    // bit 0: every env starts from the reset frame; bits 8..: ablation switches of the measurement probe (TA_PUSH_DBG;
    // 1 = no stack loads, 2 = no bulk store, 4 = no column decode) -- results are only defined with them clear
    const int init_all = init_flags & 1, dbg = init_flags >> 8;
    // Every global load of the CTA is ISSUED before the first shared-memory store, so the CTA pays one DRAM round trip:
    // the records (two words per thread), the agent position / previous done flag (threads < cnt), the position stack
    // and the six 16-byte chunks of the previous frame stacks.  (Staging the records first, as the other feature
    // kernels do, put three dependent round trips in front of the stack loads -- LDG -> STS pairs in program order.)
    constexpr int NREC = (FEAT_ENVS * REC_WORDS + FEAT_THREADS - 1) / FEAT_THREADS;   // 2
    uint32_t recw[NREC];
#pragma unroll
    for (int k = 0; k < NREC; k++) {
        const int i = k * FEAT_THREADS + threadIdx.x;
        recw[k] = i < cnt * REC_WORDS ? __ldg(grid + e0 * REC_WORDS + i) : 0u;
    }
    uint32_t my_xy = 0u, my_done = (uint32_t)(init_all != 0);
    if (threadIdx.x < cnt) {
        my_xy = __ldg(reinterpret_cast<const uint32_t *>(sc0 + e0 + threadIdx.x)) & 0xFFFFu;  // x | y << 8
        if (!init_all && prev_done) my_done = __ldg(prev_done + e0 + threadIdx.x);
    }
    const bool pact = p_out && threadIdx.x < cnt * 10;
    float pprev = 0.0f;   // data_env rows 1..4 of the previous position stack (unused for frame 4 / restarted envs)
    if (pact && !init_all && (threadIdx.x % 10) < 8) pprev = __ldg(p_prev + e0 * 10 + threadIdx.x + 2);
    // 1. every byte of frames 0..3 of the slice <- the byte one frame (289) further on in the previous stacks: aligned
    //    16-byte loads at +288 plus the 17th byte (the first byte of the next lane's chunk -- one shuffle instead of a
    //    second load).  Chunks that lie wholly inside a new frame (18 of an env's 90) are not loaded at all; chunks of
    //    restarted envs receive garbage here and are overwritten in step 2.
    constexpr int NIT = (FEAT_ENVS * STACK_ELEMS + FEAT_THREADS * 16 - 1) / (FEAT_THREADS * 16);   // 6
    uint4 av[NIT];
    uint32_t lastv[NIT];
    const uint8_t *in = s_prev + e0 * STACK_ELEMS;
    const long long in_bytes = (n - e0) * STACK_ELEMS;  // bytes of s_prev from `in` to its end
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int it = 0; it < NIT; it++) { av[it] = make_uint4(0u, 0u, 0u, 0u); lastv[it] = 0u; }
    if (!init_all && !(dbg & 1)) {
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int q0 = (it * FEAT_THREADS + threadIdx.x) * 16;
            av[it] = make_uint4(0u, 0u, 0u, 0u);
            lastv[it] = 0u;
            const bool full = push_chunk_needed(q0) && q0 < total && (long long)q0 + NCELL + 16 <= in_bytes;
            const bool nfull = push_chunk_needed(q0 + 16) && q0 + 16 < total && (long long)q0 + NCELL + 32 <= in_bytes;
            if (full) av[it] = __ldg(reinterpret_cast<const uint4 *>(in + q0 + NCELL - 1));
            if (full && (lane == 31 || !nfull)) lastv[it] = __ldg(in + q0 + NCELL + 15);
        }
    }
    // which envs restart: a 16-bit mask for the whole CTA (the flags sit in the first 16 lanes of warp 0)
    if (threadIdx.x < 32) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, threadIdx.x < cnt && my_done != 0);
        if (threadIdx.x == 0) sdone_bits = b;
    }
#pragma unroll
    for (int k = 0; k < NREC; k++) {
        const int i = k * FEAT_THREADS + threadIdx.x;
        if (i < cnt * REC_WORDS) sg[i] = recw[k];
    }
    if (threadIdx.x < cnt) sa[threadIdx.x] = my_xy;
    if (!init_all) {
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int q0 = (it * FEAT_THREADS + threadIdx.x) * 16;
            const bool need = push_chunk_needed(q0);
            const bool full = need && q0 < total && (long long)q0 + NCELL + 16 <= in_bytes;
            const bool nfull = push_chunk_needed(q0 + 16) && q0 + 16 < total && (long long)q0 + NCELL + 32 <= in_bytes;
            // the next lane's first loaded byte (its chunk starts 16 bytes further on); lane 31 and lanes whose neighbour
            // loaded nothing fetched the byte themselves above
            const uint32_t nb = __shfl_down_sync(0xFFFFFFFFu, av[it].x, 1) & 0xFFu;
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (full) {
                const uint32_t last = (lane == 31 || !nfull) ? lastv[it] : nb;
                const uint4 a = av[it];
                o = make_uint4(__funnelshift_r(a.x, a.y, 8), __funnelshift_r(a.y, a.z, 8), __funnelshift_r(a.z, a.w, 8),
                               (a.w >> 8) | (last << 24));
            } else if (need && q0 < total) {  // the last chunks of the last env: stay inside the array
                uint32_t w[4] = {0, 0, 0, 0};
                for (int k = 0; k < 16; k++)
                    if ((long long)q0 + NCELL + k < in_bytes) w[k >> 2] |= (uint32_t)in[q0 + NCELL + k] << (8 * (k & 3));
                o = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (need && q0 < total && q0 + 16 <= FEAT_ENVS * STACK_ELEMS) *reinterpret_cast<uint4 *>(tile + q0) = o;
        }
    }
    __syncthreads();
    // 2. the new frame of every env; the tiled reset frame for envs whose episode has just restarted
    if (!(dbg & 4)) write_frame_columns(sg, sa, tile, cnt, STACK_ELEMS, SHIFT_ELEMS);
    const uint32_t done_bits = sdone_bits;
    for (uint32_t m = done_bits; m; m &= m - 1) {   // (_gen_grid is a formula: no table, no memory traffic)
        const int e = __ffs(m) - 1;
        for (int j = threadIdx.x; j < SHIFT_ELEMS; j += FEAT_THREADS) tile[e * STACK_ELEMS + j] = (uint8_t)reset_frame_code(j % NCELL);
    }
    // 3. out (in place is fine: every read of s_prev / p_prev that matters happened before the
    //    barrier inside tile_out; reads beyond this CTA's slice only feed positions overwritten in 2)
    float pv = 0.0f;
    if (pact) {  // data_env: (y, x) rows; reset position (15, 3)
        const int e = threadIdx.x / 10, r = threadIdx.x - 10 * e, f = r >> 1, comp = r & 1;
        if (f == 4) pv = comp ? (float)(sa[e] & 0xFFu) : (float)(sa[e] >> 8);
        else if ((done_bits >> e) & 1u) pv = comp ? 3.0f : 15.0f;
        else pv = pprev;
    }
    if (dbg & 2) { __syncthreads(); if (tile[threadIdx.x] == 0xEE) s_out[e0] = 1; }
    else tile_out(s_out + e0 * STACK_ELEMS, tile, total);
    if (pact) p_out[e0 * 10 + threadIdx.x] = pv;
}

}  // namespace ta
