// ta_push_tma.cuh -- the uint8 frame-stack push (ta_feat.cuh: stack_push_codes_tile_kernel) as a persistent,
// software-pipelined kernel: bulk-copy loads (TMA, cp.async.bulk) several tiles ahead, shift + decode in shared memory,
// one bulk store per tile.
//
//   out[env] = [ prev[env] frames 1..4 | current frame ]      np.delete(x,0,0); np.append(x,[new],0)
//   (soa/train_ppo.py:116-121), or the tiled reset frame for envs whose episode has just restarted
//   (Env_transact.reset, soa/env_buffer.py:420-423)
//
// A tile is 16 envs = 23120 bytes of [env][5][289] (16-byte aligned in the array).  Frames 0..3 of the output are the
// input bytes 289 = 18 x 16 + 1 further on.  The register kernel keeps 6 x LDG.128 per thread in flight (48 registers,
// 5 CTAs per SM) and pays a DRAM round trip per CTA for the records before anything else: 43.6 us per 65536 envs, of
// which 21 us remain when loads, decode and store are all switched off (TA_PUSH_DBG).  Here a CTA stays resident and
// walks its tiles through a ring of PT_STAGES shared-memory stages:
//     stage = [tile 23168 | records 16 x 80 | sc0 16 x 16 | done 16]
//   * warp 0, lane 0 fills a stage with FOUR bulk copies on one mbarrier: the 22832 bytes from byte 288 of the tile's slice
//     (16-byte aligned; the slice's frames 1..4 of every env plus what lies between), the records, the agent words and
//     the restart flags -- no register holds stack data, and the loads of the next tiles are in flight while this one
//     is processed;
//   * all 288 threads (= 16 envs x 17 grid columns + 16) shift the stage IN PLACE by the remaining byte (5 chunks per
//     thread: LDS.128 + the next lane's first byte, barrier, STS.128), decode the new frames as one column item per
//     thread and overwrite restarted envs with the reset frame;
//   * thread 0 sends the stage out with one bulk store; the stage is refilled once that store has read it.
// Measured on the way (B200, 65536 envs): letting the TMA do the odd byte of the shift through a 1-D uint8 tensor map
// (box loads start at an ELEMENT coordinate) raises "illegal instruction" -- the box's global address must be 16-byte
// aligned like any bulk copy; 90 aligned 256-byte boxes per tile run at 44.8 us (the copy engine's box rate, ~30 cycles
// per box, is the limit), one 22832-byte bulk copy per tile at 34 us.
#pragma once
#include "ta_feat.cuh"

namespace ta {

constexpr int PT_THREADS = 288;
constexpr int PT_TILE = FEAT_ENVS * STACK_ELEMS;                      // 23120
constexpr int PT_LOAD = PT_TILE - (NCELL - 1);                        // 22832 bytes from byte 288 of the slice to its end
constexpr int PT_CHUNKS_PER_THREAD = 5;                               // 288 x 5 = 1440 chunks; the slice's last 5 lie in a new frame
constexpr int PT_TILE_PAD = (PT_TILE + 127) / 128 * 128;              // 23168
constexpr int PT_REC_BYTES = FEAT_ENVS * REC_WORDS * 4;               // 1280
constexpr int PT_SC_BYTES = FEAT_ENVS * 16;                           // 256
constexpr int PT_STAGE = PT_TILE_PAD + PT_REC_BYTES + PT_SC_BYTES + 128;   // + done flags (16 used)
static_assert(PT_LOAD % 16 == 0 && PT_THREADS * PT_CHUNKS_PER_THREAD * 16 + 16 <= PT_TILE_PAD, "chunk reads stay inside the stage");
static_assert(PT_THREADS * PT_CHUNKS_PER_THREAD * 16 >= (FEAT_ENVS - 1) * STACK_ELEMS + SHIFT_ELEMS, "every byte of frames 0..3 is shifted");
static_assert(FEAT_ENVS * GS <= PT_THREADS, "one column item per thread");

__device__ __forceinline__ void pt_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pt_mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a wrong byte count must not turn into a hung GPU
__device__ __forceinline__ bool pt_mbar_wait(uint64_t *bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); spin++) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void pt_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the new frame of the tile's 16 envs, one (env, grid column) per thread: thread reads the 34 bits of column x once and
// scatters its 17 codes (ta_feat.cuh: write_frame_columns, here without the loop)
__device__ __forceinline__ void pt_frame_columns(const uint32_t *sg, const uint32_t *sxy, uint8_t *tile) {
    const int i = threadIdx.x;
    if (i >= FEAT_ENVS * GS) return;
    const int e = i / GS, x = i - GS * e;
    const uint32_t *rec = sg + e * REC_WORDS;
    const int bit = 2 * GS * x, w = bit >> 5, sh = bit & 31;
    const uint32_t w1 = w + 1 < REC_WORDS ? rec[w + 1] : 0u, w2 = w + 2 < REC_WORDS ? rec[w + 2] : 0u;
    const uint32_t lo = __funnelshift_r(rec[w], w1, sh);            // rows 0..15
    const uint32_t hi = __funnelshift_r(w1, w2, sh);                // row 16 in its low 2 bits
    uint8_t *dst = tile + e * STACK_ELEMS + SHIFT_ELEMS + x;
#pragma unroll
    for (int y = 0; y < GS; y++) {
        const uint32_t cell = y < 16 ? (lo >> (2 * y)) & 3u : hi & 3u;
        dst[y * GS] = (uint8_t)((0x0210u >> (4 * cell)) & 0xFu);    // empty 0, wall 1, ball 2, goal 0 (0.9 like empty)
    }
    const uint32_t a = sxy[e];
    if ((int)(a & 0xFFu) == x && (a >> 8) < (uint32_t)GS) dst[(a >> 8) * GS] = 4;  // the agent's cell (0.3)
}

// n % 16 == 0, s_prev / s_out / prev_done 16-byte aligned (the launcher checks; otherwise the register
// kernel runs).  prev_done nullable.  In place (s_out == s_prev, p_out == p_prev) is fine: every byte a tile needs lies in
// the tile's own slice, which only its own store overwrites, after its own loads have completed.
template <int STAGES>
__global__ void __launch_bounds__(PT_THREADS) stack_push_tma_kernel(const uint32_t *grid, const uint4 *sc0, const uint8_t *s_prev,
                                                                    uint8_t *s_out, const float *p_prev, float *p_out,
                                                                    const uint8_t *prev_done, long long n, int *fail) {
    extern __shared__ __align__(128) uint8_t pt_smem[];
    __shared__ uint64_t full[STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long ntiles = n / FEAT_ENVS;
    const long long my_tiles = ((long long)blockIdx.x < ntiles) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t tx_bytes = PT_LOAD + PT_REC_BYTES + PT_SC_BYTES + (prev_done ? 16 : 0);
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) pt_mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // warp 0 fills stage (k % STAGES) with the CTA's k-th tile
    auto issue = [&](long long k) {
        const int s = (int)(k % STAGES);
        const long long e0 = ((long long)blockIdx.x + k * gridDim.x) * FEAT_ENVS;
        uint8_t *st = pt_smem + (size_t)s * PT_STAGE;
        if (lane == 0) {
            pt_mbar_expect(&full[s], tx_bytes);
            pt_bulk_g2s(st, s_prev + e0 * STACK_ELEMS + (NCELL - 1), PT_LOAD, &full[s]);
            pt_bulk_g2s(st + PT_TILE_PAD, grid + e0 * REC_WORDS, PT_REC_BYTES, &full[s]);
            pt_bulk_g2s(st + PT_TILE_PAD + PT_REC_BYTES, sc0 + e0, PT_SC_BYTES, &full[s]);
            if (prev_done) pt_bulk_g2s(st + PT_TILE_PAD + PT_REC_BYTES + PT_SC_BYTES, prev_done + e0, 16, &full[s]);
        }
        __syncwarp();
    };
    if (warp == 0)
        for (long long k = 0; k < STAGES - 1 && k < my_tiles; k++) issue(k);
    // which of this thread's chunks (16 * (tid + 288 i)) touch frames 0..3 of some env: the same for every tile
    uint32_t need_mask = 0;
#pragma unroll
    for (int i = 0; i < PT_CHUNKS_PER_THREAD; i++)
        if (push_chunk_needed((tid + PT_THREADS * i) * 16)) need_mask |= 1u << i;
    bool dead = false;
    for (long long k = 0; k < my_tiles; k++) {
        const int s = (int)(k % STAGES);
        const long long e0 = ((long long)blockIdx.x + k * gridDim.x) * FEAT_ENVS;
        uint8_t *st = pt_smem + (size_t)s * PT_STAGE;
        uint8_t *tile = st;
        const uint32_t *sg = reinterpret_cast<const uint32_t *>(st + PT_TILE_PAD);
        const uint4 *ssc = reinterpret_cast<const uint4 *>(st + PT_TILE_PAD + PT_REC_BYTES);
        const uint8_t *sdn = st + PT_TILE_PAD + PT_REC_BYTES + PT_SC_BYTES;
        // the position stack of this tile (160 floats): issued now, used after the decode
        const bool pact = p_out && tid < FEAT_ENVS * 10;
        float pprev = 0.0f;
        if (pact && (tid % 10) < 8) pprev = __ldg(p_prev + e0 * 10 + tid + 2);
        if (!dead && !pt_mbar_wait(&full[s], (uint32_t)((k / STAGES) & 1))) dead = true;
        // agent positions (x | y << 8) and the restart mask of the 16 envs, from the stage
        __shared__ uint32_t sxy[FEAT_ENVS];
        __shared__ uint32_t sdone_bits;
        if (warp == 1) {
            const bool dn = prev_done && lane < FEAT_ENVS && sdn[lane] != 0;
            const uint32_t bts = __ballot_sync(0xFFFFFFFFu, dn);
            if (lane == 0) sdone_bits = bts;
            if (lane < FEAT_ENVS) sxy[lane] = ssc[lane].x & 0xFFFFu;
        }
        // the stage holds in[288 + j] at byte j; frames 0..3 want in[289 + j]: shift by one byte, in place (all reads, barrier,
        // all writes).  A chunk's 17th byte is the next lane's first.
        uint4 av[PT_CHUNKS_PER_THREAD];
        uint32_t nb[PT_CHUNKS_PER_THREAD];
#pragma unroll
        for (int i = 0; i < PT_CHUNKS_PER_THREAD; i++) {
            const int q0 = (tid + PT_THREADS * i) * 16;
            av[i] = *reinterpret_cast<const uint4 *>(tile + q0);
            nb[i] = __shfl_down_sync(0xFFFFFFFFu, av[i].x, 1) & 0xFFu;
            if (lane == 31) nb[i] = tile[q0 + 16];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PT_CHUNKS_PER_THREAD; i++) {
            const int q0 = (tid + PT_THREADS * i) * 16;
            const uint4 a = av[i];
            if ((need_mask >> i) & 1u)
                *reinterpret_cast<uint4 *>(tile + q0) = make_uint4(__funnelshift_r(a.x, a.y, 8), __funnelshift_r(a.y, a.z, 8),
                                                                    __funnelshift_r(a.z, a.w, 8), (a.w >> 8) | (nb[i] << 24));
        }
        __syncthreads();
        const uint32_t done_bits = sdone_bits;
        pt_frame_columns(sg, sxy, tile);
        for (uint32_t m = done_bits; m; m &= m - 1) {   // restarted envs: the tiled reset frame (_gen_grid is a formula)
            const int e = __ffs(m) - 1;
            for (int j = tid; j < SHIFT_ELEMS; j += PT_THREADS) tile[e * STACK_ELEMS + j] = (uint8_t)reset_frame_code(j % NCELL);
        }
        float pv = 0.0f;
        if (pact) {  // data_env: (y, x) rows; reset position (15, 3)
            const int e = tid / 10, r = tid - 10 * e, f = r >> 1, comp = r & 1;
            if (f == 4) pv = comp ? (float)(sxy[e] & 0xFFu) : (float)(sxy[e] >> 8);
            else if ((done_bits >> e) & 1u) pv = comp ? 3.0f : 15.0f;
            else pv = pprev;
        }
        fence_proxy_async();
        __syncthreads();
        if (pact) p_out[e0 * 10 + tid] = pv;
        if (warp == 0) {
            if (lane == 0) {
                bulk_s2g(s_out + e0 * STACK_ELEMS, tile, (uint32_t)PT_TILE);
                bulk_commit();
                bulk_wait_read<1>();   // the store issued one tile ago has read its stage: that stage is free
            }
            __syncwarp();
            if (k + STAGES - 1 < my_tiles) issue(k + STAGES - 1);
        }
    }
    if (tid == 0) bulk_wait_read<0>();
    if (dead && fail) atomicExch(fail, 1);
}

// ---- matrix_env / data_env of every env (ta_feat.cuh: frame_codes_tile_kernel) in the same pipelined form -----------
// Env_transact.matrix_env / data_env, soa/env_buffer.py:300-334.  The input of a 16-env tile is 1.5 KB (records + agent
// words), its output 4.6 KB of codes (+ 18 KB of float32 LUT values): the one-tile-per-CTA kernel is a DRAM round trip
// followed by a decode per CTA, 15 us per 65536 envs for 25 MB.  Here FP_STAGES input stages are prefetched by bulk copies,
// a tile costs ONE CTA barrier, and the codes leave with a bulk store from a ring of FP_OUT output tiles.
constexpr int FP_STAGES = 8;
constexpr int FP_OUT = 3;
constexpr int FP_IN = PT_REC_BYTES + PT_SC_BYTES;                    // 1536
constexpr int FP_TILE = FEAT_ENVS * NCELL;                           // 4624
constexpr int FP_TILE_PAD = (FP_TILE + 127) / 128 * 128;             // 4736

__global__ void __launch_bounds__(PT_THREADS) frame_codes_pipe_kernel(const uint32_t *grid, const uint4 *sc0, uint8_t *codes,
                                                                      float *matrix, float *place, long long n, int *fail) {
    __shared__ __align__(128) uint8_t s_in[FP_STAGES * FP_IN];
    __shared__ __align__(128) uint8_t s_out[FP_OUT * FP_TILE_PAD];
    __shared__ uint64_t full[FP_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long ntiles = n / FEAT_ENVS;
    const long long my_tiles = ((long long)blockIdx.x < ntiles) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < FP_STAGES; s++) pt_mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long k) {   // one thread: stage k % FP_STAGES <- the CTA's k-th tile
        const int s = (int)(k % FP_STAGES);
        const long long e0 = ((long long)blockIdx.x + k * gridDim.x) * FEAT_ENVS;
        pt_mbar_expect(&full[s], FP_IN);
        pt_bulk_g2s(s_in + s * FP_IN, grid + e0 * REC_WORDS, PT_REC_BYTES, &full[s]);
        pt_bulk_g2s(s_in + s * FP_IN + PT_REC_BYTES, sc0 + e0, PT_SC_BYTES, &full[s]);
    };
    if (tid == 0)
        for (long long k = 0; k < FP_STAGES && k < my_tiles; k++) issue(k);
    bool dead = false;
    for (long long k = 0; k < my_tiles; k++) {
        const int s = (int)(k % FP_STAGES), o = (int)(k % FP_OUT);
        const long long e0 = ((long long)blockIdx.x + k * gridDim.x) * FEAT_ENVS;
        const uint32_t *sg = reinterpret_cast<const uint32_t *>(s_in + s * FP_IN);
        const uint4 *ssc = reinterpret_cast<const uint4 *>(s_in + s * FP_IN + PT_REC_BYTES);
        uint8_t *tile = s_out + o * FP_TILE_PAD;
        if (!dead && !pt_mbar_wait(&full[s], (uint32_t)((k / FP_STAGES) & 1))) dead = true;
        // (tile `o` is free: thread 0 waited for the store issued FP_OUT tiles ago before the previous tile's barrier)
        if (tid < FEAT_ENVS * GS) {   // one (env, grid column) per thread
            const int e = tid / GS, x = tid - GS * e;
            const uint32_t *rec = sg + e * REC_WORDS;
            const int bit = 2 * GS * x, w = bit >> 5, sh = bit & 31;
            const uint32_t w1 = w + 1 < REC_WORDS ? rec[w + 1] : 0u, w2 = w + 2 < REC_WORDS ? rec[w + 2] : 0u;
            const uint32_t lo = __funnelshift_r(rec[w], w1, sh), hi = __funnelshift_r(w1, w2, sh);
            uint8_t *dst = tile + e * NCELL + x;
#pragma unroll
            for (int y = 0; y < GS; y++) {
                const uint32_t cell = y < 16 ? (lo >> (2 * y)) & 3u : hi & 3u;
                dst[y * GS] = (uint8_t)((0x0210u >> (4 * cell)) & 0xFu);
            }
            const uint32_t a = ssc[e].x & 0xFFFFu;
            if ((int)(a & 0xFFu) == x && (a >> 8) < (uint32_t)GS) dst[(a >> 8) * GS] = 4;
        }
        if (place && tid < FEAT_ENVS * 2) {   // data_env: (y, x)
            const uint32_t a = ssc[tid >> 1].x & 0xFFFFu;
            place[e0 * 2 + tid] = (tid & 1) ? (float)(a & 0xFFu) : (float)(a >> 8);
        }
        if (tid == 0 && codes) bulk_wait_read<FP_OUT - 2>();   // frees the tile the NEXT iteration writes
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (codes) {
                bulk_s2g(codes + e0 * NCELL, tile, (uint32_t)FP_TILE);
                bulk_commit();
            }
            if (k + FP_STAGES < my_tiles) issue(k + FP_STAGES);
        }
        if (matrix) {   // the LUT applied: 4 codes -> one 16-byte store
            float4 *out = reinterpret_cast<float4 *>(matrix + e0 * NCELL);
            for (int q = tid; q < FP_TILE / 4; q += PT_THREADS) {
                const uint32_t c4 = *reinterpret_cast<const uint32_t *>(tile + 4 * q);
                out[q] = make_float4(matrix_value(c4 & 0xFFu), matrix_value((c4 >> 8) & 0xFFu), matrix_value((c4 >> 16) & 0xFFu),
                                     matrix_value(c4 >> 24));
            }
        }
    }
    if (tid == 0 && codes) bulk_wait_read<0>();
    if (dead && fail) atomicExch(fail, 1);
}

}  // namespace ta
