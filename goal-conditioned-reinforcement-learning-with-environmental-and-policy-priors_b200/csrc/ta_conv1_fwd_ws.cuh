// ta_conv1_fwd_ws.cuh -- TINet's fused first layer (ta_conv1_tc.cuh: LUT decode + UpsamplingNearest2d(4) + Conv2d(4,64,4,2)
// + bias + ReLU as D[position, (phase, channel)] = P[position, 16] x W4^T[16, 256]) as a WARP-SPECIALISED persistent kernel
// whose output leaves through TMA tensor-map stores.  Default forward kernel (TA_CONV1_TC=1 selects conv1_fwd_tc_kernel, which
// runs stage -> MMA -> epilogue serially inside 128-thread CTAs and scatters a position's four phases in 64-byte pieces); the
// two are bit-identical (same arithmetic: bf16 hi / lo operands, fp32 accumulation, bias add, cvt.rn.relu.bf16x2, the mask
// bits from the carry trick).  all_net.py:142-143,157,180-181.
//   * a tile is SEVEN ROWS of the 17-wide position grid (119 positions; rows run on across samples: linear position index
//     = (sample * 17 + m) * 17 + n), so that for a fixed output-row parity py the tile's output is seven whole image rows
//     y = 2m + py of 33 pixels x 128 bytes = 4224 contiguous bytes each;
//   * the epilogue writes the pixel of its phase (py, px) into row 2n + px of that image row's staging block in the TMA's
//     128-byte swizzle, and ONE tensor-map store per image row (box 64 channels x 33 pixels x 1 row of the [B * 33][33][64]
//     view of y) moves it: 14 stores of 4 KB per tile instead of 3808 scattered 64-byte pieces;
//   * all four phases come out of ONE N = 256 MMA per operand pair (3 MMAs per tile: hi*hi, lo*hi, hi*lo) into 256 TMEM
//     columns, double-buffered (512).
//   roles: SEVEN WARP GROUPS (896 threads) -- 0, 3, 5, 6: epilogue of phase (0,0), (1,0), (0,1), (1,1), one TMEM lane quadrant
//   per warp; 1, 4: decoders of this CTA's even / odd tiles (the tile's P operand, thread = position); 2: the MMA issuer
//   (warp 8, converged, elect.sync; its three sister warps idle).  The 896 threads start with 72 registers each; the decoder
//   and MMA groups give registers up (setmaxnreg.dec 48) and the four epilogue groups take them (setmaxnreg.inc 88) -- sixteen
//   epilogue warps do not fit otherwise (spills: 370-480 us).  The two warp groups of a row parity py share the staging of
//   its image rows (3 buffers x 7 blocks x 4352 B) and meet only inside the pair (named barriers of 256 threads); the roles meet
//   through mbarriers.  Every mbarrier wait is bounded and raises `fail`.
//
// Measured per 4096 samples (scripts/probe_conv1_fwd.py; ablation bits in `dbg`, TA_FW_DBG; conv1_fwd_tc_kernel: 148 us
// without / 164 us with the ReLU bit mask):
//   whole kernel                                     134 us without the mask, 149 us with it (three staging buffers; 135 / 156 us with two)
//   stores only (no epilogue arithmetic)              93 us = 6.1 TB/s: the store pattern reaches the write roofline
//                                                     (the same 14 x 4224-byte stores per tile as a bare probe: 6.1 TB/s, write_bw.py)
//   epilogue arithmetic only (no stores)              95 us without the mask, 115 us with it (64 extra integer ops per 32 channels)
//   neither (decode -> MMA -> TMEM read -> barriers)  66 us
// History: with eight epilogue warps (two phases each) the epilogue's dependent instruction stream was the bottleneck (150 /
// 208 us); with ONE decoder group and 64-bit index arithmetic the decoders paced everything at 116 us -- a role that is one
// warp per scheduler is a serial instruction stream and its instruction count is its time.  Stores and arithmetic still do not
// overlap fully (134 vs max(93, 95)): the staging writes conflict two ways in their banks (rows 2n + px of consecutive lanes
// have the same parity, so the swizzle only spreads them over four of the eight 16-byte bank groups) and share the
// shared-memory port with the TMA's reads and the MMA operands.  A conflict-free staging order was measured too -- one block per
// (image row, px) with lane = row, stored through a tensor map that walks x with element stride 2 (28 stores of 2 KB per tile,
// bit-identical): 143 / 157 us, the strided stores cost more than the conflicts.
#pragma once
#include <cuda.h>

#include "ta_conv1_tc.cuh"
#include "ta_dgrad_tc.cuh"

namespace ta {

constexpr int FW_THREADS = 896;   // seven warp groups: 0 epi (py0,px0), 1 dec A, 2 MMA (+3 idle warps), 3 epi (py1,px0), 4 dec B, 5 epi (py0,px1), 6 epi (py1,px1)
constexpr int FW_REGS_LOW = 48, FW_REGS_EPI = 88;   // setmaxnreg: 896 threads start with 72 registers each; 12 warps give 24 up, 16 warps take 16
constexpr int FW_ROWS = 7, FW_POS = FW_ROWS * GS;             // 119 positions per tile
constexpr int FW_SLOTS = FW_POS + TC_HALO;                    // 137 decoded positions per tile
constexpr int FW_NBUF = 3;                                    // staging buffers per epilogue group (stores of FW_NBUF - 1 tiles may be in flight)
constexpr int FW_BLOCK = 34 * 128;                            // staging of one image row: 33 x 128 B (+ one spare row); block k starts at
                                                              // row 34 k of the 1024-aligned staging area, so the TMA's address-based
                                                              // 128-byte swizzle XORs row i of block k with (34 k + i) % 8 = (2 k + i) % 8
constexpr int FW_OFF_B = 0;                                   // W4 hi / lo: 2 x 8 KB
constexpr int FW_OFF_A = FW_OFF_B + 2 * TC_N * 32;            // A hi / lo, 2 stages: 4 x 4 KB
constexpr int FW_OFF_STAGE = FW_OFF_A + 4 * TC_M * 32;        // 32768: [group 2][buffer 2][row 7] blocks
constexpr int FW_OFF_DEC = (FW_OFF_STAGE + 2 * FW_NBUF * FW_ROWS * FW_BLOCK + 15) & ~15;   // 215552
constexpr int FW_OFF_BIAS = FW_OFF_DEC + 2 * FW_SLOTS * 16;   // 180512
constexpr int FW_SMEM = FW_OFF_BIAS + TC_N * 4;               // 181536
constexpr uint32_t FW_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);   // N = 256

// raw values of linear position Q (4 frames), as tc_load_slot but with the position given explicitly and in 32-bit arithmetic
// (positions < 2^31, checked by the host: the decoders are a single dependent instruction stream per scheduler, so their
// instruction count is their time)
template <typename XT>
__device__ __forceinline__ auto fw_load(const XT *__restrict__ x, long long xstride, unsigned npos, unsigned Q, bool in_tile) {
    const bool in = in_tile && Q < npos;
    const unsigned qb = in ? Q / (unsigned)NCELL : 0u;
    const XT *xq = x + (long long)qb * xstride + (in ? Q - qb * (unsigned)NCELL : 0u);
    if constexpr (sizeof(XT) == 1) {
        uint32_t w = 0;
        if (in) w = (uint32_t)xq[0] | ((uint32_t)xq[NCELL] << 8) | ((uint32_t)xq[2 * NCELL] << 16) | ((uint32_t)xq[3 * NCELL] << 24);
        return w;
    } else {
        return in ? make_float4(xq[0], xq[NCELL], xq[2 * NCELL], xq[3 * NCELL]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <typename XT, typename RT>
__device__ __forceinline__ uint4 fw_decode(const RT &r, bool exists) {
    uint32_t hi0 = 0, hi1 = 0, lo0 = 0, lo1 = 0;
    if (exists) {
        float f0, f1, f2, f3;
        if constexpr (sizeof(XT) == 1) {
            f0 = c1_decode(r & 0xFFu); f1 = c1_decode((r >> 8) & 0xFFu); f2 = c1_decode((r >> 16) & 0xFFu); f3 = c1_decode(r >> 24);
        } else {
            f0 = r.x; f1 = r.y; f2 = r.z; f3 = r.w;
        }
        tc_split(f0, f1, hi0, lo0);
        tc_split(f2, f3, hi1, lo1);
    }
    return make_uint4(hi0, hi1, lo0, lo1);
}

template <typename XT, bool MK>
__global__ void __launch_bounds__(FW_THREADS, 1) conv1_fwd_ws_kernel(const XT *__restrict__ x, long long xstride, const float *__restrict__ w4,
                                                                     const float *__restrict__ b4, long long B,
                                                                     const __grid_constant__ CUtensorMap y_map, uint32_t *__restrict__ relu_mask,
                                                                     int *fail, int dbg) {
    // dbg (development, TA_FW_DBG): bit 0 no tensor stores, bit 1 no epilogue arithmetic / staging writes, bit 2 no mask stores
    extern __shared__ __align__(1024) uint8_t fw_smem[];
    uint8_t *sB = fw_smem + FW_OFF_B, *sA = fw_smem + FW_OFF_A, *sStage = fw_smem + FW_OFF_STAGE;
    uint4 *sDec = reinterpret_cast<uint4 *>(fw_smem + FW_OFF_DEC);
    float *sbias = reinterpret_cast<float *>(fw_smem + FW_OFF_BIAS);
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ int decoders_dead, group_dead[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // one-time setup: W4 -> bf16 (hi, lo) in the canonical K-major layout, bias, barriers, TMEM
    for (int i = tid; i < TC_N * 2; i += FW_THREADS) {
        const int n = i >> 1, c = i & 1;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; q++) tc_split(__ldg(w4 + n * 16 + c * 8 + 2 * q), __ldg(w4 + n * 16 + c * 8 + 2 * q + 1), hi[q], lo[q]);
        *reinterpret_cast<uint4 *>(sB + tc_operand_offset(n, c)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(sB + TC_N * 32 + tc_operand_offset(n, c)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int i = tid; i < TC_N; i += FW_THREADS) sbias[i] = __ldg(b4 + i);
    if (tid == 0) {
        decoders_dead = 0; group_dead[0] = 0; group_dead[1] = 0;
        for (int s = 0; s < 2; s++) {
            dg_mbar_init(&a_full[s], 128);     // the decoder threads
            dg_mbar_init(&a_empty[s], 1);      // one tcgen05.commit
            dg_mbar_init(&acc_full[s], 1);     // one tcgen05.commit
            dg_mbar_init(&acc_empty[s], 16);   // the sixteen epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const long long npos = B * NCELL, nrows = B * GS, ntiles = (nrows + FW_ROWS - 1) / FW_ROWS;
    bool dead = false;

    if ((warp >= 4 && warp < 8) || (warp >= 16 && warp < 20)) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FW_REGS_LOW));
        // ---------------- decoders: the tile's A operand (hi, lo), thread = position ----------------
        // TWO groups of 128 threads (warps 4-7: this CTA's even tiles = A stage 0, warps 16-19: the odd ones = stage 1), each with
        // its own decode scratch and named barrier.  A decoder warp is ONE dependent instruction stream on its scheduler: one
        // group with 64-bit index arithmetic needed 3.3 k cycles per tile and paced the whole kernel (the rest of the pipeline
        // runs a tile in 1.3 k without it).
        // (a thread whose wait gave up keeps going without waiting: it must reach the named barrier of its next tile, where all
        // 128 leave together)
        const int dgrp = warp >= 16 ? 1 : 0;
        const int ptid = tid - (dgrp ? 512 : 128);
        const int bar_id = dgrp ? 4 : 1;
        const unsigned np32 = (unsigned)npos;
        uint4 *dec = sDec + dgrp * FW_SLOTS;
        uint8_t *ah = sA + dgrp * 2 * TC_M * 32, *al = ah + TC_M * 32;
        const unsigned n = (unsigned)ptid % (unsigned)GS, rloc = (unsigned)ptid / (unsigned)GS;
        typename std::conditional<sizeof(XT) == 1, uint32_t, float4>::type ra, rb;
        {
            const long long t0 = (long long)blockIdx.x + (long long)dgrp * gridDim.x;
            const bool has = t0 < ntiles;
            const unsigned Q0 = has ? (unsigned)(t0 * FW_POS) : 0u;
            ra = fw_load<XT>(x, xstride, np32, Q0 + ptid, has);
            rb = fw_load<XT>(x, xstride, np32, Q0 + 128 + ptid, has && ptid + 128 < FW_SLOTS);
        }
        int it = dgrp;
        for (long long tile = (long long)blockIdx.x + (long long)dgrp * gridDim.x; tile < ntiles; tile += 2ll * gridDim.x, it += 2) {
            const unsigned Q0 = (unsigned)(tile * FW_POS);
            if (dbg & 8) {   // (ablation: no decode work at all)
                if (!dead && !tc_mbar_wait(&a_empty[dgrp], (uint32_t)(((it >> 1) & 1) ^ 1))) dead = true;
                dg_mbar_arrive(&a_full[dgrp]);
                continue;
            }
            dec[ptid] = fw_decode<XT>(ra, Q0 + ptid < np32);
            if (ptid + 128 < FW_SLOTS) dec[ptid + 128] = fw_decode<XT>(rb, Q0 + 128 + ptid < np32);
            {   // this group's next tile: its raw values are in flight during this one
                const long long tk = tile + 2ll * gridDim.x;
                const bool has = tk < ntiles;
                const unsigned Q1 = has ? (unsigned)(tk * FW_POS) : 0u;
                ra = fw_load<XT>(x, xstride, np32, Q1 + ptid, has);
                rb = fw_load<XT>(x, xstride, np32, Q1 + 128 + ptid, has && ptid + 128 < FW_SLOTS);
            }
            if (dead) decoders_dead = 1;
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // (the scratch is rewritten only after this barrier has been passed again:
            if (decoders_dead) { dead = true; break; }                  //  every thread has then read its neighbours' entries)
            const bool valid = ptid < FW_POS && Q0 + ptid < np32;
            const unsigned m = (((unsigned)tile % (unsigned)GS) * (unsigned)FW_ROWS + rloc) % (unsigned)GS;   // (tile * 7 + rloc) mod 17
            const bool rgt = valid && n < 16, bot = valid && m < 16;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const uint4 d00 = valid ? dec[ptid] : z, d01 = rgt ? dec[ptid + 1] : z;
            const uint4 d10 = bot ? dec[ptid + GS] : z, d11 = (rgt && bot) ? dec[ptid + GS + 1] : z;
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // every thread has read the scratch: the next tile may overwrite it
            if (!dead && !tc_mbar_wait(&a_empty[dgrp], (uint32_t)(((it >> 1) & 1) ^ 1))) dead = true;   // the MMAs of this stage's previous tile are done
            *reinterpret_cast<uint4 *>(ah + tc_operand_offset(ptid, 0)) = make_uint4(d00.x, d00.y, d01.x, d01.y);
            *reinterpret_cast<uint4 *>(ah + tc_operand_offset(ptid, 1)) = make_uint4(d10.x, d10.y, d11.x, d11.y);
            *reinterpret_cast<uint4 *>(al + tc_operand_offset(ptid, 0)) = make_uint4(d00.z, d00.w, d01.z, d01.w);
            *reinterpret_cast<uint4 *>(al + tc_operand_offset(ptid, 1)) = make_uint4(d10.z, d10.w, d11.z, d11.w);
            fence_proxy_async();
            dg_mbar_arrive(&a_full[dgrp]);
        }
    } else if (warp >= 8 && warp < 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FW_REGS_LOW));   // (the whole warp group; warps 9-11 have no other work)
        if (warp == 8) {
        // ---------------- MMA issuer (all 32 lanes converged, one elected lane issues) ----------------
        const uint64_t dBh = tc_smem_desc(sB), dBl = tc_smem_desc(sB + TC_N * 32);
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles && !dead; tile += gridDim.x, it++) {
            const int st = it & 1;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            const bool ok = __all_sync(0xFFFFFFFFu, tc_mbar_wait(&a_full[st], ph) && tc_mbar_wait(&acc_empty[st], ph ^ 1u));
            if (!ok) { dead = true; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t dAh = tc_smem_desc(sA + st * 2 * TC_M * 32), dAl = tc_smem_desc(sA + st * 2 * TC_M * 32 + TC_M * 32);
            const uint32_t d = tmem_base + (uint32_t)(st * TC_N);
            tc_mma_elect(d, dAh, dBh, FW_IDESC, 0u);
            tc_mma_elect(d, dAl, dBh, FW_IDESC, 1u);
            tc_mma_elect(d, dAh, dBl, FW_IDESC, 1u);
            tc_commit_elect(&a_empty[st]);
            tc_commit_elect(&acc_full[st]);
        }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FW_REGS_EPI));
        // ---------------- epilogue: one output phase (py, px) per warp group, one TMEM lane quadrant (warp % 4) per warp --------
        // group g = py (two warp groups = 256 threads) shares the staging of its image rows
        const int wg = warp >> 2;                          // 0: (0,0), 3: (1,0), 5: (0,1), 6: (1,1)
        const int g = (wg == 3 || wg == 6) ? 1 : 0, px = wg >= 5 ? 1 : 0, q = warp & 3;
        const int phase = g * 2 + px;
        const int p = q * 32 + lane;                       // TMEM lane = position of the tile
        const int r = p / GS, n = p - r * GS;              // grid row of the tile, column
        const bool leader = px == 0 && q == 0 && lane == 0;   // issues the group's stores
        const uint64_t map_u = reinterpret_cast<uint64_t>(&y_map);
        const int bar_id = 2 + g;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int buf = it & 1, sbuf = it % FW_NBUF;
            const int blk0 = (g * FW_NBUF + sbuf) * FW_ROWS;   // index of the group's first staging block of this tile
            uint8_t *stage = sStage + blk0 * FW_BLOCK;
            const int sw = (2 * n + px + 2 * (blk0 + r)) & 7;
            const long long P = tile * FW_POS + p;
            const bool valid = p < FW_POS && P < npos;
            if (!dead && !(__all_sync(0xFFFFFFFFu, tc_mbar_wait(&acc_full[buf], (uint32_t)((it >> 1) & 1))))) dead = true;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // the stores issued FW_NBUF tiles ago have read this staging buffer
            if (leader) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(FW_NBUF - 1) : "memory");
            if (dead) group_dead[g] = 1;
            asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
            if (group_dead[g]) { dead = true; break; }
            {
                uint8_t *srow = stage + r * FW_BLOCK + (2 * n + px) * 128;
                const bool wr = valid && !(px && n == 16);   // (pixel x = 33 does not exist)
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t rr[32];
                    if (!(dbg & 16)) {   // (ablation bit 4: no TMEM reads)
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_N + phase * 64 + half * 32);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]),
                              "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15]), "=r"(rr[16]),
                              "=r"(rr[17]), "=r"(rr[18]), "=r"(rr[19]), "=r"(rr[20]), "=r"(rr[21]), "=r"(rr[22]), "=r"(rr[23]), "=r"(rr[24]),
                              "=r"(rr[25]), "=r"(rr[26]), "=r"(rr[27]), "=r"(rr[28]), "=r"(rr[29]), "=r"(rr[30]), "=r"(rr[31])
                            : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                    if (half == 1) {   // this warp's share of the accumulator buffer is in registers
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) dg_mbar_arrive(&acc_empty[buf]);
                    }
                    if (dbg & 2) continue;   // (ablation bit 1: no arithmetic / staging writes)
                    const int c0 = half * 32;
                    uint32_t mbits = 0;   // bit c: channel c0 + c of this output pixel is non-zero (the ReLU mask the backward needs)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float4 b0 = *reinterpret_cast<const float4 *>(sbias + phase * 64 + c0 + 8 * j);
                        const float4 b1 = *reinterpret_cast<const float4 *>(sbias + phase * 64 + c0 + 8 * j + 4);
                        const uint32_t o0 = tc_relu_pack(__uint_as_float(rr[8 * j]) + b0.x, __uint_as_float(rr[8 * j + 1]) + b0.y);
                        const uint32_t o1 = tc_relu_pack(__uint_as_float(rr[8 * j + 2]) + b0.z, __uint_as_float(rr[8 * j + 3]) + b0.w);
                        const uint32_t o2 = tc_relu_pack(__uint_as_float(rr[8 * j + 4]) + b1.x, __uint_as_float(rr[8 * j + 5]) + b1.y);
                        const uint32_t o3 = tc_relu_pack(__uint_as_float(rr[8 * j + 6]) + b1.z, __uint_as_float(rr[8 * j + 7]) + b1.w);
                        // 16-byte chunk half * 4 + j of the pixel's 128 bytes, in the TMA's 128-byte swizzle (chunk ^ row % 8)
                        if (wr) *reinterpret_cast<uint4 *>(srow + (((half * 4 + j) ^ sw) << 4)) = make_uint4(o0, o1, o2, o3);
                        if constexpr (MK) {
                            // a ReLU output half is >= +0, so half + 0x7FFF has bit 15 set exactly when it is non-zero (no carry
                            // into the other half); word 4j+i contributes bits 4j+i and 16+4j+i
                            mbits |= (((o0 + 0x7FFF7FFFu) & 0x80008000u) >> (15 - 4 * j)) | (((o1 + 0x7FFF7FFFu) & 0x80008000u) >> (14 - 4 * j)) |
                                     (((o2 + 0x7FFF7FFFu) & 0x80008000u) >> (13 - 4 * j)) | (((o3 + 0x7FFF7FFFu) & 0x80008000u) >> (12 - 4 * j));
                        }
                    }
                    if constexpr (MK) {
                        if (valid && !(dbg & 4)) relu_mask[(P * 4 + phase) * 2 + half] = mbits;
                    }
                }
            }
            fence_proxy_async();
            asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
            if (leader && !(dbg & 1)) {
                // one store per image row y = 2m + py of the tile's seven grid rows (row 33 of a sample does not exist; rows past
                // the last sample are not stored)
#pragma unroll 1
                for (int rr7 = 0; rr7 < FW_ROWS; rr7++) {
                    const long long R = tile * FW_ROWS + rr7;
                    if (R >= nrows) break;
                    const long long b = R / GS;
                    const int m = (int)(R - b * GS);
                    if (g && m == 16) continue;
                    const int yrow = (int)(b * C1_OUT + 2 * m + g);
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map_u), "r"(0), "r"(0),
                                 "r"(yrow), "r"(smem_u32(stage + rr7 * FW_BLOCK))
                                 : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    if (dead && fail) atomicExch(fail, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

}  // namespace ta
