// ta_stem_bwd_tc.cuh -- conv2's data gradient AND conv1's weight / bias gradient in ONE kernel: the gradient of conv1's
// output (the four parity planes of ta_dgrad_tc.cuh, 605 MB per 4096 samples) goes from TMEM through shared memory
// straight into the second GEMM and never touches HBM.  Replaces conv2_dgrad_planes_ws_kernel + conv1_bwd_tc_kernel
// (all_net.py:142-145 backward).
//
//   GEMM 1 (ta_dgrad_tc.cuh)   D_c[position, ci] = sum_taps A_shift(tap)[position, co] x W_tap[ci, co]^T     per parity class c
//   mask                       dz1 = D_c where conv1's ReLU was active (the forward kernel's bit mask), bf16
//   GEMM 2 (ta_conv1_tc.cuh)   dW4[(c, ci), k] += sum_positions dz1[position, (c, ci)] x P[position, k],  P = the decoded 2x2
//                              input patch (bf16 hi / lo pair) and a ones column for the bias gradient
//
// One persistent CTA of 544 threads per SM, warp-specialised, no CTA-wide barrier in the tile loop:
//   warps 4-7  producers   cp.async of the tile's four shifted dz tiles into a ring of SIX 16 KB slots (a tile's slots are
//                          queued in the order sh3, sh2, sh1, sh0 and released in that order: the first class pair is the
//                          last user of sh3 / sh2, the second of sh1 / sh0)
//   warps 9-12 decoders    the tile's P operand (thread = position: LUT decode of the 4 frames, 2x2 patch gathered through a
//                          double-buffered shared array and one named barrier of these 128 threads); as part of the producer
//                          role it doubled the producers' time per tile and starved the MMA issuer
//   warp 8     MMA issuer  per class PAIR (classes 2h, 2h+1 = 128 rows of dW4): GEMM 1 into one of THREE 128-column TMEM
//                          slots, then GEMM 2 of the PREVIOUS pair (its operand is ready by then) into 2 x 32 resident columns
//   warps 0-3, 13-16  epilogue  tcgen05.ld of one class of the pair (lane = position; the two groups take the even / odd
//                          class), mask, bf16, 16-byte chunks written in the MN-major operand layout of GEMM 2 (32 lanes =
//                          512 contiguous bytes: conflict-free)
// TMEM: 3 x 128 (GEMM 1) + 2 x 64 (dW4, db4: hi and lo operand halves) = all 512 columns.  Shared memory: W 72 KB + ring 96 KB + dz1 operand
// 32 KB + P 16 KB + decode scratch 4.6 KB = 221 KB.  Every mbarrier wait is bounded and raises `fail`.
#pragma once
#include "ta_dgrad_tc.cuh"

namespace ta {

constexpr int SB_THREADS = 544;   // warps 0-3 and 13-16 epilogue, 4-7 producers, 8 MMA issuer, 9-12 decoders
constexpr int SB_SLOTS = 6;
constexpr int SB_OFF_RING = DG_W_BYTES;                          // 73728 (1024-aligned)
constexpr int SB_OFF_A2 = SB_OFF_RING + SB_SLOTS * DG_A_BYTES;   // 172032
constexpr int SB_OFF_P = SB_OFF_A2 + TCB_A_BYTES;                // 205056
constexpr int SB_OFF_DEC = SB_OFF_P + 2 * TCB_B_BYTES;           // 221568
constexpr int SB_SMEM = SB_OFF_DEC + 2 * (TC_M + TC_HALO) * 16;  // 226240
constexpr int SB_COLS = 512, SB_DW_COL = 384;

__device__ __forceinline__ void sb_bar_producers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// bounded wait by all 32 lanes with ONE verdict for the warp (a lane that gave up alone would leave the others at the next
// __syncwarp)
__device__ __forceinline__ bool sb_wait_warp(uint64_t *bar, uint32_t parity) { return __all_sync(0xFFFFFFFFu, tc_mbar_wait(bar, parity) ? 1 : 0) != 0; }

// tcgen05.mma with an explicit instruction descriptor.  The tensor core spends ~75 cycles on an M = 128, K = 16 step almost
// regardless of N <= 128 (measured through the issuer's busy time: 36 N = 64 steps = 2.7 k cycles per tile), so taps that
// share an A tile and write neighbouring classes go out as ONE N = 128 step, and GEMM 2's hi / lo operands as one N = 64 step.
// The MMA warp runs CONVERGED: all 32 lanes execute the issue loop, one elected lane issues (tc_mma_elect, ta_conv1_tc.cuh).
__device__ __forceinline__ void sb_mma(uint32_t tmem, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
    tc_mma_elect(tmem, descA, descB, idesc, accumulate);
}
__device__ __forceinline__ void sb_commit(uint64_t *bar) { tc_commit_elect(bar); }
// D f32, A / B bf16 K-major, M = 128
constexpr uint32_t SB_IDESC_N64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
constexpr uint32_t SB_IDESC_N128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
// GEMM 2: both operands MN-major, N = 64 = (16 taps + ones + 15 zero columns) of the hi operand, then the same of the lo operand
constexpr uint32_t SB_IDESC_W = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

template <typename XT>
__global__ void __launch_bounds__(SB_THREADS, 1) conv2_dgrad_conv1_wgrad_kernel(const __nv_bfloat16 *__restrict__ dz, const uint4 *__restrict__ wimg,
                                                                                const uint32_t *__restrict__ relu_mask, const XT *__restrict__ x,
                                                                                long long xstride, long long B, float *__restrict__ dw4,
                                                                                float *__restrict__ db4, int *fail, long long *prof) {
    // prof (nullable, development; 16 int64 filled by CTA 0): cycles [0] producer thread 0 waits for ring slots, [1] for p_empty,
    // [2] its total; [3] MMA issuer waits slot_full, [4] acc_empty, [5] a2_full / p_full, [6] total; [7] epilogue warp 0 waits
    // acc_full, [8] a2_empty, [9] total
    const bool pr = prof != nullptr && blockIdx.x == (gridDim.x > 64 ? 64 : 0);   // (CTA 0 sits on one of the few fast SMs)
    long long t_a = 0, t_b = 0, t_c = 0;
    const long long t_start = clock64();
    unsigned long long g_start = 0;
    if (prof != nullptr) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_start));
    extern __shared__ __align__(1024) uint8_t sb_smem[];
    uint8_t *sW = sb_smem, *sRing = sb_smem + SB_OFF_RING, *sA2 = sb_smem + SB_OFF_A2, *sP = sb_smem + SB_OFF_P;
    uint4 *sDec = reinterpret_cast<uint4 *>(sb_smem + SB_OFF_DEC);
    __shared__ __align__(8) uint64_t slot_full[SB_SLOTS], slot_empty[SB_SLOTS], acc_full[3], acc_empty[3], a2_full, a2_empty, p_full, p_empty, done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int producers_dead;   // a producer thread that gave up a wait: all 128 leave together, right after their named barrier
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) producers_dead = 0;

    for (int i = tid; i < DG_W_BYTES / 16; i += SB_THREADS) reinterpret_cast<uint4 *>(sW)[i] = __ldg(wimg + i);
    for (int i = tid; i < 2 * TCB_B_BYTES / 16; i += SB_THREADS) reinterpret_cast<uint4 *>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int s = 0; s < SB_SLOTS; s++) {
            dg_mbar_init(&slot_full[s], 128);   // noinc arrivals of the producer threads' cp.async groups
            dg_mbar_init(&slot_empty[s], 1);    // one tcgen05.commit
        }
        for (int s = 0; s < 3; s++) {
            dg_mbar_init(&acc_full[s], 1);      // one tcgen05.commit
            dg_mbar_init(&acc_empty[s], 8);     // the eight epilogue warps
        }
        dg_mbar_init(&a2_full, 8);              // the eight epilogue warps
        dg_mbar_init(&a2_empty, 1);
        dg_mbar_init(&p_full, 128);             // the producer threads
        dg_mbar_init(&p_empty, 1);
        dg_mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(SB_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const long long npos = B * (DG_P * DG_P), ntiles = (npos + TC_M - 1) / TC_M;
    const long long my_tiles = (long long)blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    bool dead = false;

    if (warp >= 4 && warp < 8) {
        // ---------------- producers ----------------
        const int pw = warp - 4, ptid = tid - 128;
        int it = 0;
        const int rsub = lane >> 3, ch = lane & 7;   // lane = (row within a quad, 16-byte chunk): 4 rows x 128 contiguous bytes per instruction
        for (long long tile = blockIdx.x; tile < ntiles && !dead; tile += gridDim.x, it++) {
            // this thread's 8 rows: element offset of the pixel's chunk in dz and which of the four shifts exist
            unsigned roff[8], rok = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = pw * 32 + 4 * i + rsub;
                const unsigned P = (unsigned)(tile * TC_M) + (unsigned)row;     // (batch * 289 < 2^31, checked by the host)
                const bool valid = (long long)P < npos;
                const unsigned b = P / (unsigned)(DG_P * DG_P), pos = P - b * (unsigned)(DG_P * DG_P), m = pos / (unsigned)DG_P, nn = pos - m * (unsigned)DG_P;
                roff[i] = ((b * DG_OH + m) * DG_OH + nn) * 64u + ch * 8u;    // (< 2^31 elements: batch * 256 * 64)
#pragma unroll
                for (int sh = 0; sh < 4; sh++) {
                    const unsigned dy = sh >> 1, dx = sh & 1;
                    const bool ok = valid && m >= dy && m - dy < (unsigned)DG_OH && nn >= dx && nn - dx < (unsigned)DG_OH;
                    rok |= (ok ? 1u : 0u) << (4 * i + sh);
                }
            }
            // ring entries 4 it .. 4 it + 3 hold shifts 3, 2, 1, 0: each is filled as soon as ITS slot is free (the first two
            // are released half a tile earlier than the others)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int seq = it * 4 + j, slot = seq % SB_SLOTS, n = seq / SB_SLOTS, sh = 3 - j;
                const long long tw = clock64();
                if (!dead && !tc_mbar_wait(&slot_empty[slot], (uint32_t)((n & 1) ^ 1))) dead = true;   // (passes at once on a slot's first use)
                t_a += clock64() - tw;
                const uint32_t sbase = smem_u32(sRing + slot * DG_A_BYTES);
                const int back = ((sh >> 1) * DG_OH + (sh & 1)) * 64;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const bool ok = (rok >> (4 * i + sh)) & 1u;
                    dg_cp_async16(sbase + dg_row_chunk(pw * 32 + 4 * i + rsub, ch), ok ? dz + roff[i] - back : dz, ok ? 16u : 0u);
                }
                // this lane's copies of the slot arrive on its barrier when they have landed
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&slot_full[slot])) : "memory");
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (pr && ptid == 0) { prof[0] = t_a; prof[2] = clock64() - t_start; }
    } else if (warp >= 9 && warp < 13) {
        // ---------------- decoders: the tile's P operand (thread = position) ----------------
        // (a thread whose wait gave up keeps going without waiting -- it must reach the named barrier of the next tile,
        // where all 128 leave together)
        const int ptid = tid - 288;
        TcRaw<XT> raw;
        if (my_tiles > 0) tc_load_raw<XT>(x, xstride, npos, blockIdx.x, ptid, raw);
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            // LUT decode of the 4 frames, 2x2 patch through a double-buffered shared array, MN-major chunks (conv1_bwd_tc_kernel's layout)
            uint4 *dec = sDec + (it & 1) * (TC_M + TC_HALO);
            tc_decode_raw<XT>(raw, npos, tile, ptid, dec);
            if (tile + gridDim.x < ntiles) tc_load_raw<XT>(x, xstride, npos, tile + gridDim.x, ptid, raw);
            if (dead) producers_dead = 1;
            sb_bar_producers();   // (the other buffer is rewritten only after every thread has passed this barrier again)
            if (producers_dead) { dead = true; break; }
            const long long Pq = tile * TC_M + ptid;
            const bool valid = Pq < npos;
            const long long bq = valid ? Pq / NCELL : 0;
            const int pos = valid ? (int)(Pq - bq * NCELL) : 0, m = pos / GS, n = pos - GS * m;
            const bool rgt = valid && n < 16, bot = valid && m < 16;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const uint4 d00 = valid ? dec[ptid] : z, d01 = rgt ? dec[ptid + 1] : z;
            const uint4 d10 = bot ? dec[ptid + GS] : z, d11 = (rgt && bot) ? dec[ptid + GS + 1] : z;
            const long long tw = clock64();
            if (!dead && !tc_mbar_wait(&p_empty, (uint32_t)((it & 1) ^ 1))) dead = true;   // GEMM 2 of the previous tile has read P
            t_b += clock64() - tw;
            const uint32_t off = (ptid >> 3) * 128 + (ptid & 7) * 16;
            *reinterpret_cast<uint4 *>(sP + off) = make_uint4(d00.x, d00.y, d01.x, d01.y);
            *reinterpret_cast<uint4 *>(sP + TCB_SBO + off) = make_uint4(d10.x, d10.y, d11.x, d11.y);
            *reinterpret_cast<uint4 *>(sP + 2 * TCB_SBO + off) = make_uint4(valid ? 0x3F80u : 0u, 0u, 0u, 0u);  // bf16 1.0: bias column
            *reinterpret_cast<uint4 *>(sP + TCB_B_BYTES + off) = make_uint4(d00.z, d00.w, d01.z, d01.w);
            *reinterpret_cast<uint4 *>(sP + TCB_B_BYTES + TCB_SBO + off) = make_uint4(d10.z, d10.w, d11.z, d11.w);
            fence_proxy_async();
            dg_mbar_arrive(&p_full);
        }
        if (pr && ptid == 0) { prof[1] = t_b; prof[10] = clock64() - t_start; }
    } else if (warp == 8) {
        // ---------------- MMA issuer ----------------
        if (my_tiles > 0) {   // (all 32 lanes, converged: see sb_mma)
            uint64_t dW[DG_TAPS];
#pragma unroll
            for (int t = 0; t < DG_TAPS; t++) dW[t] = dg_smem_desc(sW + t * DG_WTAP_BYTES);
            const uint64_t descA2 = tcb_smem_desc(sA2, 128u, TCB_SBO), descPh = tcb_smem_desc(sP, 128u, TCB_SBO);
            // GEMM 2 of pair k (k counts this CTA's class pairs): 8 K steps of 16 positions x (hi, lo)
            auto wgrad = [&](long long k) -> bool {
                const int pair = (int)(k & 1);
                const long long tw = clock64();
                if (pair == 0 && !sb_wait_warp(&p_full, (uint32_t)((k >> 1) & 1))) return false;
                if (!sb_wait_warp(&a2_full, (uint32_t)(k & 1))) return false;
                t_c += clock64() - tw;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                fence_proxy_async();
                const uint32_t d = tmem_base + (uint32_t)(SB_DW_COL + pair * 64);
#pragma unroll
                for (int ks = 0; ks < TC_M / 16; ks++) {   // columns 0..31: hi operand, 32..63: lo operand (adjacent row groups of sP)
                    const uint64_t koff = (uint64_t)((ks * 256) >> 4);
#ifndef SB_EXPERIMENT_SKIP_W
                    sb_mma(d, descA2 + koff, descPh + koff, SB_IDESC_W, (k >= 2 || ks) ? 1u : 0u);
#endif
                }
                sb_commit(&a2_empty);
                if (pair == 1) sb_commit(&p_empty);
                return true;
            };
            long long k = 0;
            for (long long itl = 0; itl < my_tiles && !dead; itl++) {
                const int it = (int)itl;
                uint64_t dA[4];
                auto ring_wait = [&](int j) -> bool {   // ring entry j of the tile = shift 3 - j
                    const int seq = it * 4 + j, slot = seq % SB_SLOTS, n = seq / SB_SLOTS;
                    const long long tw = clock64();
                    const bool ok = sb_wait_warp(&slot_full[slot], (uint32_t)(n & 1));
                    t_a += clock64() - tw;
                    dA[3 - j] = dg_smem_desc(sRing + slot * DG_A_BYTES);
                    return ok;
                };
                auto ring_release = [&](int j) {
                    sb_commit(&slot_empty[(it * 4 + j) % SB_SLOTS]);
                };
#pragma unroll
                for (int pair = 0; pair < 2; pair++, k++) {
                    const int aslot = (int)(k % 3);
                    const long long tw = clock64();
                    if (!sb_wait_warp(&acc_empty[aslot], (uint32_t)(((k / 3) & 1) ^ 1))) { dead = true; break; }
                    t_b += clock64() - tw;
                    // pair 0 = classes (0,0), (0,1): taps 6|7 from shift 2 and 0|1 from shift 0 feed both classes (N = 128: the
                    // second tap's rows follow the first's in the weight image), taps 8 (shift 3) and 2 (shift 1) class (0,0) only;
                    // pair 1 = classes (1,0), (1,1): taps 3|4 from shift 0, tap 5 (shift 1) class (1,0) only.
                    // Shifts 3 and 2 go first and their ring slots are released before the rest of the pair is issued.
                    const uint32_t dcol = tmem_base + (uint32_t)(aslot * 128);
                    if (pair == 0) {
                        if (!ring_wait(1) || !ring_wait(0)) { dead = true; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        fence_proxy_async();   // the producers' writes, acquired through the barriers -> the tensor core's async-proxy reads
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t koff = (uint64_t)(ks * DG_KSTEP);
                            sb_mma(dcol, dA[2] + koff, dW[6] + koff, SB_IDESC_N128, ks ? 1u : 0u);
                            sb_mma(dcol, dA[3] + koff, dW[8] + koff, SB_IDESC_N64, 1u);
                        }
                        ring_release(0);
                        ring_release(1);
                        if (!ring_wait(2) || !ring_wait(3)) { dead = true; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        fence_proxy_async();
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t koff = (uint64_t)(ks * DG_KSTEP);
                            sb_mma(dcol, dA[0] + koff, dW[0] + koff, SB_IDESC_N128, 1u);
                            sb_mma(dcol, dA[1] + koff, dW[2] + koff, SB_IDESC_N64, 1u);
                        }
                    } else {
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t koff = (uint64_t)(ks * DG_KSTEP);
                            sb_mma(dcol, dA[0] + koff, dW[3] + koff, SB_IDESC_N128, ks ? 1u : 0u);
                            sb_mma(dcol, dA[1] + koff, dW[5] + koff, SB_IDESC_N64, 1u);
                        }
                        ring_release(2);
                        ring_release(3);
                    }
                    sb_commit(&acc_full[aslot]);
                    if (k >= 1 && !wgrad(k - 1)) { dead = true; break; }
                }
            }
            if (!dead && k >= 1 && !wgrad(k - 1)) dead = true;
            sb_commit(&done_bar);
            if (pr && lane == 0) { prof[3] = t_a; prof[4] = t_b; prof[5] = t_c; prof[6] = clock64() - t_start; }
        }
    } else {
        // ---------------- epilogue (warps 0-3: the even class of every pair, warps 13-16: the odd one) ----------------
        // lane = position 32 * q + lane of the tile; a warp may only read the TMEM lane quadrant warp % 4
        const int q = warp & 3, cc = warp >= 13 ? 1 : 0;
        long long k = 0;
        const int p = q * 32 + lane;
        uint8_t *a2row = sA2 + (p >> 3) * 128 + (p & 7) * 16 + cc * 8 * TCB_SBO;
        for (long long tile = blockIdx.x; tile < ntiles && !dead; tile += gridDim.x) {
            const long long P = tile * TC_M + p;
            // conv1's ReLU mask words of this position and this warp's two classes (cc, 2 + cc), fetched before the waits
            uint2 mw[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
            if (P < npos) {
                mw[0] = __ldg(reinterpret_cast<const uint2 *>(relu_mask + P * 8 + cc * 2));
                mw[1] = __ldg(reinterpret_cast<const uint2 *>(relu_mask + P * 8 + 4 + cc * 2));
            }
#pragma unroll
            for (int pair = 0; pair < 2; pair++, k++) {
                const int aslot = (int)(k % 3);
                long long tw = clock64();
                if (!sb_wait_warp(&acc_full[aslot], (uint32_t)((k / 3) & 1))) { dead = true; break; }
                t_a += clock64() - tw;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t r[2][32];
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(aslot * 128 + cc * 64 + half * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[half][0]), "=r"(r[half][1]), "=r"(r[half][2]), "=r"(r[half][3]), "=r"(r[half][4]), "=r"(r[half][5]),
                          "=r"(r[half][6]), "=r"(r[half][7]), "=r"(r[half][8]), "=r"(r[half][9]), "=r"(r[half][10]), "=r"(r[half][11]),
                          "=r"(r[half][12]), "=r"(r[half][13]), "=r"(r[half][14]), "=r"(r[half][15]), "=r"(r[half][16]), "=r"(r[half][17]),
                          "=r"(r[half][18]), "=r"(r[half][19]), "=r"(r[half][20]), "=r"(r[half][21]), "=r"(r[half][22]), "=r"(r[half][23]),
                          "=r"(r[half][24]), "=r"(r[half][25]), "=r"(r[half][26]), "=r"(r[half][27]), "=r"(r[half][28]), "=r"(r[half][29]),
                          "=r"(r[half][30]), "=r"(r[half][31])
                        : "r"(taddr));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // this warp's part of the accumulator slot is in registers: it may be overwritten
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) dg_mbar_arrive(&acc_empty[aslot]);
                uint4 o[8];   // the position's 64 channels of this class: 8 chunks of 8 bf16, masked
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const uint32_t mb = half ? mw[pair].y : mw[pair].x;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t w4[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const __nv_bfloat162 pk =
                                __floats2bfloat162_rn(__uint_as_float(r[half][8 * j + 2 * i]), __uint_as_float(r[half][8 * j + 2 * i + 1]));
                            w4[i] = *reinterpret_cast<const uint32_t *>(&pk) & (((mb >> (4 * j + i)) & 0x00010001u) * 0xFFFFu);
                        }
                        o[half * 4 + j] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    }
                }
                // GEMM 2 of the previous pair must have read the operand block before it is rewritten
                tw = clock64();
                if (!sb_wait_warp(&a2_empty, (uint32_t)((k & 1) ^ 1))) dead = true;
                t_b += clock64() - tw;
                // row group g = (class of the pair, 8 channels): the 32 lanes of a store are 512 contiguous bytes
#pragma unroll
                for (int j = 0; j < 8; j++) *reinterpret_cast<uint4 *>(a2row + j * TCB_SBO) = o[j];
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) dg_mbar_arrive(&a2_full);
                if (dead) break;
            }
        }
        if (pr && tid == 0) { prof[7] = t_a; prof[8] = t_b; prof[9] = clock64() - t_start; }
        // ---- dW4 / db4: lane = row (class-pair member * 64 + channel), columns = 16 taps + bias; one atomicAdd per value and CTA
        if (my_tiles > 0 && !dead && cc == 0) {
            if (!sb_wait_warp(&done_bar, 0u)) dead = true;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (!dead) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    float acc[17];
#pragma unroll
                    for (int part = 0; part < 2; part++) {   // the hi operand's 32 columns, then the lo operand's
                        uint32_t r[32];
                        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(SB_DW_COL + h * 64 + part * 32);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                            : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int kk = 0; kk < 17; kk++) acc[kk] = part ? acc[kk] + __uint_as_float(r[kk]) : __uint_as_float(r[kk]);
                    }
                    const int row = (2 * h + (p >> 6)) * C1_CH + (p & 63);
#pragma unroll
                    for (int kk = 0; kk < 16; kk++) atomicAdd(dw4 + row * 16 + kk, acc[kk]);
                    atomicAdd(db4 + row, acc[16]);
                }
            }
        }
    }
    if (dead && fail) atomicExch(fail, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(SB_COLS));
    if (prof != nullptr && tid == 0) {   // [11] CTA 0's lifetime in ns, [12] / [13] the longest CTA lifetime in cycles / ns
        unsigned long long g_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_end));
        if (blockIdx.x == 0) prof[11] = (long long)(g_end - g_start);
        atomicMax(reinterpret_cast<unsigned long long *>(prof + 12), (unsigned long long)(clock64() - t_start));
        atomicMax(reinterpret_cast<unsigned long long *>(prof + 13), g_end - g_start);
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        prof[16 + 3 * blockIdx.x] = clock64() - t_start;   // per CTA: cycles, ns, SM id (the buffer holds 16 + 3 * grid values)
        prof[17 + 3 * blockIdx.x] = (long long)(g_end - g_start);
        prof[18 + 3 * blockIdx.x] = smid;
    }
}

}  // namespace ta
