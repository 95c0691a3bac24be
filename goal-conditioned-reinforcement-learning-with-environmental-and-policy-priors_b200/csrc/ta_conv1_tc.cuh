// ta_conv1_tc.cuh -- TINet's fused first layer (see ta_conv1.cuh) on the 5th-generation tensor cores.
//
// The folded layer is a GEMM with a tiny reduction: D[position, (phase, channel)] = P[position, 16] x W4^T[16, 256].
// tcgen05.mma (M = 128 positions, N = 64 = one output phase, K = 16, bf16 inputs, fp32 accumulators in TMEM)
// covers 128 input positions x 64 channels; the FP32-FMA version of the same work is bound by
// the FMA pipe (4.6 GFMA per 4096 samples: 326 us), this one by its 570 MB of output and the serial stage -> MMA ->
// epilogue chain of a CTA (151 us, 3.8 TB/s).  Every operand is a bf16 (hi, lo) pair and a product is three MMAs
// (hi*hi + lo*hi + hi*lo), so the result matches the FP32 kernel to the rounding of the bf16 output.
//
//   CTA = 128 threads = 4 warps; thread r owns row r of the tile = one input position (sample b, m, n):
//     build   the tile's 128 + 18 input positions are loaded / LUT-decoded once (4 frames each) into shared
//             memory; row r then gathers its 2x2 patch (positions r, r+1, r+17, r+18) -> two 16-byte chunks of
//             the A tile, canonical K-major no-swizzle layout (8-row x 16-byte core matrices)
//     mma     thread 0: tcgen05.mma [tmem], descA, descB, idesc; tcgen05.commit -> mbarrier
//     epilog  warp w reads TMEM lanes 32w..32w+31 (tcgen05.ld 32x32b.x32): thread r gets row r, 32 channels at
//             a time; + bias, ReLU, bf16 pack -> a swizzled 32 x 64-byte staging block in shared memory (warp
//             private, __syncwarp only) -> read back transposed so that one store instruction covers 8 output
//             pixels x 64 contiguous bytes (per-thread rows would touch 32 cache lines per instruction)
//   W4 (bf16 hi/lo, canonical layout, 2 x 8 KB) and the bias stay in shared memory for the CTA's lifetime.
//   TMEM: 64 columns per CTA (one output phase at a time); registers (80) and 38 KB of shared memory allow 6 CTAs per SM.
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "ta_conv1.cuh"

namespace ta {

constexpr int TC_THREADS = 128;
constexpr int TC_M = 128, TC_N = 256;
constexpr int TC_HALO = GS + 1;  // a row's patch reaches 18 positions ahead
constexpr int TC_NT = 64;  // columns per MMA = TMEM columns per CTA: one output phase
// shared-memory descriptor of a K-major, no-swizzle operand with K = 16 bf16 (two 16-byte chunks per row):
// element (row r, chunk c) at byte (r / 8) * 256 + c * 128 + (r % 8) * 16, i.e. LBO = 128 B, SBO = 256 B
__device__ __forceinline__ uint64_t tc_smem_desc(const void *smem) {
    const uint64_t addr = (uint64_t)(smem_u32(smem) >> 4) & 0x3FFFu;
    return addr | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);  // version = 1, layout = none
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, N = TC_NT, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_NT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ uint32_t tc_operand_offset(int row, int chunk) { return (row >> 3) * 256 + chunk * 128 + (row & 7) * 16; }

__device__ __forceinline__ void tc_mbar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: a descriptor mistake must not turn into a hung GPU
__device__ __forceinline__ bool tc_mbar_wait(uint64_t *bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

// (v0, v1) -> packed bf16 pairs hi = bf16(v), lo = bf16(v - hi)
__device__ __forceinline__ void tc_split(float v0, float v1, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __low2float(h), v1 - __high2float(h));
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
// (max(a, 0), max(b, 0)) -> packed bf16 pair in one instruction
__device__ __forceinline__ uint32_t tc_relu_pack(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// D[tmem] (+)= A[smem] x B[smem]^T, M = 128, N = TC_NT, K = 16; issued by one thread
__device__ __forceinline__ void tc_mma(uint32_t tmem, uint64_t descA, uint64_t descB, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                 "l"(descA), "l"(descB), "r"(TC_IDESC), "r"(accumulate)
                 : "memory");
}

// The same, executed by a CONVERGED warp: one elected lane issues.  Inside an `if (tid == 0)` the descriptors live in
// vector registers and ptxas wraps every tcgen05.mma in R2UR moves and an ELECT / BRA.U.ANY loop (~15 instructions, ~100
// cycles per MMA in the issuing thread); with warp-uniform control flow they stay in uniform registers.
__device__ __forceinline__ void tc_mma_elect(uint32_t tmem, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
        "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t *bar) {
    asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n}\n" ::"r"(smem_u32(bar))
                 : "memory");
}

template <typename XT>
__device__ __forceinline__ float tc_value(XT v) {
    if constexpr (sizeof(XT) == 1) return c1_decode((uint32_t)v);
    else return (float)v;
}

// The tile's 128 + 18 input positions (4 frames each) as raw values: slot tid and, for tid < TC_HALO, slot 128 + tid.
// Loaded one tile AHEAD into registers (the loads of tile k+1 are issued as soon as tile k's values have been decoded and
// complete under tile k's MMA / epilogue): ncu showed 40 % of the forward kernel's stall samples on these four small loads.
template <typename XT>
struct TcRaw {   // uint8 codes: the 4 frames of a position packed in one register; float frames: 4 registers (scalars, not an
    typename std::conditional<sizeof(XT) == 1, uint32_t, float4>::type a, b;   // array: ptxas must keep them in registers)
};
template <typename XT>
__device__ __forceinline__ auto tc_load_slot(const XT *__restrict__ x, long long xstride, long long npos, long long tile, int slot) {
    const long long Q = tile * TC_M + slot;
    const bool in = slot < TC_M + TC_HALO && Q < npos;
    const long long qb = in ? Q / NCELL : 0;
    const XT *xq = x + qb * xstride + (in ? (int)(Q - qb * NCELL) : 0);
    if constexpr (sizeof(XT) == 1) {
        uint32_t w = 0;
        if (in) w = (uint32_t)xq[0] | ((uint32_t)xq[NCELL] << 8) | ((uint32_t)xq[2 * NCELL] << 16) | ((uint32_t)xq[3 * NCELL] << 24);
        return w;
    } else {
        return in ? make_float4(xq[0], xq[NCELL], xq[2 * NCELL], xq[3 * NCELL]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <typename XT>
__device__ __forceinline__ void tc_load_raw(const XT *__restrict__ x, long long xstride, long long npos, long long tile, int tid, TcRaw<XT> &raw) {
    raw.a = tc_load_slot<XT>(x, xstride, npos, tile, tid);
    raw.b = tc_load_slot<XT>(x, xstride, npos, tile, tid + TC_M);
}
template <typename XT, typename RT>
__device__ __forceinline__ void tc_decode_slot(const RT &r, long long npos, long long tile, int slot, uint4 *sDec) {
    if (slot < TC_M + TC_HALO) {
        uint32_t hi0 = 0, hi1 = 0, lo0 = 0, lo1 = 0;
        if (tile * TC_M + slot < npos) {
            float f0, f1, f2, f3;
            if constexpr (sizeof(XT) == 1) {
                f0 = c1_decode(r & 0xFFu); f1 = c1_decode((r >> 8) & 0xFFu); f2 = c1_decode((r >> 16) & 0xFFu); f3 = c1_decode(r >> 24);
            } else {
                f0 = r.x; f1 = r.y; f2 = r.z; f3 = r.w;
            }
            tc_split(f0, f1, hi0, lo0);
            tc_split(f2, f3, hi1, lo1);
        }
        sDec[slot] = make_uint4(hi0, hi1, lo0, lo1);
    }
}
template <typename XT>
__device__ __forceinline__ void tc_decode_raw(const TcRaw<XT> &raw, long long npos, long long tile, int tid, uint4 *sDec) {
    tc_decode_slot<XT>(raw.a, npos, tile, tid, sDec);
    tc_decode_slot<XT>(raw.b, npos, tile, tid + TC_M, sDec);
}

// MK: also write the layer's ReLU mask as bits (relu_mask uint32 [position][4 phases][2 halves of 32 channels]; in a
// word, bit q = channel 2q of the half is non-zero, bit 16 + q = channel 2q + 1), 8 bytes per output pixel: what the
// weight-gradient kernel reads instead of the 128 bytes of y.
template <typename XT, bool MK>
__global__ void __launch_bounds__(TC_THREADS) conv1_fwd_tc_kernel(const XT *__restrict__ x, long long xstride,
                                                                 const float *__restrict__ w4, const float *__restrict__ b4,
                                                                 long long B, __nv_bfloat16 *__restrict__ y, uint32_t *__restrict__ relu_mask,
                                                                 int *fail) {
    // every operand is kept as a bf16 (hi, lo) pair, v = hi + lo to 2^-17: three MMAs (hi*hi + lo*hi + hi*lo) give
    // fp32-grade products, so the layer matches the FP32 kernel / cuDNN fp32 to rounding of the bf16 output
    __shared__ __align__(1024) uint8_t sA[2][TC_M * 32];   // 2 x 4 KB
    __shared__ __align__(1024) uint8_t sB[2][TC_N * 32];   // 2 x 8 KB
    __shared__ __align__(16) float sbias[TC_N];
    __shared__ __align__(16) uint8_t sOut[TC_M * 64];      // epilogue staging: 32 channels (64 B) per row
    __shared__ uint4 sDec[TC_M + TC_HALO];                 // decoded inputs of the tile's positions + halo
    __shared__ long long rowinfo[TC_M];                    // (output pixel of phase 0) * 4 | (m == 16) * 2 | (n == 16); -1 = no row
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *wout = sOut + warp * 2048;

    // one-time setup: W4 -> bf16 canonical layout, bias, barrier, TMEM
    for (int i = tid; i < TC_N * 2; i += TC_THREADS) {
        const int n = i >> 1, c = i & 1;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; q++) tc_split(__ldg(w4 + n * 16 + c * 8 + 2 * q), __ldg(w4 + n * 16 + c * 8 + 2 * q + 1), hi[q], lo[q]);
        *reinterpret_cast<uint4 *>(sB[0] + tc_operand_offset(n, c)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(sB[1] + tc_operand_offset(n, c)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int i = tid; i < TC_N; i += TC_THREADS) sbias[i] = __ldg(b4 + i);
    if (tid == 0) tc_mbar_init(&bar);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TC_NT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint64_t descA_hi = tc_smem_desc(sA[0]), descA_lo = tc_smem_desc(sA[1]);
    const uint64_t descB_hi = tc_smem_desc(sB[0]), descB_lo = tc_smem_desc(sB[1]);
    const long long npos = B * NCELL, ntiles = (npos + TC_M - 1) / TC_M;
    uint32_t parity = 0;
    bool dead = false;
    TcRaw<XT> raw;
    if ((long long)blockIdx.x < ntiles) tc_load_raw<XT>(x, xstride, npos, blockIdx.x, tid, raw);

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long P = tile * TC_M + tid;
        const bool valid = P < npos;
        const long long b = valid ? P / NCELL : 0;
        const int pos = valid ? (int)(P - b * NCELL) : 0, m = pos / GS, n = pos - GS * m;
        // ---- A tile, k = (dy*2+dx)*4 + c -------------------------------------------------------------------
        // stage 1: positions tile*128 .. +127+18 -> decoded (hi, lo) bf16 of the 4 frames, 16 bytes per position
        // (each input cell feeds up to four rows of the tile, so it is loaded and decoded once); the raw values were
        // loaded during the previous tile, the next tile's loads are issued now
        tc_decode_raw<XT>(raw, npos, tile, tid, sDec);
        if (tile + gridDim.x < ntiles) tc_load_raw<XT>(x, xstride, npos, tile + gridDim.x, tid, raw);
        __syncthreads();
        // stage 2: row r = the 2x2 patch at positions r, r+1, r+17, r+18 (zero past the right / bottom edge)
        {
            const bool rgt = valid && n < 16, bot = valid && m < 16;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const uint4 d00 = valid ? sDec[tid] : z, d01 = rgt ? sDec[tid + 1] : z;
            const uint4 d10 = bot ? sDec[tid + GS] : z, d11 = (rgt && bot) ? sDec[tid + GS + 1] : z;
            *reinterpret_cast<uint4 *>(sA[0] + tc_operand_offset(tid, 0)) = make_uint4(d00.x, d00.y, d01.x, d01.y);
            *reinterpret_cast<uint4 *>(sA[0] + tc_operand_offset(tid, 1)) = make_uint4(d10.x, d10.y, d11.x, d11.y);
            *reinterpret_cast<uint4 *>(sA[1] + tc_operand_offset(tid, 0)) = make_uint4(d00.z, d00.w, d01.z, d01.w);
            *reinterpret_cast<uint4 *>(sA[1] + tc_operand_offset(tid, 1)) = make_uint4(d10.z, d10.w, d11.z, d11.w);
        }
        rowinfo[tid] = valid ? (((b * (C1_OUT * C1_OUT) + 2 * m * C1_OUT + 2 * n) << 2) | (m == 16 ? 2 : 0) | (n == 16 ? 1 : 0)) : -1ll;
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's (async proxy) reads
        __syncthreads();
        // the 4 rows this lane stores in the transposed read-out: address of its 16 bytes in the phase-0 pixel, edge flags
        __nv_bfloat16 *ybase[4];
        uint32_t edge = 0;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const long long inf = dead ? -1ll : rowinfo[warp * 32 + it * 8 + (lane >> 2)];
            ybase[it] = inf >= 0 ? y + (inf >> 2) * C1_CH + (lane & 3) * 8 : nullptr;
            edge |= (uint32_t)(inf & 3) << (2 * it);
        }
        // ---- one MMA per output phase (TC_NT = 64 columns), epilogue straight out of TMEM -------------------
#pragma unroll 1
        for (int phase = 0; phase < TC_N / TC_NT; phase++) {
            if (warp == 0) {   // (converged: one elected lane issues, see tc_mma_elect)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t boff = (uint64_t)((phase * TC_NT / 8 * 256) >> 4);  // rows phase*64.. of W4
                tc_mma_elect(tmem_base, descA_hi, descB_hi + boff, TC_IDESC, 0u);
                tc_mma_elect(tmem_base, descA_lo, descB_hi + boff, TC_IDESC, 1u);
                tc_mma_elect(tmem_base, descA_hi, descB_lo + boff, TC_IDESC, 1u);
                tc_commit_elect(&bar);
            }
            if (!dead && !tc_mbar_wait(&bar, parity)) dead = true;
            parity ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int py = phase >> 1, px = phase & 1, poff = (py * C1_OUT + px) * C1_CH;
#pragma unroll
            for (int half = 0; half < TC_NT / 32; half++) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int c0 = half * 32;
                // row `lane` of the warp's 32 x 64-byte staging block, 16-byte chunks XOR-swizzled by (row / 2) % 4
                uint32_t mbits = 0;  // bit c: channel c0 + c of this row's output pixel is non-zero (the ReLU mask the backward needs)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(sbias + phase * 64 + c0 + 8 * j);
                    const float4 b1 = *reinterpret_cast<const float4 *>(sbias + phase * 64 + c0 + 8 * j + 4);
                    const uint32_t o0 = tc_relu_pack(__uint_as_float(r[8 * j]) + b0.x, __uint_as_float(r[8 * j + 1]) + b0.y);
                    const uint32_t o1 = tc_relu_pack(__uint_as_float(r[8 * j + 2]) + b0.z, __uint_as_float(r[8 * j + 3]) + b0.w);
                    const uint32_t o2 = tc_relu_pack(__uint_as_float(r[8 * j + 4]) + b1.x, __uint_as_float(r[8 * j + 5]) + b1.y);
                    const uint32_t o3 = tc_relu_pack(__uint_as_float(r[8 * j + 6]) + b1.z, __uint_as_float(r[8 * j + 7]) + b1.w);
                    *reinterpret_cast<uint4 *>(wout + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(o0, o1, o2, o3);
                    if constexpr (MK) {
                        // a ReLU output half is >= +0, so half + 0x7FFF has bit 15 set exactly when it is non-zero (no carry
                        // into the other half); word 4j+i contributes bits 4j+i and 16+4j+i
                        mbits |= (((o0 + 0x7FFF7FFFu) & 0x80008000u) >> (15 - 4 * j)) | (((o1 + 0x7FFF7FFFu) & 0x80008000u) >> (14 - 4 * j)) |
                                 (((o2 + 0x7FFF7FFFu) & 0x80008000u) >> (13 - 4 * j)) | (((o3 + 0x7FFF7FFFu) & 0x80008000u) >> (12 - 4 * j));
                    }
                }
                if constexpr (MK) {
                    if (valid) relu_mask[(P * 4 + phase) * 2 + half] = mbits;
                }
                __syncwarp();
                // transposed read-out: 4 lanes per row -> a store instruction covers 8 rows x 64 contiguous bytes
#pragma unroll
                for (int it = 0; it < 4; it++) {
                    const int row = it * 8 + (lane >> 2), j = lane & 3;
                    const uint4 v = *reinterpret_cast<const uint4 *>(wout + row * 64 + ((j ^ ((row >> 1) & 3)) << 4));
                    if (ybase[it] && !(py && (edge >> (2 * it + 1) & 1)) && !(px && (edge >> (2 * it) & 1))) *reinterpret_cast<uint4 *>(ybase[it] + poff + c0) = v;
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();  // TMEM is free for the next phase's MMA
        }
    }
    if (dead && fail) atomicExch(fail, 1);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_NT));
}

// ---- weight / bias gradient of the same layer on the tensor cores -------------------------------------------
// dW4[(phase, o), k] = sum over positions of dz[position, (phase, o)] * P[position, k], db4 = the same with P = 1:
// a GEMM whose reduction runs over the positions.  Both operands are therefore "MN-major" in shared memory
// (8 consecutive channels / taps of one position form a 16-byte chunk, consecutive positions follow at 16 bytes,
// 8 positions = one core matrix of 128 bytes):
//   A (dz)  half a tile at a time = 2 output phases x 64 channels = 128 rows, K = 128 positions; chunk (group g of
//           8 rows, position p) at g * TCB_SBO + (p / 8) * 128 + (p % 8) * 16
//   B (P)   N = 32 columns: 16 taps (bf16 hi / lo pair, as in the forward kernel), column 16 = 1 for the bias
//           gradient, 15 zero columns; same chunk rule
// dz = dy where y > 0 is formed on the way from global to shared memory (8 lanes read the 128 contiguous bytes of a
// pixel).  The two fp32 accumulators (128 lanes x 32 columns each) stay in TMEM for the CTA's whole lifetime: no
// per-tile epilogue, one atomicAdd per value and CTA at the end.
constexpr int TCB_SBO = 128 * 16 + 16;                 // + 16: the 8 lanes of a quarter warp hit 8 different bank groups
constexpr int TCB_A_BYTES = 16 * TCB_SBO, TCB_B_BYTES = 4 * TCB_SBO;
constexpr int TCB_N = 32, TCB_COLS = 64, TCB_BATCH = 4;
constexpr uint32_t TCB_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TCB_N >> 3) << 17) |
                               ((uint32_t)(TC_M >> 4) << 24);  // D f32, A/B bf16, both MN-major, N = 32, M = 128

__device__ __forceinline__ uint64_t tcb_smem_desc(const void *smem, uint32_t lbo, uint32_t sbo) {
    const uint64_t addr = (uint64_t)(smem_u32(smem) >> 4) & 0x3FFFu;
    return addr | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tcb_mma(uint32_t tmem, uint64_t descA, uint64_t descB, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                 "l"(descA), "l"(descB), "r"(TCB_IDESC), "r"(accumulate)
                 : "memory");
}
// 16-byte read-only load as a volatile asm: stays above the compiler barrier that separates a batch's loads from its stores
__device__ __forceinline__ uint4 tcb_ldg(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// 8 bf16 of dy kept where the matching bf16 of y (a ReLU output, so >= +0) is non-zero
__device__ __forceinline__ uint4 tcb_mask(uint4 g, uint4 yv) {
    auto m = [](uint32_t gw, uint32_t yw) { return gw & (((yw & 0xFFFFu) ? 0xFFFFu : 0u) | ((yw >> 16) ? 0xFFFF0000u : 0u)); };
    return make_uint4(m(g.x, yv.x), m(g.y, yv.y), m(g.z, yv.z), m(g.w, yv.w));
}

// dL/dy either as one channels-last tensor [B][33][33][64] (PL = false, `dy`) or as merged parity planes (PL = true):
// planes bf16 [B][17][17][4][64], block py*2+px of position (m, n) = the gradient of output pixel (2m+py, 2n+px) -- the
// form in which conv2's data gradient falls out of ONE stride-1 convolution of its dz (ta_conv1.cuh,
// parity_class_weights_kernel; conv1.py: _Stem).  It is exactly this kernel's phase structure, so no interleaving
// pass is needed and a position's four phases are 512 contiguous bytes.

// MK (with PL): the ReLU mask comes from the forward kernel's bit mask (uint32 [position][4 phases][2], see
// conv1_fwd_tc_kernel; 8 bytes per pixel instead of the 128 bytes of y) and y is not read at all.
template <typename XT, bool PL, bool MK>
__global__ void __launch_bounds__(TC_THREADS) conv1_bwd_tc_kernel(const XT *__restrict__ x, long long xstride,
                                                                 const __nv_bfloat16 *__restrict__ y,
                                                                 const __nv_bfloat16 *__restrict__ dy, const __nv_bfloat16 *__restrict__ planes,
                                                                 const uint32_t *__restrict__ relu_mask, long long B,
                                                                 float *__restrict__ dw4, float *__restrict__ db4,
                                                                 uint32_t zero, int *fail, long long pl_pos_stride, long long pl_cls_stride) {
    // planes addressing (elements): block c of position P at P * pl_pos_stride + c * pl_cls_stride -- position-major
    // [P][4][64] (256, 64: cuDNN's merged planes) or class-major [4][P][64] (64, npos * 64: conv2_dgrad_planes_ws_kernel)
    extern __shared__ __align__(128) uint8_t tcb_smem[];
    uint8_t *sG = tcb_smem;                               // A: TCB_A_BYTES
    uint8_t *sP = tcb_smem + TCB_A_BYTES;                 // B hi, B lo: 2 x TCB_B_BYTES
    __shared__ uint4 sDec[TC_M + TC_HALO];
    __shared__ long long rowinfo[TC_M];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) tc_mbar_init(&bar);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TCB_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < 2 * TCB_B_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4 *>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    // K-direction (positions) core-matrix stride 128 B, MN-direction group stride TCB_SBO
    const uint32_t lbo = 128u, sbo = TCB_SBO;  // (measured: the fields are not interchangeable -- swapped, the result is garbage)
    const uint64_t descA = tcb_smem_desc(sG, lbo, sbo), descBh = tcb_smem_desc(sP, lbo, sbo),
                   descBl = tcb_smem_desc(sP + TCB_B_BYTES, lbo, sbo);
    const long long npos = B * NCELL, ntiles = (npos + TC_M - 1) / TC_M;
    uint32_t parity = 0;
    bool dead = false, any = false;
    const int g8 = tid & 7, psub = tid >> 3;
    TcRaw<XT> raw;
    if ((long long)blockIdx.x < ntiles) tc_load_raw<XT>(x, xstride, npos, blockIdx.x, tid, raw);

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- B tile: decoded inputs (as in the forward kernel, raw values loaded one tile ahead), then the 2x2 patch
        tc_decode_raw<XT>(raw, npos, tile, tid, sDec);
        if (tile + gridDim.x < ntiles) tc_load_raw<XT>(x, xstride, npos, tile + gridDim.x, tid, raw);
        const long long P = tile * TC_M + tid;
        const bool valid = P < npos;
        const long long b = valid ? P / NCELL : 0;
        const int pos = valid ? (int)(P - b * NCELL) : 0, m = pos / GS, n = pos - GS * m;
        rowinfo[tid] = valid ? (((b * (C1_OUT * C1_OUT) + 2 * m * C1_OUT + 2 * n) << 2) | (m == 16 ? 2 : 0) | (n == 16 ? 1 : 0)) : -1ll;
        __syncthreads();
        {
            const bool rgt = valid && n < 16, bot = valid && m < 16;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const uint4 d00 = valid ? sDec[tid] : z, d01 = rgt ? sDec[tid + 1] : z;
            const uint4 d10 = bot ? sDec[tid + GS] : z, d11 = (rgt && bot) ? sDec[tid + GS + 1] : z;
            const uint32_t off = (tid >> 3) * 128 + (tid & 7) * 16;
            *reinterpret_cast<uint4 *>(sP + off) = make_uint4(d00.x, d00.y, d01.x, d01.y);
            *reinterpret_cast<uint4 *>(sP + TCB_SBO + off) = make_uint4(d10.x, d10.y, d11.x, d11.y);
            *reinterpret_cast<uint4 *>(sP + 2 * TCB_SBO + off) = make_uint4(valid ? 0x3F80u : 0u, 0u, 0u, 0u);  // bf16 1.0: bias column
            *reinterpret_cast<uint4 *>(sP + TCB_B_BYTES + off) = make_uint4(d00.z, d00.w, d01.z, d01.w);
            *reinterpret_cast<uint4 *>(sP + TCB_B_BYTES + TCB_SBO + off) = make_uint4(d10.z, d10.w, d11.z, d11.w);
        }
        // ---- two halves: output phases (py = h, px = 0 / 1) ---------------------------------------------------
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            // dz chunks: lane group of 8 = the 128 bytes of one pixel; 16 positions per pass
            constexpr int NB = MK ? 8 : TCB_BATCH;  // with the bit mask a whole half tile (16 x 16-byte + 16 x 1-byte loads) is one batch
#pragma unroll 1
            for (int it0 = 0; it0 < 8; it0 += NB) {  // NB x 2 (x 2 without the bit mask) 16-byte loads in flight per thread
                uint4 gv[NB][2], yv[NB][2];
                uint32_t mbv[NB][2];
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    const long long inf = rowinfo[(it0 + u) * 16 + psub];
                    // position index = (b * 17 + m) * 17 + n: the merged planes / the bit mask are indexed by it directly
                    const long long Pq = tile * TC_M + (it0 + u) * 16 + psub;
#pragma unroll
                    for (int px = 0; px < 2; px++) {
                        // a pixel that does not exist reads pixel 0 and is zeroed through its y / mask (unconditional
                        // loads keep all of the batch in flight)
                        const bool ok = inf >= 0 && !(h && (inf & 2)) && !(px && (inf & 1));
                        const long long e = ok ? ((inf >> 2) + h * C1_OUT + px) * C1_CH + g8 * 8 : 0ll;
                        if constexpr (PL) {
                            const long long eg = ok ? (long long)Pq * pl_pos_stride + (h * 2 + px) * pl_cls_stride + g8 * 8 : 0ll;
                            gv[u][px] = tcb_ldg(planes + eg);
                        } else {
                            gv[u][px] = tcb_ldg(dy + e);
                        }
                        if constexpr (MK) {
                            // the mask word of the 32-channel half that holds channels 8*g8 .. 8*g8+7
                            mbv[u][px] = ok ? __ldg(relu_mask + (Pq * 4 + (h * 2 + px)) * 2 + (g8 >> 2)) : 0u;
                        } else {
                            yv[u][px] = tcb_ldg(y + e);
                            if (!ok) yv[u][px] = make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                }
                // every store's address depends on every load of the batch, so all of them are in flight together
                // (ptxas otherwise interleaves loads and stores to save registers: 3 dependent round trips per batch)
                uint32_t live = 0;
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    live |= gv[u][0].x | gv[u][1].x;
                    if constexpr (MK) live |= mbv[u][0] | mbv[u][1];
                    else live |= yv[u][0].x | yv[u][1].x;
                }
                const uint32_t nudge = live & zero;  // zero == 0 at run time
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    const int p = (it0 + u) * 16 + psub;
#pragma unroll
                    for (int px = 0; px < 2; px++) {
                        uint4 dzv;
                        if constexpr (MK) {  // mask bits -> all-ones per selected bf16 half: (bit | bit << 16) * 0xFFFF
                            // words 4*(g8&3) .. +3 of the half: bit q = even channel, bit 16+q = odd channel of word q
                            const uint32_t mb = mbv[u][px] >> (4 * (g8 & 3));
                            const uint4 g = gv[u][px];
                            dzv = make_uint4(g.x & ((mb & 0x00010001u) * 0xFFFFu), g.y & (((mb >> 1) & 0x00010001u) * 0xFFFFu),
                                             g.z & (((mb >> 2) & 0x00010001u) * 0xFFFFu), g.w & (((mb >> 3) & 0x00010001u) * 0xFFFFu));
                        } else {
                            dzv = tcb_mask(gv[u][px], yv[u][px]);
                        }
                        *reinterpret_cast<uint4 *>(sG + (px * 8 + g8) * TCB_SBO + (p >> 3) * 128 + (p & 7) * 16 + nudge) = dzv;
                    }
                }
            }
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) {   // (converged: one elected lane issues, see tc_mma_elect)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)(h * TCB_N);
#pragma unroll
                for (int ks = 0; ks < TC_M / 16; ks++) {  // 16 positions = 2 core matrices of 128 bytes per step
                    const uint64_t koff = (uint64_t)((ks * 256) >> 4);
                    tc_mma_elect(d, descA + koff, descBh + koff, TCB_IDESC, (any || ks) ? 1u : 0u);
                    tc_mma_elect(d, descA + koff, descBl + koff, TCB_IDESC, 1u);
                }
                tc_commit_elect(&bar);
            }
            if (!dead && !tc_mbar_wait(&bar, parity)) dead = true;   // sG / sP are free again
            parity ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        any = true;
        __syncthreads();
    }
    // ---- accumulators -> global: lane = row (phase pair member * 64 + channel), columns = 16 taps + bias -------
    if (any && !dead) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * TCB_N);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = (2 * h + (tid >> 6)) * C1_CH + (tid & 63);
#pragma unroll
            for (int k = 0; k < 16; k++) atomicAdd(dw4 + row * 16 + k, __uint_as_float(r[k]));
            atomicAdd(db4 + row, __uint_as_float(r[16]));
        }
    }
    if (dead && fail) atomicExch(fail, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCB_COLS));
}

}  // namespace ta
