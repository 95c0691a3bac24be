// ta_dgrad_tc.cuh -- data gradient of TINet's SECOND convolution (Conv2d(64, 64, 3, stride 2), all_net.py:144-145) on the
// 5th-generation tensor cores, one parity class at a time with its REAL tap list.
//
// The data gradient of a stride-2 convolution splits by the parity (pa, pb) of the input pixel (ta_conv1.cuh):
//     dx[2m+pa, 2n+pb, ci] = sum over taps ky = pa, kx = pb (mod 2), over co:  dz[m - ky/2, n - kx/2, co] * w[co, ci, ky, kx]
// For k = 3 the four classes have 4 / 2 / 2 / 1 taps: 9 tap-GEMMs per position instead of the 16 the merged 2x2 stride-1
// convolution (cuDNN, parity_class_weights_kernel) spends -- and the output is exactly the "merged planes" tensor
// [B][17][17][4][64] whose position-major rows conv1_bwd_tc_kernel reads in place.
//
//   GEMM       D_c[position, ci] = sum_taps A_shift(tap)[position, co] x W_tap[ci, co]^T      (M = 128, N = 64, K = 64 per tap)
//   A tiles    dz rows of the tile's 128 positions for the four shifts (dy, dx) in {0,1}^2 (zero rows outside the 16x16
//              map): a row = one pixel's 64 channels = 128 contiguous bytes of the channels-last dz; fetched with cp.async
//              (16-byte LDGSTS, zero-fill for missing pixels) straight into the K-major SWIZZLE_128B operand layout,
//              DOUBLE-BUFFERED: tile k+1's rows land while tile k is multiplied and written out
//   W          all 9 taps resident in shared memory for the CTA's lifetime (72 KB, pre-arranged by conv2_dgrad_prep_kernel)
//   D          four fp32 accumulators of 64 columns in TMEM (256 columns), tcgen05.mma issued by one thread, completion
//              through tcgen05.commit -> mbarrier (bounded wait, as in ta_conv1_tc.cuh)
//   epilogue   tcgen05.ld -> the ReLU mask of the layer below (conv1's bit mask, ta_conv1_fwd_mask) -> bf16 -> warp-private
//              swizzled staging -> stores that cover 8 positions x 64 contiguous bytes each
// One persistent CTA of 128 threads per SM (208 KB of shared memory).  Bounded by the 606 MB it writes per 4096 samples.
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no -lcuda)

#include "ta_conv1_tc.cuh"

namespace ta {

constexpr int DG_THREADS = 128;
constexpr int DG_TAPS = 9;
constexpr int DG_A_BYTES = 128 * 128;             // one shift of one tile: 128 rows x 64 bf16
constexpr int DG_WTAP_BYTES = 64 * 128;           // one tap: 64 rows (ci) x 64 bf16 (co)
constexpr int DG_W_BYTES = DG_TAPS * DG_WTAP_BYTES;
constexpr int DG_SMEM = DG_W_BYTES + 2 * 4 * DG_A_BYTES + 4 * 2048;   // 212992
constexpr int DG_COLS = 256;                      // TMEM columns: 4 classes x 64
constexpr int DG_OH = 16, DG_P = 17;              // dz map 16x16, planes 17x17 positions per sample

// K-major operands with the 128-byte swizzle (UMMA LayoutType::SWIZZLE_128B = 2, descriptor bits [61,64)): a row = 64 bf16
// = 128 contiguous bytes, 8-row groups of 1024 bytes (SBO), and inside a group the 16-byte chunk index is XORed with the
// row index: element (row r, chunk c) at r * 128 + ((c ^ (r % 8)) * 16).  A row's 128 bytes stay together, so a warp's
// cp.async of 4 rows x 8 chunks reads 4 full lines of global memory and writes 512 contiguous bytes of shared memory
// (the no-swizzle layout puts a row's chunks 128 bytes apart: one 16-byte piece per line, 32 lines per warp instruction --
// measured: the producer warp then needs 6.4k cycles per tile just to issue its copies).  The tile base is 1024-byte
// aligned; a K step of 16 elements advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t dg_smem_desc(const void *smem) {
    const uint64_t addr = (uint64_t)(smem_u32(smem) >> 4) & 0x3FFFu;
    return addr | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
constexpr int DG_KSTEP = 32 >> 4;   // descriptor start-address units (16 bytes) per K step of 16 bf16
__device__ __forceinline__ uint32_t dg_row_chunk(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// conv weight bf16 [co][ci][3][3] (any element strides) -> the shared-memory image of the 9 B operands (row = ci, K = co):
// out[tap = ky*3+kx][ci * 128 + (((co / 8) ^ (ci % 8)) * 16) + (co % 8) * 2 bytes] = w[co][ci][ky][kx]
__global__ void __launch_bounds__(256) conv2_dgrad_prep_kernel(const __nv_bfloat16 *__restrict__ w, long long so, long long si, long long sy,
                                                              long long sx, __nv_bfloat16 *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DG_TAPS * 64 * 64; i += gridDim.x * blockDim.x) {
        const int tap = i >> 12, rem = i & 4095;
        const int ci = rem >> 6, slot = (rem >> 3) & 7, e = rem & 7;   // image order: row (ci), stored chunk slot, element
        const int co = (slot ^ (ci & 7)) * 8 + e, ky = tap / 3, kx = tap - 3 * ky;
        out[i] = w[co * so + ci * si + ky * sy + kx * sx];
    }
}

__device__ __forceinline__ void dg_cp_async16(uint32_t dst_smem, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(DG_THREADS, 1) conv2_dgrad_planes_tc_kernel(const __nv_bfloat16 *__restrict__ dz, const uint4 *__restrict__ wimg,
                                                                          const uint32_t *__restrict__ relu_mask, long long B,
                                                                          __nv_bfloat16 *__restrict__ planes, int *fail) {
    extern __shared__ __align__(1024) uint8_t dg_smem[];
    uint8_t *sW = dg_smem, *sA = dg_smem + DG_W_BYTES, *sOut = sA + 2 * 4 * DG_A_BYTES;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *wout = sOut + warp * 2048;

    for (int i = tid; i < DG_W_BYTES / 16; i += DG_THREADS) reinterpret_cast<uint4 *>(sW)[i] = __ldg(wimg + i);
    if (tid == 0) tc_mbar_init(&bar);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(DG_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const long long npos = B * (DG_P * DG_P), ntiles = (npos + TC_M - 1) / TC_M;

    // the four shifted dz rows of this thread's position of `tile` -> sA[buf] (asynchronously)
    auto stage = [&](long long tile, int buf) {
        const long long P = tile * TC_M + tid;
        const bool valid = P < npos;
        const long long b = valid ? P / (DG_P * DG_P) : 0;
        const int pos = valid ? (int)(P - b * (DG_P * DG_P)) : 0, m = pos / DG_P, n = pos - DG_P * m;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int oy = m - (s >> 1), ox = n - (s & 1);
            const bool ok = valid && (unsigned)oy < (unsigned)DG_OH && (unsigned)ox < (unsigned)DG_OH;
            const __nv_bfloat16 *src = ok ? dz + ((b * DG_OH + oy) * DG_OH + ox) * 64 : dz;
            const uint32_t dst = smem_u32(sA + (buf * 4 + s) * DG_A_BYTES);
#pragma unroll
            for (int c = 0; c < 8; c++) dg_cp_async16(dst + dg_row_chunk(tid, c), src + c * 8, ok ? 16u : 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint32_t parity = 0;
    bool dead = false;
    int it = 0;
    if ((long long)blockIdx.x < ntiles) stage(blockIdx.x, 0);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int buf = it & 1;
        const bool more = tile + gridDim.x < ntiles;
        if (more) stage(tile + gridDim.x, buf ^ 1);   // (its buffer was last read by the MMAs of the previous iteration: completed)
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async();   // generic-proxy writes (cp.async) -> visible to the tensor core's async-proxy reads
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint64_t dA[4], dW[DG_TAPS];
#pragma unroll
            for (int s = 0; s < 4; s++) dA[s] = dg_smem_desc(sA + (buf * 4 + s) * DG_A_BYTES);
#pragma unroll
            for (int t = 0; t < DG_TAPS; t++) dW[t] = dg_smem_desc(sW + t * DG_WTAP_BYTES);
#pragma unroll
            for (int c = 0; c < 4; c++) {  // class (pa, pb): taps ky = pa, kx = pb (mod 2); tap (ky, kx) reads the rows shifted by (ky / 2, kx / 2)
                const int pa = c >> 1, pb = c & 1;
                bool first = true;
#pragma unroll
                for (int ky = 0; ky < 3; ky++)
#pragma unroll
                    for (int kx = 0; kx < 3; kx++) {
                        if ((ky & 1) != pa || (kx & 1) != pb) continue;
                        const int s = (ky >> 1) * 2 + (kx >> 1), t = ky * 3 + kx;
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {  // K = 16 per instruction: 32 bytes further on inside the swizzle atom
                            const uint64_t koff = (uint64_t)(ks * DG_KSTEP);
                            tc_mma(tmem_base + (uint32_t)(c * 64), dA[s] + koff, dW[t] + koff, first ? 0u : 1u);
                            first = false;
                        }
                    }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        if (!dead && !tc_mbar_wait(&bar, parity)) dead = true;
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: lane = position row of the tile; 4 classes x 2 halves of 32 channels ----------------------------
        const long long P = tile * TC_M + tid;
        const bool valid = P < npos;
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 64 + half * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // conv1's ReLU mask of output pixel (2m+pa, 2n+pb): word (P*4 + c)*2 + half, bit q / 16+q = channel 2q / 2q+1
                const uint32_t mb = relu_mask ? ((valid && !dead) ? __ldg(relu_mask + (P * 4 + c) * 2 + half) : 0u) : 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t o[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(r[8 * j + 2 * i]), __uint_as_float(r[8 * j + 2 * i + 1]));
                        o[i] = *reinterpret_cast<const uint32_t *>(&pk) & (((mb >> (4 * j + i)) & 0x00010001u) * 0xFFFFu);
                    }
                    *reinterpret_cast<uint4 *>(wout + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                __syncwarp();
                // transposed read-out: 4 lanes per row -> a store instruction covers 8 rows x 64 contiguous bytes
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int row = q * 8 + (lane >> 2), j = lane & 3;
                    const uint4 v = *reinterpret_cast<const uint4 *>(wout + row * 64 + ((j ^ ((row >> 1) & 3)) << 4));
                    const long long Pr = tile * TC_M + warp * 32 + row;
                    if (Pr < npos && !dead) *reinterpret_cast<uint4 *>(planes + ((Pr * 4 + c) * 64 + half * 32 + j * 8)) = v;
                }
                __syncwarp();
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // TMEM and sA[buf] are free again
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (dead && fail) atomicExch(fail, 1);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(DG_COLS));
}

// ---- the same computation, WARP-SPECIALISED ------------------------------------------------------------------------------
// The kernel above is correct but latency-bound: one CTA of four warps per SM runs stage -> MMA -> epilogue strictly one
// after the other (measured 401 us per 4096 samples against 188 us for cuDNN's merged-plane convolution).  Here the three
// activities are three roles that only meet through mbarriers -- no CTA-wide barrier inside the tile loop:
//   warps 4-7   producers: cp.async of the next tile's four shifted A tiles into a 2-stage ring (32 rows per warp; one warp
//               alone spends 7k cycles per tile on the address arithmetic of its 4096 copies); completion is reported with
//               cp.async.mbarrier.arrive.noinc on a_full[stage], the stage is reused after a_empty[stage]
//   warp 8      MMA issuer (one lane): waits a_full[stage] and acc_empty[buf], issues the 36 tcgen05.mma of the tile into
//               TMEM accumulator buffer buf (2 x 256 columns), then tcgen05.commit -> a_empty[stage] and acc_full[buf]
//   warps 0-3   epilogue: warp w owns TMEM lanes 32w.. = 32 consecutive positions; per class it reads the 64 columns, applies
//               the ReLU mask, packs to bf16 and writes its rows (128 bytes each) into a 4 KB staging block in the TMA's
//               128-byte swizzle, which leaves with ONE tensor-map store (cp.async.bulk.tensor); then arrives on acc_empty[buf].
//               (Measured alternatives: a linear staging block + plain bulk store, or 16-byte stores straight from the
//               registers of eight epilogue warps -- both put the 32 lanes of every store instruction on 32 different 128-byte
//               rows, i.e. 32 bank conflicts / 32 cache lines per instruction, ~4k cycles per tile: 253 / 267 us.)
// so tile k+1 is multiplied while tile k is written out and tile k+2 is being fetched.  Every wait is bounded.
// Output layout: CLASS-major planes [4][B*289 positions][64] (a warp's 32 rows of one class are 4 KB contiguous), which
// conv1_bwd_tc_kernel reads through its plane-stride arguments.
constexpr int DGW_THREADS = 288;   // warps 0-3 epilogue, 4-7 producers, 8 MMA issuer
constexpr int DGW_SMEM = DG_W_BYTES + 2 * 4 * DG_A_BYTES + 4 * 4096;   // 221184
constexpr int DGW_COLS = 512;

__device__ __forceinline__ void dg_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dg_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(DGW_THREADS, 1) conv2_dgrad_planes_ws_kernel(const __nv_bfloat16 *__restrict__ dz, const uint4 *__restrict__ wimg,
                                                                              const uint32_t *__restrict__ relu_mask, long long B,
                                                                              const __grid_constant__ CUtensorMap planes_map, int *fail,
                                                                              long long *prof) {
    // prof (nullable, development): cycles CTA 0's roles spend waiting -- [0] producer on a_empty, [1] producer total, [2] MMA on
    // a_full, [3] MMA on acc_empty, [4] MMA total, [5] epilogue warp 0 on acc_full, [6] on its staging block, [7] total
    extern __shared__ __align__(1024) uint8_t dg_smem[];
    uint8_t *sW = dg_smem, *sA = dg_smem + DG_W_BYTES, *sOut = sA + 2 * 4 * DG_A_BYTES;
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2];
    const bool pr = prof != nullptr && blockIdx.x == 0;
    long long t_w0 = 0, t_w1 = 0;
    const long long t_start = clock64();
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < DG_W_BYTES / 16; i += DGW_THREADS) reinterpret_cast<uint4 *>(sW)[i] = __ldg(wimg + i);
    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            dg_mbar_init(&a_full[s], 128);     // the four producer warps' lanes (noinc arrivals of their cp.async groups)
            dg_mbar_init(&a_empty[s], 1);      // one tcgen05.commit
            dg_mbar_init(&acc_full[s], 1);     // one tcgen05.commit
            dg_mbar_init(&acc_empty[s], 4);    // the four epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(DGW_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const long long npos = B * (DG_P * DG_P), ntiles = (npos + TC_M - 1) / TC_M;
    bool dead = false;

    if (warp >= 4 && warp < 8) {
        // ---------------- producers: warp 4 + w fetches rows 32w .. 32w+31 of every tile ----------------
        const int pw = warp - 4;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles && !dead; tile += gridDim.x, it++) {
            const int s = it & 1, n = it >> 1;
            const long long tw = clock64();
            if (!tc_mbar_wait(&a_empty[s], (uint32_t)((n & 1) ^ 1))) { dead = true; break; }   // (passes at once on the first use)
            t_w0 += clock64() - tw;
            // lane = (row within a quad, 16-byte chunk): one warp instruction copies 4 rows x 128 contiguous bytes
            const int rsub = lane >> 3, ch = lane & 7;
            const uint32_t sbase = smem_u32(sA + s * 4 * DG_A_BYTES);
#pragma unroll 4
            for (int i = 0; i < 8; i++) {
                const int row = pw * 32 + 4 * i + rsub;
                const unsigned P = (unsigned)(tile * TC_M) + (unsigned)row;     // (batch * 289 < 2^31, checked by the host)
                const bool valid = (long long)P < npos;
                const unsigned b = P / (unsigned)(DG_P * DG_P), pos = P - b * (unsigned)(DG_P * DG_P), m = pos / (unsigned)DG_P, nn = pos - m * (unsigned)DG_P;
                // the pixel (m, n) of the 16x16 map (it may lie outside); shift (dy, dx) reads the pixel dy rows / dx columns before it
                const __nv_bfloat16 *base = dz + ((long long)((b * DG_OH + m) * DG_OH + nn)) * 64 + ch * 8;
                const uint32_t doff = dg_row_chunk(row, ch);
#pragma unroll
                for (int sh = 0; sh < 4; sh++) {
                    const int dy = sh >> 1, dx = sh & 1;
                    const bool ok = valid && m >= (unsigned)dy && m - dy < (unsigned)DG_OH && nn >= (unsigned)dx && nn - dx < (unsigned)DG_OH;
                    dg_cp_async16(sbase + sh * DG_A_BYTES + doff, ok ? base - (dy * DG_OH + dx) * 64 : dz, ok ? 16u : 0u);
                }
            }
            // this lane's copies arrive on a_full[s] when they have landed
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&a_full[s])) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (pr && pw == 0 && lane == 0) { prof[0] = t_w0; prof[1] = clock64() - t_start; }
    } else if (warp == 8) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            uint64_t dW[DG_TAPS];
#pragma unroll
            for (int t = 0; t < DG_TAPS; t++) dW[t] = dg_smem_desc(sW + t * DG_WTAP_BYTES);
            int it = 0;
            for (long long tile = blockIdx.x; tile < ntiles && !dead; tile += gridDim.x, it++) {
                const int s = it & 1, n = it >> 1;
                long long tw = clock64();
                if (!tc_mbar_wait(&a_full[s], (uint32_t)(n & 1))) { dead = true; break; }
                t_w0 += clock64() - tw;
                tw = clock64();
                if (!tc_mbar_wait(&acc_empty[s], (uint32_t)((n & 1) ^ 1))) { dead = true; break; }
                t_w1 += clock64() - tw;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                fence_proxy_async();   // the producer's generic-proxy writes, acquired through the barrier -> the tensor core's async-proxy reads
                uint64_t dA[4];
#pragma unroll
                for (int sh = 0; sh < 4; sh++) dA[sh] = dg_smem_desc(sA + (s * 4 + sh) * DG_A_BYTES);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int pa = c >> 1, pb = c & 1;
                    bool first = true;
#pragma unroll
                    for (int ky = 0; ky < 3; ky++)
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) {
                            if ((ky & 1) != pa || (kx & 1) != pb) continue;
                            const int sh = (ky >> 1) * 2 + (kx >> 1), t = ky * 3 + kx;
#pragma unroll
                            for (int ks = 0; ks < 4; ks++) {
                                const uint64_t koff = (uint64_t)(ks * DG_KSTEP);
                                tc_mma(tmem_base + (uint32_t)(s * 256 + c * 64), dA[sh] + koff, dW[t] + koff, first ? 0u : 1u);
                                first = false;
                            }
                        }
                }
                // both fire when the MMAs above have completed: the A stage may be refilled, the accumulators may be read
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&a_empty[s])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&acc_full[s])) : "memory");
            }
            if (pr) { prof[2] = t_w0; prof[3] = t_w1; prof[4] = clock64() - t_start; }
        }
    } else {
        // ---------------- epilogue (warps 0-3): warp w reads TMEM lanes 32w.. = 32 consecutive positions, all four classes ----------------
        uint8_t *stage = sOut + warp * 4096;
        const uint32_t stage_u = smem_u32(stage);
        const uint64_t map_u = reinterpret_cast<uint64_t>(&planes_map);
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles && !dead; tile += gridDim.x, it++) {
            const int s = it & 1, n = it >> 1;
            const long long P0 = tile * TC_M + warp * 32, P = P0 + lane;
            // conv1's ReLU mask words of this lane's position (4 classes x 2 halves = 32 contiguous bytes), fetched BEFORE the wait
            // for the accumulators
            uint4 mw[2] = {make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu)};
            if (relu_mask && P < npos) {
                mw[0] = __ldg(reinterpret_cast<const uint4 *>(relu_mask + P * 8));
                mw[1] = __ldg(reinterpret_cast<const uint4 *>(relu_mask + P * 8) + 1);
            }
            long long tw = clock64();
            if (!tc_mbar_wait(&acc_full[s], (uint32_t)(n & 1))) { dead = true; break; }
            t_w0 += clock64() - tw;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t r[2][32];   // both halves of the class (64 columns) are requested before the wait
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * 256 + c * 64 + half * 32);
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
                uint4 o[8];   // the lane's row of class c: 64 bf16 = 8 chunks of 16 bytes
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const uint4 m4 = mw[c >> 1];
                    const uint32_t mb = (c & 1) ? (half ? m4.w : m4.z) : (half ? m4.y : m4.x);
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
                // the staging block is free once the previous class's store has READ it
                tw = clock64();
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                t_w1 += clock64() - tw;
                // 32 rows x 128 bytes in the TMA's SWIZZLE_128B order (chunk ^ row % 8): one warp store instruction touches all
                // 32 banks four times -- the minimum for 512 bytes -- where the linear order (row stride 128 bytes) would put all 32
                // lanes on the same four banks
#pragma unroll
                for (int k = 0; k < 8; k++) *reinterpret_cast<uint4 *>(stage + dg_row_chunk(lane, k)) = o[k];
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    // rows beyond the last position are clipped by the tensor map's bounds
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map_u), "r"(0),
                                 "r"((int)P0), "r"(c), "r"(stage_u)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) dg_mbar_arrive(&acc_empty[s]);   // this warp has read its quarter of accumulator buffer s
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (pr && warp == 0 && lane == 0) { prof[5] = t_w0; prof[6] = t_w1; prof[7] = clock64() - t_start; }
    }
    if (dead && fail) atomicExch(fail, 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(DGW_COLS));
}

}  // namespace ta
