// ta_common.cuh -- shared device-side definitions for the Twoarmy sm_100a kernels.
//
// HBM layout of one env batch (struct of arrays, owned by the handle):
//   grid  uint32 [Npad][20]   the 17x17 grid packed 2 bits per cell into 80 bytes (packed uint8,
//                             16 cells per word): cell codes 0 empty 1 wall 2 ball 3 goal,
//                             COLUMN-major inside an env (cell c = x*17+y sits at bits
//                             [2c,2c+1]; cells 289..319 are padding and always hold the wall code) so that one column of the
//                             egocentric view is a contiguous bit run.  Records are 16-byte
//                             aligned and move to / from shared memory with TMA bulk copies.
//   sc0   uint4 [Npad]        .x = agent_x | agent_y<<8 | flags<<16 | risk_count<<24
//                             .y = step_count  .z = step_move  .w = t (steps since creation)
//   sc1   uint4 [Npad]        ball positions, 10 bits each (x | y<<5, 0x3FF = cur_pos None):
//                             .x = obstacles[0..2]  .y = obstacles1[0..2]  .z = obstacles2[0..2]
//                             .w = obstacles2[3] | error<<16
// Npad = n_envs rounded up to 32, so state tiles are always whole.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ta {

constexpr int GS = 17;
constexpr int NCELL = 289;
constexpr int TILE = 32;        // envs per tile
constexpr int REC_WORDS = 20;   // packed grid record: 20 x uint32 = 80 B = 320 cell slots
constexpr int REC_CELLS = 320;
constexpr int REC_BYTES = 80;

constexpr uint32_t C_EMPTY = 0, C_WALL = 1, C_BALL = 2, C_GOAL = 3;
constexpr uint32_t NOPOS = 0x3FFu;

// flag bits (same values as TA_F_* in include/twoarmy_b200.h)
constexpr uint32_t F_PONE = 1, F_PATROL = 2, F_UP1 = 4, F_RIGHT2 = 8, F_UPD_H = 16, F_UPD_L = 32,
                   F_FIRST = 64;
// Twoarmy_v4.__init__ (twoarmy_v4.py:14-24): right2, Update_longitudinal, first_to_room2 set
constexpr uint32_t FLAGS_INIT = F_RIGHT2 | F_UPD_L | F_FIRST;

constexpr uint32_t ERR_BAD_ACTION = 1, ERR_NONE_POS = 2, ERR_OOB_MOVE = 4;

// reward literals of twoarmy_v4.py:180,229,240,284,295 cast to float32
enum { R_STEP = 0, R_RISK = 1, R_HIT = 2, R_ROOM2 = 3, R_GOAL = 4 };
__device__ __forceinline__ float reward_value(int idx) {
    // bit patterns of float32(-0.01, -0.1, -0.9, 0.2, 0.9); selected without a jump table
    uint32_t b = 0xBC23D70Au;
    b = idx == R_RISK ? 0xBDCCCCCDu : b;
    b = idx == R_HIT ? 0xBF666666u : b;
    b = idx == R_ROOM2 ? 0x3E4CCCCDu : b;
    b = idx == R_GOAL ? 0x3F666666u : b;
    return __uint_as_float(b);
}

__host__ __device__ __forceinline__ uint32_t pack_pos(int x, int y) { return (uint32_t)x | ((uint32_t)y << 5); }
__host__ __device__ __forceinline__ uint32_t ball_get(uint32_t w, int k) { return (w >> (10 * k)) & 0x3FFu; }
__host__ __device__ __forceinline__ int pos_x(uint32_t p) { return (int)(p & 31u); }
__host__ __device__ __forceinline__ int pos_y(uint32_t p) { return (int)(p >> 5); }
__host__ __device__ __forceinline__ bool inb(int x, int y) { return (unsigned)x < (unsigned)GS && (unsigned)y < (unsigned)GS; }

// packed-cell accessors on one env record (20 words)
__host__ __device__ __forceinline__ uint32_t cell_get(const uint32_t *rec, int x, int y) {
    const int c = x * GS + y;
    return (rec[c >> 4] >> (2 * (c & 15))) & 3u;
}
__host__ __device__ __forceinline__ void cell_set(uint32_t *rec, int x, int y, uint32_t code) {
    const int c = x * GS + y, sh = 2 * (c & 15);
    rec[c >> 4] = (rec[c >> 4] & ~(3u << sh)) | (code << sh);
}

// _gen_grid (twoarmy_v4.py:38-80) as a pure function of the cell
__host__ __device__ __forceinline__ uint32_t initial_cell(int x, int y) {
    if (x == 0 || y == 0 || x == GS - 1 || y == GS - 1) return C_WALL;
    if (y == 8) {
        if (x <= 5 || x >= 11) return C_WALL;
        if (x >= 7 && x <= 9) return C_BALL;
        return C_EMPTY;
    }
    if (x == 14 && y == 2) return C_GOAL;
    return C_EMPTY;
}
constexpr uint32_t MID_INIT = (7u | (8u << 5)) | ((8u | (8u << 5)) << 10) | ((9u | (8u << 5)) << 20);
constexpr uint32_t ALL_NONE3 = 0x3FFFFFFFu;

// Philox4x32-10 (Random123); one block per (env, step). Draw contract:
//   key = (seed lo, seed hi), ctr = (global env id lo, hi, t, 0)
//   slot 0 (choice of 10) = mulhi(w0,10); slots 1,2,3 (choice of 4) = (w1 >> 0,2,4)&3 + lo;
//   slots 5,6 (choice of 2) = (w2 >> 0,1)&1.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---- PTX wrappers: TMA bulk stores (cp.async.bulk -> SASS UBLKCP) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared -> global, completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
// programmatic dependent launch (PTX griddepcontrol; no-ops when the launch has no such edge)
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace ta
