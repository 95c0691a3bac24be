// ta_gae.cuh -- advantage / returns over a [T][n] rollout (fp32).
//
// Replaces soa/agent/PPO.py:112-115
//     target_v = r + gamma * V(s');  adv = target_v - V(s)      (no done mask, no normalisation)
// which is the lambda = 0 / use_mask = 0 mode and is reproduced BIT-EXACTLY (separate
// multiply and add, no FMA contraction, like the two torch ops).  General mode:
//     ret_t = r_t + gamma*nd_t*(V_{t+1} + lambda*A_{t+1}),   A_t = ret_t - V_t
//
// Parallelisation: the recurrence A_t = delta_t + c_t*A_{t+1} (c_t = gamma*lambda*nd_t) is a
// linear scan over time.  A CTA owns 32 envs (threadIdx.x, coalesced across the env axis) and
// splits time into chunks of 32 steps (threadIdx.y): every thread scans its chunk backwards
// in registers with A_in = 0 while keeping the chunk's carry product, the chunk summaries are
// combined through shared memory, and the thread then adds (product so far)*A_in to its 32
// values.  Every input is read exactly once and every load is independent of the recurrence.
#pragma once
#include "ta_common.cuh"

namespace ta {

constexpr int GAE_L = 32;   // steps per thread
constexpr int GAE_CH = 8;   // max chunks per CTA pass (256 steps per pass)

__global__ void __launch_bounds__(32 * GAE_CH)
gae_kernel(const float *__restrict__ r, const float *__restrict__ v, const float *__restrict__ v_next,
           const float *__restrict__ last_v, const uint8_t *__restrict__ done, float gamma, float lam, int use_mask,
           int T, long long n, float *__restrict__ adv, float *__restrict__ ret) {
    __shared__ float sP[GAE_CH][32], sA[GAE_CH][32];
    const int lx = threadIdx.x, cy = threadIdx.y, CH = blockDim.y;
    const long long env = (long long)blockIdx.x * 32 + lx;
    const bool valid = env < n;
    const int span = GAE_L * CH;
    const float gl = __fmul_rn(gamma, lam);
    float carry = 0.0f;
    for (int base = ((T - 1) / span) * span; base >= 0; base -= span) {
        const int t0 = base + cy * GAE_L;
        float tv[GAE_L], cc[GAE_L], A[GAE_L];
        // loads (all independent of the scan)
        float rr[GAE_L], vv[GAE_L], vn[GAE_L];
        uint32_t dn = 0;
#pragma unroll
        for (int j = 0; j < GAE_L; j++) {
            const int t = t0 + j;
            const bool in = valid && t < T;
            const long long idx = (long long)t * n + env;
            rr[j] = in ? r[idx] : 0.0f;
            vv[j] = in ? v[idx] : 0.0f;
            if (v_next) vn[j] = in ? v_next[idx] : 0.0f;
            else vn[j] = in ? (t == T - 1 ? last_v[env] : v[idx + n]) : 0.0f;
            if (use_mask && in && done[idx]) dn |= 1u << j;
        }
        float acc = 0.0f, P = 1.0f;
#pragma unroll
        for (int j = GAE_L - 1; j >= 0; j--) {
            const bool in = valid && (t0 + j) < T;
            const bool nd = !((dn >> j) & 1u);
            // r + gamma*V' as two rounded ops (torch: mul kernel then add kernel)
            const float tvj = __fadd_rn(rr[j], nd ? __fmul_rn(gamma, vn[j]) : 0.0f);
            const float cj = in ? (nd ? gl : 0.0f) : 1.0f;
            const float dj = in ? __fsub_rn(tvj, vv[j]) : 0.0f;
            acc = __fadd_rn(dj, __fmul_rn(cj, acc));
            P *= cj;
            tv[j] = tvj; cc[j] = cj; A[j] = acc;
        }
        sP[cy][lx] = P;
        sA[cy][lx] = acc;
        __syncthreads();
        float ain = carry;
        for (int k = CH - 1; k > cy; k--) ain = sA[k][lx] + sP[k][lx] * ain;
        float cnext = ain;
        for (int k = cy; k >= 0; k--) cnext = sA[k][lx] + sP[k][lx] * cnext;
        float Pj = 1.0f, anext = ain;
#pragma unroll
        for (int j = GAE_L - 1; j >= 0; j--) {
            const int t = t0 + j;
            if (valid && t < T) {
                Pj *= cc[j];
                const float at = __fadd_rn(A[j], __fmul_rn(Pj, ain));
                const long long idx = (long long)t * n + env;
                adv[idx] = at;
                ret[idx] = __fadd_rn(tv[j], __fmul_rn(cc[j], anext));
                anext = at;
            }
        }
        __syncthreads();
        carry = cnext;
    }
}

// (sum, sum of squares, count) of adv in float64 -- PPO.py:115's mean / std ingredients.
__global__ void __launch_bounds__(256) adv_stats_kernel(const float *__restrict__ adv, long long count, double *stats) {
    double s = 0.0, ss = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const double x = (double)adv[i];
        s += x;
        ss += x * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    }
    __shared__ double ws[8], wss[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { ws[w] = s; wss[w] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < 8; k++) { a += ws[k]; b += wss[k]; }
        atomicAdd(&stats[0], a);
        atomicAdd(&stats[1], b);
        if (blockIdx.x == 0) atomicAdd(&stats[2], (double)count);
    }
}

// adv <- (adv - mean) / (std + 1e-8), unbiased std (torch.Tensor.std default)
__global__ void __launch_bounds__(256) adv_normalize_kernel(float *__restrict__ adv, long long count, const double *stats) {
    const double cnt = stats[2], mean = stats[0] / cnt;
    double var = (stats[1] - stats[0] * mean) / (cnt > 1.0 ? cnt - 1.0 : 1.0);
    var = var < 0.0 ? 0.0 : var;
    const float m = (float)mean, inv = 1.0f / ((float)sqrt(var) + 1e-8f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        adv[i] = (adv[i] - m) * inv;
}

}  // namespace ta
