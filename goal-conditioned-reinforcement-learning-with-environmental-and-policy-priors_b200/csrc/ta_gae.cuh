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

// Vectorised form (n % 4 == 0, 16-byte aligned arrays): a thread owns FOUR adjacent envs (one
// float4 column) and GV_L consecutive steps, a CTA 128 envs x (GV_L x blockDim.y) steps.  Per
// element it keeps only tv_t = r_t + gamma*nd_t*V_{t+1} and delta_t = tv_t - V_t (+ the nd bits);
// pass 1 reduces the chunk to (prod c, A with A_in = 0), the chunk summaries are chained through
// shared memory, pass 2 re-runs the recurrence A_t = delta_t + c_t*A_{t+1} from the true A_in and
// writes adv_t = A_t, ret_t = tv_t + c_t*A_{t+1} as 16-byte stores.  V_{t+1} comes from the row
// already in registers, so every input byte is read once.
constexpr int GV_MAX_THREADS = 1024;

__device__ __forceinline__ float f4get(const float4 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// (sum, sum of squares) of a CTA in a fixed order: every thread's pair goes to shared memory, the first (up to) 32 threads
// add the entries tid, tid + 32, ... in that order, then a xor-shuffle tree (full first warp) or thread 0 alone (CTAs of fewer
// than 32 threads) finishes; the result is left in red[0], red[1] (valid for thread 0).
__device__ __forceinline__ void gae_cta_moments(double s, double ss, double *red, int tid, int nth) {
    __syncthreads();   // (red may still be read from an earlier use)
    red[2 * tid] = s;
    red[2 * tid + 1] = ss;
    __syncthreads();
    if (nth >= 32) {
        if (tid < 32) {
            double a = 0.0, b = 0.0;
            for (int k = tid; k < nth; k += 32) { a += red[2 * k]; b += red[2 * k + 1]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
                b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
            }
            if (tid == 0) { red[0] = a; red[1] = b; }
        }
    } else if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < nth; k++) { a += red[2 * k]; b += red[2 * k + 1]; }
        red[0] = a;
        red[1] = b;
    }
}

// GV_X = float4 columns per CTA (threadIdx.x): 32 for large batches; 8 (one 128-byte line per row) when the rollout is
// small -- BASELINE configs[3] is 128 x 16384 = 36 MB -- so that there are several CTAs per SM and one pass covers all of
// T (16 chunks x 8 steps): the launch is then bound by one DRAM round trip instead of two dependent ones.
// stats (nullable): the launch also accumulates (sum, sum of squares, count) of adv in float64 -- the moments of
// PPO.py:115's normalisation -- so that normalising costs one more pass over adv instead of two.
// COOP: the single-launch normalising form (sync = the ticket word, followed at sync + 10 by two float64 slots per CTA).
template <int GV_L, int GV_CH, int GV_X, bool COOP = false>
__global__ void __launch_bounds__(GV_X * GV_CH, 512 / (GV_X * GV_CH))
gae_vec4_kernel(const float *__restrict__ r, const float *__restrict__ v, const float *__restrict__ v_next,
                const float *__restrict__ last_v, const uint8_t *__restrict__ done, float gamma, float lam, int use_mask,
                int T, long long n, float *__restrict__ adv, float *__restrict__ ret, double *__restrict__ stats,
                unsigned int *__restrict__ sync) {
    // chunk summaries [chunk][env of the quad][quad] (+ GV_PAD floats per chunk row: the warps of a narrow CTA hold
    // several chunks, whose rows would otherwise fall on the same banks)
    constexpr int GV_PAD = GV_X < 32 ? GV_X : 0, GV_ROW = 4 * GV_X + GV_PAD;
    __shared__ float sP[GV_CH * GV_ROW], sA[GV_CH * GV_ROW], sC[4 * GV_X];
    __shared__ double red[2 * GV_X * GV_CH];   // per-thread partial moments (stats launches)
    const int lx = threadIdx.x, cy = threadIdx.y, CH = blockDim.y;
    const int tid = cy * GV_X + lx, nth = GV_X * CH;
    for (int sidx = tid; sidx < 4 * GV_X; sidx += nth) sC[sidx] = 0.0f;   // (same thread, same slots as in the scan below)
    const long long col = ((long long)blockIdx.x * GV_X + lx) * 4;  // first of this thread's 4 envs
    double st_s = 0.0, st_ss = 0.0;
    const bool valid = col < n;
    const int span = GV_L * CH;
    const float gl = __fmul_rn(gamma, lam);
    for (int base = ((T - 1) / span) * span; base >= 0; base -= span) {
        const int t0 = base + cy * GV_L;
        float tv[GV_L][4], dl[GV_L][4];
        uint32_t ndbits = 0;  // bit 4*j+i: element (j, i) is not done (or mask off)
        {
            float4 vv[GV_L + 1];
            float4 rr[GV_L];
            uint32_t dd[GV_L];
#pragma unroll
            for (int j = 0; j < GV_L; j++) {
                const int t = t0 + j;
                const bool in = valid && t < T;
                const long long idx = (long long)t * n + col;
                rr[j] = in ? __ldg(reinterpret_cast<const float4 *>(r + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
                vv[j] = in ? __ldg(reinterpret_cast<const float4 *>(v + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
                dd[j] = (use_mask && in) ? __ldg(reinterpret_cast<const uint32_t *>(done + idx)) : 0u;
            }
            float4 vn[GV_L];
            if (v_next) {
#pragma unroll
                for (int j = 0; j < GV_L; j++) {
                    const int t = t0 + j;
                    vn[j] = (valid && t < T) ? __ldg(reinterpret_cast<const float4 *>(v_next + (long long)t * n + col))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                const int tl = t0 + GV_L;  // the row after this chunk
                vv[GV_L] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && tl < T) vv[GV_L] = __ldg(reinterpret_cast<const float4 *>(v + (long long)tl * n + col));
                const float4 lv = valid ? __ldg(reinterpret_cast<const float4 *>(last_v + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < GV_L; j++) vn[j] = (t0 + j == T - 1) ? lv : vv[j + 1];
            }
#pragma unroll
            for (int j = 0; j < GV_L; j++) {
                const bool in = valid && (t0 + j) < T;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const bool nd = !((dd[j] >> (8 * i)) & 0xFFu);
                    // r + gamma*V' as two rounded ops (torch: mul kernel then add kernel)
                    const float tvj = __fadd_rn(f4get(rr[j], i), nd ? __fmul_rn(gamma, f4get(vn[j], i)) : 0.0f);
                    tv[j][i] = tvj;
                    dl[j][i] = in ? __fsub_rn(tvj, f4get(vv[j], i)) : 0.0f;
                    if (nd) ndbits |= 1u << (4 * j + i);
                }
            }
        }
        // pass 1: chunk summary
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float acc = 0.0f, P = 1.0f;
#pragma unroll
            for (int j = GV_L - 1; j >= 0; j--) {
                const bool in = valid && (t0 + j) < T;
                const float cj = in ? (((ndbits >> (4 * j + i)) & 1u) ? gl : 0.0f) : 1.0f;
                acc = __fadd_rn(dl[j][i], __fmul_rn(cj, acc));
                P *= cj;
            }
            sP[cy * GV_ROW + i * GV_X + lx] = P;
            sA[cy * GV_ROW + i * GV_X + lx] = acc;
        }
        __syncthreads();
        // the 4 * GV_X independent suffix scans over the chunks (one per env column of the CTA), each by ONE thread:
        // chunk k's slot receives A entering chunk k from the later chunks (conflict-free rows, no re-reading by
        // every thread of the column -- that serial re-fold was a third of the small-rollout launch)
        for (int sidx = tid; sidx < 4 * GV_X; sidx += nth) {   // (one scan per thread unless T is very short)
            float a = sC[sidx];
#pragma unroll 8
            for (int k = GV_CH - 1; k >= 0; k--) {
                if (k < CH) {
                    const float pk = sP[k * GV_ROW + sidx], ak = sA[k * GV_ROW + sidx];
                    sA[k * GV_ROW + sidx] = a;
                    a = ak + pk * a;
                }
            }
            sC[sidx] = a;   // A carried into the next pass (earlier time steps)
        }
        __syncthreads();
        float ain[4];
#pragma unroll
        for (int i = 0; i < 4; i++) ain[i] = sA[cy * GV_ROW + i * GV_X + lx];
        // pass 2: the recurrence from the true A_in, outputs
        float A[4] = {ain[0], ain[1], ain[2], ain[3]};
#pragma unroll
        for (int j = GV_L - 1; j >= 0; j--) {
            const int t = t0 + j;
            const bool in = valid && t < T;
            float ao[4], ro[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float cj = in ? (((ndbits >> (4 * j + i)) & 1u) ? gl : 0.0f) : 1.0f;
                const float ca = __fmul_rn(cj, A[i]);
                ro[i] = __fadd_rn(tv[j][i], ca);
                ao[i] = __fadd_rn(dl[j][i], ca);
                A[i] = ao[i];
            }
            if (COOP) {   // normalising launch: adv stays in registers until the grid's moments are known
#pragma unroll
                for (int i = 0; i < 4; i++) dl[j][i] = ao[i];
            }
            if (in) {
                const long long idx = (long long)t * n + col;
                if (!COOP) *reinterpret_cast<float4 *>(adv + idx) = make_float4(ao[0], ao[1], ao[2], ao[3]);
                *reinterpret_cast<float4 *>(ret + idx) = make_float4(ro[0], ro[1], ro[2], ro[3]);
                if (stats) {
                    st_s += (double)ao[0] + (double)ao[1] + (double)ao[2] + (double)ao[3];
                    st_ss += (double)ao[0] * ao[0] + (double)ao[1] * ao[1] + (double)ao[2] * ao[2] + (double)ao[3] * ao[3];
                }
            }
        }
        if (COOP) {
            // ---- single-launch normalisation (the launcher guarantees: one pass, every CTA of the grid resident) --------
            // per-CTA moments -> the CTA's own slot (no same-address float64 atomics: 512 CTAs finishing together spent
            // ~5 us queueing on them) -> grid barrier on a ticket -> EVERY CTA adds all slots in the same fixed order ->
            // (adv - mean) / (std + 1e-8) from the registers (adv_normalize_kernel's arithmetic, PPO.py:115): adv is
            // written once, never re-read, and the result does not depend on the order CTAs finish in
            double *slots = reinterpret_cast<double *>(sync) + 5;   // (sync = work + 3: slots start at work + 8)
            gae_cta_moments(st_s, st_ss, red, tid, nth);
            __shared__ float s_norm[2];
            if (tid == 0) {
                __stcg(slots + 2 * blockIdx.x, red[0]);
                __stcg(slots + 2 * blockIdx.x + 1, red[1]);
                __threadfence();
                atomicAdd(sync, 1u);
                for (int spin = 0; spin < (1 << 24); spin++)   // bounded: a grid that is not co-resident must not hang
                    if (*reinterpret_cast<volatile unsigned int *>(sync) >= gridDim.x) break;
                __threadfence();
            }
            __syncthreads();
            double gs = 0.0, gss = 0.0;
            for (int k = tid; k < (int)gridDim.x; k += nth) { gs += __ldcg(slots + 2 * k); gss += __ldcg(slots + 2 * k + 1); }
            gae_cta_moments(gs, gss, red, tid, nth);
            if (tid == 0) {
                const double s0 = red[0], s1 = red[1], cnt = (double)T * (double)n;
                const double mean = s0 / cnt;
                double var = (s1 - s0 * mean) / (cnt > 1.0 ? cnt - 1.0 : 1.0);
                var = var < 0.0 ? 0.0 : var;
                s_norm[0] = (float)mean;
                s_norm[1] = 1.0f / ((float)sqrt(var) + 1e-8f);
                if (blockIdx.x == 0) { stats[0] = s0; stats[1] = s1; stats[2] = cnt; }
            }
            __syncthreads();
            const float m = s_norm[0], inv = s_norm[1];
#pragma unroll
            for (int j = 0; j < GV_L; j++) {
                const int t = t0 + j;
                if (valid && t < T)
                    *reinterpret_cast<float4 *>(adv + (long long)t * n + col) =
                        make_float4((dl[j][0] - m) * inv, (dl[j][1] - m) * inv, (dl[j][2] - m) * inv, (dl[j][3] - m) * inv);
            }
        }
        __syncthreads();
    }
    if (stats && !COOP) {  // CTA-wide sums, then one atomic pair per CTA
        gae_cta_moments(st_s, st_ss, red, tid, nth);
        if (tid == 0) {
            atomicAdd(&stats[0], red[0]);
            atomicAdd(&stats[1], red[1]);
            if (blockIdx.x == 0) atomicAdd(&stats[2], (double)T * (double)n);
        }
    }
}

// (sum, sum of squares, count) of adv in float64 -- PPO.py:115's mean / std ingredients.
// 16-byte loads over the 16-byte aligned body, scalar head / tail.
__global__ void __launch_bounds__(256) adv_stats_kernel(const float *__restrict__ adv, long long count, double *stats) {
    double s = 0.0, ss = 0.0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    long long head = ((16 - ((uintptr_t)adv & 15u)) & 15u) / 4;
    head = head < count ? head : count;
    const long long n4 = (count - head) / 4;
    const float4 *a4 = reinterpret_cast<const float4 *>(adv + head);
    for (long long i = tid; i < n4; i += nth) {
        const float4 x = __ldg(a4 + i);
        // per-thread partial sums in fp32 pairs would lose bits on 1e8 elements: accumulate in fp64
        s += (double)x.x + (double)x.y + (double)x.z + (double)x.w;
        ss += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
    }
    for (long long i = tid; i < head + (count - head - 4 * n4); i += nth) {
        const long long j = i < head ? i : head + 4 * n4 + (i - head);
        const double x = (double)adv[j];
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
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    long long head = ((16 - ((uintptr_t)adv & 15u)) & 15u) / 4;
    head = head < count ? head : count;
    const long long n4 = (count - head) / 4;
    float4 *a4 = reinterpret_cast<float4 *>(adv + head);
    for (long long i = tid; i < n4; i += nth) {
        float4 x = a4[i];
        x.x = (x.x - m) * inv; x.y = (x.y - m) * inv; x.z = (x.z - m) * inv; x.w = (x.w - m) * inv;
        a4[i] = x;
    }
    for (long long i = tid; i < head + (count - head - 4 * n4); i += nth) {
        const long long j = i < head ? i : head + 4 * n4 + (i - head);
        adv[j] = (adv[j] - m) * inv;
    }
}

}  // namespace ta
