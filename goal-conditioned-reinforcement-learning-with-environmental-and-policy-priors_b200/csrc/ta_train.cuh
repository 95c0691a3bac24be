// ta_train.cuh -- the small kernels of the hand-scheduled PPO optimiser step (fused_step.py): everything around the
// convolutions / GEMMs that PyTorch's autograd + optimizer would otherwise launch as ~250 tiny library kernels per step
// (weight casts, gradient accumulation, ReLU masks, bias sums, softmax / Categorical / clamp / min / mean, Adam).
//
// Reference arithmetic restated here (paths relative to the reference root):
//   ppo_actor_loss_kernel   soa/agent/PPO.py:124-132   Categorical(probs).entropy / log_prob, ratio, clipped surrogate
//   ppo_critic_loss_kernel  soa/agent/PPO.py:133        F.smooth_l1_loss(V(s), target_v)
//   adam_shadow_kernel      soa/agent/PPO.py:57-58,136-144   torch.optim.Adam(lr, eps=1e-5) step (no weight decay, no amsgrad)
//   tinet_prep_kernel       soa/agent/net/all_net.py:142-143,157   Upsample(4) + Conv2d(4,64,4,2) folded to the 2x2-patch form
// All reductions are deterministic: fixed-order trees inside a CTA, per-CTA partials summed in index order by the last
// CTA to arrive (the order does not depend on which CTA that is).
#pragma once
#include <cuda_bf16.h>

#include "ta_common.cuh"

namespace ta {

// ---- dz = dy * [y > 0] (bf16) and its column sums (the bias gradient), one pass --------------------------------------
// dy [rows][C] with row stride ld_dy elements, y / dz [rows][C] dense; C a multiple of 8, C / 8 <= 256 and a divisor of 256.
// y == nullptr: no mask and no dz (plain column sum of dy).  scratch: float [grid][C] followed by one uint32 counter (zero
// before the first launch; the kernel leaves it zero).
constexpr int RB_THREADS = 256;
constexpr int RB_UNROLL = 4;   // rows in flight per thread
// PH > 0: dy is not a dense matrix but the merged parity planes of a stride-2 convolution's data gradient
// (bf16 [B][(PH+1)/2][(PW+1)/2][4][C], ta_conv1.cuh): row r = pixel (b, h, w) of the [B][PH][PW][C] map reads plane block
// (h & 1) * 2 + (w & 1) of position (h >> 1, w >> 1) -- planes_to_dense_relu_kernel's interleave, fused with the bias sum.
__global__ void __launch_bounds__(RB_THREADS) relu_bwd_bias_kernel(const __nv_bfloat16 *__restrict__ dy, long long ld_dy,
                                                                   const __nv_bfloat16 *__restrict__ y, __nv_bfloat16 *__restrict__ dz,
                                                                   long long rows, int C, float *__restrict__ db, float *scratch,
                                                                   int PH, int PW) {
    __shared__ float red[RB_THREADS][8];
    __shared__ bool is_last;
    const int cg = C >> 3, tid = threadIdx.x;
    const int col = tid % cg, rsub = tid / cg, rpp = RB_THREADS / cg;  // rows per pass
    const long long per_cta = (rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * per_cta, r1 = r0 + per_cta < rows ? r0 + per_cta : rows;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto mk = [](uint32_t gw, uint32_t yw) {  // y is a ReLU output (>= +0): non-zero bits <=> y > 0
        return gw & (((yw & 0x7FFFu) ? 0xFFFFu : 0u) | ((yw & 0x7FFF0000u) ? 0xFFFF0000u : 0u));
    };
    for (long long rb = r0 + rsub; rb < r1; rb += (long long)rpp * RB_UNROLL) {
        uint4 g[RB_UNROLL], yv[RB_UNROLL];
#pragma unroll
        for (int u = 0; u < RB_UNROLL; u++) {  // all loads of the batch first: RB_UNROLL (x 2) 16-byte loads in flight per thread
            const long long r = rb + (long long)u * rpp;
            g[u] = make_uint4(0u, 0u, 0u, 0u);
            yv[u] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            if (r < r1) {
                long long src = r * ld_dy;
                if (PH > 0) {   // (rows < 2^32, checked by the host: 32-bit divisions -- the 64-bit ones cost more than the row's loads)
                    const unsigned bh = (unsigned)r / (unsigned)PW, b = bh / (unsigned)PH;
                    const unsigned w = (unsigned)r - bh * (unsigned)PW, h = bh - b * (unsigned)PH;
                    src = (long long)(((b * (unsigned)((PH + 1) / 2) + (h >> 1)) * (unsigned)((PW + 1) / 2) + (w >> 1)) * 4u + ((h & 1u) * 2u + (w & 1u))) *
                          (long long)C;
                }
                g[u] = __ldg(reinterpret_cast<const uint4 *>(dy + src + col * 8));
                if (y) yv[u] = __ldg(reinterpret_cast<const uint4 *>(y + r * (long long)C + col * 8));
            }
        }
#pragma unroll
        for (int u = 0; u < RB_UNROLL; u++) {
            const long long r = rb + (long long)u * rpp;
            if (y) {
                g[u] = make_uint4(mk(g[u].x, yv[u].x), mk(g[u].y, yv[u].y), mk(g[u].z, yv[u].z), mk(g[u].w, yv[u].w));
                if (r < r1) *reinterpret_cast<uint4 *>(dz + r * (long long)C + col * 8) = g[u];
            }
            const uint32_t w[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
            for (int q = 0; q < 4; q++) {  // fixed order: rows ascending within a thread
                acc[2 * q] += __uint_as_float(w[q] << 16);
                acc[2 * q + 1] += __uint_as_float(w[q] & 0xFFFF0000u);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) red[tid][k] = acc[k];
    __syncthreads();
    float *part = scratch + (long long)blockIdx.x * C;
    if (tid < cg) {  // fixed order over the rpp row groups
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float s = 0.f;
            for (int j = 0; j < rpp; j++) s += red[j * cg + tid][k];
            part[tid * 8 + k] = s;
        }
    }
    __threadfence();
    __syncthreads();
    unsigned int *counter = reinterpret_cast<unsigned int *>(scratch + (long long)gridDim.x * C);
    if (tid == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // the last CTA to arrive sums the per-CTA partials in CTA index order (the result does not depend on which CTA that
    // is): C / 4 threads cover a partial row as float4, the remaining thread dimension splits the CTA range into
    // contiguous slices whose sums are then added slice 0 first
    const int c4n = C >> 2, nsl = RB_THREADS / c4n > 0 ? RB_THREADS / c4n : 1;
    float4 *red4 = reinterpret_cast<float4 *>(&red[0][0]);   // RB_THREADS float4 fit in red
    for (int cbase = 0; cbase < c4n; cbase += RB_THREADS) {   // (C / 4 > 256 never happens for the supported C <= 1024; kept general)
        const int c4 = cbase + (tid % (c4n < RB_THREADS ? c4n : RB_THREADS)), sl = c4n < RB_THREADS ? tid / c4n : 0;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < c4n && sl < nsl) {
            const unsigned int per = (gridDim.x + nsl - 1) / nsl, b0 = sl * per, b1 = b0 + per < gridDim.x ? b0 + per : gridDim.x;
#pragma unroll 8
            for (unsigned int b = b0; b < b1; b++) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(scratch + (long long)b * C) + c4);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        red4[tid] = s;
        __syncthreads();
        if (c4 < c4n && sl == 0) {
            float4 t = red4[tid];
            for (int j = 1; j < nsl; j++) {
                const float4 v = red4[j * c4n + (tid % c4n)];
                t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
            }
            reinterpret_cast<float4 *>(db)[c4] = t;
        }
        __syncthreads();
    }
    if (tid == 0) *counter = 0u;
}

// ---- PPO losses: one CTA, fixed-order reductions ----------------------------------------------------------------------
constexpr int LOSS_THREADS = 1024;
__device__ __forceinline__ float block_sum_ordered(float v, float *buf) {  // every thread returns the same sum
    const int tid = threadIdx.x;
    buf[tid] = v;
    __syncthreads();
    for (int s = LOSS_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) buf[tid] += buf[tid + s];
        __syncthreads();
    }
    const float r = buf[0];
    __syncthreads();
    return r;
}

// logits bf16 [B][8] (columns 0..4 used), act int32 [B], old_logp / adv float [B] ->
//   dlogits bf16 [B][8] = d(mean loss)/d logits (columns 5..7 zero), out[0] = mean loss, db_head float [5] = column sums
//   of dlogits in fp32, step_counter[0] += 1 (the optimiser step this backward belongs to; read by adam_shadow_kernel).
// loss_i = -min(ratio*A, clamp(ratio, 1-clip, 1+clip)*A) - ent_coef*H,  ratio = exp(log p(a) - old_logp)   (PPO.py:124-132)
__global__ void __launch_bounds__(LOSS_THREADS) ppo_actor_loss_kernel(const __nv_bfloat16 *__restrict__ logits, const int *__restrict__ act,
                                                                      const float *__restrict__ old_logp, const float *__restrict__ adv,
                                                                      int B, float clip, float ent_coef, __nv_bfloat16 *__restrict__ dlogits,
                                                                      float *__restrict__ out, float *__restrict__ db_head,
                                                                      float *step_counter) {
    __shared__ float buf[LOSS_THREADS];
    const float invB = 1.0f / (float)B;
    float loss = 0.f, dbs[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < B; i += LOSS_THREADS) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(logits + (long long)i * 8);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        float l[5];
#pragma unroll
        for (int j = 0; j < 5; j++) l[j] = __uint_as_float((j & 1) ? (w[j >> 1] & 0xFFFF0000u) : (w[j >> 1] << 16));
        float m = l[0];
#pragma unroll
        for (int j = 1; j < 5; j++) m = fmaxf(m, l[j]);
        float e[5], Z = 0.f;
#pragma unroll
        for (int j = 0; j < 5; j++) { e[j] = expf(l[j] - m); Z += e[j]; }
        const float logZ = logf(Z), invZ = 1.0f / Z;
        float p[5], lp[5], H = 0.f;
#pragma unroll
        for (int j = 0; j < 5; j++) { p[j] = e[j] * invZ; lp[j] = l[j] - m - logZ; H -= p[j] * lp[j]; }
        const int a = act[i];
        float lpa = lp[0];
#pragma unroll
        for (int j = 1; j < 5; j++) lpa = (a == j) ? lp[j] : lpa;
        const float A = adv[i], ratio = expf(lpa - old_logp[i]);
        const float s1 = ratio * A, s2 = fminf(fmaxf(ratio, 1.0f - clip), 1.0f + clip) * A;
        loss += -fminf(s1, s2) - ent_coef * H;
        // d(-min)/d log p(a): through surr1 when it is the smaller (or equal: inside the clip range both branches carry it)
        const float g = (s1 <= s2) ? -s1 : 0.0f;
        uint32_t o[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const float d = (g * ((a == j ? 1.0f : 0.0f) - p[j]) + ent_coef * p[j] * (lp[j] + H)) * invB;
            const __nv_bfloat16 db16 = __float2bfloat16_rn(d);
            dbs[j] += __bfloat162float(db16);   // the bias gradient is the column sum of what the GEMMs consume
            o[j >> 1] |= (uint32_t)__bfloat16_as_ushort(db16) << ((j & 1) * 16);
        }
        *reinterpret_cast<uint4 *>(dlogits + (long long)i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    const float total = block_sum_ordered(loss, buf);
    float bsum[5];
#pragma unroll
    for (int j = 0; j < 5; j++) bsum[j] = block_sum_ordered(dbs[j], buf);
    if (threadIdx.x == 0) {
        out[0] = total * invB;
#pragma unroll
        for (int j = 0; j < 5; j++) db_head[j] = bsum[j];
        if (step_counter) step_counter[0] += 1.0f;
    }
}

// v bf16 [B][8] (column 0 used), target float [B]: smooth_l1 (beta = 1), mean reduction (PPO.py:133)
__global__ void __launch_bounds__(LOSS_THREADS) ppo_critic_loss_kernel(const __nv_bfloat16 *__restrict__ v, const float *__restrict__ target, int B,
                                                                       __nv_bfloat16 *__restrict__ dv, float *__restrict__ out,
                                                                       float *__restrict__ db_head, float *step_counter) {
    __shared__ float buf[LOSS_THREADS];
    const float invB = 1.0f / (float)B;
    float loss = 0.f, dbs = 0.f;
    for (int i = threadIdx.x; i < B; i += LOSS_THREADS) {
        const float d = __bfloat162float(v[(long long)i * 8]) - target[i];
        const float ad = fabsf(d);
        loss += ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
        const float gd = (ad < 1.0f ? d : (d > 0.f ? 1.0f : -1.0f)) * invB;
        const __nv_bfloat16 g16 = __float2bfloat16_rn(gd);
        dbs += __bfloat162float(g16);
        *reinterpret_cast<uint4 *>(dv + (long long)i * 8) = make_uint4((uint32_t)__bfloat16_as_ushort(g16), 0u, 0u, 0u);
    }
    const float total = block_sum_ordered(loss, buf);
    const float bsum = block_sum_ordered(dbs, buf);
    if (threadIdx.x == 0) {
        out[0] = total * invB;
        db_head[0] = bsum;
        if (step_counter) step_counter[0] += 1.0f;
    }
}

// ---- Adam on the flat parameter buffer + the bf16 shadow the kernels read ----------------------------------------------
// torch.optim.Adam's update (weight_decay 0, amsgrad off):  m <- m + (g - m)(1 - b1);  v <- b2 v + (1 - b2) g^2;
//   p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),   t = step_counter[0] (already incremented)
// g is first scaled by grad_scale (1 / world size after the all-reduce SUM).
__global__ void __launch_bounds__(256) adam_shadow_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                          float *__restrict__ v, __nv_bfloat16 *__restrict__ p16, long long n,
                                                          const float *__restrict__ step_counter, float lr, double beta1, double beta2, float eps,
                                                          float grad_scale) {
    // the scalar factors in double, as torch.optim.Adam forms them on the host (1 - beta2 and 1 - beta2^t cancel badly in fp32)
    const double t = (double)step_counter[0];
    const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
    const float step_size = (float)((double)lr / bc1), bc2s = (float)sqrt(bc2);
    const float b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = m[i] + (gi - m[i]) * omb1;
        const float vi = b2 * v[i] + omb2 * gi * gi;
        const float pi = p[i] - step_size * (mi / (sqrtf(vi) / bc2s + eps));
        m[i] = mi;
        v[i] = vi;
        p[i] = pi;
        p16[i] = __float2bfloat16_rn(pi);
    }
}

// ---- per-step weight forms derived from the master copy -----------------------------------------------------------------
// Sel[p][d][k]: kernel tap k of one axis lands on input offset d for output phase p (even outputs see one input pixel
// through all 4 taps, odd outputs two pixels through 2 taps each)
__device__ __forceinline__ float fold_sel(int p, int d, int k) { return p == 0 ? (d == 0 ? 1.f : 0.f) : ((k >> 1) == d ? 1.f : 0.f); }

struct PrepArgs {
    const float *w1;            // conv1 weight, element (o, c, ky, kx) at o*s_o + c*s_c + ky*s_y + kx*s_x
    long long s_o, s_c, s_y, s_x;
    const float *b1;            // [64]
    float *w4, *b4;             // folded [256][16], [256]
    const __nv_bfloat16 *fc0;   // shadow [256][2304], feature f = c*9 + o (Flatten of [256,3,3])
    __nv_bfloat16 *fc0p;        // [256][9][256]: feature o*256 + c (the order conv4's GEMM produces)
    const __nv_bfloat16 *pos;   // shadow [128][10]
    __nv_bfloat16 *pos16;       // [128][16], columns 10..15 zero
    const __nv_bfloat16 *head, *head_b;  // shadow [nh][512], [nh]
    __nv_bfloat16 *head8, *head_b8;      // [8][512], [8], rows nh..7 zero
    int nh;
};
__global__ void __launch_bounds__(256) tinet_prep_kernel(const PrepArgs a) {
    const long long n_fold = 256 * 16, n_b4 = 256, n_fc0 = 256ll * 2304, n_pos = 128 * 16, n_head = 8 * 512, n_hb = 8;
    const long long total = n_fold + n_b4 + n_fc0 + n_pos + n_head + n_hb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long j = i;
        if (j < n_fold) {  // w4[(py*2+px)*64 + o][(dy*2+dx)*4 + c]
            const int row = (int)(j >> 4), colk = (int)(j & 15);
            const int ph = row >> 6, o = row & 63, py = ph >> 1, px = ph & 1, dd = colk >> 2, c = colk & 3, dy = dd >> 1, dx = dd & 1;
            float s = 0.f;
#pragma unroll
            for (int ky = 0; ky < 4; ky++)
#pragma unroll
                for (int kx = 0; kx < 4; kx++)
                    s += a.w1[o * a.s_o + c * a.s_c + ky * a.s_y + kx * a.s_x] * fold_sel(py, dy, ky) * fold_sel(px, dx, kx);
            a.w4[j] = s;
            continue;
        }
        j -= n_fold;
        if (j < n_b4) { a.b4[j] = a.b1[j & 63]; continue; }
        j -= n_b4;
        if (j < n_fc0) {  // fc0p[n][o][c] = fc0[n][c*9 + o]
            const int n = (int)(j / 2304), r = (int)(j - (long long)n * 2304), o = r >> 8, c = r & 255;
            a.fc0p[j] = a.fc0[(long long)n * 2304 + c * 9 + o];
            continue;
        }
        j -= n_fc0;
        if (j < n_pos) { const int r = (int)(j >> 4), c = (int)(j & 15); a.pos16[j] = c < 10 ? a.pos[r * 10 + c] : __float2bfloat16_rn(0.f); continue; }
        j -= n_pos;
        if (j < n_head) { const int r = (int)(j >> 9); a.head8[j] = r < a.nh ? a.head[j] : __float2bfloat16_rn(0.f); continue; }
        j -= n_head;
        a.head_b8[j] = j < a.nh ? a.head_b[j] : __float2bfloat16_rn(0.f);
    }
}

// ---- gradients back into the flat fp32 buffer (the layout of the master parameters) ------------------------------------
struct GradArgs {
    // conv1: folded gradients -> the conv's own weight / bias gradient (transpose of the fold)
    const float *dw4, *db4;     // [256][16], [256]
    float *g_w1;                // same element strides as PrepArgs::w1
    long long s_o, s_c, s_y, s_x;
    float *g_b1;
    // bf16 weight gradients as the GEMMs / cuDNN leave them -> fp32, dense copies (n elements each)
    const __nv_bfloat16 *src[4];
    float *dst[4];
    long long n[4];
    const __nv_bfloat16 *fc0p;  // [256][9][256] -> g_fc0 [256][2304] (f = c*9 + o)
    float *g_fc0;
    const __nv_bfloat16 *pos16; // [128][16] -> g_pos [128][10]
    float *g_pos;
    const __nv_bfloat16 *head8; // [8][512] -> g_head [nh][512]
    float *g_head;
    int nh;
};
// (every group is optional -- dw4 / fc0p / pos16 / head8 == nullptr, n[k] == 0 -- so that the late layers' gradients can be
// finalised, and their all-reduce started, before the convolution stem's backward has run)
__global__ void __launch_bounds__(256) tinet_grad_kernel(const GradArgs a) {
    const long long n_w1 = a.dw4 ? 64 * 4 * 4 * 4 : 0, n_b1 = a.dw4 ? 64 : 0, n_fc0 = a.fc0p ? 256ll * 2304 : 0, n_pos = a.pos16 ? 128 * 10 : 0;
    const long long n_head = a.head8 ? (long long)a.nh * 512 : 0;
    const long long total = n_w1 + n_b1 + a.n[0] + a.n[1] + a.n[2] + a.n[3] + n_fc0 + n_pos + n_head;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long j = i;
        if (j < n_w1) {
            const int o = (int)(j >> 6), c = (int)((j >> 4) & 3), ky = (int)((j >> 2) & 3), kx = (int)(j & 3);
            float s = 0.f;
#pragma unroll
            for (int ph = 0; ph < 4; ph++)
#pragma unroll
                for (int dd = 0; dd < 4; dd++)
                    s += a.dw4[(ph * 64 + o) * 16 + dd * 4 + c] * fold_sel(ph >> 1, dd >> 1, ky) * fold_sel(ph & 1, dd & 1, kx);
            a.g_w1[o * a.s_o + c * a.s_c + ky * a.s_y + kx * a.s_x] = s;
            continue;
        }
        j -= n_w1;
        if (j < n_b1) { a.g_b1[j] = a.db4[j] + a.db4[64 + j] + a.db4[128 + j] + a.db4[192 + j]; continue; }
        j -= n_b1;
        bool done = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!done) {
                if (j < a.n[k]) { a.dst[k][j] = __bfloat162float(a.src[k][j]); done = true; }
                else j -= a.n[k];
            }
        }
        if (done) continue;
        if (j < n_fc0) {  // g_fc0[n][c*9 + o] = fc0p[n][o][c]
            const int n = (int)(j / 2304), f = (int)(j - (long long)n * 2304), c = f / 9, o = f - 9 * c;
            a.g_fc0[j] = __bfloat162float(a.fc0p[(long long)n * 2304 + o * 256 + c]);
            continue;
        }
        j -= n_fc0;
        if (j < n_pos) { const int r = (int)(j / 10), c = (int)(j - 10 * r); a.g_pos[j] = __bfloat162float(a.pos16[r * 16 + c]); continue; }
        j -= n_pos;
        a.g_head[j] = __bfloat162float(a.head8[j]);
    }
}

// ---- minibatch gather ----------------------------------------------------------------------------------------------------
// sample smp = idx[i] reads record rec = src ? src[smp] : smp (hindsight-relabelled samples copy another record's frames):
// frames 0..3 of s [.,5,289] -> sb [bs][4][289]; positions 0..3 of p [.,5,2] + goal g[smp] -> pg16 bf16 [bs][16] (columns
// 10..15 zero); a, old_logp, adv, target_v (all per sample) -> dense minibatch arrays.  One CTA per sample.
__global__ void __launch_bounds__(128) gather_minibatch_kernel(const uint8_t *__restrict__ s, const float *__restrict__ p, const float *__restrict__ g,
                                                              const long long *__restrict__ a, const float *__restrict__ old_logp,
                                                              const float *__restrict__ adv, const float *__restrict__ tv,
                                                              const long long *__restrict__ idx, const long long *__restrict__ src, int bs,
                                                              uint8_t *__restrict__ sb, __nv_bfloat16 *__restrict__ pg16, int *__restrict__ a_mb,
                                                              float *__restrict__ old_mb, float *__restrict__ adv_mb, float *__restrict__ tv_mb) {
    const int i = blockIdx.x;
    if (i >= bs) return;
    const long long smp = idx[i], rec = src ? src[smp] : smp;
    const uint8_t *sp = s + rec * (5ll * NCELL);
    uint8_t *dp = sb + (long long)i * (4 * NCELL);
    // 1156 bytes: the destination is 4-byte aligned, the source is not in general
    for (int w = threadIdx.x; w < NCELL; w += blockDim.x) {
        const uint8_t *q = sp + 4 * w;
        reinterpret_cast<uint32_t *>(dp)[w] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    }
    if (threadIdx.x < 16) {
        const int c = threadIdx.x;
        float val = 0.f;
        if (c < 8) val = p[rec * 10 + c];
        else if (c < 10) val = g[smp * 2 + (c - 8)];
        pg16[(long long)i * 16 + c] = __float2bfloat16_rn(val);
    }
    if (threadIdx.x == 32) {
        a_mb[i] = (int)a[smp];
        old_mb[i] = old_logp[smp];
        adv_mb[i] = adv[smp];
        tv_mb[i] = tv[smp];
    }
}

// ---- LSTM cell gates (the frozen frame predictor of BASELINE configs[4]: soa/agent/net/all_net.py:53-98, nn.LSTM(1024, 1024, 3)) ----
// torch.nn.LSTM's cell with the two gate GEMMs done by the caller:
//     (i, f, g, o) = gx + gh + bias         gate order of torch.nn.LSTM: input, forget, cell, output
//     c' = sigmoid(f) * c + sigmoid(i) * tanh(g);   h' = sigmoid(o) * tanh(c')
// gx, gh: float32 [B][4H] pre-activations (x W_ih^T and h W_hh^T; gh nullable when one GEMM over [x, h] produced both),
// bias float32 [4H] = b_ih + b_hh, c float32 [B][H] updated in place, h_out bf16 [B][ld_h] (the next GEMM's operand).
// One thread = 4 hidden units of one sample: four 16-byte gate loads (x 2), one 16-byte cell load / store, one 8-byte h store.
__global__ void __launch_bounds__(256) lstm_gates_kernel(const float *__restrict__ gx, const float *__restrict__ gh,
                                                         const float *__restrict__ bias, float *__restrict__ c,
                                                         __nv_bfloat16 *__restrict__ h_out, long long ld_h, long long B, int H) {
    const int hq = H >> 2;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= B * hq) return;
    const long long b = q / hq;
    const int j = (int)(q - b * hq) * 4;
    float4 g4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const long long off = b * 4ll * H + (long long)k * H + j;
        float4 a = __ldcs(reinterpret_cast<const float4 *>(gx + off));
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + k * H + j));
        a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
        if (gh) {
            const float4 hh = __ldcs(reinterpret_cast<const float4 *>(gh + off));
            a.x += hh.x; a.y += hh.y; a.z += hh.z; a.w += hh.w;
        }
        g4[k] = a;
    }
    float4 cv = *reinterpret_cast<const float4 *>(c + b * H + j);
    auto sig = [](float x) { return 1.0f / (1.0f + __expf(-x)); };
    auto cell = [&](float i, float f, float g, float o, float &cc) {
        cc = sig(f) * cc + sig(i) * tanhf(g);
        return sig(o) * tanhf(cc);
    };
    const float h0 = cell(g4[0].x, g4[1].x, g4[2].x, g4[3].x, cv.x), h1 = cell(g4[0].y, g4[1].y, g4[2].y, g4[3].y, cv.y);
    const float h2 = cell(g4[0].z, g4[1].z, g4[2].z, g4[3].z, cv.z), h3 = cell(g4[0].w, g4[1].w, g4[2].w, g4[3].w, cv.w);
    *reinterpret_cast<float4 *>(c + b * H + j) = cv;
    __nv_bfloat162 lo = __floats2bfloat162_rn(h0, h1), hi = __floats2bfloat162_rn(h2, h3);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t *>(&lo);
    pk.y = *reinterpret_cast<uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(h_out + b * ld_h + j) = pk;
}

}  // namespace ta
