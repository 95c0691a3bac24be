// ta_pred.cuh -- the frozen frame predictor's convolution stacks (BASELINE configs[4]) as two fused inference kernels.
//
//   pred_encoder_kernel   Net_Encoder  soa/agent/net/all_net.py:7-51   UpsamplingNearest2d(4) -> Conv2d(1,16,4,2)+BN+ReLU
//                                      -> Conv2d(16,16,5,4)+BN+ReLU -> Conv2d(16,64,2,2)+BN+ReLU       17x17 -> 64x4x4
//   pred_decoder_kernel   Net_Decoder  soa/agent/net/all_net.py:100-137 ConvTranspose2d(64,16,2,2)+ReLU -> ConvTranspose2d(16,16,5,4)
//                                      +ReLU -> ConvTranspose2d(16,1,4,2) -> AvgPool2d(4,4)             64x4x4 -> 17x17
//
// PPO_Predictor.pred_states (soa/agent/PPO_Predictor.py:70-83) runs both in eval mode on 4 frames per env inside every
// select_action and over the whole buffer in update; only the POOLED decoder output is ever used.  Through cuDNN the two
// stacks cost 19 + 49 ms per 8192-env step (1- and 16-channel convolutions on 68x68 / 33x33 maps: legacy kernels, NCHW <->
// NHWC conversions, a strided data-gradient kernel for each transposed convolution) against 4 ms for the 3-layer LSTM
// between them.  Here one CTA carries an image through a whole stack with every activation in shared memory (the
// largest, 16 x 33 x 33, is 70 KB in fp32): HBM sees 289 bytes / 1.2 KB in and 2 KB / 1.2 KB out per image.
//   * eval-mode BatchNorm is an affine map per channel: the host folds it (and the convolution bias) into scale / shift;
//   * the upsampling is an index (u[y][x] = in[y / 4][x / 4]), never a tensor;
//   * the last transposed convolution followed by the 4x4 average pool is ONE linear map: a 3x3, stride-2, padding-1
//     convolution of the 33x33 map whose taps are sums of the 4x4 kernel's (host: predictor.fold_decoder_tail) -- the
//     68x68 image is never formed.
// fp32 arithmetic throughout (the networks are tiny: 2.3 MFLOP per image), weights read through shared memory.
#pragma once
#include <cuda_bf16.h>

#include "ta_aux.cuh"

namespace ta {

constexpr int PR_THREADS = 512;
constexpr int PR_A1 = 16 * 33 * 33;   // the 16 x 33 x 33 activation both stacks pass through
constexpr int PR_TAP = 16 * 16 + 8;   // floats per tap of the decoder's [tap][ci][co] weights in shared memory

struct PredEncArgs {
    const float *w1, *s1, *t1;   // [16][4][4]; scale / shift [16]   (y = relu(conv * s + t), t includes the bias)
    const float *w2, *s2, *t2;   // [25][16 ci][16 co] (tap-major, output channel fastest); [16]
    const float *w3, *s3, *t3;   // [4][16 ci][64 co]; [64]
};

// x: uint8 matrix codes (CODES) or float32 LUT values, [M][289]; z: bf16 [M][1024] = (channel, y, x) of the 64 x 4 x 4 map
template <bool CODES>
__global__ void __launch_bounds__(PR_THREADS) pred_encoder_kernel(const void *__restrict__ x_in, PredEncArgs a, __nv_bfloat16 *__restrict__ z,
                                                                  long long M) {
    extern __shared__ __align__(16) float pe_smem[];
    float *sw2 = pe_smem;                 // 6400
    float *sw3 = sw2 + 6400;              // 4096
    float *sw1 = sw3 + 4096;              // 256
    float *sst = sw1 + 256;               // s1 t1 s2 t2 (64) + s3 t3 (128) = 192
    float *a1 = sst + 192;                // 17424
    float *a2 = a1 + PR_A1;               // 1024
    float *sin = a2 + 1024;               // 292
    const int tid = threadIdx.x;
    for (int i = tid; i < 6400; i += PR_THREADS) sw2[i] = a.w2[i];
    for (int i = tid; i < 4096; i += PR_THREADS) sw3[i] = a.w3[i];
    // layer 1 folded: an output pixel of parity (py, px) sees a 2 x 2 patch of the 17 x 17 frame; the 16 taps fall on patch row r
    // through ky in {0..3} / {} (even Y) or {0,1} / {2,3} (odd Y), the same for columns: sw1[phase][c][2r + s] = the taps' sum
    for (int i = tid; i < 256; i += PR_THREADS) {
        const int ph = i >> 6, c = (i >> 2) & 15, r = (i >> 1) & 1, q = i & 1, py = ph >> 1, px = ph & 1;
        const int ky0 = py ? 2 * r : (r ? 4 : 0), ky1 = py ? 2 * r + 2 : (r ? 4 : 4), kx0 = px ? 2 * q : (q ? 4 : 0), kx1 = px ? 2 * q + 2 : 4;
        float sum = 0.0f;
        for (int ky = ky0; ky < ky1; ky++)
            for (int kx = kx0; kx < kx1; kx++) sum += a.w1[c * 16 + ky * 4 + kx];
        sw1[i] = sum;
    }
    if (tid < 16) { sst[tid] = a.s1[tid]; sst[16 + tid] = a.t1[tid]; sst[32 + tid] = a.s2[tid]; sst[48 + tid] = a.t2[tid]; }
    if (tid < 64) { sst[64 + tid] = a.s3[tid]; sst[128 + tid] = a.t3[tid]; }
    for (long long m = blockIdx.x; m < M; m += gridDim.x) {
        __syncthreads();   // (weights staged; the previous image's a2 / sin are free)
        for (int i = tid; i < NCELL; i += PR_THREADS) {
            if (CODES) sin[i] = matrix_value(reinterpret_cast<const uint8_t *>(x_in)[m * NCELL + i]);
            else sin[i] = reinterpret_cast<const float *>(x_in)[m * NCELL + i];
        }
        __syncthreads();
        // layer 1: Conv2d(1,16,4,2) of the 4x upsampled 68x68 image -> 16 x 33 x 33.  Output row Y reads upsampled rows
        // 2Y..2Y+3 = input rows (2Y + ky) >> 2: one or two of them, so the 16 taps collapse to a 2 x 2 patch.
        for (int p = tid; p < 33 * 33; p += PR_THREADS) {
            const int Y = p / 33, X = p - 33 * Y;
            const int y0 = (2 * Y) >> 2, y1 = (2 * Y + 3) >> 2, x0 = (2 * X) >> 2, x1 = (2 * X + 3) >> 2;
            const float v00 = sin[y0 * GS + x0], v01 = sin[y0 * GS + x1], v10 = sin[y1 * GS + x0], v11 = sin[y1 * GS + x1];
            const float4 *wp = reinterpret_cast<const float4 *>(sw1 + (((Y & 1) << 1) | (X & 1)) * 64);
#pragma unroll 4
            for (int c = 0; c < 16; c++) {
                const float4 w = wp[c];
                const float acc = fmaf(w.x, v00, fmaf(w.y, v01, fmaf(w.z, v10, w.w * v11)));
                a1[c * 1089 + p] = fmaxf(fmaf(acc, sst[c], sst[16 + c]), 0.0f);
            }
        }
        __syncthreads();
        // layer 2: Conv2d(16,16,5,4) -> 16 x 8 x 8.  Thread = (pixel, 2 output channels): 64 x 8 = 512 threads.
        {
            // a warp = 8 pixels of one output row x 4 channel pairs: its activation loads touch 8 banks (4-lane broadcasts),
            // its weight loads 8 consecutive floats (with (pixel, channel pair) = (tid & 63, tid >> 6) the 4 rows of a warp fell
            // on the same 8 banks: four-way conflicts on every activation load)
            const int X = tid & 7, Y = (tid >> 5) & 7, cg = ((tid >> 8) << 2) | ((tid >> 3) & 3), p = Y * 8 + X;
            float acc0 = 0.0f, acc1 = 0.0f;
            for (int ci = 0; ci < 16; ci++) {
                const float *ap = a1 + ci * 1089 + (4 * Y) * 33 + 4 * X;
#pragma unroll
                for (int ky = 0; ky < 5; ky++) {
#pragma unroll
                    for (int kx = 0; kx < 5; kx++) {
                        const float v = ap[ky * 33 + kx];
                        const float2 w = *reinterpret_cast<const float2 *>(sw2 + ((ky * 5 + kx) * 16 + ci) * 16 + 2 * cg);
                        acc0 = fmaf(w.x, v, acc0);
                        acc1 = fmaf(w.y, v, acc1);
                    }
                }
            }
            a2[(2 * cg) * 64 + p] = fmaxf(fmaf(acc0, sst[32 + 2 * cg], sst[48 + 2 * cg]), 0.0f);
            a2[(2 * cg + 1) * 64 + p] = fmaxf(fmaf(acc1, sst[32 + 2 * cg + 1], sst[48 + 2 * cg + 1]), 0.0f);
        }
        __syncthreads();
        // layer 3: Conv2d(16,64,2,2) -> 64 x 4 x 4.  Thread = (pixel, 2 output channels): 16 x 32 = 512 threads.
        {
            const int p = tid & 15, cg = tid >> 4, Y = p >> 2, X = p & 3;
            float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 4
            for (int ci = 0; ci < 16; ci++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float v = a2[ci * 64 + (2 * Y + (k >> 1)) * 8 + 2 * X + (k & 1)];
                    const float2 w = *reinterpret_cast<const float2 *>(sw3 + (k * 16 + ci) * 64 + 2 * cg);
                    acc0 = fmaf(w.x, v, acc0);
                    acc1 = fmaf(w.y, v, acc1);
                }
            }
            const int c0 = 2 * cg;
            z[m * 1024 + c0 * 16 + p] = __float2bfloat16(fmaxf(fmaf(acc0, sst[64 + c0], sst[128 + c0]), 0.0f));
            z[m * 1024 + (c0 + 1) * 16 + p] = __float2bfloat16(fmaxf(fmaf(acc1, sst[64 + c0 + 1], sst[128 + c0 + 1]), 0.0f));
        }
    }
}

struct PredDecArgs {
    const float *w1, *b1;   // [4 taps][64 ci][16 co], [16]      ConvTranspose2d(64,16,2,2)
    const float *w2, *b2;   // [25 taps][16 ci][16 co], [16]     ConvTranspose2d(16,16,5,4)
    const float *w3;        // [9 taps][16 ci]                   the folded ConvTranspose2d(16,1,4,2) + AvgPool2d(4): 3x3, stride 2, padding 1
    float b3;
};

// z: bf16 [M][1024] (64 x 4 x 4); out: float32 [M][289] (the pooled 17 x 17 prediction)
__global__ void __launch_bounds__(PR_THREADS) pred_decoder_kernel(const __nv_bfloat16 *__restrict__ z, PredDecArgs a, float *__restrict__ out,
                                                                  long long M) {
    extern __shared__ __align__(16) float pd_smem[];
    float *sw1 = pd_smem;                 // 4096
    float *sw2 = sw1 + 4096;              // 25 taps x PR_TAP (256 weights + 8 floats of padding: lanes on different taps, different banks)
    float *sw3 = sw2 + 25 * PR_TAP;       // 144
    float *sb = sw3 + 144;                // b1 (16) b2 (16)
    float *a1 = sb + 32;                  // 17424   16 x 33 x 33
    float *a0 = a1 + PR_A1;               // 1024    16 x 8 x 8
    float *sz = a0 + 1024;                // 1024    64 x 4 x 4
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += PR_THREADS) sw1[i] = a.w1[i];
    for (int i = tid; i < 6400; i += PR_THREADS) sw2[(i >> 8) * PR_TAP + (i & 255)] = a.w2[i];
    for (int i = tid; i < 144; i += PR_THREADS) sw3[i] = a.w3[i];
    if (tid < 16) { sb[tid] = a.b1[tid]; sb[16 + tid] = a.b2[tid]; }
    for (long long m = blockIdx.x; m < M; m += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < 1024; i += PR_THREADS) sz[i] = __bfloat162float(z[m * 1024 + i]);
        __syncthreads();
        // ConvTranspose2d(64,16,2,2): out[c][2i+ky][2j+kx] = b + sum_ci in[ci][i][j] w[ci][c][ky][kx].  Thread = (pixel of the
        // 8 x 8 map, 2 output channels): 64 x 8 = 512 threads.
        {
            const int p = tid & 63, cg = tid >> 6, y = p >> 3, x = p & 7;
            const int k = (y & 1) * 2 + (x & 1), ip = (y >> 1) * 4 + (x >> 1);
            float acc0 = sb[2 * cg], acc1 = sb[2 * cg + 1];
#pragma unroll 8
            for (int ci = 0; ci < 64; ci++) {
                const float v = sz[ci * 16 + ip];
                const float2 w = *reinterpret_cast<const float2 *>(sw1 + (k * 64 + ci) * 16 + 2 * cg);
                acc0 = fmaf(w.x, v, acc0);
                acc1 = fmaf(w.y, v, acc1);
            }
            a0[(2 * cg) * 64 + p] = fmaxf(acc0, 0.0f);
            a0[(2 * cg + 1) * 64 + p] = fmaxf(acc1, 0.0f);
        }
        __syncthreads();
        // ConvTranspose2d(16,16,5,4): out[c][y][x] = b + sum over (i, ky) with 4i + ky = y, (j, kx) with 4j + kx = x.  A pixel has
        // one such pair per axis, or two where y % 4 == 0 inside the map (ky = 0 of row y/4 and ky = 4 of row y/4 - 1).
        // Work item = (pixel, 8 output channels): 1089 x 2; one activation load and two 16-byte weight loads per 8 FMAs.
        for (int it = tid; it < 1089 * 2; it += PR_THREADS) {
            const int cg = it >= 1089, p = it - 1089 * cg, y = p / 33, x = p - 33 * y;
            // first (always, except on the last row / column) the tap of row y >> 2, then the overlap tap of the row before
            const int ny = ((y < 32) ? 1 : 0) + (((y & 3) == 0 && y > 0) ? 1 : 0), nx = ((x < 32) ? 1 : 0) + (((x & 3) == 0 && x > 0) ? 1 : 0);
            float acc[8];
#pragma unroll
            for (int c = 0; c < 8; c++) acc[c] = sb[16 + 8 * cg + c];
            for (int u = 0; u < ny; u++) {
                const bool firsty = (u == 0 && y < 32);
                const int iy = firsty ? (y >> 2) : (y >> 2) - 1, ky = firsty ? (y & 3) : 4;
                for (int v2 = 0; v2 < nx; v2++) {
                    const bool firstx = (v2 == 0 && x < 32);
                    const int jx = firstx ? (x >> 2) : (x >> 2) - 1, kx = firstx ? (x & 3) : 4;
                    const float *ap = a0 + iy * 8 + jx;
                    const float *wp = sw2 + (ky * 5 + kx) * PR_TAP + 8 * cg;
#pragma unroll 4
                    for (int ci = 0; ci < 16; ci++) {
                        const float v = ap[ci * 64];
                        const float4 w0 = *reinterpret_cast<const float4 *>(wp + ci * 16), w1 = *reinterpret_cast<const float4 *>(wp + ci * 16 + 4);
                        acc[0] = fmaf(w0.x, v, acc[0]); acc[1] = fmaf(w0.y, v, acc[1]); acc[2] = fmaf(w0.z, v, acc[2]); acc[3] = fmaf(w0.w, v, acc[3]);
                        acc[4] = fmaf(w1.x, v, acc[4]); acc[5] = fmaf(w1.y, v, acc[5]); acc[6] = fmaf(w1.z, v, acc[6]); acc[7] = fmaf(w1.w, v, acc[7]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 8; c++) a1[(8 * cg + c) * 1089 + p] = fmaxf(acc[c], 0.0f);
        }
        __syncthreads();
        // the folded tail: 3x3, stride 2, padding 1 over the 16 x 33 x 33 map -> 17 x 17
        for (int p = tid; p < NCELL; p += PR_THREADS) {
            const int Y = p / GS, X = p - GS * Y;
            float acc = a.b3;
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
                const int y = 2 * Y - 1 + dy;
                if (y < 0 || y > 32) continue;
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const int x = 2 * X - 1 + dx;
                    if (x < 0 || x > 32) continue;
#pragma unroll 4
                    for (int ci = 0; ci < 16; ci++) acc = fmaf(sw3[(dy * 3 + dx) * 16 + ci], a1[ci * 1089 + y * 33 + x], acc);
                }
            }
            out[m * NCELL + p] = acc;
        }
    }
}

constexpr int PR_ENC_SMEM = (6400 + 4096 + 256 + 192 + PR_A1 + 1024 + 292) * 4;
constexpr int PR_DEC_SMEM = (4096 + 25 * PR_TAP + 144 + 32 + PR_A1 + 1024 + 1024) * 4;

}  // namespace ta
