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
    for (int i = tid; i < 256; i += PR_THREADS) sw1[i] = a.w1[i];
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
            const int ny0 = (Y & 1) ? 2 : 4, nx0 = (X & 1) ? 2 : 4;   // taps ky < ny0 fall on row y0, the rest on y1
#pragma unroll 4
            for (int c = 0; c < 16; c++) {
                float acc = 0.0f;
#pragma unroll
                for (int ky = 0; ky < 4; ky++) {
#pragma unroll
                    for (int kx = 0; kx < 4; kx++) {
                        const float v = ky < ny0 ? (kx < nx0 ? v00 : v01) : (kx < nx0 ? v10 : v11);
                        acc = fmaf(sw1[c * 16 + ky * 4 + kx], v, acc);
                    }
                }
                a1[c * 1089 + p] = fmaxf(fmaf(acc, sst[c], sst[16 + c]), 0.0f);
            }
        }
        __syncthreads();
        // layer 2: Conv2d(16,16,5,4) -> 16 x 8 x 8.  Thread = (pixel, 2 output channels): 64 x 8 = 512 threads.
        {
            const int p = tid & 63, cg = tid >> 6, Y = p >> 3, X = p & 7;
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
    float *sw2 = sw1 + 4096;              // 6400
    float *sw3 = sw2 + 6400;              // 144
    float *sb = sw3 + 144;                // b1 (16) b2 (16)
    float *a1 = sb + 32;                  // 17424   16 x 33 x 33
    float *a0 = a1 + PR_A1;               // 1024    16 x 8 x 8
    float *sz = a0 + 1024;                // 1024    64 x 4 x 4
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += PR_THREADS) sw1[i] = a.w1[i];
    for (int i = tid; i < 6400; i += PR_THREADS) sw2[i] = a.w2[i];
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
        // Work item = (pixel, 4 output channels): 1089 x 4.
        for (int it = tid; it < 1089 * 4; it += PR_THREADS) {
            const int p = it % 1089, cg = it / 1089, y = p / 33, x = p - 33 * y;
            int iy[2], ky[2], ny = 0, jx[2], kx[2], nx = 0;
            if (y < 32) { iy[ny] = y >> 2; ky[ny] = y & 3; ny++; }
            if ((y & 3) == 0 && y > 0) { iy[ny] = (y >> 2) - 1; ky[ny] = 4; ny++; }
            if (x < 32) { jx[nx] = x >> 2; kx[nx] = x & 3; nx++; }
            if ((x & 3) == 0 && x > 0) { jx[nx] = (x >> 2) - 1; kx[nx] = 4; nx++; }
            float acc[4] = {sb[16 + 4 * cg], sb[16 + 4 * cg + 1], sb[16 + 4 * cg + 2], sb[16 + 4 * cg + 3]};
            for (int u = 0; u < ny; u++)
                for (int v2 = 0; v2 < nx; v2++) {
                    const float *ap = a0 + iy[u] * 8 + jx[v2];
                    const float *wp = sw2 + ((ky[u] * 5 + kx[v2]) * 16) * 16 + 4 * cg;
#pragma unroll 4
                    for (int ci = 0; ci < 16; ci++) {
                        const float v = ap[ci * 64];
                        const float4 w = *reinterpret_cast<const float4 *>(wp + ci * 16);
                        acc[0] = fmaf(w.x, v, acc[0]); acc[1] = fmaf(w.y, v, acc[1]);
                        acc[2] = fmaf(w.z, v, acc[2]); acc[3] = fmaf(w.w, v, acc[3]);
                    }
                }
#pragma unroll
            for (int c = 0; c < 4; c++) a1[(4 * cg + c) * 1089 + p] = fmaxf(acc[c], 0.0f);
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
constexpr int PR_DEC_SMEM = (4096 + 6400 + 144 + 32 + PR_A1 + 1024 + 1024) * 4;

}  // namespace ta
