// ta_conv1.cuh -- TINet's first layer, fused: LUT decode + UpsamplingNearest2d(4) + Conv2d(4, 64, k=4, s=2)
// + bias + ReLU (soa/agent/net/all_net.py:142-143, 157, 180-181; soa/env_buffer.py:300-318 for the LUT).
//
// The layer is folded exactly: on a x4 nearest-upsampled 17x17 frame a
// 4x4 stride-2 kernel at output (2m+py, 2n+px) only ever sees the four input pixels (m+dy, n+dx),
// dy, dx in {0,1} -- through sums of kernel taps that depend on the output phase (py, px).  So
//     y[b, 2m+py, 2n+px, o] = relu( b4[(py,px,o)] + sum_{dy,dx,c} w4[(py,px,o)][(dy,dx,c)] * x[b, c, m+dy, n+dx] )
// with K = 16.  The host folds the 64x4x4x4 conv weight into w4 [256][16] (an einsum, differentiable),
// these kernels do the rest: no 68x68 tensor, no separate bias / ReLU / layout passes.  The layer is
// HBM-bound on its 139 KB (bf16, NHWC) of output per sample; the 1.1 MFLOP per sample run on the FMA pipe.
//
// Thread roles: role = (phase, channel group of 4) -- 64 roles, weights of a role live in 64 registers;
// a 256-thread CTA is 4 pixel lanes x 64 roles and walks the samples of its share of the batch.
#pragma once
#include <cuda_bf16.h>

#include "ta_common.cuh"

namespace ta {

constexpr int C1_THREADS = 256;
constexpr int C1_OUT = 33;    // output height = width
constexpr int C1_CH = 64;

// matrix_env LUT over the featuriser's codes (ppo.MATRIX_LUT): 0 -> 0.9, 1 -> -0.9, 2 -> -0.5, 4 -> 0.3
__device__ __forceinline__ float c1_decode(uint32_t code) {
    return code == 0u ? 0.9f : (code == 1u ? -0.9f : (code == 2u ? -0.5f : (code == 4u ? 0.3f : 0.0f)));
}

// one sample's four frames -> shared memory as [18][18] float4 (c = frame), zero row / column 17
template <typename XT>
__device__ __forceinline__ void c1_stage_input(const XT *x, float4 *sx) {
    for (int i = threadIdx.x; i < 18 * 18; i += C1_THREADS) {
        const int m = i / 18, n = i - 18 * m;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < 17 && n < 17) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if constexpr (sizeof(XT) == 1) v[c] = c1_decode((uint32_t)x[c * NCELL + m * GS + n]);
                else v[c] = (float)x[c * NCELL + m * GS + n];
            }
        }
        sx[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__device__ __forceinline__ void c1_patch(const float4 *sx, int m, int n, float (&p)[16]) {
    const float4 a = sx[m * 18 + n], b = sx[m * 18 + n + 1], c = sx[(m + 1) * 18 + n], d = sx[(m + 1) * 18 + n + 1];
    p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w;      // (dy,dx) = (0,0), c = 0..3
    p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;      // (0,1)
    p[8] = c.x; p[9] = c.y; p[10] = c.z; p[11] = c.w;    // (1,0)
    p[12] = d.x; p[13] = d.y; p[14] = d.z; p[15] = d.w;  // (1,1)
}

// x: XT [B] samples of 4 frames x 289, `xstride` elements apart; w4 float [256][16], b4 float [256]
// (row = (py*2+px)*64 + o); y: bf16 [B][33][33][64].
template <typename XT>
__global__ void __launch_bounds__(C1_THREADS) conv1_fwd_kernel(const XT *__restrict__ x, long long xstride,
                                                              const float *__restrict__ w4, const float *__restrict__ b4,
                                                              long long B, __nv_bfloat16 *__restrict__ y,
                                                              const __nv_bfloat16 *__restrict__ addend = nullptr, int relu = 1) {
    // addend (nullable, bf16 like y): added to the pre-activation; relu == 0: the pre-activation itself is written.  Together they
    // make a first layer with MORE than four input channels out of passes over four channels each (the layer is linear in
    // its input channels): ta_conv1_fwd_add.
    __shared__ float4 sx[2][18 * 18];
    const int role = threadIdx.x & 63, plane = threadIdx.x >> 6;
    const int phase = role >> 4, cg = role & 15, py = phase >> 1, px = phase & 1;
    float w[4][16], bias[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = phase * 64 + cg * 4 + i;
        bias[i] = __ldg(b4 + row);
#pragma unroll
        for (int k = 0; k < 16; k++) w[i][k] = __ldg(w4 + row * 16 + k);
    }
    const int M = 17 - py, N = 17 - px, npix = M * N;
    int buf = 0;
    for (long long b = blockIdx.x; b < B; b += gridDim.x, buf ^= 1) {
        c1_stage_input<XT>(x + b * xstride, sx[buf]);
        __syncthreads();  // (the other buffer is still being read by slower threads: double buffered)
        __nv_bfloat16 *yb = y + b * (long long)(C1_OUT * C1_OUT * C1_CH);
        // pixel idx = plane + 4j of this phase's M x N grid, walked without divisions (N >= 16), four pixels per trip: with an
        // addend their four 8-byte loads are in flight together before the first pixel's 64 FMAs start (one load per pixel
        // issued where it is needed left the pass latency-bound: 702 us against 343 us without the addend, B = 4096)
        int m = 0, n = plane;
        for (int idx0 = plane; idx0 < npix; idx0 += 16) {
            int mm[4], nn[4];
            uint2 av[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const bool ok = idx0 + 4 * u < npix;
                if (n >= N) { n -= N; m++; }
                mm[u] = ok ? m : -1;
                nn[u] = n;
                n += 4;
                av[u] = make_uint2(0u, 0u);
                if (addend && ok)
                    av[u] = __ldg(reinterpret_cast<const uint2 *>(addend + b * (long long)(C1_OUT * C1_OUT * C1_CH) +
                                                                  ((2 * m + py) * C1_OUT + (2 * nn[u] + px)) * C1_CH + cg * 4));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (mm[u] < 0) continue;
                float p[16];
                c1_patch(sx[buf], mm[u], nn[u], p);
                const float2 a01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&av[u].x));
                const float2 a23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&av[u].y));
                const float add[4] = {a01.x, a01.y, a23.x, a23.y};
                // Blackwell's packed fp32 FMA (FFMA2): even / odd k accumulate in the two halves of a float2
                float acc[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float2 a2 = make_float2(bias[i], 0.f);
#pragma unroll
                    for (int k = 0; k < 16; k += 2)
                        a2 = __ffma2_rn(make_float2(w[i][k], w[i][k + 1]), make_float2(p[k], p[k + 1]), a2);
                    const float a = a2.x + a2.y + add[i];
                    acc[i] = (a > 0.f || !relu) ? a : 0.f;
                }
                const __nv_bfloat162 lo = __floats2bfloat162_rn(acc[0], acc[1]), hi = __floats2bfloat162_rn(acc[2], acc[3]);
                uint2 out;
                out.x = *reinterpret_cast<const uint32_t *>(&lo);
                out.y = *reinterpret_cast<const uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(yb + ((2 * mm[u] + py) * C1_OUT + (2 * nn[u] + px)) * C1_CH + cg * 4) = out;
            }
        }
    }
}

// dw4 [256][16] += sum over the batch of dz (x) patch, db4 [256] += sum dz, dz = dy where y > 0.
template <typename XT>
__global__ void __launch_bounds__(C1_THREADS, 2) conv1_bwd_kernel(const XT *__restrict__ x, long long xstride,
                                                              const __nv_bfloat16 *__restrict__ y,
                                                              const __nv_bfloat16 *__restrict__ dy, long long B,
                                                              float *__restrict__ dw4, float *__restrict__ db4) {
    __shared__ float4 sx[2][18 * 18];
    __shared__ float sacc[64][17 * 4 + 1];
    const int role = threadIdx.x & 63, plane = threadIdx.x >> 6;
    const int phase = role >> 4, cg = role & 15, py = phase >> 1, px = phase & 1;
    float dw[4][16], db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int k = 0; k < 16; k++) dw[i][k] = 0.f;
    const int M = 17 - py, N = 17 - px, npix = M * N;
    int buf = 0;
    for (long long b = blockIdx.x; b < B; b += gridDim.x, buf ^= 1) {
        c1_stage_input<XT>(x + b * xstride, sx[buf]);
        __syncthreads();
        const long long base = b * (long long)(C1_OUT * C1_OUT * C1_CH);
        // four pixels per trip, their eight 8-byte loads in flight together (the loop is latency-bound otherwise)
        int m = 0, n = plane;
        for (int idx0 = plane; idx0 < npix; idx0 += 16) {
            uint2 yv[4], gv[4];
            int mm[4], nn[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = idx0 + 4 * u;
                const bool ok = idx < npix;
                if (n >= N) { n -= N; m++; }
                mm[u] = ok ? m : 0;
                nn[u] = ok ? n : 0;
                n += 4;
                const long long off = base + ((2 * mm[u] + py) * C1_OUT + (2 * nn[u] + px)) * C1_CH + cg * 4;
                yv[u] = ok ? __ldg(reinterpret_cast<const uint2 *>(y + off)) : make_uint2(0u, 0u);
                gv[u] = ok ? __ldg(reinterpret_cast<const uint2 *>(dy + off)) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&yv[u].x));
                const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&yv[u].y));
                const float2 g01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&gv[u].x));
                const float2 g23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&gv[u].y));
                const float dz[4] = {y01.x > 0.f ? g01.x : 0.f, y01.y > 0.f ? g01.y : 0.f, y23.x > 0.f ? g23.x : 0.f,
                                     y23.y > 0.f ? g23.y : 0.f};  // (a pixel past the end contributes zeros)
                float p[16];
                c1_patch(sx[buf], mm[u], nn[u], p);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    db[i] += dz[i];
                    const float2 d2 = make_float2(dz[i], dz[i]);
#pragma unroll
                    for (int k = 0; k < 16; k += 2) {
                        const float2 r = __ffma2_rn(d2, make_float2(p[k], p[k + 1]), make_float2(dw[i][k], dw[i][k + 1]));
                        dw[i][k] = r.x;
                        dw[i][k + 1] = r.y;
                    }
                }
            }
        }
    }
    // the four pixel lanes of a role -> shared memory -> one atomicAdd per value and CTA
    for (int pl = 0; pl < 4; pl++) {
        __syncthreads();
        if (plane == pl) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    float *s = &sacc[role][i * 17 + k];
                    *s = pl == 0 ? dw[i][k] : *s + dw[i][k];
                }
                float *s = &sacc[role][i * 17 + 16];
                *s = pl == 0 ? db[i] : *s + db[i];
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 64 * 68; j += C1_THREADS) {
        const int r = j / 68, q = j - 68 * r, i = q / 17, k = q - 17 * i;
        const int ph = r >> 4, g = r & 15, row = ph * 64 + g * 4 + i;
        const float v = sacc[r][q];
        if (k < 16) atomicAdd(dw4 + row * 16 + k, v);
        else atomicAdd(db4 + row, v);
    }
}

// ---- data gradient of a stride-2, unpadded convolution in channels-last bf16: col2im ---------------
// dcols bf16 [B*OH*OW][KS*KS*C] (one row per output pixel, columns (ky,kx,c)) is the GEMM dY x W; every
// input element gathers the <= ceil(KS/2)^2 patch entries that cover it (no atomics, each dcols byte is
// read exactly once, 16-byte loads / stores over 8 channels).  dx bf16 [B][H][W][C].
template <int KS>
__global__ void __launch_bounds__(256) col2im_s2_kernel(const __nv_bfloat16 *__restrict__ dcols, __nv_bfloat16 *__restrict__ dx,
                                                       int H, int W, int C, int OH, int OW) {
    // one CTA per input row (b, y); a thread = 8 channels of one pixel of that row
    const int c8n = C >> 3;
    const long long b = blockIdx.x / H;
    const int yy = blockIdx.x - (int)b * H;
    const long long rowlen = (long long)KS * KS * C;
    for (int t = threadIdx.x; t < W * c8n; t += blockDim.x) {
    const int xx = t / c8n, c8 = t - xx * c8n;
    const long long i = ((long long)blockIdx.x * W + xx) * c8n + c8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < KS; ky++) {
        const int ty = yy - ky;
        if (ty < 0 || (ty & 1) || (ty >> 1) >= OH) continue;
#pragma unroll
        for (int kx = 0; kx < KS; kx++) {
            const int tx = xx - kx;
            if (tx < 0 || (tx & 1) || (tx >> 1) >= OW) continue;
            const long long row = (b * OH + (ty >> 1)) * OW + (tx >> 1);
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(dcols + row * rowlen + (ky * KS + kx) * C + c8 * 8));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[q]));
                acc[2 * q] += f.x;
                acc[2 * q + 1] += f.y;
            }
        }
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
        o[q] = *reinterpret_cast<const uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(dx + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---- patches of a stride-2, unpadded convolution in channels-last bf16: im2col (col2im's forward) -----------
// x bf16 [B][H][W][C] -> cols bf16 [B*OH*OW][KS*KS*C], columns (ky,kx,c): one 16-byte chunk (8 channels) per
// thread and trip, grid-strided; consecutive threads write consecutive chunks.
__global__ void __launch_bounds__(256) im2col_s2_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ cols, unsigned total,
                                                       int H, int W, int c8n, int OH, int OW, int KS) {
    const unsigned row_chunks = (unsigned)(KS * c8n), patch_chunks = (unsigned)KS * row_chunks;  // one patch row (kx, c) is contiguous in x
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned pix = i / patch_chunks, in_patch = i - pix * patch_chunks;   // output pixel (b, oy, ox)
        const unsigned ky = in_patch / row_chunks, in_row = in_patch - ky * row_chunks;
        const unsigned bo = pix / (unsigned)OW, ox = pix - bo * (unsigned)OW;
        const unsigned b = bo / (unsigned)OH, oy = bo - b * (unsigned)OH;
        cols[i] = __ldg(x + ((size_t)(b * H + 2 * oy + ky) * W + 2 * ox) * c8n + in_row);
    }
}

// ---- the four parity-class kernels of a stride-2 convolution's data gradient as ONE conv2d weight ------------
// The data gradient of a k x k stride-2 unpadded convolution (k = 3, 4) splits by the parity (pa, pb) of the input
// pixel: dx[2i+pa, 2j+pb] = sum over the taps ky = pa, kx = pb (mod 2) of dz[i - ky/2, j - kx/2] . w[:, :, ky, kx],
// a stride-1 convolution of dz with a <= 2x2 sub-kernel.  All four are computed by one stride-1 convolution with
// 4*cin output channels, a 2x2 kernel and padding 1 ("merged planes": output [B][OH+1][OW+1][4][cin], class
// c = pa*2+pb in channel block c; a class with a single tap along an axis gets a zero in the other slot, and its
// extra last row / column comes out as zero).  This kernel builds that weight:
// w bf16 [cout][cin][k][k] (any strides) -> out = conv2d weight [4*cin][cout][2][2] in channels-last memory
//   out[c*cin + ci][u][v][co] = w[co][ci][ky][kx],  ky = taps(pa) == 2 ? 2*(1-u)+pa : (u == 1 ? pa : none), kx alike
__global__ void __launch_bounds__(256) parity_class_weights_kernel(const __nv_bfloat16 *__restrict__ w, long long so, long long si,
                                                                  long long sy, long long sx, int cout, int cin, int k,
                                                                  __nv_bfloat16 *__restrict__ out) {
    const int total = 4 * cin * 4 * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r = i;
        const int co = r % cout; r /= cout;
        const int v = r & 1, u = (r >> 1) & 1; r >>= 2;
        const int ci = r % cin, c = r / cin, pa = c >> 1, pb = c & 1;
        const int kh = (k - pa + 1) / 2, kw = (k - pb + 1) / 2;
        const int ky = kh == 2 ? 2 * (1 - u) + pa : (u == 1 ? pa : -1);
        const int kx = kw == 2 ? 2 * (1 - v) + pb : (v == 1 ? pb : -1);
        out[i] = (ky >= 0 && kx >= 0) ? w[co * so + ci * si + ky * sy + kx * sx] : __float2bfloat16(0.f);
    }
}

// ---- merged parity planes -> dense, fused with the ReLU backward of the layer below ------------------------------
// planes bf16 [B][OH+1][OW+1][4][C] (see above) -> the dense channels-last gradient, masked with the ReLU of the
// layer that produced the convolution's input (y [B][H][W][C]):
//   dz_out[b][h][w][c] = y > 0 ? planes[b][h>>1][w>>1][(h&1)*2 + (w&1)][c] : 0
// = aten::threshold_backward's traffic, so the interleave comes for free.
__global__ void __launch_bounds__(256) planes_to_dense_relu_kernel(const uint4 *__restrict__ planes, const uint4 *__restrict__ y,
                                                                  uint4 *__restrict__ out, unsigned total, int H, int W, int c8n,
                                                                  int OH, int OW) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned pix = i / (unsigned)c8n, c8 = i - pix * (unsigned)c8n;
        const unsigned bh = pix / (unsigned)W, w = pix - bh * (unsigned)W;
        const unsigned b = bh / (unsigned)H, h = bh - b * (unsigned)H;
        const unsigned cls = (h & 1u) * 2u + (w & 1u), hp = h >> 1, wp = w >> 1;
        const uint4 g = __ldg(planes + (((size_t)(b * (OH + 1) + hp) * (OW + 1) + wp) * 4u + cls) * c8n + c8);
        const uint4 yv = __ldg(y + i);
        auto m = [](uint32_t gw, uint32_t yw) {  // bf16 pairs; y is a ReLU output (>= +0)
            return gw & (((yw & 0xFFFFu) ? 0xFFFFu : 0u) | ((yw >> 16) ? 0xFFFF0000u : 0u));
        };
        out[i] = make_uint4(m(g.x, yv.x), m(g.y, yv.y), m(g.z, yv.z), m(g.w, yv.w));
    }
}

// ---- per-channel sum of a channels-last bf16 tensor (the bias gradient of a convolution) -------------
// x bf16 [rows][C], C in {64, 128, ..., 2048}: a thread owns 8 channels (16-byte loads) and strides over the
// rows of its CTA's slab; row lanes are combined through shared memory, one atomicAdd per channel and CTA.
// (aten::sum over (N,H,W) of a channels-last tensor runs far below the HBM rate; this is one streaming pass.)
__global__ void __launch_bounds__(256) channel_sum_bf16_kernel(const __nv_bfloat16 *__restrict__ x, long long rows, int C,
                                                              float *__restrict__ out) {
    __shared__ float sacc[256][9];
    const int c8n = C >> 3, c8 = threadIdx.x % c8n, rl = threadIdx.x / c8n, nrl = 256 / c8n;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long r = (long long)blockIdx.x * nrl + rl; r < rows; r += (long long)gridDim.x * nrl) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(x + r * C + c8 * 8));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[q]));
            acc[2 * q] += f.x;
            acc[2 * q + 1] += f.y;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) sacc[threadIdx.x][q] = acc[q];
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += 256) {  // channel ch: sum over the row lanes
        const int g = ch >> 3, q = ch & 7;
        float s = 0.f;
        for (int l = 0; l < nrl; l++) s += sacc[l * c8n + g][q];
        atomicAdd(out + ch, s);
    }
}

}  // namespace ta
