// HRNet-W40 stage 4 glue kernels (reference lib/models/backbones/hrnet.py:38-67,217-234,272-277).
// Activations live as NHWC bf16 with the channel count padded to a multiple of 64 (one SWIZZLE_128B atom), so
// every 3x3 / 1x1 convolution is an implicit GEMM of gemm_bf16_tc_kernel (TMA gathers the shifted image rows, zero
// fill = padding); BatchNorm (eval) is folded into the weights/bias at pack time.
#pragma once
#include "common.cuh"

namespace poem {

// (N, C, H*W) fp32 -> (N, H*W, Cp) bf16, channels >= C zero-filled.  32x32 smem transpose tiles.
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, int Cp,
                                             int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < C && p < HW) ? in[((size_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (c < Cp && p < HW) out[((size_t)n * HW + p) * Cp + c] = __float2bfloat16(tile[tx][i]);
  }
}

// (N, H*W, Cp) bf16 -> (N, C, H*W) fp32 (only the C real channels)
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int C, int Cp,
                                             int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (c < Cp && p < HW) ? __bfloat162float(in[((size_t)n * HW + p) * Cp + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (c < C && p < HW) out[((size_t)n * C + c) * HW + p] = tile[tx][i];
  }
}

// Stem conv1 (hrnet.py:245-246, 388-390): 3x3 stride-2 convolution 3 -> 64 channels on the NCHW fp32 image, BatchNorm
// folded, ReLU; writes NHWC bf16 (64 channels = one swizzle atom) for the implicit-GEMM convolutions that follow.
// K = 27 is far too small for the tensor core: one thread per output pixel, weights broadcast from smem.
//   w: fp32 [64][27] with k = (ky*3 + kx)*3 + c ; b: fp32 [64]
__global__ void __launch_bounds__(128)
stem_conv1_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b,
                  __nv_bfloat16* __restrict__ out, int N, int H, int W) {
  __shared__ float sw[64 * 27];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 64 * 27; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)N * Ho * Wo) return;
  const int xo = (int)(pix % Wo);
  const int yo = (int)((pix / Wo) % Ho);
  const size_t n = pix / ((size_t)Wo * Ho);
  float in[27];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int y = 2 * yo + ky - 1, x = 2 * xo + kx - 1;
      const bool ok = (y >= 0 && y < H && x >= 0 && x < W);
#pragma unroll
      for (int c = 0; c < 3; ++c) in[(ky * 3 + kx) * 3 + c] = ok ? img[((n * 3 + c) * H + y) * W + x] : 0.f;
    }
  __nv_bfloat16* o = out + pix * 64;
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = sb[c0 + j];
      const float* wr = sw + (c0 + j) * 27;
#pragma unroll
      for (int k = 0; k < 27; ++k) a = fmaf(wr[k], in[k], a);
      acc[j] = fmaxf(a, 0.f);
    }
    uint4 pk;
    pk.x = pack_bf16x2(acc[0], acc[1]);
    pk.y = pack_bf16x2(acc[2], acc[3]);
    pk.z = pack_bf16x2(acc[4], acc[5]);
    pk.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(o + c0) = pk;
  }
}

// F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) of an fp32 NHWC map (padded channels) written as
// the fp32 NCHW tensor (N, C, 2H, 2W) the head consumes (POEM.py:200).  Source index = (dst + 0.5) / 2 - 0.5 clamped at 0.
__global__ void upsample2x_nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W,
                                               int Cp, int C) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ho = 2 * H, Wo = 2 * W;
  if (gid >= (size_t)N * C * Ho * Wo) return;
  const int x = (int)(gid % Wo);
  const int y = (int)((gid / Wo) % Ho);
  const int c = (int)((gid / ((size_t)Wo * Ho)) % C);
  const size_t n = gid / ((size_t)Wo * Ho * C);
  const float sy = fmaxf((y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((x + 0.5f) * 0.5f - 0.5f, 0.f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float* b = in + n * H * W * Cp + c;
  const float v00 = b[((size_t)y0 * W + x0) * Cp], v01 = b[((size_t)y0 * W + x1) * Cp];
  const float v10 = b[((size_t)y1 * W + x0) * Cp], v11 = b[((size_t)y1 * W + x1) * Cp];
  out[gid] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// Fuse layer sum (hrnet.py:225-233): out[n,y,x,c] = relu(sum_j in_j[n, y >> s_j, x >> s_j, c]); in_j has resolution
// (H >> s_j, W >> s_j) — nearest-neighbour upsampling by 2^s_j of the 1x1-conv terms, s_j = 0 for the others.
struct FuseSumArgs {
  const __nv_bfloat16* in[4];
  int shift[4];
  int n_in;
};
__global__ void fuse_sum_relu_kernel(FuseSumArgs a, __nv_bfloat16* __restrict__ out, int H, int W, int Cp, size_t total8) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total8) return;
  const int c8 = Cp / 8;
  const int c = (int)(gid % c8) * 8;
  size_t pix = gid / c8;
  const int x = (int)(pix % W);
  pix /= W;
  const int y = (int)(pix % H);
  const size_t n = pix / H;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j = 0; j < a.n_in; ++j) {
    const int s = a.shift[j];
    const int hj = H >> s, wj = W >> s;
    const uint4 v = *reinterpret_cast<const uint4*>(a.in[j] + ((n * hj + (y >> s)) * wj + (x >> s)) * Cp + c);
    const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(v2[i]);
      acc[2 * i] += f.x;
      acc[2 * i + 1] += f.y;
    }
  }
  uint4 o;
  o.x = pack_bf16x2(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f));
  o.y = pack_bf16x2(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
  o.z = pack_bf16x2(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f));
  o.w = pack_bf16x2(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
  *reinterpret_cast<uint4*>(out + gid * 8) = o;
}

}  // namespace poem
