// HRNet-W40 glue kernels (reference lib/models/backbones/hrnet.py, lib/models/POEM.py:189-229): layout changes, stem,
// fuse-layer sums, upsampling / concat, heatmap soft-argmax, DLT triangulation.
// Activations live as NHWC op16.  Inside the image half a pixel keeps only its live channels rounded up to 16
// (48 / 80 / 160 / 320: "compact storage"); every convolution is a tcgen05 GEMM whose A operand TMA gathers (conv3x3.cuh
// for 3x3 stride 1, the ConvOperand mode of gemm.cuh otherwise) — a 64-channel box that runs past a pixel's channels is
// zero-filled by the hardware; BatchNorm (eval) and conv biases are folded into the weights at pack time.
#pragma once
#include "common.cuh"

namespace poem {

// (N, C, H*W) fp32 -> (N, H*W, Cp) op16, channels >= C zero-filled.  32x32 smem transpose tiles.
__global__ void nchw_f32_to_nhwc_op16_kernel(const float* __restrict__ in, op16* __restrict__ out, int C, int Cp,
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
    if (c < Cp && p < HW) out[((size_t)n * HW + p) * Cp + c] = f2op16(tile[tx][i]);
  }
}

// (N, H*W, Cp) op16 -> (N, C, H*W) fp32 (only the C real channels)
__global__ void nhwc_op16_to_nchw_f32_kernel(const op16* __restrict__ in, float* __restrict__ out, int C, int Cp,
                                             int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (c < Cp && p < HW) ? op16_to_f(in[((size_t)n * HW + p) * Cp + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (c < C && p < HW) out[((size_t)n * C + c) * HW + p] = tile[tx][i];
  }
}

// Stem conv1 (hrnet.py:245-246, 388-390): 3x3 stride-2 convolution 3 -> 64 channels on the NCHW fp32 image, BatchNorm
// folded, ReLU; writes NHWC op16 (64 channels = one swizzle atom) for the implicit-GEMM convolutions that follow.
// K = 27 is far too small for the tensor core: one thread per output pixel, weights broadcast from smem.
//   w: fp32 [64][27] with k = (ky*3 + kx)*3 + c ; b: fp32 [64]
__global__ void __launch_bounds__(128)
stem_conv1_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b,
                  op16* __restrict__ out, int N, int H, int W) {
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
  op16* o = out + pix * 64;
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
    pk.x = pack_op16x2(acc[0], acc[1]);
    pk.y = pack_op16x2(acc[2], acc[3]);
    pk.z = pack_op16x2(acc[4], acc[5]);
    pk.w = pack_op16x2(acc[6], acc[7]);
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

// uv_decode step input (POEM.py:208-210): cat(F.interpolate(hi, x2, bilinear, align_corners=False), lo) along channels,
// NHWC op16, written with the channel count padded to 64 (zeros).  C_hi and C_lo are multiples of 8, so an 8-channel
// group never straddles the two sources.  hi: (N, H, W, Cp_hi), lo: (N, 2H, 2W, Cp_lo), out: (N, 2H, 2W, Cp_out).
__global__ void upsample2x_concat_kernel(const op16* __restrict__ hi, const op16* __restrict__ lo,
                                         op16* __restrict__ out, int H, int W, int Cp_hi, int C_hi, int Cp_lo,
                                         int C_lo, int Cp_out, size_t total8) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total8) return;
  const int c8 = Cp_out / 8;
  const int c = (int)(gid % c8) * 8;
  size_t pix = gid / c8;
  const int Wo = 2 * W, Ho = 2 * H;
  const int x = (int)(pix % Wo);
  pix /= Wo;
  const int y = (int)(pix % Ho);
  const size_t n = pix / Ho;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (c < C_hi) {
    const float sy = fmaxf((y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((x + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const op16* b = hi + n * H * W * Cp_hi + c;
    const uint4 q00 = *reinterpret_cast<const uint4*>(b + ((size_t)y0 * W + x0) * Cp_hi);
    const uint4 q01 = *reinterpret_cast<const uint4*>(b + ((size_t)y0 * W + x1) * Cp_hi);
    const uint4 q10 = *reinterpret_cast<const uint4*>(b + ((size_t)y1 * W + x0) * Cp_hi);
    const uint4 q11 = *reinterpret_cast<const uint4*>(b + ((size_t)y1 * W + x1) * Cp_hi);
    const op16x2* p00 = reinterpret_cast<const op16x2*>(&q00);
    const op16x2* p01 = reinterpret_cast<const op16x2*>(&q01);
    const op16x2* p10 = reinterpret_cast<const op16x2*>(&q10);
    const op16x2* p11 = reinterpret_cast<const op16x2*>(&q11);
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = op16x2_to_f2(p00[i]), bq = op16x2_to_f2(p01[i]);
      const float2 cq = op16x2_to_f2(p10[i]), d = op16x2_to_f2(p11[i]);
      const float vx = (1.f - ly) * ((1.f - lx) * a.x + lx * bq.x) + ly * ((1.f - lx) * cq.x + lx * d.x);
      const float vy = (1.f - ly) * ((1.f - lx) * a.y + lx * bq.y) + ly * ((1.f - lx) * cq.y + lx * d.y);
      r[i] = pack_op16x2(vx, vy);
    }
    o = make_uint4(r[0], r[1], r[2], r[3]);
  } else if (c < C_hi + C_lo) {
    o = *reinterpret_cast<const uint4*>(lo + ((n * Ho + y) * Wo + x) * Cp_lo + (c - C_hi));
  }
  *reinterpret_cast<uint4*>(out + gid * 8) = o;
}

// Heatmap head (POEM.py:213-229): 2x2 max-pool of the (N, 2R, 2R, Cp) map, 1x1 convolution C -> J + sigmoid,
// pdf normalisation (sum + 1e-6) and the integral soft-argmax (lib/models/integal_pose.py:196-220), scaled to pixels.
// One block per image; thread t owns pooled pixels t, t + 256, ...; per-joint sums reduced through smem.
//   w: fp32 [J][C], b: fp32 [J]; uv_px: (N, J, 2) = (u * img_w, v * img_h); heat: optional (N, J, R, R) fp32
constexpr int HEAT_MAX_J = 24;
__global__ void __launch_bounds__(256)
heatmap_uv_kernel(const op16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                  float* __restrict__ uv_px, float* __restrict__ heat, int R, int Cp, int C, int J, float img_w,
                  float img_h) {
  extern __shared__ float sm[];
  float* sw = sm;                 // [J][C]
  float* sb = sw + J * C;         // [J]
  float* red = sb + J;            // [8 warps][J][3]
  for (int i = threadIdx.x; i < J * C; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < J; i += blockDim.x) sb[i] = b[i];
  __syncthreads();
  const size_t n = blockIdx.x;
  const int R2 = 2 * R;
  float s0[HEAT_MAX_J], su[HEAT_MAX_J], sv[HEAT_MAX_J];
#pragma unroll
  for (int j = 0; j < HEAT_MAX_J; ++j) s0[j] = su[j] = sv[j] = 0.f;
  for (int p = threadIdx.x; p < R * R; p += blockDim.x) {
    const int py = p / R, px = p - py * R;
    float acc[HEAT_MAX_J];
#pragma unroll
    for (int j = 0; j < HEAT_MAX_J; ++j) acc[j] = (j < J) ? sb[j] : 0.f;
    const op16* base = x + ((n * R2 + 2 * py) * R2 + 2 * px) * Cp;
    for (int c = 0; c < C; c += 8) {
      float m[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 q = *reinterpret_cast<const uint4*>(base + ((size_t)(k >> 1) * R2 + (k & 1)) * Cp + c);
        const op16x2* q2 = reinterpret_cast<const op16x2*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = op16x2_to_f2(q2[i]);
          m[2 * i] = k ? fmaxf(m[2 * i], f.x) : f.x;
          m[2 * i + 1] = k ? fmaxf(m[2 * i + 1], f.y) : f.y;
        }
      }
#pragma unroll
      for (int j = 0; j < HEAT_MAX_J; ++j)
        if (j < J) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j] = fmaf(sw[j * C + c + i], m[i], acc[j]);
        }
    }
    const float fu = (float)px / (float)R, fv = (float)py / (float)R;
#pragma unroll
    for (int j = 0; j < HEAT_MAX_J; ++j)
      if (j < J) {
        const float h = 1.f / (1.f + __expf(-acc[j]));
        if (heat != nullptr) heat[((n * J + j) * R + py) * R + px] = h;
        s0[j] += h, su[j] += h * fu, sv[j] += h * fv;
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < HEAT_MAX_J; ++j)
    if (j < J) {
      float a0 = s0[j], a1 = su[j], a2 = sv[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      if (lane == 0) red[(warp * J + j) * 3] = a0, red[(warp * J + j) * 3 + 1] = a1, red[(warp * J + j) * 3 + 2] = a2;
    }
  __syncthreads();
  if ((int)threadIdx.x < J) {
    const int j = threadIdx.x;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi)
      a0 += red[(wi * J + j) * 3], a1 += red[(wi * J + j) * 3 + 1], a2 += red[(wi * J + j) * 3 + 2];
    const float inv = 1.f / (a0 + 1e-6f);
    uv_px[(n * J + j) * 2] = a1 * inv * img_w;
    uv_px[(n * J + j) * 2 + 1] = a2 * inv * img_h;
  }
}

// DLT triangulation of J joints from the views of each sample (lib/utils/triangulation.py:5-45 as called from
// POEM.py:284-299): per view M = K * inv(cam_extr)[:3, :]; rows u*M[2] - M[0], v*M[2] - M[1]; X = the right singular
// vector of the smallest singular value, dehomogenised with (w + 1e-7).  The reference calls cuSOLVER's batched SVD in
// a per-sample Python loop; here one block per sample, one thread per joint runs a one-sided (Hestenes) Jacobi SVD of
// the 2V x 4 system in fp64 (works on A itself, so the conditioning is that of A, not of A^T A).
constexpr int DLT_MAX_VIEWS = 16;
__global__ void __launch_bounds__(32)
dlt_triangulate_kernel(const float* __restrict__ uv_px, const float* __restrict__ cam_intr,
                       const float* __restrict__ cam_extr, const int* __restrict__ view_counts, float* __restrict__ out,
                       int J) {
  __shared__ double sM[DLT_MAX_VIEWS][12];
  const int s = blockIdx.x;
  int base = 0;
  for (int i = 0; i < s; ++i) base += view_counts[i];
  const int V = min(view_counts[s], DLT_MAX_VIEWS);
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    // inverse of the 4x4 extrinsic by Gauss-Jordan with partial pivoting (fp64)
    double a[4][8];
    const float* E = cam_extr + (size_t)(base + v) * 16;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) a[r][c] = (double)E[r * 4 + c], a[r][4 + c] = (r == c) ? 1.0 : 0.0;
    for (int col = 0; col < 4; ++col) {
      int piv = col;
      for (int r = col + 1; r < 4; ++r)
        if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
      if (piv != col)
        for (int c = 0; c < 8; ++c) {
          const double t = a[col][c];
          a[col][c] = a[piv][c];
          a[piv][c] = t;
        }
      const double d = 1.0 / a[col][col];
      for (int c = 0; c < 8; ++c) a[col][c] *= d;
      for (int r = 0; r < 4; ++r)
        if (r != col) {
          const double f = a[r][col];
          for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
        }
    }
    const float* K = cam_intr + (size_t)(base + v) * 9;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        double acc = 0.0;
        for (int k = 0; k < 3; ++k) acc += (double)K[r * 3 + k] * a[k][4 + c];
        sM[v][r * 4 + c] = acc;
      }
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j >= J) return;
  double A[2 * DLT_MAX_VIEWS][4];
  double Vm[4][4];
  for (int v = 0; v < V; ++v) {
    const double u = (double)uv_px[((size_t)(base + v) * J + j) * 2], w = (double)uv_px[((size_t)(base + v) * J + j) * 2 + 1];
    for (int c = 0; c < 4; ++c) {
      A[2 * v][c] = u * sM[v][8 + c] - sM[v][c];
      A[2 * v + 1][c] = w * sM[v][8 + c] - sM[v][4 + c];
    }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) Vm[r][c] = (r == c) ? 1.0 : 0.0;
  const int rows = 2 * V;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int r = 0; r < rows; ++r) al += A[r][p] * A[r][p], be += A[r][q] * A[r][q], ga += A[r][p] * A[r][q];
        if (fabs(ga) <= 1e-15 * sqrt(al * be) || ga == 0.0) continue;
        off = fmax(off, fabs(ga) / sqrt(al * be));
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int r = 0; r < rows; ++r) {
          const double x = A[r][p], y = A[r][q];
          A[r][p] = cs * x - sn * y;
          A[r][q] = sn * x + cs * y;
        }
        for (int r = 0; r < 4; ++r) {
          const double x = Vm[r][p], y = Vm[r][q];
          Vm[r][p] = cs * x - sn * y;
          Vm[r][q] = sn * x + cs * y;
        }
      }
    if (off < 1e-14) break;
  }
  int best = 0;
  double best_n = 1e300;
  for (int c = 0; c < 4; ++c) {
    double nn = 0.0;
    for (int r = 0; r < rows; ++r) nn += A[r][c] * A[r][c];
    if (nn < best_n) best_n = nn, best = c;
  }
  // torch.linalg.svd leaves the sign of the singular vector free; X = v[:3] / (v[3] + 1e-7) only sees it through the
  // 1e-7 guard, so orient v with v[3] >= 0 (|v[3]| ~ 0.5-1 for points in front of the cameras)
  double sg = Vm[3][best] >= 0.0 ? 1.0 : -1.0;
  const double wv = sg * Vm[3][best] + 1e-7;
  for (int c = 0; c < 3; ++c) out[((size_t)s * J + j) * 3 + c] = (float)(sg * Vm[c][best] / wv);
}

// Fuse layer sum (hrnet.py:225-233): out[n,y,x,c] = relu(sum_j in_j[n, y >> s_j, x >> s_j, c]); in_j has resolution
// (H >> s_j, W >> s_j) — nearest-neighbour upsampling by 2^s_j of the 1x1-conv terms, s_j = 0 for the others.
struct FuseSumArgs {
  const op16* in[4];
  int shift[4];
  int n_in;
};
// 16 channels (one 256-bit access per term) per thread
__global__ void __launch_bounds__(256)
fuse_sum_relu_kernel(FuseSumArgs a, op16* __restrict__ out, int H, int W, int Cp, size_t total16) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total16) return;
  const int c16 = Cp / 16;
  const int c = (int)(gid % c16) * 16;
  size_t pix = gid / c16;
  const int x = (int)(pix % W);
  pix /= W;
  const int y = (int)(pix % H);
  const size_t n = pix / H;
  uint32_t v[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {   // all loads in flight before the first use
    if (j < a.n_in) {
      const int s = a.shift[j];
      const int hj = H >> s, wj = W >> s;
      ldg_nc_256(a.in[j] + ((n * hj + (y >> s)) * wj + (x >> s)) * Cp + c, v[j]);
    }
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < a.n_in) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f = op16x2_to_f2(*reinterpret_cast<const op16x2*>(&v[j][i]));
        acc[2 * i] += f.x;
        acc[2 * i + 1] += f.y;
      }
    }
  }
  uint32_t o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = pack_op16x2(fmaxf(acc[2 * i], 0.f), fmaxf(acc[2 * i + 1], 0.f));
  stg_256(out + gid * 16, o);
}

}  // namespace poem
