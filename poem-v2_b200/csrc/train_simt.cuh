// SIMT kernels of the training path (fp32): everything around the TF32 GEMMs of tgemm.cuh — activations, LayerNorm,
// row softmax of the BERT cross-attention, the per-edge pieces of the vector attention (32 neighbours per query), the
// bilinear sampler and the cross-view merge, each with its backward.  Reference: the autograd graph of
// lib/models/heads/ptEmb_head.py:745-771,825-964, lib/models/bricks/pt_metro_transformer.py:34-91,
// lib/models/bricks/point_transformers.py:70-156 (eval-mode arithmetic restated in oracle/poem_oracle.py).
#pragma once
#include "common.cuh"

namespace poem {

constexpr int TR_NBR = 32;   // neighbours per query (N_NEIGHBOR == N_NEIGHBOR_QUERY == 32 in every release config)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// round to TF32 (nearest, ties away): producers of tensors that are only ever GEMM operands store them rounded, so the GEMM
// can skip its shared-memory rounding pass for that operand (include/poem_train.h: round_ops / round_out)
__device__ __forceinline__ float tf32r(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------ dropout
// Counter-based: element `idx` of dropout site `site` in the step whose seed is seed[0] is kept iff
// mix(seed[0] + site * K1 + idx) >= p * 2^32 (splitmix64 finaliser).  Nothing is stored: the backward regenerates the
// mask from the same (seed, site).  The seed lives in device memory so that a captured CUDA graph draws fresh masks on
// every replay (the caller bumps it).  Reference: nn.Dropout at pt_metro_transformer.py:117,185-186 and inside the HF
// BertSelfAttention / BertSelfOutput / BertOutput the layers are built from (hidden 0.1, attention probabilities 0.1).
__device__ __forceinline__ bool tr_keep(unsigned long long seed, unsigned long long site, unsigned long long idx, uint32_t thr) {
  unsigned long long z = seed + site * 0xD1B54A32D192ED03ull + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32) >= thr;
}
__device__ __forceinline__ uint32_t tr_drop_threshold(float p) { return (uint32_t)fminf(p * 4294967296.0f, 4294967040.0f); }
// y = keep ? x / (1 - p) : 0     (in place allowed; applied to a gradient with the same seed / site it is the backward)
__global__ void tr_dropout_kernel(const float* x, float* y, long long n, float p, const unsigned long long* seed,
                                  unsigned long long site) {
  const unsigned long long sd = seed[0];
  const uint32_t thr = tr_drop_threshold(p);
  const float sc = 1.0f / (1.0f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = tr_keep(sd, site, (unsigned long long)i, thr) ? x[i] * sc : 0.f;
}

// ------------------------------------------------------------------------------------------------ elementwise
__global__ void tr_relu_kernel(float* y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaxf(y[i], 0.f);
}
__global__ void tr_relu_bwd_kernel(float* dy, const float* y, long long n) {   // dy *= (y > 0), y = ReLU output
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(y[i] > 0.f)) dy[i] = 0.f;
}
__global__ void tr_gelu_kernel(const float* x, float* y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  }
}
__global__ void tr_gelu_bwd_kernel(float* dy, const float* x, long long n) {   // dy *= d/dx [x Phi(x)]
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * v * v);
    dy[i] *= cdf + v * pdf;
  }
}
// y = x rounded to TF32 (nearest, ties away): the weights' operand copy, refreshed once per step, so that the GEMMs need
// to round only their activation operand in shared memory
__global__ void tr_round_tf32_kernel(const float* x, float* y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x[i]));
    y[i] = __uint_as_float(r);
  }
}
__global__ void tr_axpy_kernel(float* y, const float* x, float a, long long n) {   // y += a * x
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}
// out[r, :] = a * x[r, :] + off[(r / rows_per_group), :]   (cols <= 4; the metric <-> normalised coordinate maps)
__global__ void tr_affine_rows_kernel(const float* x, const float* off, float a, float* out, long long rows,
                                      int rows_per_group, int n_groups, int cols) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * cols; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    const int g = (int)((r / rows_per_group) % n_groups);
    out[i] = a * x[i] + (off ? off[g * cols + c] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ column sums (bias gradients)
// out[n] += sum_m dy[m, n]     grid (ceil(N/32), row slabs), block (32, 8)
__global__ void tr_colsum_kernel(const float* dy, long long ld, long long M, int N, float* out) {
  __shared__ float part[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < N)
    for (long long m = blockIdx.y * 8 + threadIdx.y; m < M; m += (long long)gridDim.y * 8) acc += dy[m * ld + n];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x];
    atomicAdd(out + n, s);
  }
}
// out[row % group] += sum_c x[row, c]     (bias gradient of an NCHW 1x1 conv: rows = (image, channel), cols = pixels)
__global__ void tr_rowsum_groups_kernel(const float* x, long long rows, int cols, int group, float* out) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += x[row * cols + c];
  s = warp_sum(s);
  if (lane == 0) atomicAdd(out + (row % group), s);
}
// out[r, :] += sum_b x[b, r, :]    (gradient of a table broadcast over the batch: query_feat_embedding)
__global__ void tr_sum_batch_kernel(const float* x, int B, long long n, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += x[b * n + i];
    out[i] += s;
  }
}
__global__ void tr_bcast_batch_kernel(const float* x, int B, long long n, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * B; i += (long long)gridDim.x * blockDim.x)
    out[i] = x[i % n];
}

// ------------------------------------------------------------------------------------------------ LayerNorm (+ residual)
// y = LN(x + res) * gamma + beta ; saves xhat and rstd.  One warp per row, D <= 1024.
template <int kMaxPerLane>
__global__ void tr_layernorm_fwd_kernel(const float* x, const float* res, const float* gamma, const float* beta,
                                        float eps, float* y, float* xhat, float* rstd_out, long long M, int D) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float v[kMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < D) ? x[row * D + c] + (res ? res[row * D + c] : 0.f) : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    const float d = (c < D) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    if (c < D) {
      const float h = (v[i] - mean) * rstd;
      xhat[row * D + c] = h;
      y[row * D + c] = h * gamma[c] + beta[c];
    }
  }
  if (lane == 0) rstd_out[row] = rstd;
}
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma ;  dgamma += sum dy * xhat ; dbeta += sum dy
// one warp per row; each block folds its rows' dgamma/dbeta in shared memory, then one atomic per column.
template <int kMaxPerLane>
__global__ void tr_layernorm_bwd_kernel(const float* dy, const float* xhat, const float* rstd, const float* gamma,
                                        float* dx, float* dgamma, float* dbeta, long long M, int D) {
  extern __shared__ float sh[];   // [2 * D]
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long row = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
    float g[kMaxPerLane], h[kMaxPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        const float d = dy[row * D + c];
        h[i] = xhat[row * D + c];
        g[i] = d * gamma[c];
        atomicAdd(&sh[c], d * h[i]);
        atomicAdd(&sh[D + c], d);
      } else {
        g[i] = 0.f, h[i] = 0.f;
      }
      s1 += g[i];
      s2 += g[i] * h[i];
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    const float r = rstd[row];
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < D) dx[row * D + c] = r * (g[i] - s1 - h[i] * s2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, sh[i]);
    atomicAdd(dbeta + i, sh[D + i]);
  }
}

// ------------------------------------------------------------------------------------------------ row softmax (BERT attention)
// block-wide reductions for the row kernels (256 threads)
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = is_max ? fmaxf(t, red[w]) : t + red[w];
  return t;
}
// P = softmax(S * scale) in place; one block (256 threads) per row of length L.  kV4 > 0: the row lives in registers
// (L == 256 * 4 * kV4: one read, one write); kV4 == 0: generic three-pass version.
template <int kV4>
__global__ void tr_softmax_rows_kernel(float* S, int L, float scale, float* Pd, float p_drop, const unsigned long long* seed,
                                       unsigned long long site) {
  // Pd == nullptr: P (TF32-rounded: a GEMM operand) over S.  Pd != nullptr (attention-probability dropout): S keeps the
  // un-dropped P in full precision (the softmax backward needs it), Pd gets keep ? P / (1 - p) : 0, TF32-rounded.
  __shared__ float red[8];
  float* row = S + (long long)blockIdx.x * L;
  float* drow = Pd ? Pd + (long long)blockIdx.x * L : nullptr;
  const unsigned long long sd = Pd ? seed[0] : 0ull;
  const uint32_t thr = tr_drop_threshold(p_drop);
  const float sc = 1.0f / (1.0f - p_drop);
  const unsigned long long base = (unsigned long long)blockIdx.x * (unsigned long long)L;
  if (kV4 > 0) {
    float4 v[kV4 > 0 ? kV4 : 1];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kV4; ++i) {
      v[i] = reinterpret_cast<const float4*>(row)[threadIdx.x + 256 * i];
      m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
    m = block_reduce(m, true, red);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV4; ++i) {
      v[i].x = __expf((v[i].x - m) * scale), v[i].y = __expf((v[i].y - m) * scale);
      v[i].z = __expf((v[i].z - m) * scale), v[i].w = __expf((v[i].w - m) * scale);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float inv = 1.0f / block_reduce(s, false, red);
#pragma unroll
    for (int i = 0; i < kV4; ++i) {
      const float4 pv = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
      if (drow == nullptr) {
        reinterpret_cast<float4*>(row)[threadIdx.x + 256 * i] = make_float4(tf32r(pv.x), tf32r(pv.y), tf32r(pv.z), tf32r(pv.w));
      } else {
        const unsigned long long e = base + 4ull * (threadIdx.x + 256 * i);
        reinterpret_cast<float4*>(row)[threadIdx.x + 256 * i] = pv;
        reinterpret_cast<float4*>(drow)[threadIdx.x + 256 * i] =
            make_float4(tr_keep(sd, site, e, thr) ? tf32r(pv.x * sc) : 0.f, tr_keep(sd, site, e + 1, thr) ? tf32r(pv.y * sc) : 0.f,
                        tr_keep(sd, site, e + 2, thr) ? tf32r(pv.z * sc) : 0.f, tr_keep(sd, site, e + 3, thr) ? tf32r(pv.w * sc) : 0.f);
      }
    }
    return;
  }
  float m = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) m = fmaxf(m, row[i]);
  m = block_reduce(m, true, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float e = __expf((row[i] - m) * scale);
    row[i] = e;
    s += e;
  }
  const float inv = 1.0f / block_reduce(s, false, red);
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float pv = row[i] * inv;
    if (drow == nullptr) {
      row[i] = tf32r(pv);
    } else {
      row[i] = pv;
      drow[i] = tr_keep(sd, site, base + (unsigned long long)i, thr) ? tf32r(pv * sc) : 0.f;
    }
  }
}
// dS = P * (d - sum(P * d)) * scale, written over dP, with d = dP (no dropout) or d = keep ? dP / (1 - p) : 0 (dP is
// then the gradient w.r.t. the DROPPED probabilities and P the un-dropped ones; mask regenerated from seed / site)
template <int kV4>
__global__ void tr_softmax_rows_bwd_kernel(const float* P, float* dP, int L, float scale, float p_drop,
                                           const unsigned long long* seed, unsigned long long site) {
  __shared__ float red[8];
  const float* prow = P + (long long)blockIdx.x * L;
  float* drow = dP + (long long)blockIdx.x * L;
  const bool drop = seed != nullptr;
  const unsigned long long sd = drop ? seed[0] : 0ull;
  const uint32_t thr = tr_drop_threshold(p_drop);
  const float sc = 1.0f / (1.0f - p_drop);
  const unsigned long long base = (unsigned long long)blockIdx.x * (unsigned long long)L;
  if (kV4 > 0) {
    float4 p[kV4 > 0 ? kV4 : 1], d[kV4 > 0 ? kV4 : 1];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV4; ++i) {
      p[i] = reinterpret_cast<const float4*>(prow)[threadIdx.x + 256 * i];
      d[i] = reinterpret_cast<const float4*>(drow)[threadIdx.x + 256 * i];
      if (drop) {
        const unsigned long long e = base + 4ull * (threadIdx.x + 256 * i);
        d[i].x = tr_keep(sd, site, e, thr) ? d[i].x * sc : 0.f, d[i].y = tr_keep(sd, site, e + 1, thr) ? d[i].y * sc : 0.f;
        d[i].z = tr_keep(sd, site, e + 2, thr) ? d[i].z * sc : 0.f, d[i].w = tr_keep(sd, site, e + 3, thr) ? d[i].w * sc : 0.f;
      }
      s += (p[i].x * d[i].x + p[i].y * d[i].y) + (p[i].z * d[i].z + p[i].w * d[i].w);
    }
    const float dot = block_reduce(s, false, red);
#pragma unroll
    for (int i = 0; i < kV4; ++i)
      reinterpret_cast<float4*>(drow)[threadIdx.x + 256 * i] =
          make_float4(tf32r(p[i].x * (d[i].x - dot) * scale), tf32r(p[i].y * (d[i].y - dot) * scale),
                      tf32r(p[i].z * (d[i].z - dot) * scale), tf32r(p[i].w * (d[i].w - dot) * scale));
    return;
  }
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float dv = drow[i];
    if (drop) dv = tr_keep(sd, site, base + (unsigned long long)i, thr) ? dv * sc : 0.f;
    s += prow[i] * dv;
  }
  const float dot = block_reduce(s, false, red);
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float dv = drow[i];
    if (drop) dv = tr_keep(sd, site, base + (unsigned long long)i, thr) ? dv * sc : 0.f;
    drow[i] = tf32r(prow[i] * (dv - dot) * scale);
  }
}

// ------------------------------------------------------------------------------------------------ vector attention: edges
// Edge e = (query i, neighbour slot j) = i * 32 + j.  gidx[e] = global row of the neighbour in the key/value tables.
// local idx (B, Q, 32) -> global rows b * R + idx ; anchors (32) broadcast to every query (block 0)
__global__ void tr_va_make_idx_kernel(const int* local_idx, const int* anchor_idx, int B, int Q, int R, int* gidx) {
  const long long n = (long long)B * Q * TR_NBR;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / ((long long)Q * TR_NBR));
    const int li = anchor_idx ? anchor_idx[e % TR_NBR] : local_idx[e];
    gidx[e] = b * R + li;
  }
}
// rel[e] = q_xyz[i] - (anchor_xyz[j] | ref_xyz[gidx[e]])
__global__ void tr_va_rel_kernel(const float* q_xyz, const float* ref_xyz, const float* anchor_xyz, const int* gidx,
                                 long long E, float* rel) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / TR_NBR;
    const float* r = anchor_xyz ? anchor_xyz + 3 * (e % TR_NBR) : ref_xyz + 3 * (long long)gidx[e];
#pragma unroll
    for (int k = 0; k < 3; ++k) rel[3 * e + k] = q_xyz[3 * i + k] - r[k];
  }
}
// h[e, :] = relu(W[:, 0:3] . rel[e] + b)      (fc_delta.0: Linear(3, D) + ReLU)
// thread = 4 consecutive channels of one edge; blockDim.x = D / 4 lanes, blockDim.y edges per block
__global__ void tr_lin3_relu_kernel(const float* rel, const float* W, const float* b, float* h, long long E, int D) {
  const int c = threadIdx.x * 4;
  float w[4][3], bb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    w[i][0] = W[3 * (c + i)], w[i][1] = W[3 * (c + i) + 1], w[i][2] = W[3 * (c + i) + 2];
    bb[i] = b[c + i];
  }
  for (long long e = blockIdx.x * (long long)blockDim.y + threadIdx.y; e < E; e += (long long)gridDim.x * blockDim.y) {
    const float r0 = rel[3 * e], r1 = rel[3 * e + 1], r2 = rel[3 * e + 2];
    float4 o;
    o.x = tf32r(fmaxf(w[0][0] * r0 + w[0][1] * r1 + w[0][2] * r2 + bb[0], 0.f));
    o.y = tf32r(fmaxf(w[1][0] * r0 + w[1][1] * r1 + w[1][2] * r2 + bb[1], 0.f));
    o.z = tf32r(fmaxf(w[2][0] * r0 + w[2][1] * r1 + w[2][2] * r2 + bb[2], 0.f));
    o.w = tf32r(fmaxf(w[3][0] * r0 + w[3][1] * r1 + w[3][2] * r2 + bb[3], 0.f));
    *reinterpret_cast<float4*>(h + e * D + c) = o;
  }
}
// backward of the above given dh (already masked by the ReLU), ONE pass over dh:
//   dW[c, k] += sum_e dh[e, c] rel[e, k] ; db[c] += sum_e dh[e, c] ; drel[e, k] = sum_c dh[e, c] W[c, k] (optional)
// warp per edge (lane = channels lane + 32 i), per-lane register accumulators for dW / db, folded per block in shared
// memory and added to global with one atomic per entry.
template <int kPerLane>
__global__ void tr_lin3_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ rel, const float* __restrict__ W,
                                   float* dW, float* db, float* drel, long long E, int D) {
  extern __shared__ float sh[];          // [4 * D]
  for (int i = threadIdx.x; i < 4 * D; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float w0[kPerLane], w1[kPerLane], w2[kPerLane], a0[kPerLane], a1[kPerLane], a2[kPerLane], ab[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const int c = lane + 32 * i;
    w0[i] = c < D ? W[3 * c] : 0.f, w1[i] = c < D ? W[3 * c + 1] : 0.f, w2[i] = c < D ? W[3 * c + 2] : 0.f;
    a0[i] = a1[i] = a2[i] = ab[i] = 0.f;
  }
  for (long long e = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); e < E; e += (long long)gridDim.x * wpb) {
    const float r0 = rel[3 * e], r1 = rel[3 * e + 1], r2 = rel[3 * e + 2];
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      const int c = lane + 32 * i;
      const float g = c < D ? dh[e * D + c] : 0.f;
      a0[i] += g * r0, a1[i] += g * r1, a2[i] += g * r2, ab[i] += g;
      d0 += g * w0[i], d1 += g * w1[i], d2 += g * w2[i];
    }
    if (drel != nullptr) {
      d0 = warp_sum(d0), d1 = warp_sum(d1), d2 = warp_sum(d2);
      if (lane == 0) drel[3 * e] = d0, drel[3 * e + 1] = d1, drel[3 * e + 2] = d2;
    }
  }
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const int c = lane + 32 * i;
    if (c < D) {
      atomicAdd(&sh[c], a0[i]);
      atomicAdd(&sh[D + c], a1[i]);
      atomicAdd(&sh[2 * D + c], a2[i]);
      atomicAdd(&sh[3 * D + c], ab[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(dW + 3 * c, sh[c]);
    atomicAdd(dW + 3 * c + 1, sh[D + c]);
    atomicAdd(dW + 3 * c + 2, sh[2 * D + c]);
    atomicAdd(db + c, sh[3 * D + c]);
  }
}
// t[e, :] = q[i, :] - ktab[gidx[e], :] + pos[e, :]        thread = 4 channels of one edge (blockDim.x = D / 4)
__global__ void tr_va_gather_t_kernel(const float* q, const float* ktab, const int* gidx, const float* pos, float* t,
                                      long long E, int D) {
  const int c = threadIdx.x * 4;
  for (long long e = blockIdx.x * (long long)blockDim.y + threadIdx.y; e < E; e += (long long)gridDim.x * blockDim.y) {
    const float4 qq = *reinterpret_cast<const float4*>(q + (e / TR_NBR) * D + c);
    const float4 kk = *reinterpret_cast<const float4*>(ktab + (long long)gidx[e] * D + c);
    const float4 pp = *reinterpret_cast<const float4*>(pos + e * D + c);
    *reinterpret_cast<float4*>(t + e * D + c) = make_float4(tf32r(qq.x - kk.x + pp.x), tf32r(qq.y - kk.y + pp.y), tf32r(qq.z - kk.z + pp.z), tf32r(qq.w - kk.w + pp.w));
  }
}
// w = softmax_j(a * scale) per (query, channel), written over a ; res[i, c] = sum_j w * (vtab[gidx] + pos)
// every load of the item is issued before its first store (the pointers may alias as far as the compiler knows: a store
// in the middle of the neighbour loop would serialise the 64 remaining loads behind it)
__global__ void tr_va_softmax_agg_kernel(float* a, const float* __restrict__ vtab, const float* __restrict__ pos,
                                         const int* __restrict__ gidx, float scale, float* __restrict__ res, long long NQ,
                                         int D) {
  for (long long i = blockIdx.x; i < NQ; i += gridDim.x)
  for (int c = threadIdx.x; c < D; c += blockDim.x) {          // blockDim.x = min(D, 256)
    float v[TR_NBR], vp[TR_NBR];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < TR_NBR; ++j) {
      const long long e = i * TR_NBR + j;
      v[j] = a[e * D + c] * scale;
      vp[j] = vtab[(long long)gidx[e] * D + c] + pos[e * D + c];
      m = fmaxf(m, v[j]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < TR_NBR; ++j) {
      v[j] = __expf(v[j] - m);
      s += v[j];
    }
    const float inv = 1.0f / s;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < TR_NBR; ++j) {
      const float w = v[j] * inv;
      a[(i * TR_NBR + j) * D + c] = w;
      acc += w * vp[j];
    }
    res[i * D + c] = acc;
  }
}
// given dres: da (over w) = w * (dw - sum_j w dw) * scale with dw = dres * (v + pos) ; dvp = w * dres
__global__ void tr_va_softmax_agg_bwd_kernel(const float* dres, float* w_da, const float* vtab, const float* pos,
                                             const int* gidx, float scale, float* dvp, long long NQ, int D) {
  for (long long i = blockIdx.x; i < NQ; i += gridDim.x)
  for (int c = threadIdx.x; c < D; c += blockDim.x) {          // blockDim.x = min(D, 256)
    const long long x = i * D + c;
    const float g = dres[x];
    float w[TR_NBR], dw[TR_NBR];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < TR_NBR; ++j) {
      const long long e = i * TR_NBR + j;
      w[j] = w_da[e * D + c];
      dw[j] = g * (vtab[(long long)gidx[e] * D + c] + pos[e * D + c]);
      dot += w[j] * dw[j];
    }
#pragma unroll
    for (int j = 0; j < TR_NBR; ++j) {
      const long long e = i * TR_NBR + j;
      w_da[e * D + c] = tf32r(w[j] * (dw[j] - dot) * scale);
      dvp[e * D + c] = w[j] * g;
    }
  }
}
// dq[i] += sum_j dt ; dktab[gidx] -= dt ; dvtab[gidx] += dvp ; dpos = dt + dvp (over dt)
__global__ void tr_va_scatter_kernel(float* dt_dpos, const float* dvp, const int* gidx, float* dq, float* dktab,
                                     float* dvtab, long long NQ, int D) {
  for (long long i = blockIdx.x; i < NQ; i += gridDim.x)
  for (int c = threadIdx.x; c < D; c += blockDim.x) {          // blockDim.x = min(D, 256)
    const long long x = i * D + c;
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < TR_NBR; ++j) {
      const long long e = i * TR_NBR + j;
      const float d = dt_dpos[e * D + c];
      const float p = dvp[e * D + c];
      acc += d;
      const long long r = (long long)gidx[e] * D + c;
      atomicAdd(dktab + r, -d);
      atomicAdd(dvtab + r, p);
      dt_dpos[e * D + c] = tf32r(d + p);
    }
    dq[x] += acc;
  }
}
// dxyz_q[i] += sum_j drel[e] ; dxyz_ref[gidx[e]] -= drel[e] (when the neighbour coordinates are differentiable)
__global__ void tr_va_drel_scatter_kernel(const float* drel, const int* gidx, float* dxyz_q, float* dxyz_ref, long long NQ) {
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < NQ * 3; x += (long long)gridDim.x * blockDim.x) {
    const long long i = x / 3;
    const int k = (int)(x % 3);
    float acc = 0.f;
    for (int j = 0; j < TR_NBR; ++j) {
      const long long e = i * TR_NBR + j;
      const float d = drel[3 * e + k];
      acc += d;
      if (dxyz_ref) atomicAdd(dxyz_ref + 3 * (long long)gidx[e] + k, -d);
    }
    atomicAdd(dxyz_q + x, acc);
  }
}

// ------------------------------------------------------------------------------------------------ reg_branch.2: Linear(D, 3)
// y[m, :] = x[m, :] . W^T + b + base[m, :]     one warp per row
__global__ void tr_lin_n3_kernel(const float* x, const float* W, const float* b, const float* base, float* y, long long M, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = x[row * D + c];
    a0 += v * W[c], a1 += v * W[D + c], a2 += v * W[2 * D + c];
  }
  a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2);
  if (lane == 0) {
    y[3 * row] = a0 + b[0] + (base ? base[3 * row] : 0.f);
    y[3 * row + 1] = a1 + b[1] + (base ? base[3 * row + 1] : 0.f);
    y[3 * row + 2] = a2 + b[2] + (base ? base[3 * row + 2] : 0.f);
  }
}
// dx[m, c] = sum_k dy[m, k] W[k, c] ; dW[k, c] += sum_m dy[m, k] x[m, c] ; db[k] += sum_m dy[m, k]
// x_is_relu: x is a ReLU output and dx is wanted w.r.t. the ReLU's input (dx = 0 where x <= 0)
__global__ void tr_lin_n3_bwd_kernel(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db,
                                     long long M, int D, int x_is_relu) {
  // thread = channel c (blockDim.x >= D handled by loop), block strides over row slabs
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float w0 = W[c], w1 = W[D + c], w2 = W[2 * D + c];
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    for (long long m = blockIdx.x; m < M; m += gridDim.x) {
      const float d0 = dy[3 * m], d1 = dy[3 * m + 1], d2 = dy[3 * m + 2];
      const float v = x[m * D + c];
      dx[m * D + c] = (x_is_relu && !(v > 0.f)) ? 0.f : d0 * w0 + d1 * w1 + d2 * w2;
      g0 += d0 * v, g1 += d1 * v, g2 += d2 * v;
    }
    atomicAdd(dW + c, g0);
    atomicAdd(dW + D + c, g1);
    atomicAdd(dW + 2 * D + c, g2);
  }
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (long long m = blockIdx.x; m < M; m += gridDim.x) s += dy[3 * m + threadIdx.x];
    atomicAdd(db + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------------------------------------ projection + bilinear sampler
// uv[img, p] = pixel coordinates of BPS point p of the image's sample in the image's camera, in grid_sample units
// (ptEmb_head.py:873-883, lib/utils/transform.py:898-930): T = inverse(cam_extr), q = K (R x + t), uv = q.xy / z / inp_res * 2 - 1
__global__ void tr_project_kernel(const float* bps, const float* centre, const float* cam_intr, const float* cam_extr,
                                  const int* img_sample, int NV, int P, float inp_w, float inp_h, float* grid) {
  const long long n = (long long)NV * P;
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n; x += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(x / P);
    const int pidx = (int)(x % P);
    const float* E = cam_extr + 16 * img;
    const float* K = cam_intr + 9 * img;
    // inverse of the affine [A | t] (rows 0..2 of the 4x4): A^-1 by cofactors, t' = -A^-1 t
    const float a00 = E[0], a01 = E[1], a02 = E[2], a10 = E[4], a11 = E[5], a12 = E[6], a20 = E[8], a21 = E[9], a22 = E[10];
    const float c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const float c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
    const float c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
    const float idet = 1.0f / (a00 * c00 + a01 * c10 + a02 * c20);
    const int b = img_sample[img];
    const float X = bps[3 * pidx] + centre[3 * b] - E[3], Y = bps[3 * pidx + 1] + centre[3 * b + 1] - E[7],
                Z = bps[3 * pidx + 2] + centre[3 * b + 2] - E[11];
    const float xc = (c00 * X + c01 * Y + c02 * Z) * idet, yc = (c10 * X + c11 * Y + c12 * Z) * idet,
                zc = (c20 * X + c21 * Y + c22 * Z) * idet;
    const float qx = K[0] * xc + K[1] * yc + K[2] * zc, qy = K[3] * xc + K[4] * yc + K[5] * zc;
    float qz = K[6] * xc + K[7] * yc + K[8] * zc;
    if (fabsf(qz) < 1e-7f) qz = 1e-7f;
    grid[2 * x] = qx / qz / inp_w * 2.f - 1.f;
    grid[2 * x + 1] = qy / qz / inp_h * 2.f - 1.f;
  }
}
__device__ __forceinline__ void bilinear_taps(float gx, float gy, int W, int H, int (&off)[4], float (&wt)[4]) {
  const float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;   // align_corners = False
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float tx = ix - fx, ty = iy - fy;
  const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
  const float ws[4] = {(1.f - tx) * (1.f - ty), tx * (1.f - ty), (1.f - tx) * ty, tx * ty};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool in = xs[k] >= 0 && xs[k] < W && ys[k] >= 0 && ys[k] < H;   // padding_mode = zeros
    off[k] = in ? ys[k] * W + xs[k] : -1;
    wt[k] = in ? ws[k] : 0.f;
  }
}
// S[img, d, p] = bilinear(planes[img, d, :, :], grid[img, p])      planes NCHW (H x W = hw x hw)
__global__ void tr_sample_fwd_kernel(const float* planes, const float* grid, float* S, int NV, int D, int P, int hw) {
  const long long n = (long long)NV * P;
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n; x += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(x / P);
    const int pidx = (int)(x % P);
    int off[4];
    float wt[4];
    bilinear_taps(grid[2 * x], grid[2 * x + 1], hw, hw, off, wt);
    const float* pl = planes + (long long)img * D * hw * hw;
    float* out = S + (long long)img * D * P + pidx;
    for (int d = 0; d < D; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (off[k] >= 0) acc += wt[k] * pl[(long long)d * hw * hw + off[k]];
      out[(long long)d * P] = acc;
    }
  }
}
// block = (image, chunk of kSbCh channels): the chunk's planes are accumulated in shared memory (the 4 x P x kSbCh
// scatter-adds of an image collide on 256 pixels: global atomics were 3.6 ms of a 95 ms step), then added to dplanes.
constexpr int kSbCh = 64;
__global__ void tr_sample_bwd_kernel(const float* __restrict__ dS, const float* __restrict__ grid, float* dplanes, int NV, int D,
                                     int P, int hw) {
  extern __shared__ float acc[];          // [kSbCh][hw * hw]
  const int HW = hw * hw;
  const int img = blockIdx.x, d0 = blockIdx.y * kSbCh;
  const int nch = min(kSbCh, D - d0);
  for (int i = threadIdx.x; i < kSbCh * HW; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  for (int pidx = threadIdx.x; pidx < P; pidx += blockDim.x) {
    int off[4];
    float wt[4];
    const long long x = (long long)img * P + pidx;
    bilinear_taps(grid[2 * x], grid[2 * x + 1], hw, hw, off, wt);
    const float* in = dS + ((long long)img * D + d0) * P + pidx;
    for (int d = 0; d < nch; ++d) {
      const float g = in[(long long)d * P];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (off[k] >= 0) atomicAdd(&acc[d * HW + off[k]], wt[k] * g);
    }
  }
  __syncthreads();
  float* pl = dplanes + ((long long)img * D + d0) * HW;
  for (int i = threadIdx.x; i < nch * HW; i += blockDim.x) pl[i] += acc[i];
}

// ------------------------------------------------------------------------------------------------ cross-view merge
// X rows of sample b: row0[b] + p * n[b] + v   (the raw `.view(1,-1,N,D)` regroup, ptEmb_head.py:745-771,910-926)
// m = mlp0(X) [rows, Dm].  n > 1: agg[b*P+p] = sum_{v>=1} m_v * (m_v . m_0) ;  n == 1: agg = m_0
__global__ void tr_merge_agg_kernel(const float* m, const int* row0, const int* nviews, int P, int Dm, float* agg,
                                    long long total) {
  // one warp per (b, p)
  const int lane = threadIdx.x & 31;
  const long long gp = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);   // b * P + p
  if (gp >= total) return;
  const int b = (int)(gp / P);
  const int p = (int)(gp % P);
  const int n = nviews[b];
  const float* base = m + ((long long)row0[b] + (long long)p * n) * Dm;
  constexpr int kMax = 8;   // Dm <= 256 (D <= 512)
  float mast[kMax], acc[kMax];
  _Pragma("unroll") for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    mast[i] = (c < Dm) ? base[c] : 0.f;
    acc[i] = (n == 1) ? mast[i] : 0.f;
  }
  for (int v = 1; v < n; ++v) {
    float o[kMax];
    float dot = 0.f;
    _Pragma("unroll") for (int i = 0; i < kMax; ++i) {
      const int c = lane + 32 * i;
      o[i] = (c < Dm) ? base[(long long)v * Dm + c] : 0.f;
      dot += o[i] * mast[i];
    }
    dot = warp_sum(dot);
    _Pragma("unroll") for (int i = 0; i < kMax; ++i) acc[i] += o[i] * dot;
  }
  _Pragma("unroll") for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < Dm) agg[gp * Dm + c] = acc[i];
  }
}
// dm from dagg:  n > 1: dm_v = w_v dagg + (dagg . m_v) m_0 (v >= 1), dm_0 = sum_v (dagg . m_v) m_v ; n == 1: dm_0 = dagg
__global__ void tr_merge_agg_bwd_kernel(const float* dagg, const float* m, const int* row0, const int* nviews, int P,
                                        int Dm, float* dm, long long total) {
  const int lane = threadIdx.x & 31;
  const long long gp = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gp >= total) return;
  const int b = (int)(gp / P);
  const int p = (int)(gp % P);
  const int n = nviews[b];
  const long long r0 = ((long long)row0[b] + (long long)p * n) * Dm;
  constexpr int kMax = 8;
  float mast[kMax], g[kMax], dmast[kMax];
  _Pragma("unroll") for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    mast[i] = (c < Dm) ? m[r0 + c] : 0.f;
    g[i] = (c < Dm) ? dagg[gp * Dm + c] : 0.f;
    dmast[i] = (n == 1) ? g[i] : 0.f;
  }
  for (int v = 1; v < n; ++v) {
    float o[kMax];
    float w = 0.f, go = 0.f;
    _Pragma("unroll") for (int i = 0; i < kMax; ++i) {
      const int c = lane + 32 * i;
      o[i] = (c < Dm) ? m[r0 + (long long)v * Dm + c] : 0.f;
      w += o[i] * mast[i];
      go += o[i] * g[i];
    }
    w = warp_sum(w), go = warp_sum(go);
    _Pragma("unroll") for (int i = 0; i < kMax; ++i) {
      const int c = lane + 32 * i;
      if (c < Dm) dm[r0 + (long long)v * Dm + c] = w * g[i] + go * mast[i];
      dmast[i] += go * o[i];
    }
  }
  _Pragma("unroll") for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < Dm) dm[r0 + c] = dmast[i];
  }
}
// out[b*P+p, :] = X[row0[b] + p * n, :] + y[b*P+p, :] / n          (n == 1: divisor 1 as well: q + mlp1(mlp0(q)))
__global__ void tr_merge_out_kernel(const float* X, const float* y, const int* row0, const int* nviews, int P, int D,
                                    float* out, long long total) {
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
    const long long gp = x / D;
    const int c = (int)(x % D);
    const int b = (int)(gp / P);
    const int p = (int)(gp % P);
    const int n = nviews[b];
    out[x] = X[((long long)row0[b] + (long long)p * n) * D + c] + y[x] / n;
  }
}
// dX[row0[b] + p * n, :] += dout ; dy = dout / n
__global__ void tr_merge_out_bwd_kernel(const float* dout, const int* row0, const int* nviews, int P, int D, float* dX,
                                        float* dy, long long total) {
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
    const long long gp = x / D;
    const int c = (int)(x % D);
    const int b = (int)(gp / P);
    const int p = (int)(gp % P);
    const int n = nviews[b];
    const float g = dout[x];
    dX[((long long)row0[b] + (long long)p * n) * D + c] += g;
    dy[x] = g / n;
  }
}

// ------------------------------------------------------------------------------------------------ fused gradient norm + clip
// sumsq[0] += sum g^2 over one tensor ; then every tensor is scaled by min(1, max_norm / (sqrt(sumsq) + 1e-6))
__global__ void tr_sumsq_kernel(const float* g, long long n, float* sumsq) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(sumsq, t);
  }
}
__global__ void tr_clip_scale_kernel(float* g, long long n, const float* sumsq, float max_norm) {
  const float coef = fminf(1.0f, max_norm / (sqrtf(sumsq[0]) + 1e-6f));
  if (coef >= 1.0f) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) g[i] *= coef;
}

// ------------------------------------------------------------------------------------------------ flat-buffer optimiser step
// Parameters / gradients live in ONE flat buffer (segment s = elements [off[s], off[s] + len[s])): one launch each for
// the per-tensor norms, the per-tensor clip (lib/utils/net_utils.py:122-132) and Adam (torch.optim.Adam semantics,
// lib/utils/net_utils.py:57-63: L2 weight decay added to the gradient, bias-corrected moments).
__global__ void tr_seg_sumsq_kernel(const float* g, const long long* off, const long long* len, float* sumsq) {
  __shared__ float red[8];
  const float* p = g + off[blockIdx.x];
  const long long n = len[blockIdx.x];
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += p[i] * p[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    sumsq[blockIdx.x] = t;
  }
}
__global__ void tr_seg_clip_kernel(float* g, const long long* off, const long long* len, const float* sumsq, float max_norm) {
  const float coef = fminf(1.0f, max_norm / (sqrtf(sumsq[blockIdx.x]) + 1e-6f));
  if (coef >= 1.0f) return;
  float* p = g + off[blockIdx.x];
  const long long n = len[blockIdx.x];
  for (long long i = threadIdx.x; i < n; i += blockDim.x) p[i] *= coef;
}
__global__ void tr_adam_kernel(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                               float eps, float wd, float bc1, float bc2) {
  const float step = lr / bc1, isq = rsqrtf(bc2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] + wd * p[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi, v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) * isq + eps);
  }
}
// 3-D terms of compute_loss on the last block (lib/models/POEM.py:398-412, release loss types): MSE on the 21 joints, L1
// on the 778 vertices; writes d loss / d all_coords_preds (zero for the earlier blocks) and adds the loss to loss[0].
__global__ void tr_coord_loss_kernel(const float* coords, const float* gt_joints, const float* gt_verts, int NB, int B,
                                     int NJ, int NVt, float wj, float wv, float* loss, float* dcoords) {
  __shared__ float red[8];
  const int Q = NJ + NVt;
  const long long per_block = (long long)B * Q * 3, total = per_block * NB;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float d = 0.f;
    if (i >= per_block * (NB - 1)) {
      const long long r = i - per_block * (NB - 1);
      const int b = (int)(r / (Q * 3)), q = (int)((r / 3) % Q), k = (int)(r % 3);
      if (q < NJ) {
        const float e = coords[i] - gt_joints[((long long)b * NJ + q) * 3 + k];
        acc += wj * e * e / (float)(B * NJ * 3);
        d = wj * 2.f * e / (float)(B * NJ * 3);
      } else {
        const float e = coords[i] - gt_verts[((long long)b * NVt + (q - NJ)) * 3 + k];
        acc += wv * fabsf(e) / (float)(B * NVt * 3);
        d = wv * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) / (float)(B * NVt * 3);
      }
    }
    dcoords[i] = d;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(loss, t);
  }
}

// ------------------------------------------------------------------------------------------------ compute_loss (head terms)
// `PtEmbedMultiviewStereoV2.compute_loss` (lib/models/POEM.py:363-466, release loss types) on the LAST block's prediction:
//   [0] loss_3d_joints_from_mesh  MSE(openpose(J_regressor . verts_pred), openpose(J_regressor . verts_gt))
//   [1] loss_3d_joints            MSE(joints_pred, joints_gt)
//   [2] loss_3d_verts             L1(verts_pred, verts_gt)          (the parametric variant centres both on the same GT joint: identical)
//   [3] loss_2d_joints            mean over (view, joint) of |clamp(proj(joints_pred) - target_2d, +-s/2) / s|^2
//   [4] loss_2d_verts             the same for the vertices, target = proj(verts_gt)
//   [5] loss_pose, [6] loss_shape MSE on the MANO parameters (parametric heads)
//   [7] loss_recon = w_j ([1] + [0]) + w_v [2] + w_j2d [3] + w_v2d [4] + w_pose [5] + w_shape [6]
// (the heat-map term of the reference's total belongs to the image half).  Also writes d loss_recon / d coords (last block).
struct TrLossArgs {
  const float *coords, *gt_joints, *gt_verts, *j_regressor;      // [B, 799, 3], [B, 21, 3], [B, 778, 3], [16, 778]
  const float *cam_intr, *cam_extr, *target_2d;                  // [NV, 9], [NV, 16], [NV, 21, 2]
  const int* img_sample;                                         // [NV]
  int B, NV;
  float img_scale, w_j, w_v, w_j2d, w_v2d;
  float *losses, *dcoords;                                       // [8], [B, 799, 3] (zeroed by the caller)
};
constexpr int kLossJ = 21, kLossV = 778, kLossQ = kLossJ + kLossV;

// block per sample: the three 3-D terms
__global__ void tr_loss_3d_kernel(const TrLossArgs a) {
  __shared__ float dJ[16][3];
  __shared__ float red[3][8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int order[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};
  const int tips[5] = {744, 320, 443, 555, 672};        // lib/utils/misc.py:76-82 (mano_to_openpose)
  const float* pj = a.coords + (size_t)b * kLossQ * 3;
  const float* pv = pj + kLossJ * 3;
  const float* jg = a.gt_joints + (size_t)b * kLossJ * 3;
  const float* vg = a.gt_verts + (size_t)b * kLossV * 3;
  float* dj = a.dcoords + (size_t)b * kLossQ * 3;
  float* dv = dj + kLossJ * 3;
  const float nj = 1.0f / (float)(a.B * kLossJ * 3), nv = 1.0f / (float)(a.B * kLossV * 3);
  float l_jm = 0.f, l_j = 0.f, l_v = 0.f;
  // regressed joints of predicted and GT mesh (16 x 3), one warp per (joint, component)
  for (int o = warp; o < 48; o += (int)(blockDim.x >> 5)) {
    const int j = o / 3, c = o % 3;
    float acc = 0.f;
    for (int v = lane; v < kLossV; v += 32) acc += a.j_regressor[j * kLossV + v] * (pv[v * 3 + c] - vg[v * 3 + c]);
    acc = warp_sum(acc);
    if (lane == 0) {
      dJ[j][c] = a.w_j * 2.f * acc * nj;
      l_jm += acc * acc * nj;
    }
  }
  __syncthreads();
  for (int i = tid; i < kLossJ * 3; i += blockDim.x) {
    const float e = pj[i] - jg[i];
    l_j += e * e * nj;
    atomicAdd(dj + i, a.w_j * 2.f * e * nj);
  }
  for (int i = tid; i < kLossV * 3; i += blockDim.x) {
    const int v = i / 3, c = i % 3;
    const float e = pv[i] - vg[i];
    l_v += fabsf(e) * nv;
    float g = a.w_v * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * nv;
#pragma unroll
    for (int j = 0; j < 16; ++j) g += a.j_regressor[j * kLossV + v] * dJ[j][c];
#pragma unroll
    for (int t = 0; t < 5; ++t)
      if (v == tips[t]) {
        g += a.w_j * 2.f * e * nj;
        l_jm += e * e * nj;
      }
    atomicAdd(dv + i, g);
  }
  (void)order;                                            // the 21-joint re-ordering does not change a mean over all joints
  l_jm = warp_sum(l_jm), l_j = warp_sum(l_j), l_v = warp_sum(l_v);
  if (lane == 0) red[0][warp] = l_jm, red[1][warp] = l_j, red[2][warp] = l_v;
  __syncthreads();
  if (tid < 3) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[tid][w];
    atomicAdd(a.losses + tid, t);
    atomicAdd(a.losses + 7, (tid == 2 ? a.w_v : a.w_j) * t);
  }
}

// thread per (image, point): the two 2-D terms (loss_proj_to_multicam, POEM.py:335-361; camera = inverse of cam_extr)
__global__ void tr_loss_2d_kernel(const TrLossArgs a) {
  __shared__ float red[2][8];
  const long long n = (long long)a.NV * kLossQ;
  float l_j = 0.f, l_v = 0.f;
  const float s = a.img_scale, lim = 0.5f * a.img_scale;
  const float cj = 1.0f / (float)(a.NV * kLossJ), cv = 1.0f / (float)(a.NV * kLossV);
  for (long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x; x < n; x += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(x / kLossQ), q = (int)(x % kLossQ);
    const bool is_j = q < kLossJ;
    const float w = is_j ? a.w_j2d : a.w_v2d;
    if (w == 0.f) continue;
    const int b = a.img_sample[img];
    const float* E = a.cam_extr + 16 * img;
    const float* K = a.cam_intr + 9 * img;
    const float a00 = E[0], a01 = E[1], a02 = E[2], a10 = E[4], a11 = E[5], a12 = E[6], a20 = E[8], a21 = E[9], a22 = E[10];
    const float c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const float c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
    const float c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
    const float idet = 1.0f / (a00 * c00 + a01 * c10 + a02 * c20);
    const float Rm[9] = {c00 * idet, c01 * idet, c02 * idet, c10 * idet, c11 * idet, c12 * idet, c20 * idet, c21 * idet, c22 * idet};
    // M = K . R (3 x 3): q = M (p - t_e)
    float M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) M[r * 3 + c] = K[r * 3] * Rm[c] + K[r * 3 + 1] * Rm[3 + c] + K[r * 3 + 2] * Rm[6 + c];
    auto project = [&](const float* p, float& u, float& v, float& z) {
      const float X = p[0] - E[3], Y = p[1] - E[7], Z = p[2] - E[11];
      const float qx = M[0] * X + M[1] * Y + M[2] * Z, qy = M[3] * X + M[4] * Y + M[5] * Z;
      z = M[6] * X + M[7] * Y + M[8] * Z;
      if (fabsf(z) < 1e-7f) z = 1e-7f;
      u = qx / z, v = qy / z;
    };
    const float* p = a.coords + ((size_t)b * kLossQ + q) * 3;
    float u, v, z, tu, tv;
    project(p, u, v, z);
    if (is_j) {
      tu = a.target_2d[((size_t)img * kLossJ + q) * 2], tv = a.target_2d[((size_t)img * kLossJ + q) * 2 + 1];
    } else {
      float tz;
      project(a.gt_verts + ((size_t)b * kLossV + (q - kLossJ)) * 3, tu, tv, tz);
    }
    const float du = u - tu, dv = v - tv;
    const bool cu = du < -lim || du > lim, cvv = dv < -lim || dv > lim;
    const float ou = fminf(fmaxf(du, -lim), lim) / s, ov = fminf(fmaxf(dv, -lim), lim) / s;
    const float norm = is_j ? cj : cv;
    (is_j ? l_j : l_v) += (ou * ou + ov * ov) * norm;
    // d / d(u, v), then through u = qx / z, v = qy / z and q = M (p - t)
    const float gu = cu ? 0.f : w * 2.f * ou / s * norm, gv = cvv ? 0.f : w * 2.f * ov / s * norm;
    const float gqx = gu / z, gqy = gv / z, gqz = -(gu * u + gv * v) / z;
    float* d = a.dcoords + ((size_t)b * kLossQ + q) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(d + c, M[c] * gqx + M[3 + c] * gqy + M[6 + c] * gqz);
  }
  l_j = warp_sum(l_j), l_v = warp_sum(l_v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[0][warp] = l_j, red[1][warp] = l_v;
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
    atomicAdd(a.losses + 3 + threadIdx.x, t);
    atomicAdd(a.losses + 7, (threadIdx.x == 0 ? a.w_j2d : a.w_v2d) * t);
  }
}
// MSE of a parameter vector (pose / shape): loss[slot] += mean (x - y)^2 ; loss[7] += w * that ; dx = w * 2 (x - y) / n
__global__ void tr_loss_mse_kernel(const float* x, const float* y, int n, float w, float* losses, int slot, float* dx) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float e = x[i] - y[i];
    acc += e * e / (float)n;
    if (dx) dx[i] = w * 2.f * e / (float)n;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    atomicAdd(losses + slot, t);
    atomicAdd(losses + 7, w * t);
  }
}

}  // namespace poem
