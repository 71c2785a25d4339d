// HBM-bound / latency-bound SIMT kernels of the decoder path: layout conversion, camera projection +
// bilinear sampling, cross-view reduce, LayerNorm, 32-NN search, vector-attention glue, coordinate regression.
#pragma once
#include "common.cuh"

namespace poem {

// ------------------------------------------------------------------------------------------------
// (BV,C,256) f32 NCHW feature maps -> (BV*256, C) op16 rows (K-major A operand of the input projection)
// reference: input of nn.Conv2d(k=1) at lib/models/heads/ptEmb_head.py:835
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_rows_op16_kernel(const float* __restrict__ feat, op16* __restrict__ rows, int C,
                                         int HW) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < C && p < HW) ? feat[((size_t)img * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (c < C && p < HW) rows[((size_t)img * HW + p) * C + c] = f2op16(tile[tx][i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Camera preparation: P = K · inv(cam_extr)[:3,:]  (3x4), one thread per image.
// reference: torch.linalg.inv + batch_cam_extr_transf + batch_cam_intr_projection
// (lib/utils/collation.py:60-61, lib/utils/transform.py:898-930)
// ------------------------------------------------------------------------------------------------
__global__ void camera_prep_kernel(const float* __restrict__ intr, const float* __restrict__ extr,
                                   float* __restrict__ proj, int n_img) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img) return;
  float a[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      a[r][c] = extr[i * 16 + r * 4 + c];
      a[r][4 + c] = (r == c) ? 1.f : 0.f;
    }
  // Gauss-Jordan with partial pivoting
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    float best = fabsf(a[col][col]);
    for (int r = col + 1; r < 4; ++r)
      if (fabsf(a[r][col]) > best) {
        best = fabsf(a[r][col]);
        piv = r;
      }
    if (piv != col)
      for (int c = 0; c < 8; ++c) {
        const float t = a[col][c];
        a[col][c] = a[piv][c];
        a[piv][c] = t;
      }
    const float inv = 1.0f / a[col][col];
    for (int c = 0; c < 8; ++c) a[col][c] *= inv;
    for (int r = 0; r < 4; ++r)
      if (r != col) {
        const float f = a[r][col];
        for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
      }
  }
  // T = inverse (master -> camera); keep rows 0..2, then store [R|t] and K separately (two-step like the reference)
  float* o = proj + i * 24;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) o[r * 4 + c] = a[r][4 + c];
  for (int k = 0; k < 9; ++k) o[12 + k] = intr[i * 9 + k];
}

// ------------------------------------------------------------------------------------------------
// Projection + bilinear sampling, emitted directly in the reference's reinterpreted order.
//   xmap : (NV, D, 256) f32 channel-planar feature volume (input_proj + positional term)
//   X    : merge-MLP input rows. For sample b with N views and row base R_b, row R_b + r (r < N*P) holds
//          element (n, d, p) of the sampled tensor S[n,d,p] with n = r / P, d = (r / (P/D)) % D,
//          p = (r % (P/D)) * D + d'   -- the raw `.view(1,-1,N,D)` of ptEmb_head.py:914-915.
// grid = (D / CH, NV); block = 512 threads, each owning 8 of the 4096 points.
// reference: generate_grid_sample_proj (collation.py:48-65), normalisation ptEmb_head.py:880-883,
//            F.grid_sample(bilinear, zeros, align_corners=False) ptEmb_head.py:900-901
// ------------------------------------------------------------------------------------------------
constexpr int SAMPLE_CH = 32;
constexpr int SAMPLE_PITCH = SAMPLE_CH + 4;   // floats per pixel in smem
constexpr int SAMPLE_THREADS = 512;

__global__ void __launch_bounds__(SAMPLE_THREADS)
project_sample_kernel(const float* __restrict__ xmap, const float* __restrict__ proj,
                      const float* __restrict__ bps, const float* __restrict__ centre,
                      const int* __restrict__ img_sample, const int* __restrict__ img_view,
                      const int* __restrict__ sample_rowbase, op16* __restrict__ X, int D, int P, int FH,
                      int FW, float inv_w, float inv_h) {
  // Pixel-major copy of this block's 32 channels: pix[pixel][SAMPLE_PITCH], 36 floats per pixel (32 channels + pad).
  // A tap is then read as 8 x LDS.128 (4 channels each) instead of 32 scalar loads, and with the 144-byte pitch the
  // 16-byte bank group of a lane is (pixel + channel group) % 8, so random pixels spread over all banks (the
  // channel-planar layout had every lane of a warp gathering from the same 1 KB plane: 66 M bank conflicts per call).
  extern __shared__ __align__(16) float pix[];
  const int img = blockIdx.y;
  const int d0 = blockIdx.x * SAMPLE_CH;
  const int F = FH * FW;
  const int b = img_sample[img];
  const int n = img_view[img];
  for (int i = threadIdx.x; i < SAMPLE_CH * F; i += SAMPLE_THREADS) {
    const int ch = i / F, px = i - ch * F;
    pix[px * SAMPLE_PITCH + ch] = xmap[((size_t)img * D + d0) * F + i];
  }
  const float* pm = proj + img * 24;
  const float cx = centre[b * 3 + 0], cy = centre[b * 3 + 1], cz = centre[b * 3 + 2];
  constexpr int PTS = 8;
  float w00[PTS], w01[PTS], w10[PTS], w11[PTS];
  int o00[PTS], o01[PTS], o10[PTS], o11[PTS];
#pragma unroll
  for (int j = 0; j < PTS; ++j) {
    const int p = threadIdx.x * PTS + j;   // 8 consecutive points per thread -> one 16-byte store per channel
    // world point (bps + centre), then master->camera, then intrinsics (same two-step order as the reference)
    const float wx = bps[p * 3 + 0] + cx, wy = bps[p * 3 + 1] + cy, wz = bps[p * 3 + 2] + cz;
    const float X0 = pm[0] * wx + pm[1] * wy + pm[2] * wz + pm[3];
    const float Y0 = pm[4] * wx + pm[5] * wy + pm[6] * wz + pm[7];
    const float Z0 = pm[8] * wx + pm[9] * wy + pm[10] * wz + pm[11];
    const float qx = pm[12] * X0 + pm[13] * Y0 + pm[14] * Z0;
    const float qy = pm[15] * X0 + pm[16] * Y0 + pm[17] * Z0;
    float qz = pm[18] * X0 + pm[19] * Y0 + pm[20] * Z0;
    if (fabsf(qz) < 1e-7f) qz = 1e-7f;
    const float gx = (qx / qz) * inv_w * 2.f - 1.f;
    const float gy = (qy / qz) * inv_h * 2.f - 1.f;
    // align_corners=False unnormalisation
    const float ix = ((gx + 1.f) * FW - 1.f) * 0.5f;
    const float iy = ((gy + 1.f) * FH - 1.f) * 0.5f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float ax = ix - fx0, ay = iy - fy0;
    // clamp before the int conversion so far-away projections cannot overflow
    const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)FW + 1.f);
    const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)FH + 1.f);
    const bool in_range = (fx0 >= -2.f) && (fx0 <= (float)FW + 1.f) && (fy0 >= -2.f) && (fy0 <= (float)FH + 1.f);
    const bool vx0 = in_range && x0 >= 0 && x0 < FW, vx1 = in_range && x0 + 1 >= 0 && x0 + 1 < FW;
    const bool vy0 = in_range && y0 >= 0 && y0 < FH, vy1 = in_range && y0 + 1 >= 0 && y0 + 1 < FH;
    w00[j] = (vx0 && vy0) ? (1.f - ax) * (1.f - ay) : 0.f;
    w01[j] = (vx1 && vy0) ? ax * (1.f - ay) : 0.f;
    w10[j] = (vx0 && vy1) ? (1.f - ax) * ay : 0.f;
    w11[j] = (vx1 && vy1) ? ax * ay : 0.f;
    o00[j] = (vx0 && vy0) ? y0 * FW + x0 : 0;
    o01[j] = (vx1 && vy0) ? y0 * FW + x0 + 1 : 0;
    o10[j] = (vx0 && vy1) ? (y0 + 1) * FW + x0 : 0;
    o11[j] = (vx1 && vy1) ? (y0 + 1) * FW + x0 + 1 : 0;
  }
  __syncthreads();
  const int chunks = P / D;  // rows per (view, channel)
  const size_t row0 = (size_t)sample_rowbase[b] + (size_t)n * P;
  const int p0 = threadIdx.x * PTS;
#pragma unroll 1
  for (int cg = 0; cg < SAMPLE_CH / 4; ++cg) {
    float v[4][PTS];
#pragma unroll
    for (int j = 0; j < PTS; ++j) {
      // ATen accumulates the four taps in the order nw, ne, sw, se
      const float4 a = *reinterpret_cast<const float4*>(pix + o00[j] * SAMPLE_PITCH + 4 * cg);
      const float4 bq = *reinterpret_cast<const float4*>(pix + o01[j] * SAMPLE_PITCH + 4 * cg);
      const float4 c = *reinterpret_cast<const float4*>(pix + o10[j] * SAMPLE_PITCH + 4 * cg);
      const float4 d = *reinterpret_cast<const float4*>(pix + o11[j] * SAMPLE_PITCH + 4 * cg);
      float t;
      t = a.x * w00[j], t += bq.x * w01[j], t += c.x * w10[j], t += d.x * w11[j], v[0][j] = t;
      t = a.y * w00[j], t += bq.y * w01[j], t += c.y * w10[j], t += d.y * w11[j], v[1][j] = t;
      t = a.z * w00[j], t += bq.z * w01[j], t += c.z * w10[j], t += d.z * w11[j], v[2][j] = t;
      t = a.w * w00[j], t += bq.w * w01[j], t += c.w * w10[j], t += d.w * w11[j], v[3][j] = t;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const size_t rbase = row0 + (size_t)(d0 + 4 * cg + q) * chunks;
      uint4 pk;
      pk.x = pack_op16x2(v[q][0], v[q][1]);
      pk.y = pack_op16x2(v[q][2], v[q][3]);
      pk.z = pack_op16x2(v[q][4], v[q][5]);
      pk.w = pack_op16x2(v[q][6], v[q][7]);
      *reinterpret_cast<uint4*>(X + (rbase + p0 / D) * D + (p0 % D)) = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cross-view reduce of the merge network (ptEmb_head.py:755-759):
//   m (rows of D/2) for token p', views 0..N-1 at rows R_b + p'*N + n.
//   N == 1 : s = m_0                                   (merge_features_sv feeds MLP1 directly)
//   N  > 1 : w_n = <m_n, m_0>, s = sum_{n>=1} w_n m_n
// one warp per token.
// ------------------------------------------------------------------------------------------------
template <int PER>   // op16 values per lane: H = 32 * PER, lane owns columns [lane*PER, lane*PER + PER)
__global__ void merge_reduce_kernel(const op16* __restrict__ m, const int* __restrict__ sample_rowbase,
                                    const int* __restrict__ sample_views, op16* __restrict__ s,
                                    float* __restrict__ sigma, int P, int n_tokens) {
  constexpr int H = 32 * PER;
  constexpr int WORDS = PER / 2;   // 32-bit words (op16 pairs) per lane
  const int tok = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= n_tokens) return;
  const int b = tok / P, pp = tok - b * P;
  const int N = sample_views[b];
  const uint32_t* base = reinterpret_cast<const uint32_t*>(m + ((size_t)sample_rowbase[b] + (size_t)pp * N) * H) + lane * WORDS;
  auto load_row = [&](int nn, float (&v)[PER]) {
    uint32_t w[WORDS];
    const uint32_t* r = base + (size_t)nn * (H / 2);
    if constexpr (WORDS == 1) {
      w[0] = __ldg(r);
    } else if constexpr (WORDS == 2) {
      const uint2 t = __ldg(reinterpret_cast<const uint2*>(r));
      w[0] = t.x, w[1] = t.y;
    } else {
#pragma unroll
      for (int q = 0; q < WORDS / 4; ++q) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(r) + q);
        w[4 * q] = t.x, w[4 * q + 1] = t.y, w[4 * q + 2] = t.z, w[4 * q + 3] = t.w;
      }
    }
#pragma unroll
    for (int i = 0; i < WORDS; ++i) {
      v[2 * i] = op16_lo(w[i]);
      v[2 * i + 1] = op16_hi(w[i]);
    }
  };
  float m0[PER], acc[PER];
  load_row(0, m0);
#pragma unroll
  for (int i = 0; i < PER; ++i) acc[i] = (N == 1) ? m0[i] : 0.f;
  for (int nn = 1; nn < N; ++nn) {
    float mv[PER];
    load_row(nn, mv);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) dot += mv[i] * m0[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
    for (int i = 0; i < PER; ++i) acc[i] += dot * mv[i];
  }
  // s is cubic in the activations (<m_n, m_0> m_n): the row is stored as s / sigma with sigma = 2^floor(log2 max|s|)
  // (exact), so the fp16 operand of MLP1 can neither overflow nor lose its small rows; the GEMM epilogues undo it.
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) mx = fmaxf(mx, fabsf(acc[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const uint32_t ebits = __float_as_uint(mx) & 0x7f800000u;
  const float sig = (ebits == 0u || ebits == 0x7f800000u) ? 1.0f : __uint_as_float(ebits);
  const float inv = 1.0f / sig;
  if (lane == 0) sigma[tok] = sig;
  uint32_t* out = reinterpret_cast<uint32_t*>(s + (size_t)tok * H) + lane * WORDS;
#pragma unroll
  for (int i = 0; i < WORDS; ++i) out[i] = pack_op16x2(acc[2 * i] * inv, acc[2 * i + 1] * inv);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (eps 1e-12, biased variance) — HF BertSelfOutput / BertOutput.
// one warp per row; writes fp32 and op16 copies.
// ------------------------------------------------------------------------------------------------
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ y_f32,
                                 op16* __restrict__ y_op16, int rows, int D, float eps) {
  pdl_wait();
  pdl_trigger();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * D;
  float v[32];  // D <= 1024
  const int per = D / 32;
  float sum = 0.f;
  for (int i = 0; i < per; ++i) {
    v[i] = xr[lane + 32 * i];
    sum += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)D;
  float var = 0.f;
  for (int i = 0; i < per; ++i) {
    const float d = v[i] - mean;
    var += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / (float)D + eps);
  for (int i = 0; i < per; ++i) {
    const int c = lane + 32 * i;
    const float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
    if (y_f32) y_f32[(size_t)row * D + c] = o;
    if (y_op16) y_op16[(size_t)row * D + c] = f2op16(o);
  }
}

// ------------------------------------------------------------------------------------------------
// 32 nearest neighbours (squared L2, ascending, lower index wins ties) — pytorch3d `knn_points(K=32)`
// as called at lib/models/bricks/point_transformers.py:83,134.  One warp per query; the warp keeps the
// current best 32 sorted across its lanes and inserts candidates with ballot/shuffle.
// Distances are (dx*dx + dy*dy) + dz*dz with separate fp32 roundings (no FMA contraction) so that the
// neighbour sets are bit-identical to the fp32 oracle given identical coordinates.
// ------------------------------------------------------------------------------------------------
__global__ void knn32_kernel(const float* __restrict__ query, const float* __restrict__ ref, int* __restrict__ idx_out,
                             int Lq, int Lr, int n_query_total) {
  pdl_wait();
  pdl_trigger();
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= n_query_total) return;
  const int b = qi / Lq;
  const float qx = query[qi * 3 + 0], qy = query[qi * 3 + 1], qz = query[qi * 3 + 2];
  const float* rb = ref + (size_t)b * Lr * 3;
  float best_d = INFINITY;  // lane l holds the (l+1)-th smallest so far
  int best_i = -1;
  for (int base = 0; base < Lr; base += 32) {
    const int c = base + lane;
    float d = INFINITY;
    if (c < Lr) {
      const float dx = __fsub_rn(qx, rb[c * 3 + 0]);
      const float dy = __fsub_rn(qy, rb[c * 3 + 1]);
      const float dz = __fsub_rn(qz, rb[c * 3 + 2]);
      d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
    float thr = __shfl_sync(0xffffffffu, best_d, 31);
    unsigned cand = __ballot_sync(0xffffffffu, d < thr);
    while (cand) {
      const int src = __ffs(cand) - 1;
      cand &= cand - 1;
      const float cd = __shfl_sync(0xffffffffu, d, src);
      if (cd < thr) {  // warp-uniform
        // stable position: after every element <= cd (earlier index wins ties)
        const unsigned le = __ballot_sync(0xffffffffu, best_d <= cd);
        const int pos = __popc(le);
        const float up_d = __shfl_up_sync(0xffffffffu, best_d, 1);
        const int up_i = __shfl_up_sync(0xffffffffu, best_i, 1);
        if (lane > pos) {
          best_d = up_d;
          best_i = up_i;
        } else if (lane == pos) {
          best_d = cd;
          best_i = base + src;
        }
        thr = __shfl_sync(0xffffffffu, best_d, 31);
      }
    }
  }
  idx_out[(size_t)qi * 32 + lane] = best_i;
}

// ------------------------------------------------------------------------------------------------
// 32-NN against the FIXED basis-point set (cross attention, point_transformers.py:134): same result as knn32_kernel,
// but the 4096 BPS points are visited in a spatially sorted (Morton) order prepared at pack time, 32 per chunk, and a
// chunk is skipped when the distance from the query to its bounding box already exceeds the current 32nd-best
// distance.  The per-sample coordinates ((bps + c) - c) / r differ from the canonical bps / r only by rounding, so
// the boxes are grown by 1e-4 at pack time; distances are still evaluated exactly on the per-sample coordinates
// (passed in chunk order: ref_sorted[b][k] = ref[b][perm[k]]) and
// ties are broken by the ORIGINAL index, so the output is bit-identical to the brute-force kernel.
// ------------------------------------------------------------------------------------------------
__global__ void knn32_bps_kernel(const float* __restrict__ query, const float* __restrict__ ref_sorted,
                                 const int* __restrict__ perm, const float* __restrict__ boxes,
                                 int* __restrict__ idx_out, int Lq, int Lr, int n_query_total) {
  pdl_wait();
  pdl_trigger();
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= n_query_total) return;
  const int b = qi / Lq;
  const float qx = query[qi * 3 + 0], qy = query[qi * 3 + 1], qz = query[qi * 3 + 2];
  const float* rb = ref_sorted + (size_t)b * Lr * 3;   // per-sample coordinates already in chunk order
  const int n_chunks = Lr >> 5;          // Lr is a multiple of 32 (4096)
  constexpr int SLOTS = 4;               // up to 128 chunks: lane l owns chunks l, l+32, l+64, l+96
  float lb[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int c = s * 32 + lane;
    float v = INFINITY;
    if (c < n_chunks) {
      const float* bx = boxes + c * 6;
      const float dx = fmaxf(fmaxf(bx[0] - qx, qx - bx[3]), 0.f);
      const float dy = fmaxf(fmaxf(bx[1] - qy, qy - bx[4]), 0.f);
      const float dz = fmaxf(fmaxf(bx[2] - qz, qz - bx[5]), 0.f);
      v = (dx * dx + dy * dy + dz * dz) * 0.9999f;   // strict lower bound of any true distance in the chunk
    }
    lb[s] = v;
  }
  float best_d = INFINITY;   // lane l holds the (l+1)-th smallest so far, ordered by (distance, original index)
  int best_i = -1;

  auto scan_chunk = [&](int c) {
    const int k = c * 32 + lane;
    const int oi = perm[k];
    const float dx = __fsub_rn(qx, rb[k * 3 + 0]);
    const float dy = __fsub_rn(qy, rb[k * 3 + 1]);
    const float dz = __fsub_rn(qz, rb[k * 3 + 2]);
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    float thr_d = __shfl_sync(0xffffffffu, best_d, 31);
    int thr_i = __shfl_sync(0xffffffffu, best_i, 31);
    unsigned cand = __ballot_sync(0xffffffffu, d < thr_d || (d == thr_d && oi < thr_i));
    while (cand) {
      const int src = __ffs(cand) - 1;
      cand &= cand - 1;
      const float cd = __shfl_sync(0xffffffffu, d, src);
      const int ci = __shfl_sync(0xffffffffu, oi, src);
      if (cd < thr_d || (cd == thr_d && ci < thr_i)) {   // warp-uniform
        const unsigned before = __ballot_sync(0xffffffffu, best_d < cd || (best_d == cd && best_i < ci));
        const int pos = __popc(before);
        const float up_d = __shfl_up_sync(0xffffffffu, best_d, 1);
        const int up_i = __shfl_up_sync(0xffffffffu, best_i, 1);
        if (lane > pos) {
          best_d = up_d;
          best_i = up_i;
        } else if (lane == pos) {
          best_d = cd;
          best_i = ci;
        }
        thr_d = __shfl_sync(0xffffffffu, best_d, 31);
        thr_i = __shfl_sync(0xffffffffu, best_i, 31);
      }
    }
  };

  // start with the chunk whose box is closest to the query: it fills the list with nearby points
  float mn = fminf(fminf(lb[0], lb[1]), fminf(lb[2], lb[3]));
  int arg = (mn == lb[0]) ? lane : (mn == lb[1]) ? 32 + lane : (mn == lb[2]) ? 64 + lane : 96 + lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mn, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om < mn || (om == mn && oa < arg)) {
      mn = om;
      arg = oa;
    }
  }
  const int first = arg;
  scan_chunk(first);
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    float thr = __shfl_sync(0xffffffffu, best_d, 31);
    unsigned todo = __ballot_sync(0xffffffffu, lb[s] <= thr && (s * 32 + lane) != first && (s * 32 + lane) < n_chunks);
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      thr = __shfl_sync(0xffffffffu, best_d, 31);
      const float l_lb = __shfl_sync(0xffffffffu, lb[s], l);
      if (l_lb <= thr) scan_chunk(s * 32 + l);      // warp-uniform re-check against the tightened threshold
    }
  }
  idx_out[(size_t)qi * 32 + lane] = best_i;
}

// ------------------------------------------------------------------------------------------------
// Vector attention (Point-Transformer layer) glue, un-fused variant (token tensors in HBM):
//   tokens t = (b, i, j), j < 32 neighbours.  reference: point_transformers.py:86-95,139-151
// ------------------------------------------------------------------------------------------------
// h_delta[t, c] = relu(Wd1[c,:] · (xyz_i - nbr_xyz_j) + bd1[c])
__global__ void va_hdelta_kernel(const float* __restrict__ q_xyz, const float* __restrict__ ref_xyz,
                                 const int* __restrict__ idx, const float* __restrict__ anchor_xyz,
                                 const float* __restrict__ wd1, const float* __restrict__ bd1,
                                 op16* __restrict__ hdelta, int Lq, int Lr, int D, size_t n_tokens) {
  const size_t t = blockIdx.x;  // one block per token group of 8
  const int sub = threadIdx.x / (blockDim.x / 8);
  const int tl = threadIdx.x % (blockDim.x / 8);
  const size_t tok = t * 8 + sub;
  if (tok >= n_tokens) return;
  const size_t qi = tok >> 5;
  const int j = (int)(tok & 31);
  const int b = (int)(qi / Lq);
  float nx, ny, nz;
  if (anchor_xyz != nullptr) {
    nx = anchor_xyz[j * 3 + 0], ny = anchor_xyz[j * 3 + 1], nz = anchor_xyz[j * 3 + 2];
  } else {
    const int r = idx[tok];
    const float* rp = ref_xyz + ((size_t)b * Lr + r) * 3;
    nx = rp[0], ny = rp[1], nz = rp[2];
  }
  const float rx = q_xyz[qi * 3 + 0] - nx, ry = q_xyz[qi * 3 + 1] - ny, rz = q_xyz[qi * 3 + 2] - nz;
  for (int c = tl; c < D; c += blockDim.x / 8) {
    float v = wd1[c * 3 + 0] * rx;
    v += wd1[c * 3 + 1] * ry;
    v += wd1[c * 3 + 2] * rz;
    v += bd1[c];
    hdelta[tok * D + c] = f2op16(fmaxf(v, 0.f));
  }
}

// g[t, c] = relu(g[t, c] + qt[i, c] - kt[nbr_j, c])   (in place; g holds (W_g1 W_d2) h on entry)
__global__ void va_gmix_kernel(const op16* __restrict__ q, int ldq, const op16* __restrict__ ktab,
                               int ldk, const int* __restrict__ idx, const int* __restrict__ anchor_idx,
                               op16* __restrict__ g, int Lq, int Lr, int D, size_t n_tokens) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int vec = D / 8;
  const size_t tok = gid / vec;
  const int c = (int)(gid % vec) * 8;
  if (tok >= n_tokens) return;
  const size_t qi = tok >> 5;
  const int j = (int)(tok & 31);
  const int b = (int)(qi / Lq);
  const int r = (anchor_idx != nullptr) ? anchor_idx[j] : idx[tok];
  const uint4 qv = *reinterpret_cast<const uint4*>(q + qi * ldq + c);
  const uint4 kv = *reinterpret_cast<const uint4*>(ktab + ((size_t)b * Lr + r) * ldk + c);
  const uint4 gv = *reinterpret_cast<const uint4*>(g + tok * D + c);
  const op16x2* q2 = reinterpret_cast<const op16x2*>(&qv);
  const op16x2* k2 = reinterpret_cast<const op16x2*>(&kv);
  const op16x2* g2 = reinterpret_cast<const op16x2*>(&gv);
  uint4 ov;
  uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = op16x2_to_f2(q2[i]), bb = op16x2_to_f2(k2[i]), cc = op16x2_to_f2(g2[i]);
    o[i] = pack_op16x2(fmaxf(cc.x + (a.x - bb.x), 0.f), fmaxf(cc.y + (a.y - bb.y), 0.f));
  }
  *reinterpret_cast<uint4*>(g + tok * D + c) = ov;
}

// res[i, c] = sum_j softmax_j(a[t, c] * inv_sqrt_d) * (v[nbr_j, c] + pos[t, c]);  one thread per (query, channel)
__global__ void va_reduce_kernel(const op16* __restrict__ a, const op16* __restrict__ pos,
                                 const op16* __restrict__ vtab, int ldv, const int* __restrict__ idx,
                                 const int* __restrict__ anchor_idx, op16* __restrict__ res, int Lq, int Lr,
                                 int D, float inv_sqrt_d, size_t n_query) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t qi = gid / D;
  const int c = (int)(gid % D);
  if (qi >= n_query) return;
  const int b = (int)(qi / Lq);
  float av[32];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    av[j] = op16_to_f(a[(qi * 32 + j) * D + c]) * inv_sqrt_d;
    mx = fmaxf(mx, av[j]);
  }
  float sum = 0.f, acc = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float e = __expf(av[j] - mx);
    const int r = (anchor_idx != nullptr) ? anchor_idx[j] : idx[qi * 32 + j];
    const float val = op16_to_f(vtab[((size_t)b * Lr + r) * ldv + c]) + op16_to_f(pos[(qi * 32 + j) * D + c]);
    sum += e;
    acc += e * val;
  }
  res[qi * D + c] = f2op16(acc / sum);
}

// ------------------------------------------------------------------------------------------------
// Coordinate regression tail: xyz' = xyz + W2 · h + b2 (h = relu(W1 f + b1) from the GEMM), W2 is (3,D).
// Also writes the de-normalised prediction  out = nan_to_num(xyz') * r + centre  (ptEmb_head.py:944-948).
// one warp per query.
// ------------------------------------------------------------------------------------------------
__global__ void reg_out_kernel(const op16* __restrict__ h, const float* __restrict__ w2,
                               const float* __restrict__ b2, const float* __restrict__ xyz_in,
                               float* __restrict__ xyz_out, float* __restrict__ coords_out,
                               const float* __restrict__ centre, float radius, int Lq, int D, int n_query) {
  pdl_wait();
  pdl_trigger();
  const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (qi >= n_query) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float hv = op16_to_f(h[(size_t)qi * D + c]);
    a0 += hv * w2[c];
    a1 += hv * w2[D + c];
    a2 += hv * w2[2 * D + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if (lane < 3) {
    const float d = (lane == 0) ? a0 : (lane == 1) ? a1 : a2;
    float v = xyz_in[qi * 3 + lane] + d + b2[lane];
    xyz_out[qi * 3 + lane] = v;
    if (centre == nullptr) {  // transformer-only call: raw normalised coordinates
      coords_out[qi * 3 + lane] = v;
    } else {
      // torch.nan_to_num: nan -> 0, +-inf -> +-FLT_MAX
      if (isnan(v)) v = 0.f;
      else if (isinf(v)) v = (v > 0.f) ? 3.4028234663852886e38f : -3.4028234663852886e38f;
      const int b = qi / Lq;
      coords_out[qi * 3 + lane] = v * radius + centre[b * 3 + lane];
    }
  }
}

// pt_xyz = ((bps + c) - c) / r  and  q_xyz = ((c + template) - c) / r  in the reference's rounding order
// (ptEmb_head.py:808,896-897,933-934)
__global__ void normalise_points_kernel(const float* __restrict__ bps, const float* __restrict__ templ,
                                        const float* __restrict__ centre, float* __restrict__ pt_xyz,
                                        float* __restrict__ q_xyz, int P, int Q, float radius, int B,
                                        const int* __restrict__ perm, float* __restrict__ pt_xyz_sorted) {
  pdl_wait();
  pdl_trigger();
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = (P + Q) * 3;
  if (gid >= B * per) return;
  const int b = gid / per, r = gid - b * per;
  if (r < P * 3) {
    const float c = centre[b * 3 + r % 3];
    pt_xyz[(size_t)b * P * 3 + r] = __fdiv_rn(__fsub_rn(__fadd_rn(bps[r], c), c), radius);
    if (perm != nullptr) {   // the same value of point perm[k], stored at chunk-order position k (32-NN pruning)
      const int k = r / 3, a = r - k * 3;
      const float cc = centre[b * 3 + a];
      pt_xyz_sorted[(size_t)b * P * 3 + r] = __fdiv_rn(__fsub_rn(__fadd_rn(bps[perm[k] * 3 + a], cc), cc), radius);
    }
  } else {
    const int k = r - P * 3;
    const float c = centre[b * 3 + k % 3];
    q_xyz[(size_t)b * Q * 3 + k] = __fdiv_rn(__fsub_rn(__fadd_rn(c, templ[k]), c), radius);
  }
}

// centre[b] = reference_joints[b, centre_idx]
__global__ void gather_centre_kernel(const float* __restrict__ ref_joints, float* __restrict__ centre, int centre_idx,
                                     int B) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 3) centre[i] = ref_joints[(i / 3) * 63 + centre_idx * 3 + (i % 3)];
}

__global__ void f32_to_op16_kernel(const float* __restrict__ x, op16* __restrict__ y, size_t n) {
  pdl_wait();
  pdl_trigger();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = f2op16(x[i]);
}

// broadcast the (Q,D) query embedding table to (B*Q, D) fp32 + op16
__global__ void broadcast_queries_kernel(const float* __restrict__ table, float* __restrict__ out_f32,
                                         op16* __restrict__ out_op16, int QD, size_t total) {
  pdl_wait();
  pdl_trigger();
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const float v = table[gid % QD];
  out_f32[gid] = v;
  out_op16[gid] = f2op16(v);
}

// Per-image / per-sample index tables from the per-sample view counts (passed by value):
//   dev = [img_sample (NV) | img_view (NV) | img_posrow (NV) | sample_rowbase (B) | sample_views (B) | tile_start (B+1)]
//   tile_start: first row tile of every sample in the fused sampler/merge kernel (tiles of floor(128 / n) tokens)
constexpr int VIEW_PARAM_MAX = 256;
struct ViewCountsParam {
  int n[VIEW_PARAM_MAX];
};
__host__ __device__ inline int merge_tiles_of(int n_views, int P) {
  const int tok = 128 / n_views;
  return (P + tok - 1) / tok;
}
__global__ void view_tables_kernel(ViewCountsParam vc, int B, int NV, int P, int* __restrict__ dev) {
  pdl_wait();
  pdl_trigger();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int first = 0, tiles = 0;
    for (int i = 0; i < b; ++i) first += vc.n[i], tiles += merge_tiles_of(vc.n[i], P);
    const int n = vc.n[b];
    dev[3 * NV + b] = first * P;
    dev[3 * NV + B + b] = n;
    dev[3 * NV + 2 * B + b] = tiles;
    if (b == B - 1) dev[3 * NV + 2 * B + B] = tiles + merge_tiles_of(n, P);
    for (int v = 0; v < n; ++v) {
      dev[first + v] = b;
      dev[NV + first + v] = v;
      dev[2 * NV + first + v] = n * (n - 1) / 2 + v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Evaluation metrics on the device (reference lib/metrics/pa_eval.py:41-124, mean_epe.py:23-33): per sample the mean
// point distance and the Procrustes-aligned mean point distance (`PAEval.align_w_scale`: centre both sets, scale each
// to unit Frobenius norm (+1e-8), orthogonal Procrustes R = U V^T of M = A^T B = U S V^T with scale sum(S), map the
// prediction back with the ground truth's scale and centre).  The reference loops over samples around SciPy on the
// host; here one block per sample, sums in fp64, the 3x3 SVD by one-sided Jacobi in fp64.
//   gt, pred: (batch, n, 3) fp32;  out: (batch, 2) = [aligned mean distance, raw mean distance];  aligned: optional
// ------------------------------------------------------------------------------------------------
constexpr int PA_THREADS = 128;
__device__ __forceinline__ double block_sum_128(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}
__global__ void __launch_bounds__(PA_THREADS)
pa_metrics_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int n, float* __restrict__ out,
                  float* __restrict__ aligned) {
  __shared__ double red[4];
  __shared__ double sR[9];
  __shared__ double sScale;
  const float* a = gt + (size_t)blockIdx.x * n * 3;
  const float* b = pred + (size_t)blockIdx.x * n * 3;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += PA_THREADS)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += (double)a[i * 3 + c], acc[3 + c] += (double)b[i * 3 + c];
  float t1[3], t2[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    t1[c] = (float)(block_sum_128(acc[c], red) / n);       // the reference works in fp32 (numpy arrays of the tensors)
    t2[c] = (float)(block_sum_128(acc[3 + c], red) / n);
  }
  double n1 = 0, n2 = 0, m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += PA_THREADS) {
    float x[3], y[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = a[i * 3 + c] - t1[c], y[c] = b[i * 3 + c] - t2[c];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      n1 += (double)x[r] * x[r];
      n2 += (double)y[r] * y[r];
#pragma unroll
      for (int c = 0; c < 3; ++c) m[r * 3 + c] += (double)x[r] * y[c];
    }
  }
  const float s1 = (float)sqrt(block_sum_128(n1, red)) + 1e-8f;
  const float s2 = (float)sqrt(block_sum_128(n2, red)) + 1e-8f;
  double M[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = block_sum_128(m[k], red) / ((double)s1 * (double)s2);
  if (threadIdx.x == 0) {
    // one-sided Jacobi: rotate column pairs of W = M V until orthogonal; then W = U S
    double W[3][3], V[3][3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) W[r][c] = M[r * 3 + c], V[r][c] = (r == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
      double off = 0.0;
      for (int p = 0; p < 2; ++p)
        for (int q = p + 1; q < 3; ++q) {
          double al = 0, be = 0, ga = 0;
          for (int r = 0; r < 3; ++r) al += W[r][p] * W[r][p], be += W[r][q] * W[r][q], ga += W[r][p] * W[r][q];
          if (ga == 0.0 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;
          off = fmax(off, fabs(ga) / sqrt(al * be));
          const double zeta = (be - al) / (2.0 * ga);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          for (int r = 0; r < 3; ++r) {
            const double x = W[r][p], y = W[r][q];
            W[r][p] = cs * x - sn * y, W[r][q] = sn * x + cs * y;
            const double vx = V[r][p], vy = V[r][q];
            V[r][p] = cs * vx - sn * vy, V[r][q] = sn * vx + cs * vy;
          }
        }
      if (off < 1e-15) break;
    }
    double sig[3], scale = 0.0;
    for (int c = 0; c < 3; ++c) {
      sig[c] = sqrt(W[0][c] * W[0][c] + W[1][c] * W[1][c] + W[2][c] * W[2][c]);
      scale += sig[c];
    }
    // R = U V^T with U = W / sigma (a vanishing singular value leaves that column of U free: degenerate input)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double v = 0.0;
        for (int k = 0; k < 3; ++k) v += (sig[k] > 0.0 ? W[r][k] / sig[k] : 0.0) * V[c][k];
        sR[r * 3 + c] = v;
      }
    sScale = scale;
  }
  __syncthreads();
  double d_al = 0.0, d_raw = 0.0;
  const float sc = (float)sScale;
  for (int i = threadIdx.x; i < n; i += PA_THREADS) {
    float y[3], z[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = (b[i * 3 + c] - t2[c]) / s2;
#pragma unroll
    for (int r = 0; r < 3; ++r)   // (pred_t . R^T) * s, then the ground truth's scale and centre
      z[r] = ((float)sR[r * 3 + 0] * y[0] + (float)sR[r * 3 + 1] * y[1] + (float)sR[r * 3 + 2] * y[2]) * sc * s1 + t1[r];
    float e0 = 0.f, e1 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float da = z[c] - a[i * 3 + c], dr = b[i * 3 + c] - a[i * 3 + c];
      e0 += da * da, e1 += dr * dr;
      if (aligned != nullptr) aligned[((size_t)blockIdx.x * n + i) * 3 + c] = z[c];
    }
    d_al += (double)sqrtf(e0), d_raw += (double)sqrtf(e1);
  }
  const double sa = block_sum_128(d_al, red), sr = block_sum_128(d_raw, red);
  if (threadIdx.x == 0) out[blockIdx.x * 2] = (float)(sa / n), out[blockIdx.x * 2 + 1] = (float)(sr / n);
}

}  // namespace poem
