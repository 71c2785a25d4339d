// Fused multi-view feature sampler + merge network MLP0 + cross-view reduce
// (reference: F.grid_sample ptEmb_head.py:900-901, the raw `.view(1,-1,N,D)` regroup :914-915, merge_features_mv /
// merge_features_sv :745-771 with merge_net_feature.0 :701-703).
//
// Round 1 ran this stage as four kernels with three token-sized tensors round-tripping HBM (X 537 MB, H1 537 MB,
// Mm 268 MB for POEM-medium, 8 views, batch 32: ~2.7 GB of traffic for ~155 MB of algorithmic bytes).  Here a tile
// of 128 merge rows never leaves the SM:
//
//   rows r of sample b (N views): X[r, :] = S[n, d, p0 .. p0 + D)  with n = r / P, d = (r % P) / G, p0 = (r % G) * D,
//   G = P / D — i.e. ONE channel plane (n, d) sampled at D consecutive BPS points (the reference re-interprets the
//   (N, D, P) grid_sample output as (P, N, D)); token p' owns rows p' N .. p' N + N - 1.
//
//   sampler warps : bilinear taps (4 weights + 4 pixel offsets per (image, point), precomputed once per call by
//                   sample_taps_kernel, L2-resident) x the tile's channel planes (pixel-major fp32 slab in smem)
//                   -> A tile [128 rows x D] fp16, K-major SWIZZLE_128B, double buffered; the token-first rows
//                   (q1 of merge_features_mv) also go to HBM
//   MMA warp      : acc1 = A · W0a^T (TMEM, D columns) ; acc2 = H1 · W0b^T (TMEM, D/2 columns); weights streamed by
//                   a TMA ring of [128 x 64] tiles
//   epilogue warps: H1 = relu(acc1 + b0a) -> fp16 over the dead A tile ; Mm = acc2 + b0b -> fp32 staging over the dead
//                   H1 tile ; w_n = <m_n, m_0>, s = sum_{n>=1} w_n m_n (s = m_0 for one view) -> S / sigma (fp16) + sigma
//
// HBM traffic: reads the feature volume (NV * D * 1 KB) and the tap table, writes q1 (B*P*D*2) and S (B*P*D/2*2).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace poem {

constexpr int SM_P = 4096;         // BPS points
constexpr int SM_F = 256;          // feature pixels
constexpr int SM_TAP_WORDS = 24;   // per group of 4 points: w00[4] w01[4] w10[4] w11[4] (fp32), then per point two words
                                   // = four u16 byte offsets (pixel index * slab pixel pitch) of the taps nw|ne, sw|se

// ------------------------------------------------------------------------------------------------
// taps[img][p / 4][20]: projection of BPS point p into image img + bilinear weights (zero padding folded into the
// weights), same arithmetic as the reference chain collation.py:48-65 -> transform.py:898-930 -> grid_sample
// ------------------------------------------------------------------------------------------------
__global__ void sample_taps_kernel(const float* __restrict__ proj, const float* __restrict__ bps,
                                   const float* __restrict__ centre, const int* __restrict__ img_sample,
                                   uint32_t* __restrict__ taps, int n_img, int FH, int FW, float inv_w, float inv_h,
                                   int pixel_pitch_bytes) {
  pdl_wait();
  pdl_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * SM_P) return;
  const int img = idx / SM_P, p = idx - img * SM_P;
  const int b = img_sample[img];
  const float* pm = proj + img * 24;
  const float cx = centre[b * 3 + 0], cy = centre[b * 3 + 1], cz = centre[b * 3 + 2];
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
  const float w00 = (vx0 && vy0) ? (1.f - ax) * (1.f - ay) : 0.f;
  const float w01 = (vx1 && vy0) ? ax * (1.f - ay) : 0.f;
  const float w10 = (vx0 && vy1) ? (1.f - ax) * ay : 0.f;
  const float w11 = (vx1 && vy1) ? ax * ay : 0.f;
  const uint32_t o00 = (vx0 && vy0) ? y0 * FW + x0 : 0;
  const uint32_t o01 = (vx1 && vy0) ? y0 * FW + x0 + 1 : 0;
  const uint32_t o10 = (vx0 && vy1) ? (y0 + 1) * FW + x0 : 0;
  const uint32_t o11 = (vx1 && vy1) ? (y0 + 1) * FW + x0 + 1 : 0;
  uint32_t* t = taps + ((size_t)img * (SM_P / 4) + (p >> 2)) * SM_TAP_WORDS + (p & 3);
  t[0] = __float_as_uint(w00);
  t[4] = __float_as_uint(w01);
  t[8] = __float_as_uint(w10);
  t[12] = __float_as_uint(w11);
  const uint32_t pb = (uint32_t)pixel_pitch_bytes;   // 255 * 48 < 65536
  uint32_t* o = taps + ((size_t)img * (SM_P / 4) + (p >> 2)) * SM_TAP_WORDS + 16 + 2 * (p & 3);
  o[0] = (o00 * pb) | ((o01 * pb) << 16);
  o[1] = (o10 * pb) | ((o11 * pb) << 16);
}

// test helper: packed tap table -> (n_img, P, 4) pixel indices (nw, ne, sw, se; pixel pitch 1) and weights
__global__ void unpack_taps_kernel(const uint32_t* __restrict__ taps, int32_t* __restrict__ pixels, float* __restrict__ weights,
                                   int n_img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * SM_P) return;
  const int img = idx / SM_P, p = idx - img * SM_P;
  const uint32_t* t = taps + ((size_t)img * (SM_P / 4) + (p >> 2)) * SM_TAP_WORDS;
  const uint32_t oa = t[16 + 2 * (p & 3)], ob = t[17 + 2 * (p & 3)];
  pixels[idx * 4 + 0] = (int32_t)(oa & 0xffffu), pixels[idx * 4 + 1] = (int32_t)(oa >> 16);
  pixels[idx * 4 + 2] = (int32_t)(ob & 0xffffu), pixels[idx * 4 + 3] = (int32_t)(ob >> 16);
#pragma unroll
  for (int k = 0; k < 4; ++k) weights[idx * 4 + k] = __uint_as_float(t[4 * k + (p & 3)]);
}

template <int D>
struct SmCfg {
  static_assert(D == 128 || D == 256, "fused sampler/merge kernel: D = 128 or 256 (D = 512 takes the un-fused path)");
  static constexpr int H = D / 2;                    // merge MLP0 output width
  static constexpr int G = SM_P / D;                 // rows per channel plane
  static constexpr int RPI = 128 / G;                // rows (= consecutive planes) per sampler work item, tile aligned
  static constexpr int SLOTS = ((RPI + 1 + 3) / 4) * 4;   // slab planes allocated per tile (one extra when misaligned)
  static constexpr int KC = D / 8;                   // 16-byte pieces per A row
  static constexpr int KB = D / 64;                  // 64-wide K blocks of A / H1
  static constexpr int X_BYTES = 128 * D * 2;        // A / H1 tile (also the fp32 Mm staging: 128 x H x 4 bytes)
  static constexpr int W_TILE_BYTES = 128 * 64 * 2;
  static constexpr int W_STAGES = 4;
  static constexpr int SLAB_BYTES = SM_F * SLOTS * 4;
  static constexpr int N_SAMPLER = 256, N_EPI = 256;
  static constexpr int THREADS = 64 + N_EPI + N_SAMPLER;
  static constexpr int OFF_X = 0;                                  // two X buffers
  static constexpr int OFF_W = 2 * X_BYTES;
  static constexpr int OFF_SLAB = OFF_W + W_STAGES * W_TILE_BYTES;  // two slabs
  static constexpr int OFF_BIAS = OFF_SLAB + 2 * SLAB_BYTES;        // b0a [D] | b0b [H]
  static constexpr int OFF_TOK = OFF_BIAS + (D + H) * 4;             // two tables of 128 ints (token of a tile row)
  static constexpr int OFF_BARS = OFF_TOK + 2 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BARS + 256;
  static constexpr int TMEM_COLS = (D + H <= 256) ? 256 : 512;
  static constexpr int W1_STEPS = KB * (D / 128);    // weight tiles of MLP0 layer 1 per row tile ([128 n] x [64 k])
  static constexpr int W2_STEPS = KB * (H / 128 > 0 ? H / 128 : 1);
};

struct SmParams {
  const float* xmap;           // (NV, D, 256) fp32 channel-planar feature volume
  const uint32_t* taps;        // sample_taps_kernel output
  const int* tile_start;       // [B + 1] first tile of every sample (tiles of floor(128 / N) tokens)
  const int* sample_views;     // [B]
  const int* sample_rowbase;   // [B] = (first image of the sample) * P
  const float* b0a;            // [D]
  const float* b0b;            // [H]
  op16* q1;                    // [B * P, D]   token-first rows of X (residual of merge_features_mv / _sv)
  op16* s;                     // [B * P, H]   cross-view aggregate / sigma
  float* sigma;                // [B * P]
  int n_samples, n_tiles;
};

// tile -> sample lookup for a monotonically increasing tile sequence
struct SmTileCursor {
  int b = 0, lo = 0, hi = 0;   // tiles [lo, hi) belong to sample b
  int n_views = 1, img0 = 0;   // of sample b (re-read only when the sample changes)
  __device__ __forceinline__ void seek(const SmParams& p, int tile) {
    while (tile >= hi) {
      if (hi != 0) ++b;
      lo = p.tile_start[b];
      hi = p.tile_start[b + 1];
      n_views = p.sample_views[b];
      img0 = p.sample_rowbase[b] / SM_P;
    }
  }
};

template <int D>
__global__ void __launch_bounds__(SmCfg<D>::THREADS, 1)
sample_merge_kernel(const __grid_constant__ CUtensorMap tmap_w0a, const __grid_constant__ CUtensorMap tmap_w0b, SmParams p) {
  using Cfg = SmCfg<D>;
  constexpr int H = Cfg::H, G = Cfg::G, RPI = Cfg::RPI, SLOTS = Cfg::SLOTS, KC = Cfg::KC, KB = Cfg::KB;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_x = smem + Cfg::OFF_X;
  uint8_t* s_w = smem + Cfg::OFF_W;
  float* s_slab = reinterpret_cast<float*>(smem + Cfg::OFF_SLAB);
  float* s_b0a = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  float* s_b0b = s_b0a + D;
  int* s_tok = reinterpret_cast<int*>(smem + Cfg::OFF_TOK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* w_full = bars;                       // [W_STAGES]
  uint64_t* w_empty = bars + Cfg::W_STAGES;      // [W_STAGES]
  uint64_t* a_full = bars + 2 * Cfg::W_STAGES;   // [2] sampler -> MMA (count N_SAMPLER)
  uint64_t* x_free = a_full + 2;                 // [2] epilogue -> sampler (count N_EPI)
  uint64_t* acc1_full = x_free + 2;              // MMA -> epilogue
  uint64_t* h1_full = acc1_full + 1;             // epilogue -> MMA (count N_EPI)
  uint64_t* acc2_full = h1_full + 1;             // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w0a);
    tma_prefetch_desc(&tmap_w0b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::W_STAGES; ++s) {
        mbar_init(&w_full[s], 1);
        mbar_init(&w_empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&a_full[s], Cfg::N_SAMPLER);
        mbar_init(&x_free[s], Cfg::N_EPI);
      }
      mbar_init(acc1_full, 1);
      mbar_init(h1_full, Cfg::N_EPI);
      mbar_init(acc2_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  for (int c = threadIdx.x; c < D + H; c += Cfg::THREADS) s_b0a[c] = (c < D) ? p.b0a[c] : p.b0b[c - D];
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the setup above only touched weights (biases) and on-chip state, and the TMA warp only ever streams weights:
  // it runs ahead while the predecessor kernel drains, everybody else waits for it here
  if (warp != 0) pdl_wait();
  pdl_trigger();
  const uint32_t tmem_acc1 = tmem_base;          // D columns
  const uint32_t tmem_acc2 = tmem_base + D;      // H columns

  if (warp == 0) {
    // ===================== TMA weight producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int step = 0; step < Cfg::W1_STEPS + Cfg::W2_STEPS; ++step) {
          const bool l1 = step < Cfg::W1_STEPS;
          const int st = l1 ? step : step - Cfg::W1_STEPS;
          // layer 1: K block outer, 128-wide n half inner; layer 2: K block
          const int kb = l1 ? st / (D / 128) : st / (H / 128 > 0 ? H / 128 : 1);
          const int nh = l1 ? st % (D / 128) : 0;
          mbar_wait(&w_empty[stage], phase ^ 1);
          mbar_expect_tx(&w_full[stage], l1 ? Cfg::W_TILE_BYTES : (H < 128 ? H * 64 * 2 : Cfg::W_TILE_BYTES));
          tma_load_2d(s_w + stage * Cfg::W_TILE_BYTES, l1 ? &tmap_w0a : &tmap_w0b, &w_full[stage], kb * 64, nh * 128);
          if (++stage == Cfg::W_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr int N1 = 128;                     // layer 1: n halves of 128
      constexpr int N2 = (H < 128) ? H : 128;     // layer 2: H columns (64 or 128)
      constexpr uint32_t idesc1 = make_idesc_op16(128, N1);
      constexpr uint32_t idesc2 = make_idesc_op16(128, N2);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t x_addr = smem_u32(s_x + buf * Cfg::X_BYTES);
        mbar_wait(&a_full[buf], (uint32_t)(it >> 1) & 1);
        tc_fence_after_sync();
        for (int st = 0; st < Cfg::W1_STEPS; ++st) {
          const int kb = st / (D / 128), nh = st % (D / 128);
          mbar_wait(&w_full[stage], phase);
          tc_fence_after_sync();
          const uint64_t da = make_kmajor_desc<128>(x_addr + kb * (128 * 128));
          const uint64_t dw = make_kmajor_desc<128>(smem_u32(s_w) + stage * Cfg::W_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_op16(tmem_acc1 + nh * 128, da + 2 * k, dw + 2 * k, idesc1, (kb | k) != 0);
          umma_commit(&w_empty[stage]);
          if (++stage == Cfg::W_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(acc1_full);
        mbar_wait(h1_full, (uint32_t)it & 1);
        tc_fence_after_sync();
        for (int st = 0; st < Cfg::W2_STEPS; ++st) {
          const int kb = st;
          mbar_wait(&w_full[stage], phase);
          tc_fence_after_sync();
          const uint64_t da = make_kmajor_desc<128>(x_addr + kb * (128 * 128));
          const uint64_t dw = make_kmajor_desc<128>(smem_u32(s_w) + stage * Cfg::W_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_op16(tmem_acc2, da + 2 * k, dw + 2 * k, idesc2, (kb | k) != 0);
          umma_commit(&w_empty[stage]);
          if (++stage == Cfg::W_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(acc2_full);
      }
    }
  } else if (warp < 2 + Cfg::N_EPI / 32) {
    // ===================== epilogue warps (2..9) =====================
    const int et = threadIdx.x - 64;               // 0..255
    const int ew = warp - 2;                       // 0..7
    const int quarter = warp & 3;                  // TMEM lane quarter of this warp
    const int colh = ew >> 2;                      // column half
    const int row = quarter * 32 + lane;           // tile row == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    SmTileCursor cur;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      uint8_t* xb = s_x + buf * Cfg::X_BYTES;
      cur.seek(p, tile);
      const int N = cur.n_views;
      const int TOK = 128 / N;
      const int tok0 = (tile - cur.lo) * TOK;
      const int ntok = min(TOK, SM_P - tok0);
      // ---- epilogue 1: H1 = relu(acc1 + b0a) -> fp16, K-major SWIZZLE_128B, over the (dead) A tile
      mbar_wait(acc1_full, (uint32_t)it & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < D / 2; c0 += 32) {
        const int c = colh * (D / 2) + c0;
        uint32_t r[32];
        tmem_ld32(tmem_acc1 + lane_off + c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_b0a + c + 4 * i);
          pk[2 * i] = pack_op16x2_relu(__uint_as_float(r[4 * i]) + b4.x, __uint_as_float(r[4 * i + 1]) + b4.y);
          pk[2 * i + 1] = pack_op16x2_relu(__uint_as_float(r[4 * i + 2]) + b4.z, __uint_as_float(r[4 * i + 3]) + b4.w);
        }
        uint8_t* blk = xb + (c >> 6) * (128 * 128);
        const uint32_t chunk0 = (uint32_t)(c & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(blk + sw128_offset(row, chunk0 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(h1_full);
      // ---- epilogue 2: Mm = acc2 + b0b -> fp32 staging [128 rows x H], 16-byte chunks XOR-swizzled by (row & 7),
      //      over the (dead) H1 tile
      mbar_wait(acc2_full, (uint32_t)it & 1);
      tc_fence_after_sync();
      float* stg = reinterpret_cast<float*>(xb);
#pragma unroll
      for (int c0 = 0; c0 < H / 2; c0 += 32) {
        const int c = colh * (H / 2) + c0;
        uint32_t r[32];
        tmem_ld32(tmem_acc2 + lane_off + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c4 = (c >> 2) + q;           // 16-byte chunk (4 channels)
          const float4 b4 = *reinterpret_cast<const float4*>(s_b0b + c + 4 * q);
          float4 v;
          v.x = __uint_as_float(r[4 * q + 0]) + b4.x;
          v.y = __uint_as_float(r[4 * q + 1]) + b4.y;
          v.z = __uint_as_float(r[4 * q + 2]) + b4.z;
          v.w = __uint_as_float(r[4 * q + 3]) + b4.w;
          *reinterpret_cast<float4*>(stg + (size_t)row * H + ((c4 ^ (row & 7)) << 2)) = v;
        }
      }
      tc_fence_before_sync();
      asm volatile("bar.sync 2, %0;" ::"n"(Cfg::N_EPI) : "memory");
      // ---- cross-view reduce: one warp per token, lane owns CPL consecutive channels
      constexpr int CPL = H / 32;                  // 2 or 4 channels per lane
      for (int t = ew; t < ntok; t += Cfg::N_EPI / 32) {
        auto load_row = [&](int rr, float (&v)[CPL]) {
          const float* src = stg + (size_t)rr * H;
          if constexpr (CPL == 4) {
            const float4 f = *reinterpret_cast<const float4*>(src + ((lane ^ (rr & 7)) << 2));
            v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
          } else {
            const int c4 = (lane >> 1) ^ (rr & 7);
            const float2 f = *reinterpret_cast<const float2*>(src + (c4 << 2) + ((lane & 1) << 1));
            v[0] = f.x, v[1] = f.y;
          }
        };
        float m0[CPL], acc[CPL];
        load_row(t * N, m0);
#pragma unroll
        for (int i = 0; i < CPL; ++i) acc[i] = (N == 1) ? m0[i] : 0.f;
        int n = 1;
        for (; n + 1 < N; n += 2) {                 // two views per step: the two shuffle reductions interleave
          float ma[CPL], mb[CPL];
          load_row(t * N + n, ma);
          load_row(t * N + n + 1, mb);
          float da = 0.f, db = 0.f;
#pragma unroll
          for (int i = 0; i < CPL; ++i) da += ma[i] * m0[i], db += mb[i] * m0[i];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            da += __shfl_xor_sync(0xffffffffu, da, o);
            db += __shfl_xor_sync(0xffffffffu, db, o);
          }
#pragma unroll
          for (int i = 0; i < CPL; ++i) acc[i] += da * ma[i], acc[i] += db * mb[i];
        }
        if (n < N) {
          float mv[CPL];
          load_row(t * N + n, mv);
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < CPL; ++i) dot += mv[i] * m0[i];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
          for (int i = 0; i < CPL; ++i) acc[i] += dot * mv[i];
        }
        // s is cubic in the activations: stored as s / sigma, sigma = 2^floor(log2 max|s|) (see merge_reduce_kernel)
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) mx = fmaxf(mx, fabsf(acc[i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const uint32_t ebits = __float_as_uint(mx) & 0x7f800000u;
        const float sig = (ebits == 0u || ebits == 0x7f800000u) ? 1.0f : __uint_as_float(ebits);
        const float inv = 1.0f / sig;
        const size_t tok = (size_t)cur.b * SM_P + tok0 + t;
        if (lane == 0) p.sigma[tok] = sig;
        uint32_t* out = reinterpret_cast<uint32_t*>(p.s + tok * H) + lane * (CPL / 2);
#pragma unroll
        for (int i = 0; i < CPL / 2; ++i) out[i] = pack_op16x2(acc[2 * i] * inv, acc[2 * i + 1] * inv);
      }
      mbar_arrive(&x_free[buf]);     // the sampler may overwrite this X buffer (its reads above are done)
    }
  } else {
    // ===================== sampler warps (10..17) =====================
    const int stid = threadIdx.x - (64 + Cfg::N_EPI);   // 0..255
    SmTileCursor cur;
    // slab loader: thread = (pixel px, slot group of 4): 4-byte cp.async per plane value, transposed into pixel-major
    auto load_slab = [&](int tile, SmTileCursor& c, int sbuf) {
      c.seek(p, tile);
      const int N = c.n_views;
      const int TOK = 128 / N;
      const int r0 = (tile - c.lo) * TOK * N;
      const int L0 = r0 / G;
      const int planes = N * D;                        // planes of this sample
      const float* base = p.xmap + ((size_t)c.img0 * D) * SM_F;
      float* dst = s_slab + sbuf * (Cfg::SLAB_BYTES / 4);
      for (int e = stid; e < SM_F * SLOTS; e += Cfg::N_SAMPLER) {
        const int slot = e / SM_F, px = e - slot * SM_F;
        const int L = L0 + slot;
        if (slot <= RPI && L < planes) {
          const uint32_t d = smem_u32(dst + px * SLOTS + slot);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(base + (size_t)L * SM_F + px) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    SmTileCursor cur_next;
    if ((int)blockIdx.x < p.n_tiles) load_slab(blockIdx.x, cur_next, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      uint8_t* xb = s_x + buf * Cfg::X_BYTES;
      const float* slab = s_slab + buf * (Cfg::SLAB_BYTES / 4);
      cur.seek(p, tile);
      const int b = cur.b;
      const int N = cur.n_views;
      const int TOK = 128 / N;
      const int tok0 = (tile - cur.lo) * TOK;
      const int r0 = tok0 * N;                           // first row of the tile inside the sample (a multiple of N)
      const int rows_total = N * SM_P;
      const int Rb = (r0 / G) * G;                       // aligned base: item j owns rows Rb + j + G q
      const int L0 = Rb / G;
      const int QN = (Rb == r0) ? RPI : RPI + 1;
      const int img0 = cur.img0;
      // this tile's slab has landed (issued one tile ago); everyone has finished reading the other slab -> prefetch
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      // token of every tile row that is the first row of a token (q1 of merge_features_mv), -1 otherwise:
      // r = r0 + i with r0 a multiple of N, so r % N == i % N and r / N == tok0 + i / N  (one division per row and tile)
      if (stid < 128) {
        const int qd = stid / N;
        s_tok[buf * 128 + stid] = (stid - qd * N == 0 && r0 + stid < rows_total) ? tok0 + qd : -1;
      }
      asm volatile("bar.sync 3, %0;" ::"n"(Cfg::N_SAMPLER) : "memory");
      if (tile + (int)gridDim.x < p.n_tiles) load_slab(tile + gridDim.x, cur_next, buf ^ 1);
      // the X buffer must have been released by the epilogue of tile it - 2
      if (it >= 2) mbar_wait(&x_free[buf], (uint32_t)((it - 2) >> 1) & 1);
      const int* tok_of = s_tok + buf * 128;
      const bool fast = (Rb == r0) && (L0 % D) + RPI <= D && r0 + 128 <= rows_total;
      if (fast) {
        // ---- aligned tile (always the case when N divides 128): one view, slots 0 .. RPI-1, every row valid.
        // Work item = (8-point group kc, row residue j): lane = (j & 3) * 8 + (kc & 7), so a warp's 8-byte A-tile
        // stores cover four consecutive 128-byte rows (conflict-free); rows of one item are 16 apart, so the
        // swizzle XOR (row & 7) is the same for all of them and the stores differ by immediate offsets.
        const int n = L0 / D;
        const int lane_s = stid & 31, wv = stid >> 5;
        for (int wi = wv; wi < (G / 4) * (KC / 8); wi += Cfg::N_SAMPLER / 32) {
          const int j = (wi % (G / 4)) * 4 + (lane_s >> 3);
          const int kc = (wi / (G / 4)) * 8 + (lane_s & 7);
          const uint32_t* tp = p.taps + ((size_t)(img0 + n) * (SM_P / 4) + (size_t)(j * D + kc * 8) / 4) * SM_TAP_WORDS;
          uint8_t* arow = xb + (kc >> 3) * (128 * 128) + sw128_offset(j, kc & 7);    // row j + G q: + q * G * 128 bytes
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {               // points kc*8 + 4*hf .. + 3
            uint4 T[6];
#pragma unroll
            for (int w = 0; w < 6; ++w) T[w] = __ldg(reinterpret_cast<const uint4*>(tp + hf * SM_TAP_WORDS) + w);
            const float w00[4] = {__uint_as_float(T[0].x), __uint_as_float(T[0].y), __uint_as_float(T[0].z), __uint_as_float(T[0].w)};
            const float w01[4] = {__uint_as_float(T[1].x), __uint_as_float(T[1].y), __uint_as_float(T[1].z), __uint_as_float(T[1].w)};
            const float w10[4] = {__uint_as_float(T[2].x), __uint_as_float(T[2].y), __uint_as_float(T[2].z), __uint_as_float(T[2].w)};
            const float w11[4] = {__uint_as_float(T[3].x), __uint_as_float(T[3].y), __uint_as_float(T[3].z), __uint_as_float(T[3].w)};
            const uint32_t oa[4] = {T[4].x, T[4].z, T[5].x, T[5].z};     // nw | ne byte offsets of points 0..3
            const uint32_t ob[4] = {T[4].y, T[4].w, T[5].y, T[5].w};     // sw | se
            const char* sl8 = reinterpret_cast<const char*>(slab);
#pragma unroll
            for (int g4 = 0; g4 < RPI / 4; ++g4) {
              float v[4][4];                             // [slot][point]
#pragma unroll
              for (int pt = 0; pt < 4; ++pt) {
                // ATen accumulates the four taps in the order nw, ne, sw, se
                const float4 a = *reinterpret_cast<const float4*>(sl8 + (oa[pt] & 0xffffu) + 16 * g4);
                const float4 bq = *reinterpret_cast<const float4*>(sl8 + (oa[pt] >> 16) + 16 * g4);
                const float4 c = *reinterpret_cast<const float4*>(sl8 + (ob[pt] & 0xffffu) + 16 * g4);
                const float4 d = *reinterpret_cast<const float4*>(sl8 + (ob[pt] >> 16) + 16 * g4);
                float t;
                t = a.x * w00[pt], t += bq.x * w01[pt], t += c.x * w10[pt], t += d.x * w11[pt], v[0][pt] = t;
                t = a.y * w00[pt], t += bq.y * w01[pt], t += c.y * w10[pt], t += d.y * w11[pt], v[1][pt] = t;
                t = a.z * w00[pt], t += bq.z * w01[pt], t += c.z * w10[pt], t += d.z * w11[pt], v[2][pt] = t;
                t = a.w * w00[pt], t += bq.w * w01[pt], t += c.w * w10[pt], t += d.w * w11[pt], v[3][pt] = t;
              }
#pragma unroll
              for (int sl = 0; sl < 4; ++sl) {
                const int q = 4 * g4 + sl;
                const uint2 val = make_uint2(pack_op16x2(v[sl][0], v[sl][1]), pack_op16x2(v[sl][2], v[sl][3]));
                *reinterpret_cast<uint2*>(arow + q * (G * 128) + hf * 8) = val;
                const int tk = tok_of[j + G * q];
                if (tk >= 0) *reinterpret_cast<uint2*>(p.q1 + ((size_t)b * SM_P + tk) * D + kc * 8 + hf * 4) = val;
              }
            }
          }
        }
      } else {
      // ---- generic tile (view counts that do not divide 128): the tile may start inside a channel plane and straddle
      //      two views; slot groups of 4 are sampled per view and only the rows that belong to it are stored
      for (int item = stid; item < G * KC; item += Cfg::N_SAMPLER) {
        const int j = item / KC, kc = item - j * KC;     // chunk j (points j*D ..), 16-byte piece kc of the row
        // (the last tile of a sample may reach past its last view: those rows do not exist)
        const int n_first = L0 / D, n_last = min((L0 + QN - 1) / D, N - 1);
        for (int n = n_first; n <= n_last; ++n) {
          const int q_lo = max(0, n * D - L0), q_hi = min(QN, (n + 1) * D - L0);
          const uint32_t* tp = p.taps + ((size_t)(img0 + n) * (SM_P / 4) + (size_t)(j * D + kc * 8) / 4) * SM_TAP_WORDS;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {               // points kc*8 + 4*hf .. + 3
            const uint4* t4 = reinterpret_cast<const uint4*>(tp + hf * SM_TAP_WORDS);
            const uint4 W00 = __ldg(t4), W01 = __ldg(t4 + 1), W10 = __ldg(t4 + 2), W11 = __ldg(t4 + 3);
            const uint4 OA = __ldg(t4 + 4), OB = __ldg(t4 + 5);
            const float w00[4] = {__uint_as_float(W00.x), __uint_as_float(W00.y), __uint_as_float(W00.z), __uint_as_float(W00.w)};
            const float w01[4] = {__uint_as_float(W01.x), __uint_as_float(W01.y), __uint_as_float(W01.z), __uint_as_float(W01.w)};
            const float w10[4] = {__uint_as_float(W10.x), __uint_as_float(W10.y), __uint_as_float(W10.z), __uint_as_float(W10.w)};
            const float w11[4] = {__uint_as_float(W11.x), __uint_as_float(W11.y), __uint_as_float(W11.z), __uint_as_float(W11.w)};
            const uint32_t oa[4] = {OA.x, OA.z, OB.x, OB.z};     // nw | ne byte offsets into the slab
            const uint32_t ob[4] = {OA.y, OA.w, OB.y, OB.w};     // sw | se
            const char* sl8 = reinterpret_cast<const char*>(slab);
            for (int g4 = q_lo / 4; 4 * g4 < q_hi; ++g4) {
              float v[4][4];                             // [slot][point]
#pragma unroll
              for (int pt = 0; pt < 4; ++pt) {
                // ATen accumulates the four taps in the order nw, ne, sw, se
                const float4 a = *reinterpret_cast<const float4*>(sl8 + (oa[pt] & 0xffffu) + 16 * g4);
                const float4 bq = *reinterpret_cast<const float4*>(sl8 + (oa[pt] >> 16) + 16 * g4);
                const float4 c = *reinterpret_cast<const float4*>(sl8 + (ob[pt] & 0xffffu) + 16 * g4);
                const float4 d = *reinterpret_cast<const float4*>(sl8 + (ob[pt] >> 16) + 16 * g4);
                float t;
                t = a.x * w00[pt], t += bq.x * w01[pt], t += c.x * w10[pt], t += d.x * w11[pt], v[0][pt] = t;
                t = a.y * w00[pt], t += bq.y * w01[pt], t += c.y * w10[pt], t += d.y * w11[pt], v[1][pt] = t;
                t = a.z * w00[pt], t += bq.z * w01[pt], t += c.z * w10[pt], t += d.z * w11[pt], v[2][pt] = t;
                t = a.w * w00[pt], t += bq.w * w01[pt], t += c.w * w10[pt], t += d.w * w11[pt], v[3][pt] = t;
              }
#pragma unroll
              for (int sl = 0; sl < 4; ++sl) {
                const int q = 4 * g4 + sl;
                const int r = Rb + j + G * q;            // row inside the sample
                const int i = r - r0;                    // row inside the tile
                if (q >= q_lo && q < q_hi && i >= 0 && i < 128 && r < rows_total) {
                  uint2 pk;
                  pk.x = pack_op16x2(v[sl][0], v[sl][1]);
                  pk.y = pack_op16x2(v[sl][2], v[sl][3]);
                  *reinterpret_cast<uint2*>(xb + (kc >> 3) * (128 * 128) + sw128_offset(i, kc & 7) + hf * 8) = pk;
                  const int tk = tok_of[i];              // token-first row: q1 of merge_features_mv / q of _sv
                  if (tk >= 0) *reinterpret_cast<uint2*>(p.q1 + ((size_t)b * SM_P + tk) * D + kc * 8 + hf * 4) = pk;
                }
              }
            }
          }
        }
      }
      }
      fence_proxy_async_smem();      // A tile visible to the tensor-core (async) proxy
      mbar_arrive(&a_full[buf]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace poem
