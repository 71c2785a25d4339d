// Query-stream chain kernels: the small Linear layers between the attention kernels of a decoder block, fused per
// segment so that a 128-row tile of the (B*799, D) query stream goes through two (or, for the feed-forward, three)
// layers without leaving the SM.
//
//   chain2_kernel<D>:  C1 = A·W1^T + b1 (+ fp32 residual) [-> LayerNorm]          -> fp32 rows (residual stream) and/or fp16 rows
//                      C2 = act(fp16(C1)·W2^T + b2),  N2 = D, 2D, 3D ...           -> fp16 rows
//   reference call sites (one launch each in the reference, two or three kernels each in round 1):
//     embedding + attn.self.query                       pt_metro_transformer.py:180, HF BertSelfAttention
//     attn.output.dense + residual + LayerNorm + cross_attn.self.query              pt_metro_transformer.py:57-74
//     cross_attn.output.dense + residual + LayerNorm + (fc1 ∘ w_qs | w_ks | w_vs)   point_transformers.py:86-88
//     query_self_attn.fc2 + residual + query_cross_attn.w_qs                        point_transformers.py:95,139
//     query_cross_attn.fc2 + residual + reg_branch.0 + ReLU                         point_transformers.py:151, pt_metro_transformer.py:34-40
//   ffn_kernel<D>:     out = LayerNorm(f + W2·gelu(W1·f + b1) + b2)  with the 4D-wide hidden layer produced and consumed
//                      in 256-column chunks (BertIntermediate / BertOutput, pt_metro_transformer.py:86-90)
//
// Round 1 ran every layer as its own persistent GEMM (+ a LayerNorm kernel): 19 launches of 15-70 us per block for
// ~0.1 ms worth of HBM traffic per step: each launch pays its prologue, pipeline fill and tail on only ~1.35 tiles per
// CTA.  Here the second layer's A operand is the first layer's epilogue output in shared memory (K-major SWIZZLE_128B,
// like every other tile), weights stream through a TMA ring from L2, accumulators ping-pong between two TMEM regions.
// warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue (lane = row, two warps per lane quarter)
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm.cuh"

namespace poem {

template <int D>
struct QcCfg {
  static_assert(D == 128 || D == 256, "query-stream chain kernels: D = 128 or 256 (D = 512 keeps the separate GEMMs)");
  static constexpr int KB = D / 64;                        // K blocks of a D-wide operand
  static constexpr int NC = D;                             // MMA N = columns per TMEM region (stage 1 width, stage-2 chunk)
  static constexpr int A_BYTES = 128 * D * 2;              // A tile; the H tile (stage-1 output) is written over it
  static constexpr int W_TILE_BYTES = NC * 64 * 2;         // [NC rows x 64 k]
  static constexpr int W_STAGES = (D == 256) ? 4 : 6;
  static constexpr int N_EPI = 256;
  static constexpr int THREADS = 64 + N_EPI;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_H = 0;                          // alias: stage-1 MMAs have retired before the epilogue writes H
  static constexpr int OFF_W = A_BYTES;
  static constexpr int OFF_PAR = OFF_W + W_STAGES * W_TILE_BYTES;     // b1 | ln_g | ln_b  (3 x D floats), then b2 (up to 4D)
  static constexpr int OFF_XCH = OFF_PAR + (3 * D + 4 * D) * 4;       // LayerNorm partial-sum exchange: 2 x 128 float2
  static constexpr int OFF_BARS = OFF_XCH + 2 * 128 * 8;
  static constexpr int SMEM_BYTES = OFF_BARS + 256;
  static constexpr int TMEM_COLS = (2 * NC <= 256) ? 256 : 512;
};

struct Chain2Params {
  int M;                    // rows
  // stage 1
  const float* b1;          // [D]
  const float* res32;       // [M, D] fp32 residual or nullptr
  int ln;                   // LayerNorm after the residual (eps 1e-12, biased variance)
  const float* ln_g;
  const float* ln_b;
  float* out1_f32;          // [M, D] or nullptr
  op16* out1_h16;           // [M, D] or nullptr (the smem copy feeds stage 2 regardless)
  // stage 2
  int N2;                   // multiple of D
  const float* b2;          // [N2] or nullptr
  int act2;                 // GemmAct
  op16* out2;               // [M, ld2]
  int ld2;
};

// region r of the TMEM allocation: r = 1 holds the stage-1 accumulators and the odd stage-2 chunks, r = 0 the even chunks
template <int D>
__global__ void __launch_bounds__(QcCfg<D>::THREADS, 1)
chain2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w1,
              const __grid_constant__ CUtensorMap tmap_w2, Chain2Params p) {
  using Cfg = QcCfg<D>;
  constexpr int KB = Cfg::KB, NC = Cfg::NC;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_a = smem + Cfg::OFF_A;
  uint8_t* s_h = smem + Cfg::OFF_H;
  uint8_t* s_w = smem + Cfg::OFF_W;
  float* s_b1 = reinterpret_cast<float*>(smem + Cfg::OFF_PAR);
  float* s_g = s_b1 + D;
  float* s_be = s_g + D;
  float* s_b2 = s_be + D;
  float2* s_xch = reinterpret_cast<float2*>(smem + Cfg::OFF_XCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* w_full = bars;                         // [W_STAGES]
  uint64_t* w_empty = bars + Cfg::W_STAGES;        // [W_STAGES]
  uint64_t* a_full = bars + 2 * Cfg::W_STAGES;     // TMA -> MMA: A tile landed
  uint64_t* a_empty = a_full + 1;                  // MMA -> TMA: every MMA of the tile has read the A / H buffer
  uint64_t* h_full = a_empty + 1;                  // epilogue -> MMA: H tile written (count N_EPI)
  uint64_t* r_full = h_full + 1;                   // [2] MMA -> epilogue: region r holds a finished accumulator
  uint64_t* r_free = r_full + 2;                   // [2] epilogue -> MMA: region r drained (count N_EPI)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.M + 127) / 128;
  const int n_chunks = p.N2 / NC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::W_STAGES; ++s) {
        mbar_init(&w_full[s], 1);
        mbar_init(&w_empty[s], 1);
      }
      mbar_init(a_full, 1);
      mbar_init(a_empty, 1);
      mbar_init(h_full, Cfg::N_EPI);
      for (int r = 0; r < 2; ++r) {
        mbar_init(&r_full[r], 1);
        mbar_init(&r_free[r], Cfg::N_EPI);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  for (int c = threadIdx.x; c < D; c += Cfg::THREADS) {
    s_b1[c] = p.b1 ? p.b1[c] : 0.f;
    s_g[c] = p.ln ? p.ln_g[c] : 1.f;
    s_be[c] = p.ln ? p.ln_b[c] : 0.f;
  }
  for (int c = threadIdx.x; c < p.N2; c += Cfg::THREADS) s_b2[c] = p.b2 ? p.b2[c] : 0.f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: only the A operand / the residual depend on the predecessor kernel.  The TMA warp
  // first fills the weight ring (weights are constants) and waits afterwards; everybody else waits here.
  if (warp != 0) pdl_wait();
  pdl_trigger();
  auto region = [&](int r) { return tmem_base + (uint32_t)(r ? 0 : NC); };   // r = 1 -> columns [0, NC), r = 0 -> [NC, 2 NC)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      const int w_per_tile = KB * (1 + n_chunks);
      auto issue_w = [&](int idx) {                 // weight stage idx of a tile: job = idx / KB, K block = idx % KB
        const int job = idx / KB, kb = idx - job * KB;
        mbar_wait(&w_empty[stage], phase ^ 1);
        mbar_expect_tx(&w_full[stage], Cfg::W_TILE_BYTES);
        tma_load_2d(s_w + stage * Cfg::W_TILE_BYTES, job == 0 ? &tmap_w1 : &tmap_w2, &w_full[stage], kb * 64,
                    job == 0 ? 0 : (job - 1) * NC);
        if (++stage == Cfg::W_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      bool first = true;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int w0 = 0;
        if (first) {   // ring is empty: its first W_STAGES weight tiles travel while the predecessor kernel drains
          for (; w0 < Cfg::W_STAGES && w0 < w_per_tile; ++w0) issue_w(w0);
          pdl_wait();
          first = false;
        }
        mbar_wait(a_empty, a_phase ^ 1);
        a_phase ^= 1;
        mbar_expect_tx(a_full, Cfg::A_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(s_a + kb * (128 * 128), &tmap_a, a_full, kb * 64, tile * 128);
        for (int i = w0; i < w_per_tile; ++i) issue_w(i);
      }
      if (first) pdl_wait();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_op16(128, NC);
      int stage = 0;
      uint32_t phase = 0, a_phase = 0, h_phase = 0;
      uint32_t free_phase[2] = {0, 0};
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int job = 0; job < 1 + n_chunks; ++job) {
          const int r = (job == 0) ? 1 : ((job - 1) & 1);
          if (job == 0) {
            mbar_wait(a_full, a_phase);
            a_phase ^= 1;
          } else if (job == 1) {
            mbar_wait(h_full, h_phase);
            h_phase ^= 1;
          }
          mbar_wait(&r_free[r], free_phase[r] ^ 1);     // the previous accumulator of this region has been drained
          free_phase[r] ^= 1;
          tc_fence_after_sync();
          const uint32_t src = smem_u32(job == 0 ? s_a : s_h);
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&w_full[stage], phase);
            tc_fence_after_sync();
            const uint64_t da = make_kmajor_desc<128>(src + kb * (128 * 128));
            const uint64_t dw = make_kmajor_desc<128>(smem_u32(s_w) + stage * Cfg::W_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_op16(region(r), da + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
            umma_commit(&w_empty[stage]);
            if (++stage == Cfg::W_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit(&r_full[r]);
          if (job == n_chunks) umma_commit(a_empty);     // the A / H buffer may take the next tile's load
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                  // TMEM lane quarter of this warp
    const int colh = ew >> 2;                      // column half of a region this thread owns
    const int row = quarter * 32 + lane;           // tile row == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    constexpr int HC = NC / 2;                     // columns per thread and region
    uint32_t full_phase[2] = {0, 0};
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m = tile * 128 + row;
      const bool live = m < p.M;
      // ---------------- stage-1 epilogue: bias, residual, LayerNorm -> fp32 / fp16 rows + the H tile ----------------
      // The fp32 residual row goes into the accumulator registers FIRST: the loads (one 32-byte sector per lane and
      // instruction, a row per lane) travel while the A tile is loaded and the stage-1 MMAs run — issued behind the
      // accumulator wait they were 24 % of all warp samples (long scoreboard, ncu round 2)
      float v[HC];
      if (p.res32 != nullptr && live) {
        const float* rp = p.res32 + (size_t)m * D + colh * HC;
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 8) ldg_nc_256(rp + c0, reinterpret_cast<uint32_t*>(&v[c0]));
      } else {
#pragma unroll
        for (int i = 0; i < HC; ++i) v[i] = 0.f;
      }
      mbar_wait(&r_full[1], full_phase[1]);
      full_phase[1] ^= 1;
      tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(region(1) + lane_off + colh * HC + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[c0 + i] += __uint_as_float(r[i]) + s_b1[colh * HC + c0 + i];
      }
      tc_fence_before_sync();
      mbar_arrive(&r_free[1]);                      // accumulators are in registers: the region can take the next job
      if (p.ln) {
        // two-pass statistics over the whole row: this thread's half + the partner warp's half (smem exchange)
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < HC; ++i) s += v[i];
        s_xch[colh * 128 + row].x = s;
        pair_sync();
        const float mean = (s + s_xch[(colh ^ 1) * 128 + row].x) * (1.0f / (float)D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < HC; ++i) {
          const float d = v[i] - mean;
          q += d * d;
        }
        s_xch[colh * 128 + row].y = q;
        pair_sync();
        const float rstd = rsqrtf((q + s_xch[(colh ^ 1) * 128 + row].y) * (1.0f / (float)D) + 1e-12f);
#pragma unroll
        for (int i = 0; i < HC; ++i) v[i] = (v[i] - mean) * rstd * s_g[colh * HC + i] + s_be[colh * HC + i];
        pair_sync();   // the exchange slots are reused by the next tile
      }
      {
        // H tile (A operand of stage 2), K-major SWIZZLE_128B; fp16 / fp32 copies to HBM
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 8) {
          const int c = colh * HC + c0;
          uint4 pk;
          pk.x = pack_op16x2(v[c0 + 0], v[c0 + 1]);
          pk.y = pack_op16x2(v[c0 + 2], v[c0 + 3]);
          pk.z = pack_op16x2(v[c0 + 4], v[c0 + 5]);
          pk.w = pack_op16x2(v[c0 + 6], v[c0 + 7]);
          *reinterpret_cast<uint4*>(s_h + (c >> 6) * (128 * 128) + sw128_offset(row, (uint32_t)(c & 63) >> 3)) = pk;
          if (p.out1_h16 != nullptr && live) *reinterpret_cast<uint4*>(p.out1_h16 + (size_t)m * D + c) = pk;
        }
        if (p.out1_f32 != nullptr && live) {
          float* o = p.out1_f32 + (size_t)m * D + colh * HC;
#pragma unroll
          for (int c0 = 0; c0 < HC; c0 += 8) stg_256(o + c0, reinterpret_cast<const uint32_t*>(&v[c0]));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(h_full);
      // ---------------- stage-2 epilogues: bias, activation -> fp16 rows ----------------
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int r = ch & 1;
        mbar_wait(&r_full[r], full_phase[r]);
        full_phase[r] ^= 1;
        tc_fence_after_sync();
#pragma unroll
        for (int c0 = 0; c0 < HC; c0 += 32) {
          const int n = ch * NC + colh * HC + c0;
          uint32_t rr[32];
          tmem_ld32(region(r) + lane_off + colh * HC + c0, rr);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0 = __uint_as_float(rr[2 * i]) + s_b2[n + 2 * i];
            float x1 = __uint_as_float(rr[2 * i + 1]) + s_b2[n + 2 * i + 1];
            if (p.act2 == ACT_RELU) x0 = fmaxf(x0, 0.f), x1 = fmaxf(x1, 0.f);
            else if (p.act2 == ACT_GELU) x0 = gelu_erf(x0), x1 = gelu_erf(x1);
            pk[i] = pack_op16x2(x0, x1);
          }
          if (live) {
            op16* o = p.out2 + (size_t)m * p.ld2 + n;
            stg_256(o, &pk[0]);
            stg_256(o + 16, &pk[8]);
          }
        }
        tc_fence_before_sync();
        mbar_arrive(&r_free[r]);
      }
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
