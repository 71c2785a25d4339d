// Cross-view multi-head attention, 799 queries -> 4096 BPS tokens, no mask (reference: HF BertSelfAttention
// reached from lib/models/bricks/pt_metro_transformer.py:57-72; scores (B,h,799,4096) never materialised here).
//
// One CTA = 128 queries of one (sample, head). Per 128-key block:
//   S = Q·K^T   tcgen05.mma (M=128, N=128, K=HD)  -> TMEM
//   softmax warps (thread == query row): running max / sum in fp32, P = exp2(...) written as bf16 into a
//   SWIZZLE_128B K-major smem tile, then O_blk = P·V (M=128, N=HD, K=128) -> TMEM, accumulated in registers
//   with the usual online-softmax rescale.
// Q/K tiles [rows x HD] and V^T tiles [HD x 64 keys] are staged by TMA; K/V are double buffered.
// warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = softmax + output.
#pragma once
#include "common.cuh"

namespace poem {

constexpr int MHA_BQ = 128;    // queries per CTA
constexpr int MHA_BKEY = 128;  // keys per block
constexpr int MHA_THREADS = 192;

template <int HD>
struct MhaCfg {
  static constexpr int kRowBytes = (HD >= 64) ? 128 : 64;       // swizzle atom width of Q/K tiles
  static constexpr int kKBlocks = (HD * 2) / kRowBytes;          // 64-element K blocks of the head dim (1 or 2)
  static constexpr int kQBytes = MHA_BQ * HD * 2;
  static constexpr int kKBytes = MHA_BKEY * HD * 2;
  static constexpr int kVBytes = HD * MHA_BKEY * 2;              // two [HD x 64 keys] SWIZZLE_128B tiles
  static constexpr int kPBytes = MHA_BQ * MHA_BKEY * 2;          // two [128 x 64 keys] SWIZZLE_128B tiles
  static constexpr int kStages = 2;
  static constexpr int kSmemBytes = kQBytes + kStages * (kKBytes + kVBytes) + kPBytes + 256;
  static constexpr int kTmemCols = (128 + HD <= 256) ? 256 : 512;  // S: 128 columns, O_blk: HD columns
};

template <int HD>
__global__ void __launch_bounds__(MHA_THREADS, (HD <= 64) ? 2 : 1)
mha_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,   // [B*Lq rows, ldq] bf16, box [HD(or 64) x 128]
                  const __grid_constant__ CUtensorMap tmap_k,   // [B*Lk rows, ldk] bf16, box [HD(or 64) x 128]
                  const __grid_constant__ CUtensorMap tmap_vt,  // [B*vt_batch_rows, Lk] bf16, box [64 keys x HD]
                  __nv_bfloat16* __restrict__ ctx, int ld_ctx, int Lq, int Lk, int vt_batch_rows, int q_col0,
                  int k_col0, int vt_row0, float scale_log2e) {
  using Cfg = MhaCfg<HD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::kQBytes;                               // stage s: K at s*(K+V), V right after
  uint8_t* sP = sKV + Cfg::kStages * (Cfg::kKBytes + Cfg::kVBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::kPBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * MHA_BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kblocks = Lk / MHA_BKEY;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_vt);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&kv_full[s], 1);
        mbar_init(&kv_empty[s], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(p_full, 128);
      mbar_init(o_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::kQBytes);
      for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
        tma_load_2d(sQ + kb * (MHA_BQ * Cfg::kRowBytes), &tmap_q, q_full, q_col0 + head * HD + kb * 64, b * Lq + q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kblocks; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sK = sKV + stage * (Cfg::kKBytes + Cfg::kVBytes);
        uint8_t* sV = sK + Cfg::kKBytes;
        mbar_expect_tx(&kv_full[stage], Cfg::kKBytes + Cfg::kVBytes);
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
          tma_load_2d(sK + kb * (MHA_BKEY * Cfg::kRowBytes), &tmap_k, &kv_full[stage],
                      k_col0 + head * HD + kb * 64, b * Lk + j * MHA_BKEY);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(sV + kb * (HD * 128), &tmap_vt, &kv_full[stage], j * MHA_BKEY + kb * 64,
                      vt_row0 + b * vt_batch_rows + head * HD);
        if (++stage == 2) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(MHA_BQ, MHA_BKEY);
      constexpr uint32_t idesc_o = make_idesc_bf16(MHA_BQ, HD);
      auto issue_S = [&](int stage) {
        const uint32_t sK = smem_u32(sKV + stage * (Cfg::kKBytes + Cfg::kVBytes));
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb) {
          const uint64_t dq = make_kmajor_desc<Cfg::kRowBytes>(smem_u32(sQ) + kb * (MHA_BQ * Cfg::kRowBytes));
          const uint64_t dk = make_kmajor_desc<Cfg::kRowBytes>(sK + kb * (MHA_BKEY * Cfg::kRowBytes));
#pragma unroll
          for (int k = 0; k < Cfg::kRowBytes / 32; ++k) umma_bf16(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, (kb | k) != 0);
        }
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after_sync();
      issue_S(0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kblocks; ++j) {
        // P(j) is in smem and S(j) has been consumed
        mbar_wait(p_full, j & 1);
        tc_fence_after_sync();
        const uint32_t sV = smem_u32(sKV + stage * (Cfg::kKBytes + Cfg::kVBytes) + Cfg::kKBytes);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t dp = make_kmajor_desc<128>(smem_u32(sP) + kb * (MHA_BQ * 128));
          const uint64_t dv = make_kmajor_desc<128>(sV + kb * (HD * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_O, dp + 2 * k, dv + 2 * k, idesc_o, (kb | k) != 0);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[stage]);
        if (++stage == 2) {
          stage = 0;
          phase ^= 1;
        }
        if (j + 1 < n_kblocks) {
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after_sync();
          issue_S(stage);
        }
      }
    }
  } else {
    // ===================== softmax + output warps =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float o_acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) o_acc[c] = 0.f;
    float m_run = -INFINITY;   // running max of raw scores
    float l_run = 0.f;         // running sum of exp
    for (int j = 0; j < n_kblocks; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      // pass A: block max
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < MHA_BKEY; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_S + lane_off + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) m_blk = fmaxf(m_blk, __uint_as_float(r[i]));
      }
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = fast_exp2((m_run - m_new) * scale_log2e);   // 0 on the first block (m_run = -inf)
      const float m_scaled = m_new * scale_log2e;
      float l_blk = 0.f;
      // pass B: probabilities -> bf16 -> swizzled smem (A operand of P·V)
#pragma unroll 1
      for (int c0 = 0; c0 < MHA_BKEY; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_S + lane_off + c0, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = fast_exp2(__uint_as_float(r[2 * i]) * scale_log2e - m_scaled);
          const float p1 = fast_exp2(__uint_as_float(r[2 * i + 1]) * scale_log2e - m_scaled);
          l_blk += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        uint8_t* tile = sP + (c0 >> 6) * (MHA_BQ * 128);
        const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          *reinterpret_cast<uint4*>(tile + sw128_offset(row, chunk0 + q)) = v;
        }
      }
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      fence_proxy_async_smem();     // P visible to the tensor-core (async) proxy
      tc_fence_before_sync();       // our TMEM reads of S are done before the next S MMA may overwrite it
      mbar_arrive(p_full);
      // O_blk
      mbar_wait(o_full, j & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_O + lane_off + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c0 + i] = o_acc[c0 + i] * alpha + __uint_as_float(r[i]);
      }
      tc_fence_before_sync();
    }
    const int q = q0 + row;
    if (q < Lq) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* o = ctx + (size_t)(b * Lq + q) * ld_ctx + head * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 8) {
        uint4 pk;
        pk.x = pack_bf16x2(o_acc[c + 0] * inv, o_acc[c + 1] * inv);
        pk.y = pack_bf16x2(o_acc[c + 2] * inv, o_acc[c + 3] * inv);
        pk.z = pack_bf16x2(o_acc[c + 4] * inv, o_acc[c + 5] * inv);
        pk.w = pack_bf16x2(o_acc[c + 6] * inv, o_acc[c + 7] * inv);
        *reinterpret_cast<uint4*>(o + c) = pk;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace poem
