// Cross-view multi-head attention, 799 queries -> 4096 BPS tokens, no mask (reference: HF BertSelfAttention
// reached from lib/models/bricks/pt_metro_transformer.py:57-72; scores (B,h,799,4096) never materialised here).
//
// One CTA = 128 queries of one (sample, head). Per 128-key block j:
//   S(j) = Q·K(j)^T     tcgen05.mma (M=128, N=128, K=HD) -> TMEM buffer j%2          (double buffered)
//   softmax warps (thread == query row, no cross-lane traffic): pass A row max, pass B P = exp2(..) as op16 into a
//   SWIZZLE_128B K-major smem tile
//   O_blk(j) = P(j)·V(j) (M=128, N=HD, K=128) -> TMEM, written over the first HD columns of the ALREADY CONSUMED
//   S(j) buffer, so 256 TMEM columns hold two S buffers and the P·V result (2 CTAs per SM for HD <= 64)
//   O_blk(j) is folded into the register accumulator (online-softmax rescale) one iteration later, while the
//   tensor core already works on S(j+2) / P·V(j+1): the softmax warps never wait for an MMA in steady state.
// Q/K/V tiles [rows x HD] are staged by TMA as they lie in memory; V is consumed as an MN-major B operand (no
// transposed copy of the value projection is needed); K and V rings are 2 deep.
// warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = softmax + output (two threads per row).
#pragma once
#include "common.cuh"

namespace poem {

constexpr int MHA_BQ = 128;    // queries per CTA
constexpr int MHA_BKEY = 128;  // keys per block
constexpr int MHA_THREADS = 64 + 256;   // TMA warp, MMA warp, 8 softmax warps (two threads per query row)
#ifndef POEM_MHA_POLY_EVERY
#define POEM_MHA_POLY_EVERY 4           // 1 of every 4 pairs = 12.5 % of the exponentials on the FMA pipes (0 = all on the SFU).
                                        // Measured per launch (medium, B = 32): 0 -> 209.0 us, 12.5 % -> 201.6, 25 % -> 201.6, 50 % -> 223.7
#endif
constexpr int MHA_POLY_EVERY = POEM_MHA_POLY_EVERY;
#ifndef POEM_MHA_BUNDLE
#define POEM_MHA_BUNDLE 16
#endif
constexpr int MHA_BUNDLE = POEM_MHA_BUNDLE;   // (sample, head) groups scheduled together (1 MB of K / V each at head dim 64)

template <int HD>
struct MhaCfg {
  static constexpr int kRowBytes = (HD >= 64) ? 128 : 64;       // swizzle atom width of Q/K tiles
  static constexpr int kKBlocks = (HD * 2) / kRowBytes;          // 64-element K blocks of the head dim (1 or 2)
  static constexpr int kQBytes = MHA_BQ * HD * 2;
  static constexpr int kKBytes = MHA_BKEY * HD * 2;
  static constexpr int kVBytes = MHA_BKEY * HD * 2;              // kKBlocks tiles of [128 keys x kRowBytes]
  static constexpr int kPBytes = MHA_BQ * MHA_BKEY * 2;          // two [128 x 64 keys] SWIZZLE_128B tiles
  static constexpr int kXchgBytes = 2 * 128 * 2;                // bf16 row-max exchange between the two row halves (512 B: two CTAs of 113.6 KB still fit one SM)
  static constexpr int kSmemBytes = kQBytes + 2 * (kKBytes + kVBytes) + kPBytes + kXchgBytes + 160;   // 2 CTAs/SM at HD=64
  static constexpr int kTmemCols = 256;                          // two S buffers; O_blk aliases the consumed one
};

template <int HD>
__global__ void __launch_bounds__(MHA_THREADS, (HD <= 64) ? 2 : 1)
mha_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,   // [B*Lq rows, ldq] op16, box [HD(or 64) x 128]
                  const __grid_constant__ CUtensorMap tmap_k,   // [B*Lk rows, ldk] op16, box [HD(or 64) x 128]
                  const __grid_constant__ CUtensorMap tmap_v,   // [B*Lk rows, ldv] op16, box [HD(or 64) x 128]
                  op16* __restrict__ ctx, int ld_ctx, int Lq, int Lk, int q_col0, int k_col0, int v_col0,
                  float scale_log2e, int n_heads) {
  using Cfg = MhaCfg<HD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::kQBytes;               // 2 stages
  uint8_t* sV = sK + 2 * Cfg::kKBytes;           // 2 stages
  uint8_t* sP = sV + 2 * Cfg::kVBytes;
  __nv_bfloat16* s_xchg = reinterpret_cast<__nv_bfloat16*>(sP + Cfg::kPBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::kPBytes + Cfg::kXchgBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;       // [2] TMA -> MMA
  uint64_t* k_empty = bars + 3;      // [2] MMA -> TMA (S(j) retired)
  uint64_t* v_full = bars + 5;       // [2]
  uint64_t* v_empty = bars + 7;      // [2] (P·V(j) retired)
  uint64_t* s_full = bars + 9;       // [2] MMA -> softmax: S(j) in TMEM buffer j%2
  uint64_t* o_full = bars + 11;      // [2] MMA -> softmax: O_blk(j) in TMEM buffer j%2 (and P smem tile free)
  uint64_t* o_done = bars + 13;      // [2] softmax -> MMA: O_blk(j) folded, buffer j%2 may take S(j+2)   (count 256)
  uint64_t* p_full = bars + 15;      //     softmax -> MMA: P(j) in smem, S(j) consumed                   (count 256)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // 1-D grid over (sample, head) groups in bundles of MHA_BUNDLE: the bundle's full 128-query tiles first, then its
  // partial tiles (Lq = 799 leaves a 31-row tile whose idle softmax warps skip their work, so those CTAs are short).
  // Bundling keeps the K / V slices of a group in L2 between its full and partial tiles: with all partial tiles at
  // the end of the grid every K / V byte was fetched from HBM twice (ncu: 253 MB per launch for 134 MB of K / V).
  const int n_full = Lq / MHA_BQ;
  const int per_group = n_full + ((Lq % MHA_BQ) ? 1 : 0);
  const int n_groups = (int)gridDim.x / per_group;
  const int bundle0 = ((int)blockIdx.x / (MHA_BUNDLE * per_group)) * MHA_BUNDLE;   // first group of this bundle
  const int in_bundle = min(MHA_BUNDLE, n_groups - bundle0);                        // groups in it (last one may be short)
  const int r = (int)blockIdx.x - bundle0 * per_group;                              // index inside the bundle
  int tile, group;
  if (r < in_bundle * n_full) {
    tile = r % n_full;
    group = bundle0 + r / n_full;
  } else {
    tile = n_full;
    group = bundle0 + (r - in_bundle * n_full);
  }
  const int q0 = tile * MHA_BQ;
  const int head = group % n_heads;
  const int b = group / n_heads;
  const int n_kblocks = Lk / MHA_BKEY;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
        mbar_init(&s_full[s], 1);
        mbar_init(&o_full[s], 1);
        mbar_init(&o_done[s], 256);
      }
      mbar_init(p_full, 256);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;   // buffer i at columns [128*i, 128*i + 128)
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // elect.sync: ptxas knows the region is single-threaded (no per-instruction elect loop)
      mbar_expect_tx(q_full, Cfg::kQBytes);
      for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
        tma_load_2d(sQ + kb * (MHA_BQ * Cfg::kRowBytes), &tmap_q, q_full, q_col0 + head * HD + kb * 64, b * Lq + q0);
      for (int j = 0; j < n_kblocks; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::kKBytes);
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
          tma_load_2d(sK + st * Cfg::kKBytes + kb * (MHA_BKEY * Cfg::kRowBytes), &tmap_k, &k_full[st],
                      k_col0 + head * HD + kb * 64, b * Lk + j * MHA_BKEY);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], Cfg::kVBytes);
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
          tma_load_2d(sV + st * Cfg::kVBytes + kb * (MHA_BKEY * Cfg::kRowBytes), &tmap_v, &v_full[st],
                      v_col0 + head * HD + kb * 64, b * Lk + j * MHA_BKEY);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {   // elect.sync: ptxas knows the region is single-threaded (no per-instruction elect loop)
      constexpr uint32_t idesc_s = make_idesc_op16(MHA_BQ, MHA_BKEY);
      constexpr uint32_t idesc_o = make_idesc_op16(MHA_BQ, HD, /*b_mn_major=*/1);
      auto issue_S = [&](int j) {   // S(j) -> TMEM buffer j%2, from K stage j%2
        const int st = j & 1;
        mbar_wait(&k_full[st], (uint32_t)(j >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(sK + st * Cfg::kKBytes);
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb) {
          const uint64_t dq = make_kmajor_desc<Cfg::kRowBytes>(smem_u32(sQ) + kb * (MHA_BQ * Cfg::kRowBytes));
          const uint64_t dk = make_kmajor_desc<Cfg::kRowBytes>(k_addr + kb * (MHA_BKEY * Cfg::kRowBytes));
#pragma unroll
          for (int k = 0; k < Cfg::kRowBytes / 32; ++k)
            umma_op16(tmem_base + st * 128, dq + 2 * k, dk + 2 * k, idesc_s, (kb | k) != 0);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_S(0);
      if (n_kblocks > 1) issue_S(1);
      for (int j = 0; j < n_kblocks; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1;
        mbar_wait(p_full, (uint32_t)j & 1);          // P(j) in smem, S(j) consumed by every softmax thread
        mbar_wait(&v_full[st], ph);
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(sV + st * Cfg::kVBytes);
        // O_blk[128 x HD] = P[128 x 128 keys] (K-major) · V[128 keys x HD] (MN-major: each key row holds HD values)
        const uint64_t dv0 = make_mnmajor_desc<Cfg::kRowBytes>(v_addr, MHA_BKEY * Cfg::kRowBytes);
#pragma unroll
        for (int ks = 0; ks < MHA_BKEY / 16; ++ks) {     // 16 keys per MMA
          const uint64_t dp = make_kmajor_desc<128>(smem_u32(sP) + (ks >> 2) * (MHA_BQ * 128)) + 2 * (ks & 3);
          const uint64_t dv = dv0 + (uint64_t)((ks * 16 * Cfg::kRowBytes) >> 4);
          umma_op16(tmem_base + st * 128, dp, dv, idesc_o, ks != 0);
        }
        umma_commit(&o_full[st]);
        umma_commit(&v_empty[st]);
        if (j + 2 < n_kblocks) {
          mbar_wait(&o_done[st], ph);                // O_blk(j) has been read out of buffer st
          tc_fence_after_sync();
          issue_S(j + 2);
        }
      }
    }
  } else {
    // ===================== softmax + output warps =====================
    // 8 warps: two threads per query row (warps w and w+4 share a TMEM lane quarter); `half` selects which 64 of the
    // 128 S columns / which HD/2 of the O columns a thread owns.  The row maximum is exchanged through smem once per
    // key block, the row sums once at the end.
    constexpr int HH = HD / 2;
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    if (q0 + quarter * 32 >= Lq) {
      // No valid query row in this warp (and in its partner warp): keep the mbarrier protocol, skip the arithmetic.
      for (int j = 0; j < n_kblocks; ++j) {
        if (j > 0) mbar_arrive(&o_done[(j - 1) & 1]);
        mbar_arrive(p_full);
        mbar_wait(p_full, (uint32_t)j & 1);   // stay in step with the working warps: one arrival per thread and phase
      }
      mbar_arrive(&o_done[(n_kblocks - 1) & 1]);
    } else {
    // the two warps that share a lane quarter only ever exchange with each other: named barrier 1 + quarter, 64 threads
    // (a CTA-wide barrier here kept all eight warps in lock step once per key block)
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); };
    float o_acc[HH];
#pragma unroll
    for (int c = 0; c < HH; ++c) o_acc[c] = 0.f;
    float m_run = -INFINITY;   // running max of raw scores (whole row)
    float l_run = 0.f;         // running sum of exp over this thread's columns
    float alpha_prev = 0.f;    // rescale that belongs to the not-yet-folded O_blk(j-1)

    auto fold_o = [&](int j, float alpha) {         // o_acc = o_acc * alpha + O_blk(j)[:, half*HH .. +HH)
      const int st = j & 1;
      mbar_wait(&o_full[st], (uint32_t)(j >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t base = tmem_base + lane_off + st * 128 + half * HH;
      if constexpr (HH == 16) {
        uint32_t r[16];
        tmem_ld16(base, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[i] = o_acc[i] * alpha + __uint_as_float(r[i]);
      } else {
#pragma unroll
        for (int c0 = 0; c0 < HH; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(base + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c0 + i] = o_acc[c0 + i] * alpha + __uint_as_float(r[i]);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&o_done[st]);
    };

    // Per key block: pass A (row maximum), pass B (probabilities -> P tile), and the fold of the PREVIOUS block's P·V.
    // The fold sits in the MIDDLE of pass B: at the top of the iteration P·V(j-1) — issued only once the slowest thread
    // had finished pass B(j-1) — is usually still in flight (ncu round 2: 16 % of all warp samples waited on o_full
    // there).  The P tile is single-buffered and P·V(j-1) reads it until o_full, so the first half of the new
    // probabilities waits in 16 registers and is stored after the fold.
    for (int j = 0; j < n_kblocks; ++j) {
      const int st = j & 1;
      const uint32_t tmem_S = tmem_base + lane_off + st * 128 + half * 64;
      mbar_wait(&s_full[st], (uint32_t)(j >> 1) & 1);
      tc_fence_after_sync();
      // pass A: max over this thread's 64 columns, then over the row via the partner thread
      float m_blk = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_S + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) m_blk = fmaxf(m_blk, __uint_as_float(r[i]));
      }
      // Exchange the half-row maxima as bfloat16 (fp32 range, so a large raw score cannot overflow; any common reference
      // value works for the softmax and both threads of a row use the same rounded pair — the probabilities may exceed
      // 1 by the rounding, < 2^(2^-8 |m| scale), harmless in fp16).  Single buffer: the partner can only write its block
      // j+1 value after S(j+1) was issued, which (for j >= 1) waits for o_done(j-1), i.e. for every thread's fold below
      // (after its read here); S(0) and S(1) are issued up front, so block 0 syncs again.
      const __nv_bfloat16 m_mine = __float2bfloat16(m_blk);
      s_xchg[half * 128 + row] = m_mine;
      pair_sync();
      m_blk = fmaxf(__bfloat162float(m_mine), __bfloat162float(s_xchg[(half ^ 1) * 128 + row]));
      if (j == 0) pair_sync();
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = fast_exp2((m_run - m_new) * scale_log2e);   // 0 on the first block (m_run = -inf)
      const float m_scaled = m_new * scale_log2e;
      // pass B: probabilities -> fp16 -> swizzled smem (A operand of P·V); this thread fills K block `half`
      float l_blk = 0.f;
      uint8_t* tile = sP + half * (MHA_BQ * 128);
      auto pass_b32 = [&](int c0, uint32_t (&pk)[16]) {     // 32 columns -> 16 packed pairs
        uint32_t r[32];
        tmem_ld32(tmem_S + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = __uint_as_float(r[2 * i]) * scale_log2e - m_scaled;
          const float x1 = __uint_as_float(r[2 * i + 1]) * scale_log2e - m_scaled;
          const float p0 = fast_exp2(x0);
          // every MHA_POLY_EVERY-th pair takes its second exponential from the FMA pipes instead of the SFU
          const float p1 = (MHA_POLY_EVERY > 0 && (i % (MHA_POLY_EVERY > 0 ? MHA_POLY_EVERY : 1)) == 0) ? poly_exp2(x1) : fast_exp2(x1);
          l_blk += p0 + p1;
          pk[i] = pack_op16x2(p0, p1);
        }
      };
      auto store_p32 = [&](int c0, const uint32_t (&pk)[16]) {
        const int chunk0 = c0 >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(tile + sw128_offset(row, chunk0 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      };
      uint32_t pk0[16], pk1[16];
      pass_b32(0, pk0);
      // fold the previous block's P·V: by now it has had pass A + half of pass B to finish; also frees the P tile
      if (j > 0) fold_o(j - 1, alpha_prev);
      alpha_prev = alpha;
      store_p32(0, pk0);
      pass_b32(32, pk1);
      store_p32(32, pk1);
      l_run = l_run * alpha + l_blk;
      m_run = m_new;
      fence_proxy_async_smem();     // P visible to the tensor-core (async) proxy
      tc_fence_before_sync();       // our TMEM reads of S(j) are done before P·V(j) overwrites the buffer
      mbar_arrive(p_full);
    }
    fold_o(n_kblocks - 1, alpha_prev);
    // row sum = both halves
    // (every P·V has retired: the P tile is free and serves as the fp32 exchange buffer)
    float* lbuf = reinterpret_cast<float*>(sP);
    lbuf[half * 128 + row] = l_run;
    pair_sync();
    const float l_row = l_run + lbuf[(half ^ 1) * 128 + row];

    const int q = q0 + row;
    if (q < Lq) {
      const float inv = 1.0f / l_row;
      op16* o = ctx + (size_t)(b * Lq + q) * ld_ctx + head * HD + half * HH;
#pragma unroll
      for (int c = 0; c < HH; c += 8) {
        uint4 pk;
        pk.x = pack_op16x2(o_acc[c + 0] * inv, o_acc[c + 1] * inv);
        pk.y = pack_op16x2(o_acc[c + 2] * inv, o_acc[c + 3] * inv);
        pk.z = pack_op16x2(o_acc[c + 4] * inv, o_acc[c + 5] * inv);
        pk.w = pack_op16x2(o_acc[c + 6] * inv, o_acc[c + 7] * inv);
        *reinterpret_cast<uint4*>(o + c) = pk;
      }
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
