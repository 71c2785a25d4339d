// Fused Point-Transformer vector attention (reference lib/models/bricks/point_transformers.py:86-94, 139-150):
//
//   pos_ij = W_d2 relu(W_d1 (xyz_i - nbr_j) + b_d1) + b_d2
//   a_ij   = W_g2 relu(W_g1 (q_i - k_j + pos_ij) + b_g1) + b_g2
//   res_i  = sum_j softmax_j(a_ij / sqrt(D)) * (v_j + pos_ij)            (per channel, over the 32 neighbours)
//
// Algebraic form used here (exact in real arithmetic, folds done at pack time in fp64):
//   W_g1 (q_i - k_j + pos_ij) + b_g1 = (W_g1 W_d2) h_ij + qt_i - kt_j,   h_ij = relu(W_d1 (xyz_i - nbr_j) + b_d1),
//   qt_i = W_g1 q_i + W_g1 b_d2 + b_g1 (per query),  kt_j = W_g1 k_j (per reference point)
// so the gamma hidden layer and pos are BOTH plain GEMMs of h (no dependency between them) and the only per-token
// tensors are h, relu(gamma1) and the logits.
//
// The reference materialises ~10 (B,799,32,D) fp32 tensors; here a tile of NT tokens (NT/32 queries x 32 neighbours)
// goes through the three D x D GEMMs without leaving the SM:
//   * the GEMMs are computed TRANSPOSED on the tensor cores: out[c, t] = sum_k W[c,k] act[t,k]  (tcgen05.mma, M = 128
//     output channels per accumulator tile, N = NT tokens, K = D).  A = weights, streamed by TMA (SWIZZLE_128B,
//     64-wide K blocks) through a ring of smem stages; B = activations, written by the epilogue warps straight into
//     a SWIZZLE_128B K-major smem tile; accumulators live in TMEM.
//   * with channels on the TMEM lanes, one thread owns one channel for all tokens of the tile, so the per-channel
//     softmax over a query's 32 neighbours (32 consecutive TMEM columns) is thread-local register work — no
//     shuffles, no smem round trip.
//   * pos stays in TMEM (first MT*NT columns) until the final reduction; the gamma hidden layer / logits reuse the
//     other MT*NT columns.
//   * software pipeline over the tiles of a CTA: the h tile of tile i + 1 is built while the tensor core computes the
//     logits of tile i; GEMM order per tile is gamma1 -> pos -> logits so the pos GEMMs run under epilogue 2.
// warp 0 = TMA weight producer, warp 1 = MMA issuer (+TMEM alloc), warps 2.. = SETS*MT*4 "channel" warps
// (two sets of channel warps share the TMEM lanes and split the tile's queries, doubling the warps that hide the
// neighbour-gather latency).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace poem {

template <int D>
struct VaCfg {
  static constexpr int MT = D / 128;                 // accumulator tiles along the channel (M) axis
  static constexpr int KB = D / 64;                  // 64-wide K blocks
#ifndef POEM_VA_TWO_CTA
#define POEM_VA_TWO_CTA 0   // measured on B200 (medium): 0.618 ms vs 0.534 ms per launch — the 2-stage weight ring starves the MMAs
#endif
  // D <= 256: two CTAs per SM, each on its own 64-token tiles (one CTA's MMAs overlap the other's epilogues);
  // otherwise one CTA per SM with two channel-thread sets on 128-token tiles.
  static constexpr bool TWO = (POEM_VA_TWO_CTA != 0) && (D <= 256);
  static constexpr int CTAS_PER_SM = TWO ? 2 : 1;
  static constexpr int NT = (D <= 256 && !TWO) ? 128 : 64;   // tokens per tile (TMEM: 2 * MT * NT columns)
  static constexpr int QT = NT / 32;                 // queries per tile
  static constexpr int SETS = (D <= 256 && !TWO) ? 2 : 1;    // channel-thread sets; each set owns QT/SETS queries
  static constexpr int QPS = QT / SETS;              // queries per thread and tile (== 2 for every supported D)
  static constexpr int EP = MT * 128 * SETS;         // channel threads
  static constexpr int THREADS = 64 + EP;
  static constexpr int W_STAGES = TWO ? 2 : 5;
  static constexpr int W_TILE_BYTES = 128 * 64 * 2;  // [128 channels x 64 k] op16
  static constexpr int ACT_BYTES = NT * D * 2;       // KB blocks of [NT tokens x 64 k] op16
  static constexpr int TMEM_COLS = (2 * MT * NT <= 256) ? 256 : 512;
  // smem: act0 (h) | act1 (relu gamma1) | weight ring | wd1 (float4 per channel) | token rows | token rel xyz | barriers
  static constexpr int OFF_W = 2 * ACT_BYTES;
  static constexpr int OFF_WD1 = OFF_W + W_STAGES * W_TILE_BYTES;
  static constexpr int OFF_ROWS = OFF_WD1 + D * 16;
  static constexpr int OFF_REL = OFF_ROWS + 3 * NT * 4;      // token rows and rel xyz: three buffers (tile index mod 3)
  static constexpr int OFF_BARS = OFF_REL + 3 * NT * 16;
  static constexpr int SMEM_BYTES = OFF_BARS + 256;
};

struct VaParams {
  const op16* q;      // [n_query, ldq]  qt_i (gamma1-folded query term)
  const op16* ktab;   // [B*Lr, ldk]     kt_j (gamma1-folded key term)
  const op16* vtab;   // [B*Lr, ldv]
  int ldq, ldk, ldv;
  const float* q_xyz;          // [n_query, 3]
  const float* ref_xyz;        // [B*Lr, 3] (unused with anchors)
  const int* idx;              // [n_query, 32] or nullptr
  const int* anchor_idx;       // [32] or nullptr
  const float* anchor_xyz;     // [32,3] or nullptr
  const float* wd1;            // [D,3]
  const float* bd1;            // [D]
  const float* bd2;            // [D]
  op16* res;          // [n_query, D]
  int Lq, Lr, n_query;
  float softmax_scale_log2e;   // log2(e) / sqrt(D)
};

template <int D>
__global__ void __launch_bounds__(VaCfg<D>::THREADS, VaCfg<D>::CTAS_PER_SM)
va_fused_kernel(const __grid_constant__ CUtensorMap tmap_wd2, const __grid_constant__ CUtensorMap tmap_wg1,   // wg1 = W_g1 W_d2
                const __grid_constant__ CUtensorMap tmap_wg2, VaParams p) {
  using Cfg = VaCfg<D>;
  constexpr int MT = Cfg::MT, KB = Cfg::KB, NT = Cfg::NT, QT = Cfg::QT, EP = Cfg::EP;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_act = smem;                       // B operand of the pos / gamma1 GEMMs (stage A output)
  uint8_t* s_act1 = smem + Cfg::ACT_BYTES;     // B operand of the logits GEMM (epilogue 2 output)
  uint8_t* s_w = smem + Cfg::OFF_W;
  float4* s_wd1 = reinterpret_cast<float4*>(smem + Cfg::OFF_WD1);
  int* s_rows_base = reinterpret_cast<int*>(smem + Cfg::OFF_ROWS);
  float4* s_rel_base = reinterpret_cast<float4*>(smem + Cfg::OFF_REL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* w_full = bars;                        // [W_STAGES]
  uint64_t* w_empty = bars + Cfg::W_STAGES;       // [W_STAGES]
  uint64_t* act_full = bars + 2 * Cfg::W_STAGES;  // channel threads -> MMA: B operand ready (count EP)
  uint64_t* acc_full = act_full + 1;              // [MT] MMA -> channel threads of tile mt: accumulators ready
  uint64_t* pos_done = acc_full + MT;             // MMA -> channel threads: every round-0 MMA has read the h tile
  uint64_t* h_free = pos_done + 1;                // [MT] channel threads of tile mt -> MMA: logits / pos of this tile consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_free + MT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.n_query + QT - 1) / QT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_wd2);
    tma_prefetch_desc(&tmap_wg1);
    tma_prefetch_desc(&tmap_wg2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::W_STAGES; ++s) {
        mbar_init(&w_full[s], 1);
        mbar_init(&w_empty[s], 1);
      }
      mbar_init(act_full, EP);
      for (int m = 0; m < MT; ++m) mbar_init(&acc_full[m], 1);
      for (int m = 0; m < MT; ++m) mbar_init(&h_free[m], EP / MT);
      mbar_init(pos_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  // fc_delta.0 (3 -> D) as float4 (w0, w1, w2, b) per channel
  for (int c = threadIdx.x; c < D; c += Cfg::THREADS)
    s_wd1[c] = make_float4(p.wd1[c * 3 + 0], p.wd1[c * 3 + 1], p.wd1[c * 3 + 2], p.bd1[c]);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the setup above only touched weights (fc_delta.0) and on-chip state, and the TMA warp only ever streams
  // weights: it runs ahead while the predecessor kernel drains, everybody else waits for it here
  if (warp != 0) pdl_wait();
  pdl_trigger();
  const uint32_t tmem_pos = tmem_base;                 // MT tiles of NT columns
  const uint32_t tmem_h = tmem_base + MT * NT;         // MT tiles of NT columns

  if (warp == 0) {
    // ===================== TMA weight producer =====================
    if (elect_one()) {   // elect.sync: ptxas knows the region is single-threaded (no per-instruction elect loop)
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // order = consumption order of the MMA issuer: gamma1 (g = 1) of every channel tile, then pos (g = 0) of every
        // channel tile, then the logits (g = 2) of every channel tile; K blocks innermost
        for (int step = 0; step < 3 * MT; ++step) {
          const int g = (step < MT) ? 1 : (step < 2 * MT) ? 0 : 2;
          const int mt = step % MT;
          const CUtensorMap* tm = (g == 0) ? &tmap_wd2 : (g == 1) ? &tmap_wg1 : &tmap_wg2;
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&w_empty[stage], phase ^ 1);
            mbar_expect_tx(&w_full[stage], Cfg::W_TILE_BYTES);
            tma_load_2d(s_w + stage * Cfg::W_TILE_BYTES, tm, &w_full[stage], kb * 64, mt * 128);
            if (++stage == Cfg::W_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {   // elect.sync: ptxas knows the region is single-threaded (no per-instruction elect loop)
      constexpr uint32_t idesc = make_idesc_op16(128, NT);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t act_phase = 0;
      uint32_t hfree_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // Round 0 (B operand h): gamma1_pre = (W_g1 W_d2) h of every channel tile first — each committed to its own
        // barrier, so the channel warps of tile mt start epilogue 2 as early as possible — then pos = W_d2 h, which only
        // epilogue 3 reads and which therefore runs on the tensor core WHILE the channel warps are in epilogue 2.
        // Round 1 (B operand relu(gamma1)): logits = W_g2 relu(.), committed per channel tile; MMAs retire in order,
        // so that commit also covers the tile's pos accumulators.
        for (int step = 0; step < 3 * MT; ++step) {
          const int g = (step < MT) ? 1 : (step < 2 * MT) ? 0 : 2;
          const int mt = step % MT;
          if (step == 0 || step == 2 * MT) {
            mbar_wait(act_full, act_phase);
            act_phase ^= 1;
            tc_fence_after_sync();
          }
          if (step < MT) {
            // the accumulators of channel tile mt (logits in tmem_h, pos in tmem_pos) are released per channel tile: the
            // gamma1 MMAs of tile i + 1 for mt = 0 run while the channel warps of mt = 1 are still in epilogue 3 of tile i
            mbar_wait(&h_free[mt], hfree_phase ^ 1);
            tc_fence_after_sync();
          }
          const uint32_t acc = ((g == 0) ? tmem_pos : tmem_h) + mt * NT;
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t db = make_kmajor_desc<128>(smem_u32(g == 2 ? s_act1 : s_act) + kb * (NT * 128));
            mbar_wait(&w_full[stage], phase);
            tc_fence_after_sync();
            const uint64_t da = make_kmajor_desc<128>(smem_u32(s_w) + stage * Cfg::W_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_op16(acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            umma_commit(&w_empty[stage]);
            if (++stage == Cfg::W_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (g != 0) umma_commit(&acc_full[mt]);
          if (step == 2 * MT - 1) umma_commit(pos_done);   // the h tile may be overwritten by the next tile's stage A
        }
        hfree_phase ^= 1;
      }
    }
  } else {
    // ===================== channel warps =====================
    static_assert(Cfg::QPS == 2, "the epilogues are written for two queries per thread and tile");
    const int et = threadIdx.x - 64;                 // 0 .. EP-1
    const int quarter = warp & 3;
    const int e2 = (warp - 2) % (MT * 4);
    const int set = (warp - 2) / (MT * 4);
    const int mt = e2 >> 2;
    const int c = mt * 128 + quarter * 32 + lane;    // the channel this thread owns
    const int qi0 = set * Cfg::QPS;                  // first of this thread's two queries inside the tile
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float bd2 = p.bd2[c];
    // byte offset of (token row t, channel c) inside the activation tile: block c/64, 16-byte chunk (c%64)/8
    const uint32_t act_blk = (uint32_t)(c >> 6) * (NT * 128);
    const uint32_t act_chunk = (uint32_t)(c & 63) >> 3;
    const uint32_t act_byte = (uint32_t)(c & 7) * 2;
    uint32_t acc_phase = 0;

    // 32 op16 values of column c from the gather rows of one query, packed two per register (rows 2i | 2i+1).
    // Lanes 2k / 2k+1 own channels c / c+1 of the same 4-byte word: the even lane fetches that word for the even rows,
    // the odd lane for the odd rows (16 loads of 4 bytes instead of 32 of 2, half the registers in flight), then one
    // shuffle per word swaps them and a byte permute keeps this thread's half of both.
    int* s_rows = s_rows_base;        // gather rows of the current tile (one of three buffers)
    const uint32_t odd = (uint32_t)lane & 1u;
    const uint32_t prmt_sel = odd ? 0x3276u : 0x5410u;   // (own, partner) -> odd: partner.hi | own.hi << 16; even: own.lo | partner.lo << 16
    auto gather32 = [&](const op16* tab, int qi, uint32_t(&out)[16]) {
      const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tab + (c & ~1));
      // s_rows (element offsets row * ld) holds the even neighbours of a query first, then the odd ones
      const int4* rows4 = reinterpret_cast<const int4*>(s_rows + qi * 32 + odd * 16);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const int4 o = rows4[j4];
        out[4 * j4 + 0] = __ldg(t32 + ((uint32_t)o.x >> 1));
        out[4 * j4 + 1] = __ldg(t32 + ((uint32_t)o.y >> 1));
        out[4 * j4 + 2] = __ldg(t32 + ((uint32_t)o.z >> 1));
        out[4 * j4 + 3] = __ldg(t32 + ((uint32_t)o.w >> 1));
      }
    };
    auto exchange32 = [&](uint32_t(&w)[16]) {   // call once the words are needed: own word + partner's word -> rows 2i | 2i+1
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, w[i], 1);
        w[i] = __byte_perm(w[i], other, prmt_sel);
      }
    };
    auto bf_lo = [](uint32_t x) { return op16_lo(x); };
    auto bf_hi = [](uint32_t x) { return op16_hi(x); };

    // ---- tile metadata: gather row + relative position of every token, written by the first NT channel threads into
    //      one of three buffers (tile index mod 3).
    // Software pipeline over the tiles of this CTA: while the tensor core computes the logits of tile i the channel
    // warps already build the h tile of tile i + 1 (stage A), so at the end of tile i they only have to arrive and the
    // round-0 MMAs of tile i + 1 start at once.  That needs the metadata of tile i + 1 in the middle of tile i: its
    // neighbour indices are loaded one tile ahead (register), the dependent xyz loads are issued at the top of tile i
    // behind the kt gathers, and the single CTA-wide barrier of the tile (after epilogue 2, where every warp has to
    // wait for the logits anyway) publishes them.  Three buffers: the one being written was last read two barriers ago.
    auto load_index = [&](int q_first_) -> int {          // neighbour row of token `et` (element offset / ldk)
      const int qg = q_first_ + (et >> 5);
      const int j = et & 31;
      if (qg >= p.n_query) return 0;
      const int b = qg / p.Lq;
      return b * p.Lr + (p.anchor_idx != nullptr ? p.anchor_idx[j] : p.idx[(size_t)qg * 32 + j]);
    };
    auto write_meta = [&](int q_first_, int row, int mb) {
      const int qg = q_first_ + (et >> 5);
      const int j = et & 31;
      float rx = 0.f, ry = 0.f, rz = 0.f;
      if (qg < p.n_query) {
        float nx, ny, nz;
        if (p.anchor_idx != nullptr) {
          nx = p.anchor_xyz[j * 3 + 0], ny = p.anchor_xyz[j * 3 + 1], nz = p.anchor_xyz[j * 3 + 2];
        } else {
          const float* rp = p.ref_xyz + (size_t)row * 3;
          nx = rp[0], ny = rp[1], nz = rp[2];
        }
        rx = p.q_xyz[(size_t)qg * 3 + 0] - nx;
        ry = p.q_xyz[(size_t)qg * 3 + 1] - ny;
        rz = p.q_xyz[(size_t)qg * 3 + 2] - nz;
      }
      // element offset of the gather row (ldk == ldv and even, checked on the host); even neighbours first
      s_rows_base[mb * NT + (et & ~31) + (j & 1) * 16 + (j >> 1)] = row * p.ldk;
      s_rel_base[mb * NT + et] = make_float4(rx, ry, rz, 0.f);
    };
    // ---- stage A: h = relu(W_d1 rel + b_d1) -> activation tile (B operand of the pos and gamma1 GEMMs).
    //      A thread owns 8 channels x 8 tokens: the 8 weight rows stay in registers, so the tile costs 16 LDS.128 per
    //      thread (8 weight rows + 8 token offsets) instead of one per output channel (64).
    auto stage_a = [&](const float4* rel_buf) {
      constexpr int TG = NT / 8;               // token groups; thread = (token group, 8-channel group)
      static_assert(TG * (D / 8) == EP, "stage A mapping: 8 tokens x 8 channels per channel thread");
      const int tg = et % TG;
      const int ch = (et / TG) * 8;
      // packed fp16 arithmetic (HFMA2: two channels per instruction): h is rounded to fp16 for the tensor core anyway,
      // and the token-level MLPs are the precision-insensitive part of the path (scripts/precision_study2.py)
      op16x2 wx[4], wy[4], wz[4], wb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 w0 = s_wd1[ch + 2 * k], w1 = s_wd1[ch + 2 * k + 1];
        wx[k] = __floats2half2_rn(w0.x, w1.x);
        wy[k] = __floats2half2_rn(w0.y, w1.y);
        wz[k] = __floats2half2_rn(w0.z, w1.z);
        wb[k] = __floats2half2_rn(w0.w, w1.w);
      }
      uint8_t* blk = s_act + (ch >> 6) * (NT * 128);
      const uint32_t chunk = (uint32_t)(ch & 63) >> 3;
      const op16x2 zero2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int t = tg + TG * i;
        const float4 rel = rel_buf[t];
        const op16x2 rx = __float2half2_rn(rel.x), ry = __float2half2_rn(rel.y), rz = __float2half2_rn(rel.z);
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const op16x2 h2 = __hmax2(__hfma2(wx[k], rx, __hfma2(wy[k], ry, __hfma2(wz[k], rz, wb[k]))), zero2);
          pk[k] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(blk + sw128_offset(t, chunk)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
    };

    // prologue: metadata and h tile of this CTA's first tile, neighbour indices of its second one
    const int tile_step = (int)gridDim.x;
    int next_row = 0;
    if ((int)blockIdx.x < n_tiles) {
      if (et < NT) {
        write_meta(blockIdx.x * QT, load_index(blockIdx.x * QT), 0);
        if ((int)blockIdx.x + tile_step < n_tiles) next_row = load_index((blockIdx.x + tile_step) * QT);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EP) : "memory");
      stage_a(s_rel_base);
      mbar_arrive(act_full);
    }
    uint32_t pos_phase = 0;
    int mb = 0;                                                 // metadata buffer of the current tile
    for (int tile = blockIdx.x; tile < n_tiles; tile += tile_step) {
      const int q_first = tile * QT;
      const int mb_next = (mb == 2) ? 0 : mb + 1;
      s_rows = s_rows_base + mb * NT;
      const bool has_next = tile + tile_step < n_tiles;

      // ---- epilogue 2: relu((W_g1 W_d2) h + qt_i - kt_j) -> activation tile (B operand of the logits GEMM).
      //      All kt gathers of this thread's two queries are in flight before it waits for the accumulators.
      {
        uint32_t kk[2][16];
        gather32(p.ktab, qi0, kk[0]);
        gather32(p.ktab, qi0 + 1, kk[1]);
        op16x2 qv2[2];        // (qt, qt) of this channel for the thread's two queries
        const op16x2 zero2 = __floats2half2_rn(0.f, 0.f);
        uint32_t act1_q[2];   // smem address of (token row (qi0+u)*32, channel c) before the swizzle XOR
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int qg = q_first + qi0 + u;
          qv2[u] = __half2half2((qg < p.n_query) ? p.q[(size_t)qg * p.ldq + c] : __float2half(0.f));
          act1_q[u] = smem_u32(s_act1) + act_blk + (act_chunk << 4) + act_byte + (uint32_t)(qi0 + u) * (32 * 128);
        }
        // metadata of the next tile (its xyz loads queue up behind the kt gathers, all of it under the gamma1 GEMMs),
        // neighbour indices of the tile after it
        if (et < NT && has_next) {
          write_meta((tile + tile_step) * QT, next_row, mb_next);
          if (tile + 2 * tile_step < n_tiles) next_row = load_index((tile + 2 * tile_step) * QT);
        }
        mbar_wait(&acc_full[mt], acc_phase);
        acc_phase ^= 1;
        tc_fence_after_sync();
        exchange32(kk[0]);
        exchange32(kk[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[16];
            tmem_ld16(tmem_h + lane_off + mt * NT + (qi0 + u) * 32 + h * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // relu(acc + qt - kt) for token rows 2j | 2j+1 of this channel, two tokens per packed instruction
              const uint32_t kp = kk[u][h * 8 + j];
              const op16x2 d2 = __hsub2(qv2[u], *reinterpret_cast<const op16x2*>(&kp));
              const uint32_t a2 = pack_op16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
              const op16x2 v2 = __hmax2(__hadd2(*reinterpret_cast<const op16x2*>(&a2), d2), zero2);
              const uint32_t vb = *reinterpret_cast<const uint32_t*>(&v2);
              // token row t = (qi0+u)*32 + tl: SWIZZLE_128B flips address bits 4..6 with (t & 7) == (tl & 7)
              const int tl = h * 16 + 2 * j;
              asm volatile("st.shared.b16 [%0], %1;" ::"r"((act1_q[u] ^ ((uint32_t)(tl & 7) << 4)) + tl * 128),
                           "h"((unsigned short)(vb & 0xffffu)) : "memory");
              asm volatile("st.shared.b16 [%0], %1;" ::"r"((act1_q[u] ^ ((uint32_t)((tl + 1) & 7) << 4)) + (tl + 1) * 128),
                           "h"((unsigned short)(vb >> 16)) : "memory");
            }
          }
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      mbar_arrive(act_full);

      // ---- epilogue 3: per-channel softmax over the 32 neighbours, weighted sum of (v + pos).
      //      softmax((a + b_g2) / sqrt(D)): the bias is constant over the neighbours and cancels.
      {
        uint32_t vv[2][16];
#ifndef POEM_VA_V_EARLY
#define POEM_VA_V_EARLY 1   // v gathers of how many of the thread's two queries are issued before stage A of the next tile
                            // (measured, medium: 0 -> 0.3668 ms per launch, 1 -> 0.3655, 2 -> 0.3771 with 44 bytes of spills)
#endif
        if (POEM_VA_V_EARLY >= 1) gather32(p.vtab, qi0, vv[0]);
        if (POEM_VA_V_EARLY >= 2) gather32(p.vtab, qi0 + 1, vv[1]);
        // the only CTA-wide barrier of the tile: every warp is past epilogue 2 (the logits GEMM cannot start earlier
        // anyway) and the next tile's metadata is visible.  Then the h tile of the next tile, while the tensor core
        // computes this tile's logits; the round-0 MMAs of this tile must have read the old h tile (pos_done).
        asm volatile("bar.sync 1, %0;" ::"n"(EP) : "memory");
        if (has_next) {
          mbar_wait(pos_done, pos_phase);
          stage_a(s_rel_base + mb_next * NT);
          mbar_arrive(act_full);   // this thread's part of the next h tile is written (fenced in stage_a)
        }
        if (POEM_VA_V_EARLY < 1) gather32(p.vtab, qi0, vv[0]);
        if (POEM_VA_V_EARLY < 2) gather32(p.vtab, qi0 + 1, vv[1]);
        pos_phase ^= 1;
        mbar_wait(&acc_full[mt], acc_phase);
        acc_phase ^= 1;
        tc_fence_after_sync();
        exchange32(vv[0]);
        exchange32(vv[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int qg = q_first + qi0 + u;
          const uint32_t col = lane_off + mt * NT + (qi0 + u) * 32;
          // pass 1: max of the 32 logits
          float mx = -INFINITY;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t a[16];
            tmem_ld16(tmem_h + col + h * 16, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(a[j]));
          }
          // pass 2: exp, sum, weighted sum
          float sum = 0.f, acc = 0.f;
          const float mx_s = mx * p.softmax_scale_log2e;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t a[16], ps[16];
            tmem_ld16(tmem_h + col + h * 16, a);
            tmem_ld16(tmem_pos + col + h * 16, ps);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t vp = vv[u][h * 8 + j];
              const float e0 = fast_exp2(fmaf(__uint_as_float(a[2 * j]), p.softmax_scale_log2e, -mx_s));
              const float e1 = fast_exp2(fmaf(__uint_as_float(a[2 * j + 1]), p.softmax_scale_log2e, -mx_s));
              sum += e0 + e1;
              // b_d2 is constant over the neighbours: sum_j a_j (v_j + pos_j + b) = sum_j a_j (v_j + pos_j) + b
              acc = fmaf(e0, bf_lo(vp) + __uint_as_float(ps[2 * j]), acc);
              acc = fmaf(e1, bf_hi(vp) + __uint_as_float(ps[2 * j + 1]), acc);
            }
          }
          if (qg < p.n_query) p.res[(size_t)qg * D + c] = f2op16(acc / sum + bd2);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&h_free[mt]);   // this thread's TMEM reads of the tile (logits, pos of channel tile mt) are done
      mb = mb_next;
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
