// Persistent warp-specialised tcgen05 GEMM:  C[M,N] = epilogue(A[M,K] · W[N,K]^T)
//   A, W : op16, K-major (row-major with K contiguous), staged by TMA (SWIZZLE_128B, 64-element K blocks)
//   accumulate fp32 in TMEM (two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue (two per TMEM lane quarter)
//   W-stationary mode (whenever the [BN x K] weight slab fits in smem): every CTA keeps ONE n-tile for its whole
//   life, loads its weight slab once and streams only A tiles through the ring -> L2->smem traffic per tile drops
//   from (A + W) to A (the skinny-K GEMMs of this path are L2/HBM-bound, not MMA-bound).
// Every Linear / 1x1-conv of the decoder path runs through this kernel (reference call sites: cuBLAS GEMMs
// behind nn.Linear / nn.Conv2d(k=1) in lib/models/heads/ptEmb_head.py:94,755,760 and
// lib/models/bricks/pt_metro_transformer.py:180-181, point_transformers.py:86-95,139-151).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace poem {

enum GemmAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };
enum GemmRes : int {
  RES_NONE = 0,
  RES_F32 = 1,      // out += res_f32[m * res_ld + n]
  RES_POSADD = 2,   // out += res_f32[(row_tab[m / 256] * 256 + m % 256) * res_ld + n]   (positional term per view)
  RES_MERGE = 3,    // out = out * row_scale[m / P] + res_op16[(row_base[m / P] + (m % P) * row_cnt[m / P]) * res_ld + n]
  RES_BF16 = 4      // out += res_op16[m * res_ld + n]   (BasicBlock identity shortcut, NHWC)
};

// Implicit-GEMM convolution: the A operand is gathered by TMA straight from an NHWC activation tensor
// (4-D tensor map, zero fill outside the image = conv padding); K block kb = (filter tap, 64-channel block).
// An M tile is 128 consecutive NHWC output pixels (whole image rows), so the epilogue addressing is unchanged.
struct ConvOperand {
  int enabled;
  int ksize;     // 1 or 3
  int pad;       // 0 or 1
  int stride;    // 1 or 2
  int cblocks;   // padded input channels / 64
  int Hout, Wout;
};

constexpr int GEMM_MAX_STAGES = 8;

// Launch-time shape of the pipeline (computed on the host from the smem budget).
struct GemmPipe {
  int w_stationary;   // 1: weight slab resident in smem, ring holds A tiles only
  int n_stages;       // ring depth (<= GEMM_MAX_STAGES)
};

struct GemmEpilogue {
  const float* bias;   // [N] or nullptr
  int act;             // applied before the residual
  int act_after_res;   // applied after the residual (row-major path only)
  int res_mode;
  const float* res_f32;
  const op16* res_op16;
  int res_ld;
  const int* row_tab;      // RES_POSADD: per-image row of the positional table; RES_MERGE: row_base per sample
  const int* row_cnt;      // RES_MERGE: views per sample
  int rows_per_group;      // RES_MERGE: P
  // per-row power-of-two scale sigma[m] of a pre-scaled A operand (merge-net aggregate, see merge_reduce_kernel):
  //   sigma_mode 1: out = act(acc + bias / sigma[m])   (the ReLU MLP layer computed on A / sigma: positively homogeneous)
  //   sigma_mode 2: acc is multiplied by sigma[m] before the bias (undoes the scaling)
  const float* row_sigma;  // nullptr = unused
  int sigma_mode;
  float* out_f32;          // row-major [M, ld_f32] or nullptr
  int ld_f32;
  op16* out_op16; // row-major [M, ld_op16] or nullptr
  int ld_op16;
  int n_store;             // row-major outputs / residuals only exist for columns < n_store (multiple of 16; = N normally)
  // columns >= trans_from go to a per-group transposed buffer:
  //   out_t[(m / t_rows) * t_group_stride + (n - trans_from) * t_rows + (m % t_rows)]
  int trans_from;          // = N when unused
  int t_rows;
  long long t_group_stride;
  op16* out_t_op16;
  float* out_t_f32;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;   // two warps per TMEM lane quarter, interleaved over the 32-column chunks
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kWBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kBiasMax = 4096;                                   // the whole bias vector lives in smem up to this N
  static constexpr int kBiasBytes = kBiasMax * 4;
  static constexpr int kTailBytes = kBiasBytes + 256 /*barriers*/;
  static constexpr int kSmemMax = 232448;                                 // 227 KB per CTA
  static constexpr int kStagesRoom = (kSmemMax - 1024 - kTailBytes) / kStageBytes;
  static constexpr int kStages = kStagesRoom > 6 ? 6 : kStagesRoom;       // BN = 256: 4 x 48 KB
  static constexpr int kSmemBytes = kStages * kStageBytes + kTailBytes;   // streaming mode
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// kRes: the epilogue has a residual operand (ep.res_mode != RES_NONE).  Compile-time so that the residual-free
// instantiation carries none of the prefetch registers / branches.
template <int BN, bool kRes>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_op16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, int M,
                    int N, int K, GemmEpilogue ep, ConvOperand conv, GemmPipe pipe) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  // smem: [resident W slab (stationary mode)] [ring] [bias] [barriers]
  const bool wst = pipe.w_stationary != 0;
  const int n_stages = pipe.n_stages;
  const int ring_stage_bytes = wst ? Cfg::kABytes : Cfg::kStageBytes;
  uint8_t* s_wres = smem;
  uint8_t* s_ring = smem + (wst ? k_blocks * Cfg::kWBytes : 0);
  uint8_t* s_tail = s_ring + n_stages * ring_stage_bytes;
  float* s_bias = reinterpret_cast<float*>(s_tail);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tail + Cfg::kBiasBytes);
  uint64_t* full_bar = bars;                               // [GEMM_MAX_STAGES]
  uint64_t* empty_bar = bars + GEMM_MAX_STAGES;            // [GEMM_MAX_STAGES]
  uint64_t* tmem_full = bars + 2 * GEMM_MAX_STAGES;        // [2]
  uint64_t* tmem_empty = bars + 2 * GEMM_MAX_STAGES + 2;   // [2]
  uint64_t* wres_full = bars + 2 * GEMM_MAX_STAGES + 4;    // resident weight slab landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_MAX_STAGES + 5);
  // tile schedule. streaming: tile = blockIdx.x + i * gridDim.x over all (m, n) tiles, n fastest.
  // stationary: this CTA owns n-tile blockIdx.x % tiles_n and walks m-tiles blockIdx.x / tiles_n + i * (gridDim.x / tiles_n)
  const int it0 = wst ? (int)blockIdx.x / tiles_n : (int)blockIdx.x;
  const int it_step = wst ? (int)gridDim.x / tiles_n : (int)gridDim.x;
  const int it_end = wst ? tiles_m : num_tiles;
  const int my_n_tile = (int)blockIdx.x % tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < n_stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(wres_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full[s], 1);
        mbar_init(&tmem_empty[s], GEMM_EPI_WARPS);  // one arrive per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  // the bias vector goes to smem once (a per-tile global load + named barrier sat on the epilogue's critical path)
  const bool bias_resident = (N <= Cfg::kBiasMax);
  if (bias_resident)
    for (int j = threadIdx.x; j < N; j += GEMM_THREADS) s_bias[j] = ep.bias != nullptr ? __ldg(ep.bias + j) : 0.f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barriers, TMEM, bias preload = weights) is independent of the predecessor kernel; so is the
  // resident weight slab of the W-stationary mode, which the TMA warp requests before it waits
  if (warp != 0) pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // elect.sync: ptxas knows the region is single-threaded (no per-instruction elect loop)
      int stage = 0;
      uint32_t phase = 0;
      if (wst) {   // the whole [BN x K] weight slab of this CTA's n-tile, once
        mbar_expect_tx(wres_full, (uint32_t)(k_blocks * Cfg::kWBytes));
        for (int kb = 0; kb < k_blocks; ++kb)
          tma_load_2d(s_wres + kb * Cfg::kWBytes, &tmap_w, wres_full, kb * GEMM_BK, my_n_tile * BN);
      }
      pdl_wait();   // the A operand (and a residual) come from the predecessor kernel
      for (int it = it0; it < it_end; it += it_step) {
        const int m0 = (wst ? it : it / tiles_n) * GEMM_BM;
        const int n0 = (wst ? my_n_tile : it % tiles_n) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = s_ring + stage * ring_stage_bytes;
          uint8_t* sw = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], (uint32_t)ring_stage_bytes);
          if (conv.enabled) {
            const int tap = kb / conv.cblocks, cc = kb - tap * conv.cblocks;
            const int ky = tap / conv.ksize, kx = tap - ky * conv.ksize;
            const int pix = conv.Hout * conv.Wout;
            const int n_img = m0 / pix, y0 = (m0 - n_img * pix) / conv.Wout;
            tma_load_4d(sa, &tmap_a, &full_bar[stage], cc * 64, kx - conv.pad, y0 * conv.stride + ky - conv.pad, n_img);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * GEMM_BK, m0);
          }
          if (!wst) tma_load_2d(sw, &tmap_w, &full_bar[stage], kb * GEMM_BK, n0);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {   // elect.sync: ptxas knows the region is single-threaded (no per-instruction elect loop)
      constexpr uint32_t idesc = make_idesc_op16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (wst) mbar_wait(wres_full, 0);
      for (int it = it0; it < it_end; it += it_step) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(s_ring + stage * ring_stage_bytes);
          const uint32_t sw = wst ? smem_u32(s_wres + kb * Cfg::kWBytes) : sa + Cfg::kABytes;
          const uint64_t da = make_kmajor_desc<128>(sa);
          const uint64_t dw = make_kmajor_desc<128>(sw);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 op16 (32 B) along K inside the 128-byte swizzle atom: +2 in the (addr>>4) field
            umma_op16(d_tmem, da + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    // TMEM hands each lane one accumulator ROW (32 columns at a time) and the epilogue keeps that layout: a lane's 32
    // columns are 128 (fp32) / 64 (op16) contiguous bytes of its output row, moved with 256-bit accesses (one full
    // 32-byte sector per lane and instruction), bias broadcast from smem.  No shared-memory transpose: with SS-mode
    // MMAs the operand reads already take most of the SM's shared-memory bandwidth, and a staged epilogue was what
    // bounded the skinny-K GEMMs of this path.
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    const int ew = warp - 2;       // 0..7
    const int chunk_par = ew >> 2; // this warp handles 32-column chunks with (chunk index & 1) == chunk_par
    int acc = 0;
    uint32_t acc_phase = 0;
    // residual row of output row m: fp32 or op16 pointer (nullptr when absent / out of range) and the RES_MERGE scale
    auto res_row = [&](int m, const float*& pf, const op16*& pb, float& scale) {
      pf = nullptr;
      pb = nullptr;
      scale = 1.0f;
      if (!kRes || m >= M) return;
      if (ep.res_mode == RES_F32) {
        pf = ep.res_f32 + (size_t)m * ep.res_ld;
      } else if (ep.res_mode == RES_POSADD) {
        pf = ep.res_f32 + ((size_t)ep.row_tab[m >> 8] * 256 + (m & 255)) * ep.res_ld;
      } else if (ep.res_mode == RES_BF16) {
        pb = ep.res_op16 + (size_t)m * ep.res_ld;
      } else if (ep.res_mode == RES_MERGE) {
        const int g = m / ep.rows_per_group;
        const int cnt = ep.row_cnt[g];
        scale = 1.0f / (float)cnt;
        // row_tab == nullptr: the token-first rows were gathered by the producer (fused sampler/merge kernel): row m
        pb = (ep.row_tab == nullptr) ? ep.res_op16 + (size_t)m * ep.res_ld
                                     : ep.res_op16 + ((size_t)ep.row_tab[g] + (size_t)(m - g * ep.rows_per_group) * cnt) * ep.res_ld;
      }
    };
    for (int it = it0; it < it_end; it += it_step) {
      const int m0 = (wst ? it : it / tiles_n) * GEMM_BM;
      const int n0 = (wst ? my_n_tile : it % tiles_n) * BN;
      float* sb = s_bias + n0;
      if (!bias_resident) {   // very wide outputs: this tile's columns, double buffered by accumulator stage
        sb = s_bias + acc * BN;
        for (int j = threadIdx.x - 64; j < BN; j += 32 * GEMM_EPI_WARPS)
          sb[j] = (ep.bias != nullptr && n0 + j < N) ? __ldg(ep.bias + n0 + j) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * GEMM_EPI_WARPS) : "memory");
      }

      const int mt_row = m0 + quarter * 32 + lane;   // this lane's output row
      const float* resf;
      const op16* resb;
      float rscale;
      res_row(mt_row, resf, resb, rscale);
      const float row_sig = (ep.row_sigma != nullptr && mt_row < M) ? __ldg(ep.row_sigma + mt_row) : 1.0f;
      // the residual rows of this CTA's NEXT tile start travelling HBM -> L2 now (the epilogue of a skinny-K GEMM is
      // otherwise bound by the latency of these loads, not by bandwidth)
      if (kRes && it + it_step < it_end) {
        const int itn = it + it_step;
        const int m0n = (wst ? itn : itn / tiles_n) * GEMM_BM;
        const int n0n = (wst ? my_n_tile : itn % tiles_n) * BN;
        const float* pf;
        const op16* pb;
        float sc;
        res_row(m0n + quarter * 32 + lane, pf, pb, sc);
        for (int c0 = chunk_par * 32; c0 < BN && n0n + c0 < N && n0n + c0 < ep.trans_from; c0 += 64) {
          if (pf != nullptr) prefetch_l2(pf + n0n + c0);
          if (pb != nullptr) prefetch_l2(pb + n0n + c0);
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();

      size_t t_base = 0;
      if (ep.trans_from < N && mt_row < M) {
        const int g = mt_row / ep.t_rows;
        t_base = (size_t)g * ep.t_group_stride + (size_t)(mt_row - g * ep.t_rows);
      }

      // This warp's chunks are c0 = 32 * chunk_par + 64 * j.  The active ones form a prefix (columns < N that exist in
      // memory), so without a residual operand the TMEM load of chunk j + 1 is issued before chunk j is processed.
      constexpr int kChunksPerWarp = (BN / 32 + 1) / 2;
      constexpr int kRegSets = kRes ? 1 : 2;
      auto chunk_active = [&](int j) {
        const int c0 = chunk_par * 32 + 64 * j, n = n0 + c0;
        return c0 < BN && n < N && !(n < ep.trans_from && ep.n_store - n <= 0);
      };
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + chunk_par * 32;
      uint32_t rbuf[kRegSets][32];
      if (!kRes && chunk_active(0)) tmem_ld32(tmem_acc, rbuf[0]);
#pragma unroll
      for (int j = 0; j < kChunksPerWarp; ++j) {
        if (!chunk_active(j)) break;   // warp-uniform
        const int c0 = chunk_par * 32 + 64 * j;
        const int n = n0 + c0;
        // this chunk's residual values: in flight before the accumulator chunk is read
        uint32_t rr[32];
        const int valid = ep.n_store - n;   // row-major columns of this chunk that exist in memory: >= 32 or 16
        if (kRes && n < ep.trans_from) {
          if (resf != nullptr) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (valid >= 8 * (q + 1)) ldg_nc_256(resf + n + 8 * q, &rr[8 * q]);
          } else if (resb != nullptr) {
            ldg_nc_256(resb + n, &rr[0]);
            if (valid >= 32) ldg_nc_256(resb + n + 16, &rr[8]);
          }
        }
        uint32_t (&r)[32] = rbuf[kRes ? 0 : (j & 1)];
        if (kRes) {
          __syncwarp();   // lanes past the last row skip the stores below: reconverge before the aligned TMEM load
          tmem_ld32(tmem_acc + 64 * j, r);
        }
        tmem_ld_wait();
        if (!kRes && j + 1 < kChunksPerWarp && chunk_active(j + 1)) {
          __syncwarp();
          tmem_ld32(tmem_acc + 64 * (j + 1), rbuf[(j + 1) & 1]);
        }
        if (mt_row < M) {
        float v[32];
        if (ep.row_sigma == nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + 4 * j);
            v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b4.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
          }
        } else {
          const float a_sc = (ep.sigma_mode == 2) ? row_sig : 1.0f;
          const float b_sc = (ep.sigma_mode == 1) ? 1.0f / row_sig : 1.0f;   // sigma is a power of two: exact
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + 4 * j);
            v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), a_sc, b4.x * b_sc);
            v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), a_sc, b4.y * b_sc);
            v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), a_sc, b4.z * b_sc);
            v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), a_sc, b4.w * b_sc);
          }
        }
        if (ep.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (ep.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (n >= ep.trans_from) {
          // ---------------- transposed store (consecutive lanes = consecutive addresses) ----------------
          if (kRes && (ep.res_mode == RES_F32 || ep.res_mode == RES_POSADD)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(resf + n) + j);
              v[4 * j + 0] += t.x;
              v[4 * j + 1] += t.y;
              v[4 * j + 2] += t.z;
              v[4 * j + 3] += t.w;
            }
          }
          const int tc = n - ep.trans_from;
          if (ep.out_t_op16 != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) ep.out_t_op16[t_base + (size_t)(tc + j) * ep.t_rows] = f2op16(v[j]);
          }
          if (ep.out_t_f32 != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) ep.out_t_f32[t_base + (size_t)(tc + j) * ep.t_rows] = v[j];
          }
        } else {
        // ---------------- row-major store ----------------
        if (kRes && resf != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(rr[j]);
        } else if (kRes && resb != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 f = op16x2_to_f2(*reinterpret_cast<const op16x2*>(&rr[j]));
            v[2 * j] = v[2 * j] * rscale + f.x;
            v[2 * j + 1] = v[2 * j + 1] * rscale + f.y;
          }
        }
        if (ep.act_after_res == ACT_RELU) {   // BasicBlock: relu(bn2(conv2(.)) + identity)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (ep.out_f32 != nullptr) {
          float* o = ep.out_f32 + (size_t)mt_row * ep.ld_f32 + n;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (valid >= 8 * (j + 1)) stg_256(o + 8 * j, reinterpret_cast<const uint32_t*>(&v[8 * j]));
        }
        if (ep.out_op16 != nullptr) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_op16x2(v[2 * j], v[2 * j + 1]);
          op16* o = ep.out_op16 + (size_t)mt_row * ep.ld_op16 + n;
          stg_256(o, &pk[0]);
          if (valid >= 32) stg_256(o + 16, &pk[8]);
        }
        }   // row-major
        }   // mt_row < M
      }
      // all TMEM reads of this warp are complete (wait::ld above) -> hand the accumulator back
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
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
