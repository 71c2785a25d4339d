// 3x3 stride-1 convolution with halo reuse (the BasicBlock convolutions of HRNet, reference
// lib/models/backbones/hrnet.py:38-67 — 212 of the backbone's 305 convolutions and 85 % of its FLOPs).
//
// The generic implicit-GEMM path (gemm.cuh, ConvOperand) re-fetches every input pixel once per filter tap: 9 x the
// activation bytes travel L2 -> smem and that, not the tensor core, bounds it.  Here a CTA owns a 16x16 output patch
// and fetches its 18x18 input halo ONCE per 64-channel block with one 4-D TMA box (64 ch, 18 px, 18 rows; zero fill
// outside the image = padding).  The nine taps are then nine *shifted views* of that smem patch:
//   * an MMA M-tile is 16 image rows x 8 pixels, so an 8-row core-matrix group = 8 consecutive pixels of one image row
//     (128 B each, SWIZZLE_128B) and the stride between groups (SBO) is the uniform smem row pitch 18 x 128 B;
//   * tap (ky, kx) of sub-tile s starts at pixel (ky, kx + 8 s) of the patch: a 128-byte-granular start address that
//     is NOT 1024-byte aligned.  Measured on B200 (scripts/exp_halo.py): the tensor core applies the 128-byte swizzle
//     XOR to the bits of the final shared-memory address, exactly like TMA does when it writes the box, so any
//     128-byte-aligned start and any SBO address the TMA-written patch correctly with descriptor base offset 0
//     (setting base offset = (start >> 7) & 7 gives wrong results).
// Two sub-tiles (left / right 8 columns) share the patch and every weight tile -> L2 -> smem traffic per output pixel
// drops from 9 x 128 B (+ weights per 128 px) to 1.27 x 128 B (+ weights per 256 px; resident in smem when C = 64).
// What bounds the kernel after that is shared-memory bandwidth: an SS-mode MMA reads 128 x 32 B of A and C x 32 B of B
// per K = 16 step (C = 64: 192 B/clk against the 128 B/clk the SM delivers), so the epilogue keeps out of smem:
//   warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue straight from the TMEM layout
//   (lane = pixel, 32 channels = 64 contiguous bytes: 256-bit shortcut loads and stores, bias broadcast from smem);
//   accumulators double-buffered in TMEM when 4 x C columns fit (C <= 128).
// Real channel counts: HRNet-W40 has 40 / 80 / 160 channels stored padded to 64 / 128 / 192.  Multiplying the padding
// costs 2.6x / 2.6x / 1.4x the useful MMA work, so the K and N extents follow CR = the real count rounded up to 16:
// K is cut into blocks of 64 channels (SWIZZLE_128B), then 32 (SWIZZLE_64B), then 16 (SWIZZLE_32B) — each block is its
// own TMA box of the same padded NHWC tensor with the matching swizzle and smem descriptor — and N = CR.  The epilogue
// writes the CP - CR padding channels as zeros so every consumer still sees a clean padded tensor.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace poem {

// K blocks of a CR-channel operand: CR / 64 blocks of 64 channels, then the remainder as 32 and / or 16 channels
template <int CR>
struct HaloBlocks {
  static_assert(CR % 16 == 0 && CR >= 16, "real channel count is rounded up to 16");
  static constexpr int n64 = CR / 64;
  static constexpr int tail = CR % 64;   // 0, 16, 32 or 48
  static constexpr int n = n64 + (tail == 48 ? 2 : (tail ? 1 : 0));
  __host__ __device__ static constexpr int ch0(int b) { return b <= n64 ? 64 * b : 64 * n64 + 32; }
  __host__ __device__ static constexpr int nch(int b) { return b < n64 ? 64 : (b == n64 ? (tail == 16 ? 16 : 32) : 16); }
};

// CIP / CIR: padded / live input channels; CP / CR: padded / live output channels (live counts rounded up to 16)
template <int CIP, int CIR, int CP, int CR>
struct HaloCfg {
  static_assert(CP % 64 == 0 && CP <= 256 && CIP % 64 == 0, "padded channel counts");
  static_assert(CR <= CP && CR > CP - 64 && CIR <= CIP && CIR > CIP - 64, "live counts are rounded up to 16");
  using Blk = HaloBlocks<CIR>;
  static constexpr int kNB = Blk::n;                   // K blocks per tap
  static constexpr int kPitch = 18;                    // patch pixels per smem row
  static constexpr int kRows = 18;
  // patch ring: T tile slots, each holding the kNB K-block patches of one tile at fixed (1024-aligned) offsets, one
  // full/empty barrier pair per (slot, block) -> stage = step % (T * kNB)
  __host__ __device__ static constexpr int blk_bytes(int b) { return kRows * kPitch * Blk::nch(b) * 2; }
  __host__ __device__ static constexpr int blk_off(int b) {
    int off = 0;
    for (int i = 0; i < b; ++i) off += (blk_bytes(i) + 1023) / 1024 * 1024;
    return off;
  }
  static constexpr int kSlotBytes = blk_off(kNB);
  static constexpr int kTailBytes = CP * 4 + 512;      // bias + barriers
  static constexpr int kBudget = 232448 - 1024 - kTailBytes;
  // weights: all nine taps stay resident in smem when they leave room for two tile slots (CR <= 80); otherwise the
  // [CR x nch] tiles stream through a ring that takes whatever one tile slot leaves (an MMA consumes a tile in
  // ~0.3 us, an L2 round trip is ~1 us: a shallow ring starves the tensor core)
  static constexpr int kTapBytes = CR * CIR * 2;       // the kNB tiles of one tap, packed
  static constexpr bool kBResident = (kBudget - 9 * kTapBytes) >= 2 * kSlotBytes;
  static constexpr int kBStride = CR * 128;            // room for a [CR out-channel rows][64 k] tile
  static constexpr int kBStagesRoom = (kBudget - kSlotBytes) / kBStride;
  static constexpr int kBStages = kBStagesRoom > 8 ? 8 : kBStagesRoom;
  static constexpr int kBBytesTotal = ((kBResident ? 9 * kTapBytes : kBStages * kBStride) + 1023) / 1024 * 1024;
  static constexpr int kSlotsRoom = (kBudget - kBBytesTotal) / kSlotBytes;
  static constexpr int kSlots = kSlotsRoom > 4 ? 4 : kSlotsRoom;
  static constexpr int kAStages = kSlots * kNB;
  static_assert(kSlots >= 1 && kAStages <= 12 && (kBResident || kBStages >= 3), "smem plan");
  static constexpr int kAccPairs = 512 / (2 * CP) >= 4 ? 4 : (512 / (2 * CP) >= 2 ? 2 : 1);   // accumulator stages in TMEM
  static constexpr int kTmemCols = (2 * kAccPairs * CP <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kSlots * kSlotBytes + kBBytesTotal + kTailBytes;
};

constexpr int HALO_THREADS = 64 + 32 * 8;

struct HaloArgs {
  int n_images, R;                  // square maps, R % 16 == 0
  const float* bias;                // [CP]
  int relu;                         // ReLU after bias (+ residual)
  const op16* res;         // NHWC identity shortcut or nullptr
  op16* out;               // NHWC
  int cout_s;                       // channels per pixel of out / res in memory (multiple of 16, >= CR, <= CP)
};

// K-major descriptor for rows of `row_bytes` (128 / 64 / 32 -> SWIZZLE_128B / 64B / 32B) with an explicit stride
// between 8-row groups; base offset stays 0 (see above).
__device__ __forceinline__ uint64_t make_kmajor_desc_ex(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         (layout << 61);
}

struct HaloMaps {
  CUtensorMap x[3];   // activation boxes of 64 / 32 / 16 channels (SWIZZLE_128B / 64B / 32B)
  CUtensorMap w[3];   // weight boxes [CR rows][64 / 32 / 16 k]
};
__host__ __device__ constexpr int halo_map_index(int nch) { return nch == 64 ? 0 : (nch == 32 ? 1 : 2); }

template <int CIP, int CIR, int CP, int CR>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ HaloMaps maps, HaloArgs a) {
  using Cfg = HaloCfg<CIP, CIR, CP, CR>;
  using Blk = typename Cfg::Blk;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* s_a = smem;                                        // [kSlots][kSlotBytes]
  uint8_t* s_b = smem + Cfg::kSlots * Cfg::kSlotBytes;        // [kBStages][kBStride] or [9][kTapBytes]
  float* s_bias = reinterpret_cast<float*>(s_b + Cfg::kBBytesTotal);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + CP);
  uint64_t* a_full = bars;            // [12]
  uint64_t* a_empty = bars + 12;      // [12]
  uint64_t* b_full = bars + 24;       // [9]  (resident mode: b_full[0] only)
  uint64_t* b_empty = bars + 33;      // [9]
  uint64_t* tmem_full = bars + 42;    // [4]
  uint64_t* tmem_empty = bars + 46;   // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 50);

  const int tiles_x = a.R / 16;
  const int tiles_per_img = tiles_x * tiles_x;
  const int num_tiles = a.n_images * tiles_per_img;

  if (warp == 0 && lane == 0) {
    for (int b = 0; b < Cfg::kNB; ++b) {
      tma_prefetch_desc(&maps.x[halo_map_index(Blk::nch(b))]);
      tma_prefetch_desc(&maps.w[halo_map_index(Blk::nch(b))]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::kAStages; ++s) {
        mbar_init(&a_full[s], 1);
        mbar_init(&a_empty[s], 1);
      }
      for (int s = 0; s < Cfg::kAccPairs; ++s) {
        mbar_init(&tmem_full[s], 1);
        mbar_init(&tmem_empty[s], 8);
      }
      for (int s = 0; s < 9; ++s) {
        mbar_init(&b_full[s], 1);
        mbar_init(&b_empty[s], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  for (int j = threadIdx.x; j < CP; j += HALO_THREADS) s_bias[j] = a.bias ? __ldg(a.bias + j) : 0.f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // elect.sync: ptxas knows the region is single-threaded (no per-instruction elect loop)
      if (Cfg::kBResident) {
        mbar_expect_tx(&b_full[0], 9u * Cfg::kTapBytes);
        for (int tap = 0; tap < 9; ++tap)
          for (int b = 0; b < Cfg::kNB; ++b)
            tma_load_2d(s_b + tap * Cfg::kTapBytes + CR * Blk::ch0(b) * 2, &maps.w[halo_map_index(Blk::nch(b))], &b_full[0],
                        tap * CIP + Blk::ch0(b), 0);
      }
      const int my_tiles = ((int)blockIdx.x < num_tiles) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const int steps = my_tiles * Cfg::kNB;   // (tile, K block) pairs, in order
      // One thread feeds two independent rings (patches, weight tiles): it probes both "slot free" barriers without
      // blocking, so a full patch ring never delays the weight tiles of the step in flight (and vice versa).
      const int b_total = Cfg::kBResident ? 0 : steps * 9;
      int a_next = 0, b_next = 0;
      while (a_next < steps || b_next < b_total) {
        bool progress = false;
        if (a_next < steps) {
          const int st = a_next % Cfg::kAStages;
          if (mbar_test_wait(&a_empty[st], ((a_next / Cfg::kAStages) & 1) ^ 1)) {
            const int tile = (int)blockIdx.x + (a_next / Cfg::kNB) * (int)gridDim.x;
            const int b = a_next % Cfg::kNB;
            const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
            const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 16;
            mbar_expect_tx(&a_full[st], (uint32_t)Cfg::blk_bytes(b));
            tma_load_4d(s_a + (st / Cfg::kNB) * Cfg::kSlotBytes + Cfg::blk_off(b), &maps.x[halo_map_index(Blk::nch(b))],
                        &a_full[st], Blk::ch0(b), x0 - 1, y0 - 1, n);
            ++a_next;
            progress = true;
          }
        }
        if (b_next < b_total) {
          const int bs = b_next % Cfg::kBStages;
          if (mbar_test_wait(&b_empty[bs], ((b_next / Cfg::kBStages) & 1) ^ 1)) {
            const int b = (b_next / 9) % Cfg::kNB, tap = b_next % 9;
            mbar_expect_tx(&b_full[bs], (uint32_t)(CR * Blk::nch(b) * 2));
            tma_load_2d(s_b + bs * Cfg::kBStride, &maps.w[halo_map_index(Blk::nch(b))], &b_full[bs], tap * CIP + Blk::ch0(b), 0);
            ++b_next;
            progress = true;
          }
        }
        if (!progress) __nanosleep(64);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // elect.sync (not `lane == 0`): ptxas then knows the region is single-threaded and issues each UTCHMMA directly
    // instead of wrapping it in a per-instruction elect / branch loop (~40 cycles per MMA, which at N = 48 — 24 tensor
    // cycles per MMA — left the tensor pipe half idle)
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_op16(128, CR);
      int step = 0, bs = 0, acc = 0;
      uint32_t bphase = 0, acc_phase = 0;
      if (Cfg::kBResident) mbar_wait(&b_full[0], 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * 2) * CP;
#pragma unroll
        for (int b = 0; b < Cfg::kNB; ++b, ++step) {
          const int nch = Blk::nch(b);
          const uint32_t row_bytes = (uint32_t)nch * 2;
          const int st = step % Cfg::kAStages;
          mbar_wait(&a_full[st], (step / Cfg::kAStages) & 1);
          tc_fence_after_sync();
          const uint32_t pa = smem_u32(s_a + (st / Cfg::kNB) * Cfg::kSlotBytes + Cfg::blk_off(b));
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t pb;
            if (Cfg::kBResident) {
              pb = smem_u32(s_b + tap * Cfg::kTapBytes + CR * Blk::ch0(b) * 2);
            } else {
              mbar_wait(&b_full[bs], bphase);
              tc_fence_after_sync();
              pb = smem_u32(s_b + bs * Cfg::kBStride);
            }
            const int ky = tap / 3, kx = tap - 3 * ky;
            const uint64_t db = make_kmajor_desc_ex(pb, 8 * row_bytes, row_bytes);
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const uint32_t start = pa + (uint32_t)(ky * Cfg::kPitch + kx + 8 * sub) * row_bytes;
              const uint64_t da = make_kmajor_desc_ex(start, Cfg::kPitch * row_bytes, row_bytes);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nch / 16) umma_op16(d0 + sub * CP, da + 2 * k, db + 2 * k, idesc, (b | tap | k) != 0);
            }
            if (!Cfg::kBResident) {
              umma_commit(&b_empty[bs]);
              if (++bs == Cfg::kBStages) {
                bs = 0;
                bphase ^= 1;
              }
            }
          }
          umma_commit(&a_empty[st]);
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == Cfg::kAccPairs) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    // TMEM lane m = 32 q + lane is the pixel (row y0 + m / 8, column x0 + 8 sub + m % 8); a 32-channel chunk of it is 64
    // contiguous bytes of the NHWC tensor, moved as two 256-bit accesses.  Warps q and q + 4 split the chunks by parity.
    // With two warps per scheduler the epilogue is bound by the length of its own dependent instruction stream (an
    // "empty" kernel — no loads, MMAs or stores — ran at 2/3 of the full one), so the stream is kept short: the chunk
    // parity is a compile-time constant (channel masks fold away), tile -> pixel arithmetic is shifts on 32-bit values
    // (the map is a power of two wide), ReLU rides on the op16 conversion (cvt.rn.relu), and for C <= 64 the bias sits
    // in registers and both sub-tiles' TMEM loads are in flight together.
    const int quarter = warp & 3;
    constexpr int kChunks = CP / 64;
    const int tx_log2 = 31 - __clz(tiles_x);                  // R / 16 is 1, 2, 4, ...
    const uint32_t lane_pix = (uint32_t)((quarter * 4 + (lane >> 3)) * a.R + (lane & 7));
    auto tile_pix = [&](int tile) -> uint32_t {               // pixel index of the tile's first pixel
      const int n = tile >> (2 * tx_log2), rem = tile & (tiles_per_img - 1);
      return (uint32_t)((n * a.R + ((rem >> tx_log2) << 4)) * a.R + ((rem & (tiles_x - 1)) << 4));
    };
    auto body = [&](auto par_c) {
      constexpr int par = decltype(par_c)::value;
      constexpr bool kRegBias = (kChunks == 1);
      constexpr bool kBothSubs = (kChunks == 1);              // 2 x 32 accumulator registers
      float bias_r[kRegBias ? 32 : 1];
      if constexpr (kRegBias) {
#pragma unroll
        for (int q = 0; q < 32; ++q) bias_r[q] = (par * 32 + q < CR) ? s_bias[par * 32 + q] : 0.f;
      }
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const uint32_t pix0 = tile_pix(tile) + lane_pix;
        if (a.res != nullptr && tile + (int)gridDim.x < num_tiles) {
          // the shortcut pixels of this CTA's NEXT tile start travelling HBM -> L2 now
          const op16* rp = a.res + (size_t)(tile_pix(tile + (int)gridDim.x) + lane_pix) * a.cout_s;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub)
#pragma unroll
            for (int j = 0; j < kChunks; ++j)
              if ((2 * j + par) * 32 < CR) prefetch_l2(rp + (size_t)(sub * 8) * a.cout_s + (2 * j + par) * 32);
        }
        auto finish = [&](int sub, int j, const uint32_t (&r)[32], const uint32_t (&rr)[16]) {
          const int c0 = (2 * j + par) * 32;
          op16* op = a.out + (size_t)(pix0 + sub * 8) * a.cout_s + c0;
          uint32_t pk[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int c = c0 + 2 * q;
            if (c >= CR) {   // columns past N = CR were never written by the MMA (compile-time for a given chunk)
              pk[q] = 0u;
              continue;
            }
            float v0 = __uint_as_float(r[2 * q]), v1 = __uint_as_float(r[2 * q + 1]);
            if constexpr (kRegBias) {
              v0 += bias_r[2 * q], v1 += bias_r[2 * q + 1];
            } else {
              const float2 b2 = *reinterpret_cast<const float2*>(s_bias + c);
              v0 += b2.x, v1 += b2.y;
            }
            if (a.res != nullptr) {
              const float2 f = op16x2_to_f2(*reinterpret_cast<const op16x2*>(&rr[q]));
              v0 += f.x, v1 += f.y;
            }
            pk[q] = a.relu ? pack_op16x2_relu(v0, v1) : pack_op16x2(v0, v1);
          }
          stg_256(op, &pk[0]);
          if (c0 + 16 < a.cout_s) stg_256(op + 16, &pk[8]);
        };
        auto load_res = [&](int sub, int j, uint32_t (&rr)[16]) {
          const int c0 = (2 * j + par) * 32;
          const op16* rp = a.res + (size_t)(pix0 + sub * 8) * a.cout_s + c0;
          ldg_nc_256(rp, &rr[0]);
          if (c0 + 16 < a.cout_s) ldg_nc_256(rp + 16, &rr[8]);
        };
        auto zero_chunk = [&](int sub, int j) {   // pure padding chunk inside the stored channels
          const int c0 = (2 * j + par) * 32;
          op16* op = a.out + (size_t)(pix0 + sub * 8) * a.cout_s + c0;
          uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          stg_256(op, z);
          if (c0 + 16 < a.cout_s) stg_256(op + 16, z);
        };
        const uint32_t tm = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 2 * CP + par * 32);
        if constexpr (kBothSubs) {
          uint32_t rr[2][16], r[2][32];
          if (a.res != nullptr) {
            load_res(0, 0, rr[0]);
            load_res(1, 0, rr[1]);
          }
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after_sync();
          tmem_ld32(tm, r[0]);
          tmem_ld32(tm + CP, r[1]);
          tmem_ld_wait();
          finish(0, 0, r[0], rr[0]);
          finish(1, 0, r[1], rr[1]);
        } else {
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            uint32_t rr[kChunks][16];
            if (a.res != nullptr) {   // every shortcut load of this sub-tile is in flight before the accumulator is read
#pragma unroll
              for (int j = 0; j < kChunks; ++j)
                if ((2 * j + par) * 32 < CR) load_res(sub, j, rr[j]);
            }
            if (sub == 0) {
              mbar_wait(&tmem_full[acc], acc_phase);
              tc_fence_after_sync();
            }
#pragma unroll
            for (int j = 0; j < kChunks; ++j) {
              const int c0 = (2 * j + par) * 32;
              if (c0 >= CR) {
                if (c0 < a.cout_s) zero_chunk(sub, j);
                continue;
              }
              uint32_t r[32];
              tmem_ld32(tm + (uint32_t)(sub * CP + 64 * j), r);
              tmem_ld_wait();
              finish(sub, j, r, rr[j]);
            }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == Cfg::kAccPairs) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    };
    if (((warp - 2) >> 2) == 0) body(std::integral_constant<int, 0>{});
    else body(std::integral_constant<int, 1>{});
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace poem
