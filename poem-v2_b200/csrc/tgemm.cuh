// TF32 tcgen05 GEMM of the training path:  C[M,N] (+)= alpha * op(A) · op(B)^T (+ bias),  fp32 in HBM, fp32 accumulate.
//   Each operand is either K-major (stored [rows x K], K contiguous) or MN-major (stored [K x rows], rows contiguous),
//   so that the three GEMMs of a Linear layer read the SAME row-major tensors without a transpose pass:
//     forward  y  = x · W^T      A = x  (K-major)   B = W  (K-major)
//     dgrad    dx = dy · W       A = dy (K-major)   B = W  (MN-major: stored [N_out x K_in] = [K_g x N_g])
//     wgrad    dW = dy^T · x     A = dy (MN-major)  B = x  (MN-major), reduction over the token rows
//   and likewise Q·K^T, P·V, P^T·dO, dS·K, dS^T·Q of the attention backward (reference: autograd of
//   lib/models/bricks/pt_metro_transformer.py:57-91, point_transformers.py:70-156, heads/ptEmb_head.py:745-771).
//   TMA stages 128-byte (32 x fp32) SWIZZLE_128B rows; `tcgen05.mma kind::tf32` consumes 8 K-values per instruction;
//   warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM), warps 2..5 = operand rounding, then epilogue.  One 128 x BN tile
//   per CTA, 2 CTAs/SM.  The tensor core TRUNCATES fp32 operands to TF32 (low 13 mantissa bits ignored: a biased error,
//   measured 4.5e-3 on the regressed coordinates); warps 2..5 therefore round every landed stage to nearest
//   (cvt.rna.tf32.f32, in place in shared memory, any layout) before the MMA warp may read it.  These GEMMs are
//   HBM/L2-bound (K = D), the extra shared-memory pass hides behind the other resident CTA.
//   A batch (two strides per operand, rank-4 tensor maps: per-slice bounds, zero fill outside) and a K split with atomic
//   accumulation (tall reductions: wgrad) ride in blockIdx.z.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace poem {

#ifndef POEM_TG_TRACE
#define POEM_TG_TRACE 0      // 1: per-CTA phase timestamps (globaltimer, ns) into g_tg_trace (scripts/tgemm_trace.py)
#endif
#if POEM_TG_TRACE
__device__ unsigned long long g_tg_trace[8192 * 16];
__device__ __forceinline__ unsigned long long tg_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TG_MARK(slot) do { if ((threadIdx.x & 31) == 0) g_tg_trace[((blockIdx.y * gridDim.x + blockIdx.x) % 8192) * 16 + (slot)] = tg_now(); } while (0)
#else
#define TG_MARK(slot) do {} while (0)
#endif
constexpr int TG_BM = 128;
constexpr int TG_BK = 32;          // fp32 elements per 128-byte swizzle row
constexpr int TG_STAGES = 3;          // ring depth; 2 for problems with K <= 64 (three CTAs per SM instead of two)
constexpr int TG_THREADS = 192;
constexpr int TG_ATOM_BYTES = TG_BK * 128;   // one MN-major atom: 32 K rows x 128 B

enum TgMode : int { TG_STORE = 0, TG_ADD = 1, TG_ATOMIC = 2 };

struct TgParams {
  int M, N, K;
  int nb1, nb2;            // batch extents (>= 1); blockIdx.z = (split * nb2 + b2) * nb1 + b1
  int a_bc1, a_bc2;        // 1: operand A is shared along that batch axis (coordinate forced to 0)
  int b_bc1, b_bc2;
  int splits;              // K splits (>= 1); > 1 requires mode == TG_ATOMIC
  int k_per_split;         // multiple of TG_BK
  float alpha;
  const float* bias;       // nullptr or [N] (bias_on_m == 0) / [M] (bias_on_m == 1); added by split 0 only
  int bias_on_m;
  float* C;
  long long ldc, c_s1, c_s2;
  int mode;
  int relu;                // max(., 0) after the bias (forward of Linear + ReLU)
  int round_out;           // store C rounded to TF32 (nearest): C is only ever a GEMM operand again, whose rounding pass is then skipped
  int round_ops;           // bit 0: round operand A to TF32 (nearest) in shared memory before the MMA, bit 1: operand B; 0 = the tensor
                           // core truncates (gradient GEMMs: a 2^-11 relative shrink instead of an unbiased 2^-12 error, no smem pass)
  const float* relu_mask;  // nullptr or [M, N] (pitch ld_mask, same batch strides as C): C = 0 where relu_mask <= 0 (dgrad into a ReLU)
  long long ld_mask;
};

template <int BN, int STAGES = TG_STAGES>
struct TgCfg {
  static constexpr int kABytes = TG_BM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = STAGES * kStageBytes + 256 + 1024;   // + barriers + alignment slack
};

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::tf32: c_format F32 (1) [4,6); a/b_format TF32 (2) [7,10) / [10,13); a_major bit 15, b_major bit 16; N>>3 [17,23); M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Shared-memory descriptor of an MN-major fp32 (TF32) operand: 32-bit MN-major operands exist only in the
// SWIZZLE_128B_BASE32B layout (layout type 1: 32-byte chunks XOR-swizzled inside 128-byte rows, pattern period 4 rows;
// TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  Canonical form in 16-byte units
// ((8, n), (4, k)) : ((1, LBO), (8, SBO)): a K row holds 32 consecutive MN values (128 B), SBO = 4 rows = 512 B,
// LBO = distance to the next 32-value MN atom.
__device__ __forceinline__ uint64_t make_mnmajor_desc_f32(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)(512u >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}

// One 32-row x 32-column piece of the accumulator (row `lane` of the warp's TMEM quarter in r[]) -> C: alpha, bias, ReLU,
// TF32 rounding, ReLU mask, store / read-modify-write / atomic, transposed through the warp's [32 x 36] shared-memory pad
// into whole 128-byte row segments (four rows per store instruction; conflict-free 128-bit accesses on both sides).
__device__ __forceinline__ void tg_epilogue_chunk(const TgParams& p, const uint32_t (&r)[32], float* stg, int lane, int m_base,
                                                  int col0, bool with_bias, float bias_m, bool vec_ok, bool mask_vec,
                                                  float* cbase) {
  const int rr4 = lane >> 3, l8 = lane & 7;
  const int m0 = m_base, quarter = 0;       // m_base already includes the warp's quarter offset
  // ReLU mask of this chunk, fetched before the transpose and the stores (a load behind a store to C could not be
  // hoisted: the compiler must assume the two alias)
  float4 mk[8];
  if (p.relu_mask != nullptr) {
    const int mcol = col0 + 4 * l8;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int grow = m0 + quarter * 32 + 4 * it + rr4;
      mk[it] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (grow < p.M && mcol < p.N) {
        const float* mp = p.relu_mask + (long long)grow * p.ld_mask + mcol;
        if (mask_vec && mcol + 4 <= p.N) {
          mk[it] = __ldg(reinterpret_cast<const float4*>(mp));
        } else {
          mk[it].x = __ldg(mp);
          if (mcol + 1 < p.N) mk[it].y = __ldg(mp + 1);
          if (mcol + 2 < p.N) mk[it].z = __ldg(mp + 2);
          if (mcol + 3 < p.N) mk[it].w = __ldg(mp + 3);
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 o;
    o.x = p.alpha * __uint_as_float(r[4 * i]) + bias_m, o.y = p.alpha * __uint_as_float(r[4 * i + 1]) + bias_m;
    o.z = p.alpha * __uint_as_float(r[4 * i + 2]) + bias_m, o.w = p.alpha * __uint_as_float(r[4 * i + 3]) + bias_m;
    *reinterpret_cast<float4*>(stg + lane * 36 + 4 * i) = o;
  }
  __syncwarp();
  if (col0 < 32 && (threadIdx.x >> 5) == 2) TG_MARK(9);
  const int col = col0 + 4 * l8;
  float4 bn = make_float4(0.f, 0.f, 0.f, 0.f);
  if (with_bias && !p.bias_on_m) {
    if (col + 0 < p.N) bn.x = __ldg(p.bias + col);
    if (col + 1 < p.N) bn.y = __ldg(p.bias + col + 1);
    if (col + 2 < p.N) bn.z = __ldg(p.bias + col + 2);
    if (col + 3 < p.N) bn.w = __ldg(p.bias + col + 3);
  }
  if (col0 < 32 && (threadIdx.x >> 5) == 2) TG_MARK(10);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rl = 4 * it + rr4;
    const int grow = m0 + quarter * 32 + rl;
    if (grow >= p.M || col >= p.N) continue;
    float4 o = *reinterpret_cast<const float4*>(stg + rl * 36 + 4 * l8);
    o.x += bn.x, o.y += bn.y, o.z += bn.z, o.w += bn.w;
    if (p.relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
    if (p.round_out) {
      uint32_t t0, t1, t2, t3;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t0) : "f"(o.x));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t1) : "f"(o.y));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t2) : "f"(o.z));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t3) : "f"(o.w));
      o = make_float4(__uint_as_float(t0), __uint_as_float(t1), __uint_as_float(t2), __uint_as_float(t3));
    }
    if (p.relu_mask != nullptr) {
      if (!(mk[it].x > 0.f)) o.x = 0.f;
      if (!(mk[it].y > 0.f)) o.y = 0.f;
      if (!(mk[it].z > 0.f)) o.z = 0.f;
      if (!(mk[it].w > 0.f)) o.w = 0.f;
    }
    float* dst = cbase + (long long)grow * p.ldc + col;
    if (p.mode == TG_ATOMIC) {
      atomicAdd(dst, o.x);
      if (col + 1 < p.N) atomicAdd(dst + 1, o.y);
      if (col + 2 < p.N) atomicAdd(dst + 2, o.z);
      if (col + 3 < p.N) atomicAdd(dst + 3, o.w);
    } else if (vec_ok && col + 4 <= p.N) {
      if (p.mode == TG_ADD) {
        const float4 old = *reinterpret_cast<const float4*>(dst);
        o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
      }
      *reinterpret_cast<float4*>(dst) = o;
    } else {
      const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (col + k < p.N) dst[k] = (p.mode == TG_ADD ? dst[k] : 0.f) + ov[k];
    }
  }
}

template <int BN, bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(TG_THREADS, (STAGES == 2 ? 3 : 2))
tgemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, TgParams p) {
  using Cfg = TgCfg<BN, STAGES>;
  constexpr int TG_STAGES = STAGES;       // shadows the default depth inside the kernel
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TG_STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                 // [TG_STAGES]
  uint64_t* empty_bar = bars + TG_STAGES;    // [TG_STAGES]
  uint64_t* ready_bar = bars + 2 * TG_STAGES;   // [TG_STAGES] stage rounded to TF32 (128 arrivals)
  uint64_t* acc_full = bars + 3 * TG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TG_STAGES + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // n-tile fastest: the CTAs that share an A tile are launched back to back, so A comes from HBM once (measured with the
  // m-tile fastest: 838 MB of A read twice per 818176 x 256 x 256 GEMM, L2 hit rate 37 %)
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * TG_BM;
  int z = blockIdx.z;
  const int b1 = z % p.nb1;
  z /= p.nb1;
  const int b2 = z % p.nb2;
  const int split = z / p.nb2;
  const int k_begin = split * p.k_per_split;
  const int k_end = min(p.K, k_begin + p.k_per_split);
  const int n_kb = (k_end - k_begin + TG_BK - 1) / TG_BK;

  if (warp == 0) TG_MARK(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < TG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) TG_MARK(1);

  if (warp == 0) {
    if (elect_one()) {
      const int a1 = p.a_bc1 ? 0 : b1, a2 = p.a_bc2 ? 0 : b2;
      const int bb1 = p.b_bc1 ? 0 : b1, bb2 = p.b_bc2 ? 0 : b2;
      for (int kb = 0; kb < n_kb; ++kb) {
        const int st = kb % TG_STAGES;
        const uint32_t ph = (uint32_t)(kb / TG_STAGES) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        mbar_expect_tx(&full_bar[st], Cfg::kStageBytes);
        uint8_t* sa = smem + st * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        const int k0 = k_begin + kb * TG_BK;
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < TG_BM / 32; ++j) tma_load_4d(sa + j * TG_ATOM_BYTES, &tmap_a, &full_bar[st], m0 + 32 * j, k0, a1, a2);
        } else {
          tma_load_4d(sa, &tmap_a, &full_bar[st], k0, m0, a1, a2);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_4d(sb + j * TG_ATOM_BYTES, &tmap_b, &full_bar[st], n0 + 32 * j, k0, bb1, bb2);
        } else {
          tma_load_4d(sb, &tmap_b, &full_bar[st], k0, n0, bb1, bb2);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32(TG_BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      for (int kb = 0; kb < n_kb; ++kb) {
        const int st = kb % TG_STAGES;
        const uint32_t ph = (uint32_t)(kb / TG_STAGES) & 1;
        mbar_wait(p.round_ops ? &ready_bar[st] : &full_bar[st], ph);
        if (kb == 0) TG_MARK(2);
        if (kb == n_kb - 1) TG_MARK(6);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + st * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < TG_BK / 8; ++k) {      // 8 K-values per kind::tf32 instruction
          const uint64_t da = A_MN ? make_mnmajor_desc_f32(a_addr + k * 1024, TG_ATOM_BYTES) : make_kmajor_desc<128>(a_addr) + 2 * k;
          const uint64_t db = B_MN ? make_mnmajor_desc_f32(b_addr + k * 1024, TG_ATOM_BYTES) : make_kmajor_desc<128>(b_addr) + 2 * k;
          umma_tf32(tmem_base, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[st]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ===================== epilogue: TMEM lane quarter = warp % 4 =====================
    const int quarter = warp & 3;
    const int row = m0 + quarter * 32 + lane;
    const bool with_bias = (p.bias != nullptr) && split == 0;
    float* cbase = p.C + (long long)b1 * p.c_s1 + (long long)b2 * p.c_s2;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.c_s1 & 3) == 0) && ((p.c_s2 & 3) == 0);
    const float bias_m = (with_bias && p.bias_on_m && row < p.M) ? p.bias[row] : 0.f;
    const bool mask_vec = ((p.ld_mask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.relu_mask) & 15) == 0);
    // ---- main loop duty: round each landed stage to TF32 (nearest, ties away) in place
    const int et = threadIdx.x - 64;
    const int r_lo = (p.round_ops & 1) ? 0 : Cfg::kABytes / 16;
    const int r_hi = (p.round_ops & 2) ? Cfg::kStageBytes / 16 : Cfg::kABytes / 16;
    for (int kb = 0; p.round_ops != 0 && kb < n_kb; ++kb) {
      const int st = kb % TG_STAGES;
      mbar_wait(&full_bar[st], (uint32_t)(kb / TG_STAGES) & 1);
      uint4* sp = reinterpret_cast<uint4*>(smem + st * Cfg::kStageBytes);
#pragma unroll 4
      for (int i = r_lo + et; i < r_hi; i += 128) {
        uint4 v = sp[i];
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.x) : "f"(__uint_as_float(v.x)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.y) : "f"(__uint_as_float(v.y)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.z) : "f"(__uint_as_float(v.z)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.w) : "f"(__uint_as_float(v.w)));
        sp[i] = v;
      }
      fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(&ready_bar[st]);
    }
    if (n_kb > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after_sync();
    }
    if (warp == 2) TG_MARK(3);
    // every stage has been consumed (acc_full follows the last MMA): the ring doubles as the transpose buffer.
    // TMEM hands out one row per lane; a [32 x 36]-float pad per warp turns that into whole 128-byte row segments
    // (four rows per store instruction), conflict-free for the 128-bit accesses on both sides.
    float* stg = reinterpret_cast<float*>(smem) + quarter * (32 * 36);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      if (n_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (warp == 2 && c == 0) TG_MARK(8);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      const int col0 = n0 + c * 32;
      if (col0 >= p.N) break;                      // warp-uniform
      tg_epilogue_chunk(p, r, stg, lane, m0 + quarter * 32, col0, with_bias, bias_m, vec_ok, mask_vec, cbase);
      if (warp == 2 && c == 0) TG_MARK(11);
    }
  }
  if (warp == 2) TG_MARK(4);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
    TG_MARK(7);
  }
}

}  // namespace poem
