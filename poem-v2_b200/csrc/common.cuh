// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05 (UMMA/TMEM) wrappers.
// Everything here is inline PTX for Blackwell (compile with -gencode arch=compute_100a,code=sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace poem {

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still draining.  `pdl_wait` blocks until the predecessor grid has
// completed and its memory is visible (a no-op for a normal launch) and must precede every global access that depends
// on it; `pdl_trigger` lets the successor's CTAs be scheduled as soon as SM resources free up, so its prologue
// (barrier init, TMEM allocation, descriptor prefetch, weight / bias preloads) overlaps this kernel's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// 2^x on the SFU (MUFU.EX2), no range fix-up: inputs here are <= 0 after max subtraction
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes for -126 <= x <= ~1: x = n + f with n = rint(x) taken from the low mantissa bits of
// x + 1.5 * 2^23, a degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (max relative error 7.5e-5, far below the op16
// rounding of the probabilities it feeds), n added to the exponent field.  Used for a fraction of the softmax
// exponentials of the attention kernel, whose MUFU.EX2 issue rate (16 / clk / SM) is the measured limiter
// (profiles/r1_ncu_mha.md: 26 % of the warp samples sit on ex2 with stall reason mio_throttle).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.0551716648042202f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// 256-bit global accesses (sm_100: LDG.256 / STG.256), one full 32-byte sector per lane
__device__ __forceinline__ void ldg_nc_256(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ------------------------------------------------------------------------------------------------
// op16: the 16-bit tensor-core operand / activation format of the whole library = IEEE fp16.
// Round 1 used bfloat16 (8-bit significand); the north star's 1e-3 relative bound on the regressed coordinates needs
// TF32-class operands (scripts/precision_study2.py).  fp16 has the same 11-bit significand as TF32 at the full
// kind::f16 MMA rate and half of TF32's bytes.  Its range (6.1e-5 .. 65504 normal) is handled by saturating every
// float -> op16 conversion (F2FP.SATFINITE: no inf / NaN is ever produced from a finite value) and by per-row
// power-of-two scaling of the one stage whose magnitude is cubic in the activations (merge-net aggregate).
// ------------------------------------------------------------------------------------------------
typedef __half op16;
typedef __half2 op16x2;

// max(x, 0) fused into the conversion (first source -> upper half)
__device__ __forceinline__ uint32_t pack_op16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ uint32_t pack_op16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ op16 f2op16(float v) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return __ushort_as_half(h);
}
__device__ __forceinline__ float op16_to_f(op16 v) { return __half2float(v); }
__device__ __forceinline__ float2 op16x2_to_f2(op16x2 v) { return __half22float2(v); }
// halves of a packed pair held in a 32-bit register
__device__ __forceinline__ float op16_lo(uint32_t x) { return __half2float(__ushort_as_half((unsigned short)(x & 0xffffu))); }
__device__ __forceinline__ float op16_hi(uint32_t x) { return __half2float(__ushort_as_half((unsigned short)(x >> 16))); }

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires)
// instead of burning issue slots of its SM sub-partition in a polling loop.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// non-blocking probe (no suspend): for a thread that multiplexes several barriers
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifndef POEM_MBAR_MODE
#define POEM_MBAR_MODE 0   // 0: try_wait with a 10 ms suspend hint; 1: try_wait without a hint; 2: test_wait busy loop (experiments)
#endif
__device__ __forceinline__ bool mbar_try_wait_nohint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (POEM_MBAR_MODE == 0) {
    while (!mbar_try_wait(bar, parity)) {
    }
  } else if (POEM_MBAR_MODE == 1) {
    while (!mbar_try_wait_nohint(bar, parity)) {
    }
  } else {
    while (!mbar_test_wait(bar, parity)) {
    }
  }
}

// generic-proxy writes to smem -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — tensor maps are built on the host with cuTensorMapEncodeTiled
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, loads
// ------------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // whole warp
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; fp16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_op16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor for kind::f16 with FP16 A/B, FP32 accumulator, K-major A and B.
//   bits [4,6) c_format=1 (F32); [7,10) a_format=0 (F16; 1 = BF16); [10,13) b_format=0 (F16);
//   bit 15 a_major=0 (K), bit 16 b_major=0 (K); [17,23) N>>3; [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_op16(uint32_t M, uint32_t N, uint32_t b_mn_major = 0) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Shared-memory descriptor of an MN-major operand tile stored as [K rows x kRowBytes] (each K row holds kRowBytes/2
// consecutive MN elements; 128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B), rows packed back to back:
//   canonical layout ((8 x16B, n), (8, k)) : ((1, LBO), (kRowBytes/16, SBO))  ->  SBO = 8 rows, LBO = next MN atom.
template <uint32_t kRowBytes>
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  static_assert(kRowBytes == 128 || kRowBytes == 64, "row must be one swizzle atom wide");
  constexpr uint64_t layout = (kRowBytes == 128) ? 2ull : 4ull;
  constexpr uint64_t sbo = (8u * kRowBytes) >> 4;
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | (sbo << 32) |
         (1ull << 46) | (layout << 61);
}

// Shared-memory matrix descriptor, K-major operand, dense rows of `kRowBytes` (128 -> SWIZZLE_128B,
// 64 -> SWIZZLE_64B), 8-row groups packed back to back (SBO = 8 * kRowBytes). Tile base must be
// 1024-byte aligned so that the hardware XOR pattern matches what TMA wrote.
//   bits [0,14) addr>>4; [16,30) LBO>>4 (unused for swizzled K-major, set 1); [32,46) SBO>>4;
//   [46,48) version=1 (Blackwell); [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B.
template <uint32_t kRowBytes>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  static_assert(kRowBytes == 128 || kRowBytes == 64, "row must be one swizzle atom wide");
  constexpr uint64_t layout = (kRowBytes == 128) ? 2ull : 4ull;
  constexpr uint64_t sbo = (8u * kRowBytes) >> 4;
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// 32 lanes x 32 columns of fp32 from TMEM: thread i of the warp gets lane (taddr.lane + i), 32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_shared_b16(uint32_t addr, float v) {
  const unsigned short h = __half_as_ushort(f2op16(v));
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(h) : "memory");
}

// byte offset of element (row, 16-byte chunk) inside a [rows x 128B] SWIZZLE_128B K-major tile
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ (row & 7u)) << 4);
}

}  // namespace poem
