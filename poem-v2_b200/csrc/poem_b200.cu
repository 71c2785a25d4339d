// C-ABI implementation (include/poem_b200.h): host-side launch logic + whole-path orchestration.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/poem_b200.h"
#include "gemm.cuh"
#include "conv3x3.cuh"
#include "mha.cuh"
#include "hrnet.cuh"
#include "simt.cuh"
#include "sample_merge.cuh"
#include "qchain.cuh"
#include "vecattn.cuh"
#include "mano.cuh"

using namespace poem;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return fail(POEM_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)
#define POEM_TRY(expr)        \
  do {                        \
    int _r = (expr);          \
    if (_r != POEM_OK) return _r; \
  } while (0)

// ---- launch accounting + optional per-launch CUDA-event timing (poem_profile_*) ----
#include <atomic>
#include <map>
#include <string>
struct ProfRec {
  std::string tag;
  cudaEvent_t a, b;
};
static std::atomic<long long> g_launches{0};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static thread_local const char* g_tag = nullptr;   // stage label for the generic GEMM launches
static thread_local cudaEvent_t g_prof_a = nullptr;
static thread_local cudaStream_t g_prof_stream = nullptr;
static void prof_begin(cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventCreate(&g_prof_a);
  cudaEventRecord(g_prof_a, st);
  g_prof_stream = st;
}
static void prof_end(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on || g_prof_a == nullptr) return;
  ProfRec r;
  r.tag = g_tag ? (std::string(name) + ":" + g_tag) : std::string(name);
  r.a = g_prof_a;
  cudaEventCreate(&r.b);
  cudaEventRecord(r.b, g_prof_stream);
  g_prof.push_back(r);
  g_prof_a = nullptr;
}
struct TagScope {
  const char* prev;
  explicit TagScope(const char* t) : prev(g_tag) { g_tag = t; }
  ~TagScope() { g_tag = prev; }
};
#define LAUNCH_CHECK(name)                                                                       \
  do {                                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (g_launch_err != cudaSuccess) _e = g_launch_err, g_launch_err = cudaSuccess;              \
    prof_end(name);                                                                              \
    if (_e != cudaSuccess) return fail(POEM_E_CUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while its predecessor drains; every
// kernel launched this way calls pdl_wait() before its first dependent global access (common.cuh).  POEM_PDL=0 in the
// environment falls back to plain launches.
static int split_parts() {   // POEM_SPLIT=<n>: batch slices whose decoder blocks run concurrently (default 2; 1 = one stream)
  static int n = -1;
  if (n < 0) {
    const char* e = getenv("POEM_SPLIT");
    n = e ? atoi(e) : 2;
    if (n < 1) n = 1;
    if (n > 4) n = 4;
  }
  return n;
}
static bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("POEM_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
static thread_local cudaError_t g_launch_err = cudaSuccess;
template <typename... Params, typename... Args>
static void launch_pdl(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && !g_prof_on) ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
  if (e != cudaSuccess) g_launch_err = e;
}

extern "C" long long poem_kernel_launches(void) { return g_launches.load(); }
extern "C" void poem_profile_enable(int on) {
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_on = (on != 0);
}
// JSON object {"<kernel>:<stage>": {"ms": total, "n": launches}, ...}; returns bytes written (0 if buf too small)
extern "C" size_t poem_profile_summary(char* buf, size_t cap) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<double, int>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      agg[r.tag].first += ms;
      agg[r.tag].second += 1;
    }
  }
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s\"%s\": {\"ms\": %.6f, \"n\": %d}", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    out += line;
    first = false;
  }
  out += "}";
  if (out.size() + 1 > cap) return 0;
  memcpy(buf, out.c_str(), out.size() + 1);
  return out.size();
}

extern "C" int poem_abi_version(void) { return POEM_ABI_VERSION; }
extern "C" const char* poem_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point fetched at run time: no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D op16 tensor [rows, cols] with row pitch ld (elements); box = [box_cols, box_rows]; swizzle by box width.
static int make_tmap_op16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                          uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(POEM_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15))
    return fail(POEM_E_ALIGN, "TMA operand needs 16-byte aligned base and pitch (base=%p ld=%llu)", base,
                (unsigned long long)ld);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = (box_cols * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (box_cols * 2 == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                          : (box_cols * 2 == 32) ? CU_TENSOR_MAP_SWIZZLE_32B
                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
  if (sw == CU_TENSOR_MAP_SWIZZLE_NONE) return fail(POEM_E_BADDIM, "unsupported TMA box width %u", box_cols);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(POEM_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return POEM_OK;
}

static int num_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// ------------------------------------------------------------------------------------------------
// workspace bump allocator
// ------------------------------------------------------------------------------------------------
struct Bump {
  uint8_t* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    off = (off + 1023) & ~size_t(1023);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

// ------------------------------------------------------------------------------------------------
// GEMM launch
// ------------------------------------------------------------------------------------------------
template <int BN>
static int launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tw, int M, int N, int K, const GemmEpilogue& ep,
                          const ConvOperand& conv, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_op16_tc_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemMax));
    CUDA_TRY(cudaFuncSetAttribute(gemm_op16_tc_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemMax));
    configured = true;
  }
  // the epilogue moves 32-column chunks of a row with 256-bit accesses
  auto misaligned = [](const void* p, int ld, int elem) {
    return p != nullptr && ((reinterpret_cast<uintptr_t>(p) & 31) || ((size_t)ld * elem) % 32);
  };
  if (N % 32 || misaligned(ep.out_f32, ep.ld_f32, 4) || misaligned(ep.out_op16, ep.ld_op16, 2) ||
      (ep.res_mode != RES_NONE && (misaligned(ep.res_f32, ep.res_ld, 4) || misaligned(ep.res_op16, ep.res_ld, 2))))
    return fail(POEM_E_ALIGN, "gemm: N %% 32 == 0 and 32-byte aligned output / residual rows required (N=%d)", N);
  const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM, tiles_n = (N + BN - 1) / BN;
  const int tiles = tiles_m * tiles_n;
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
  // W-stationary when the [BN x K] slab plus >= 3 A stages fit, and every CTA gets
  // at least two M tiles (otherwise the slab load is not amortised)
  GemmPipe pipe;
  const long long w_slab = (long long)k_blocks * Cfg::kWBytes;
  const long long room = (long long)Cfg::kSmemMax - Cfg::kTailBytes - w_slab;
  int grid, smem;
  if (room >= 3 * Cfg::kABytes && tiles_n <= num_sms() && tiles_m >= 2 * (num_sms() / tiles_n)) {
    pipe.w_stationary = 1;
    pipe.n_stages = (int)(room / Cfg::kABytes);
    if (pipe.n_stages > GEMM_MAX_STAGES) pipe.n_stages = GEMM_MAX_STAGES;
    grid = (num_sms() / tiles_n) * tiles_n;
    smem = (int)w_slab + pipe.n_stages * Cfg::kABytes + Cfg::kTailBytes;
  } else {
    pipe.w_stationary = 0;
    pipe.n_stages = Cfg::kStages;
    grid = tiles < num_sms() ? tiles : num_sms();
    smem = Cfg::kSmemBytes;
  }
  prof_begin(st);
  if (ep.res_mode != RES_NONE)
    launch_pdl(gemm_op16_tc_kernel<BN, true>, dim3(grid), dim3(GEMM_THREADS), (size_t)(smem), st, ta, tw, M, N, K, ep, conv, pipe);
  else
    launch_pdl(gemm_op16_tc_kernel<BN, false>, dim3(grid), dim3(GEMM_THREADS), (size_t)(smem), st, ta, tw, M, N, K, ep, conv, pipe);
  LAUNCH_CHECK("gemm_op16_tc_kernel");
  return POEM_OK;
}

static GemmEpilogue epi_default(int N) {
  GemmEpilogue e;
  memset(&e, 0, sizeof(e));
  e.trans_from = N;
  e.t_rows = 1;
  e.n_store = N;
  return e;
}

static int launch_gemm(const op16* A, int lda, const op16* W, int ldw, int M, int N, int K,
                       const GemmEpilogue& ep, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return fail(POEM_E_BADDIM, "gemm: bad shape %d %d %d", M, N, K);
  if (!A || !W) return fail(POEM_E_NULL, "gemm: null operand");
  if (N % 32) return fail(POEM_E_BADDIM, "gemm: N=%d must be a multiple of 32", N);
  if ((ep.out_f32 && (ep.ld_f32 % 4)) || (ep.out_op16 && (ep.ld_op16 % 8)) ||
      (ep.res_mode == RES_F32 && (ep.res_ld % 4)))
    return fail(POEM_E_ALIGN, "gemm: output/residual leading dimensions must keep rows 16-byte aligned");
  // widest tile that divides N, narrowed while the grid would be under ~2 waves (small-M GEMMs of the query stream:
  // more, smaller tiles keep the TMEM double buffering and the TMA ring busy instead of one serial tile per CTA)
  int BN = (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : 64;
  const int tiles_m_ = (M + GEMM_BM - 1) / GEMM_BM;
  while (BN > 64 && tiles_m_ * (N / BN) < 3 * num_sms()) BN >>= 1;
  CUtensorMap ta, tw;
  POEM_TRY(make_tmap_op16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BK, GEMM_BM));
  POEM_TRY(make_tmap_op16(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, GEMM_BK, (uint32_t)BN));
  ConvOperand none;
  memset(&none, 0, sizeof(none));
  if (BN == 256) return launch_gemm_bn<256>(ta, tw, M, N, K, ep, none, st);
  if (BN == 128) return launch_gemm_bn<128>(ta, tw, M, N, K, ep, none, st);
  return launch_gemm_bn<64>(ta, tw, M, N, K, ep, none, st);
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM convolution on NHWC op16 (HRNet stage 4)
// ------------------------------------------------------------------------------------------------
// 4-D op16 tensor map over an NHWC activation tensor: dims (C, W, H, N); box (64, bw*stride, bh*stride, bn) with
// traversal stride `stride` along W and H, SWIZZLE_128B, zero fill outside (= convolution padding).
static int make_tmap_nhwc(CUtensorMap* tm, const void* base, int N, int H, int W, int Cp, int bw, int bh, int bn,
                          int stride, int box_c = 64) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(POEM_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (Cp % 8)) return fail(POEM_E_ALIGN, "conv: bad activation tensor");
  cuuint64_t gdim[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
  const CUtensorMapSwizzle swz = box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(POEM_E_CUDA, "cuTensorMapEncodeTiled (4d) failed (%d)", (int)r);
  return POEM_OK;
}

// 3x3 stride-1 C -> C convolution with halo reuse (conv3x3.cuh)
static int g_conv_mode = 0;   // 0: halo-reuse kernel, live channels only; 1: halo-reuse, all padded channels; 2: generic path
extern "C" void poem_debug_conv_mode(int mode) { g_conv_mode = mode; }

template <int CIP, int CIR, int CP, int CR>
static int launch_conv3x3_halo_cp(const op16* in, int N, int R, int cin_s, const PoemLinear& wt, const HaloArgs& a,
                                  cudaStream_t st) {
  using Cfg = HaloCfg<CIP, CIR, CP, CR>;
  using Blk = typename Cfg::Blk;
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(conv3x3_halo_kernel<CIP, CIR, CP, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    attr_done = true;
  }
  HaloMaps maps;
  bool have[3] = {false, false, false};
  for (int b = 0; b < Cfg::kNB; ++b) {
    const int nch = Blk::nch(b), mi = halo_map_index(nch);
    if (have[mi]) continue;
    POEM_TRY(make_tmap_nhwc(&maps.x[mi], in, N, R, R, cin_s, Cfg::kPitch, Cfg::kRows, 1, 1, nch));
    POEM_TRY(make_tmap_op16(&maps.w[mi], wt.w, (uint64_t)CR, (uint64_t)9 * CIP, (uint64_t)9 * CIP, (uint32_t)nch, (uint32_t)CR));
    have[mi] = true;
  }
  int first = have[0] ? 0 : (have[1] ? 1 : 2);
  for (int mi = 0; mi < 3; ++mi)
    if (!have[mi]) maps.x[mi] = maps.x[first], maps.w[mi] = maps.w[first];
  const int tiles = N * (R / 16) * (R / 16);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  prof_begin(st);
  conv3x3_halo_kernel<CIP, CIR, CP, CR><<<grid, HALO_THREADS, Cfg::kSmemBytes, st>>>(maps, a);
  LAUNCH_CHECK("conv3x3_halo_kernel");
  return POEM_OK;
}

// Shapes the halo-reuse kernel is instantiated for: (padded in, live in, padded out, live out), live counts rounded to 16
static bool halo_supported(int cip, int cir, int cop, int cor) {
  const long long key = ((long long)cip << 48) | ((long long)cir << 32) | ((long long)cop << 16) | cor;
  auto k = [](long long a, long long b, long long c, long long d) { return (a << 48) | (b << 32) | (c << 16) | d; };
  return key == k(64, 48, 64, 48) || key == k(64, 64, 64, 64) || key == k(128, 80, 128, 80) ||
         key == k(128, 128, 128, 128) || key == k(192, 160, 192, 160) || key == k(192, 192, 192, 192) ||
         key == k(256, 256, 64, 48) || key == k(256, 240, 128, 80) || key == k(128, 128, 64, 48);
}

// ci_live / co_live: live channels of the input / output (the rest of the padded count is zero); 0 = all live
static int launch_conv3x3_halo(const op16* in, int N, int R, int Cip, int ci_live, int Cop, int co_live,
                               const PoemLinear& wt, bool relu, const op16* res, op16* out,
                               cudaStream_t st, bool* handled, int cin_s, int cout_s) {
  const bool live = (g_conv_mode != 1);
  int cir = (ci_live > 0 && live) ? (ci_live + 15) / 16 * 16 : Cip;
  int cor = (co_live > 0 && live) ? (co_live + 15) / 16 * 16 : Cop;
  if (!halo_supported(Cip, cir, Cop, cor)) cir = Cip, cor = Cop;
  // compact storage keeps only the live channels (rounded to 16) of every pixel: the kernel must not touch more
  *handled = halo_supported(Cip, cir, Cop, cor) && cir <= cin_s && cor <= cout_s;
  if (!*handled) return POEM_OK;
  HaloArgs a;
  a.n_images = N, a.R = R, a.bias = wt.b, a.relu = relu ? 1 : 0, a.res = res, a.out = out, a.cout_s = cout_s;
  char tag[64];
  snprintf(tag, sizeof(tag), "conv3x3halo_c%d_of_%d_to_c%d_of_%d_r%d", cir, Cip, cor, Cop, R);
  TagScope ts(tag);
  const long long key = ((long long)Cip << 48) | ((long long)cir << 32) | ((long long)Cop << 16) | cor;
#define HALO_CASE(a_, b_, c_, d_)                                                             \
  if (key == (((long long)(a_) << 48) | ((long long)(b_) << 32) | ((long long)(c_) << 16) | (d_))) \
    return launch_conv3x3_halo_cp<a_, b_, c_, d_>(in, N, R, cin_s, wt, a, st);
  HALO_CASE(64, 48, 64, 48)
  HALO_CASE(64, 64, 64, 64)
  HALO_CASE(128, 80, 128, 80)
  HALO_CASE(128, 128, 128, 128)
  HALO_CASE(192, 160, 192, 160)
  HALO_CASE(192, 192, 192, 192)
  HALO_CASE(256, 256, 64, 48)
  HALO_CASE(256, 240, 128, 80)
  HALO_CASE(128, 128, 64, 48)
#undef HALO_CASE
  return fail(POEM_E_BADDIM, "conv3x3 halo: no instantiation");
}

// out[N, Hout, Wout, Cout_p] = act(conv(in[N, Hin, Win, Cin_p], w[Cout_p, k*k*Cin_p]) + b) (+ res)
static int launch_conv(const op16* in, int N, int Hin, int Win, int Cin_p, const PoemLinear& wt, int Cout_p,
                       int ksize, int stride, bool relu, const op16* res, op16* out,
                       cudaStream_t st, int c_real = 0, bool relu_before_res = false, float* out_f32 = nullptr,
                       int c_real_in = -1, int cin_s = 0, int cout_s = 0) {
  // cin_s / cout_s: channels per pixel of the tensors in memory (multiples of 16; 0 = the padded counts).  With
  // compact storage a 64-channel TMA box simply runs past the pixel's channels and the hardware zero-fills the rest.
  if (cin_s <= 0) cin_s = Cin_p;
  if (cout_s <= 0) cout_s = Cout_p;
  if (cin_s % 16 || cout_s % 16 || cin_s > Cin_p || cout_s > Cout_p)
    return fail(POEM_E_BADDIM, "conv: storage channels %d/%d for padded %d/%d", cin_s, cout_s, Cin_p, Cout_p);
  if (!wt.w || !wt.b) return fail(POEM_E_NULL, "conv: weight pointer missing");
  if (!(ksize == 1 || ksize == 3) || !(stride == 1 || stride == 2) || (ksize == 1 && stride != 1))
    return fail(POEM_E_BADDIM, "conv: unsupported kernel %d / stride %d", ksize, stride);
  // c_real: live output channels; c_real_in: live input channels (defaults to c_real, the C -> C BasicBlock case)
  if (g_conv_mode != 2 && ksize == 3 && stride == 1 && Hin == Win && Hin % 16 == 0 && ((Hin / 16) & (Hin / 16 - 1)) == 0 &&
      !relu_before_res &&
      out != nullptr && out_f32 == nullptr) {
    bool handled = false;
    POEM_TRY(launch_conv3x3_halo(in, N, Hin, Cin_p, c_real_in >= 0 ? c_real_in : c_real, Cout_p, c_real, wt, relu, res, out,
                                 st, &handled, cin_s, cout_s));
    if (handled) return POEM_OK;
  }
  const int Hout = Hin / stride, Wout = Win / stride;
  if (Wout < 1 || 128 % Wout || Wout > 128 || Cin_p % 64 || Cout_p % 32)
    return fail(POEM_E_BADDIM, "conv: unsupported shape %dx%d C %d -> %d", Hin, Win, Cin_p, Cout_p);
  int bh = 128 / Wout, bn = 1;
  if (bh > Hout) {
    bn = bh / Hout;
    bh = Hout;
  }
  if (bh * stride > 256 || Wout * stride > 256) return fail(POEM_E_BADDIM, "conv: TMA box too large");
  const int taps = ksize * ksize, cblocks = Cin_p / 64;
  const int M = N * Hout * Wout, K = taps * Cin_p;
  CUtensorMap ta, tw;
  POEM_TRY(make_tmap_nhwc(&ta, in, N, Hin, Win, cin_s, Wout, bh, bn, stride));
  const int BN = (Cout_p <= 256) ? Cout_p : 160;
  if (!(BN == 64 || BN == 128 || BN == 160 || BN == 192 || BN == 256) || Cout_p % BN)
    return fail(POEM_E_BADDIM, "conv: Cout_p=%d has no tile", Cout_p);
  POEM_TRY(make_tmap_op16(&tw, wt.w, (uint64_t)Cout_p, (uint64_t)K, (uint64_t)K, GEMM_BK, (uint32_t)BN));
  GemmEpilogue e = epi_default(Cout_p);
  e.bias = wt.b;
  e.act = (relu && (!res || relu_before_res)) ? ACT_RELU : ACT_NONE;
  e.act_after_res = (relu && res && !relu_before_res) ? ACT_RELU : ACT_NONE;
  if (res) {
    e.res_mode = RES_BF16;
    e.res_op16 = res;
    e.res_ld = cout_s;
  }
  e.out_op16 = out;
  e.ld_op16 = cout_s;
  e.n_store = out_f32 ? Cout_p : cout_s;   // the fp32 output (feat_in) keeps the padded row
  e.out_f32 = out_f32;
  e.ld_f32 = Cout_p;
  ConvOperand cv;
  cv.enabled = 1, cv.ksize = ksize, cv.pad = ksize / 2, cv.stride = stride, cv.cblocks = cblocks, cv.Hout = Hout,
  cv.Wout = Wout;
  char tag[48];
  snprintf(tag, sizeof(tag), "conv%dx%d%s_c%d_to_%d", ksize, ksize, stride == 2 ? "s2" : "", Cin_p, Cout_p);
  TagScope ts(tag);
  switch (BN) {
    case 64: return launch_gemm_bn<64>(ta, tw, M, Cout_p, K, e, cv, st);
    case 128: return launch_gemm_bn<128>(ta, tw, M, Cout_p, K, e, cv, st);
    case 160: return launch_gemm_bn<160>(ta, tw, M, Cout_p, K, e, cv, st);
    case 192: return launch_gemm_bn<192>(ta, tw, M, Cout_p, K, e, cv, st);
    default: return launch_gemm_bn<256>(ta, tw, M, Cout_p, K, e, cv, st);
  }
}

extern "C" int poem_conv_nhwc(const poem_op16* in, int N, int H, int W, int Cin_p, const poem_op16* w, const float* b,
                              int Cout_p, int ksize, int stride, int relu, const poem_op16* res, poem_op16* out,
                              int c_live_in, int c_live_out, void* stream) {
  if (!in || !out) return fail(POEM_E_NULL, "conv: null pointer");
  if (c_live_in < 0 || c_live_in > Cin_p || c_live_out < 0 || c_live_out > Cout_p)
    return fail(POEM_E_BADDIM, "conv: c_live=%d/%d", c_live_in, c_live_out);
  PoemLinear wt;
  wt.w = w;
  wt.b = b;
  return launch_conv(reinterpret_cast<const op16*>(in), N, H, W, Cin_p, wt, Cout_p, ksize, stride, relu != 0,
                     reinterpret_cast<const op16*>(res), reinterpret_cast<op16*>(out),
                     (cudaStream_t)stream, c_live_out, false, nullptr, c_live_in);
}

static inline int pad64(int c) { return (c + 63) / 64 * 64; }   // weight / K-block padding
static inline int pad16(int c) { return (c + 15) / 16 * 16; }   // channels per pixel kept in memory

struct HrPlan {
  op16* x[4][3];     // per branch: current / scratch / next
  op16* term[4][4];  // fuse terms (i, j != i) at the size of branch i
  op16* chain[2];    // intermediates of the stride-2 chains
};
static size_t hr_plan(int N, int R0, const int* ch, uint8_t* base, HrPlan* p) {
  Bump b{base, 0};
  size_t chain_max = 0;
  for (int i = 0; i < 4; ++i) {
    const size_t n = (size_t)N * (R0 >> i) * (R0 >> i) * pad16(ch[i]);
    for (int k = 0; k < 3; ++k) p->x[i][k] = b.take<op16>(n);
    for (int j = 0; j < 4; ++j) p->term[i][j] = (j == i) ? nullptr : b.take<op16>(n);
    if (i >= 1 && i <= 2) {
      int cmax = 0;   // chain intermediates at this resolution keep the source branch's channel count (j < i)
      for (int j = 0; j < i; ++j) cmax = ch[j] > cmax ? ch[j] : cmax;
      const size_t c = (size_t)N * (R0 >> i) * (R0 >> i) * pad16(cmax);
      chain_max = c > chain_max ? c : chain_max;
    }
  }
  for (int k = 0; k < 2; ++k) p->chain[k] = b.take<op16>(chain_max > 0 ? chain_max : 64);
  return b.off + 1024;
}

extern "C" size_t poem_hrnet_stage4_workspace_bytes(const PoemHRStage4* w, int n_images, int base_res) {
  if (!w || n_images < 1 || base_res < 8) return 0;
  HrPlan p;
  return hr_plan(n_images, base_res, w->channels, nullptr, &p);
}

// n_modules HighResolutionModules over the first nb branches (hrnet.py:217-234); cur[b] = live buffer of branch b
// Cw: channel counts padded to 64 (weight layout, K blocks); Cs: channels per pixel in memory (padded to 16)
static int run_hr_modules(const PoemHRModule* mods, int n_modules, int nb, int N, const int* R, const int* Cw,
                          const int* ch, const HrPlan& p, int* cur, cudaStream_t st) {
  int Cs[4];
  for (int i = 0; i < 4; ++i) Cs[i] = pad16(ch[i]);
  for (int m = 0; m < n_modules; ++m) {
    const PoemHRModule& mod = mods[m];
    // ---- branches: 4 BasicBlocks each (hrnet.py:38-67)
    for (int b = 0; b < nb; ++b) {
      for (int k = 0; k < 4; ++k) {
        op16* x = p.x[b][cur[b]];
        op16* t = p.x[b][(cur[b] + 1) % 3];
        op16* y = p.x[b][(cur[b] + 2) % 3];
        POEM_TRY(launch_conv(x, N, R[b], R[b], Cw[b], mod.branch[b][k][0], Cw[b], 3, 1, true, nullptr, t, st, ch[b], false, nullptr, -1,
                             Cs[b], Cs[b]));
        POEM_TRY(launch_conv(t, N, R[b], R[b], Cw[b], mod.branch[b][k][1], Cw[b], 3, 1, true, x, y, st, ch[b], false, nullptr, -1,
                             Cs[b], Cs[b]));
        cur[b] = (cur[b] + 2) % 3;
      }
    }
    // ---- fuse layers (hrnet.py:177-207, 225-233)
    for (int i = 0; i < nb; ++i) {
      FuseSumArgs fa;
      fa.n_in = 0;
      for (int j = 0; j < nb; ++j) {
        const op16* xj = p.x[j][cur[j]];
        if (j == i) {
          fa.in[fa.n_in] = xj, fa.shift[fa.n_in] = 0;
        } else if (j > i) {   // 1x1 conv + BN at the low resolution, upsampled by the sum kernel
          POEM_TRY(launch_conv(xj, N, R[j], R[j], Cw[j], mod.fuse[i][j][0], Cw[i], 1, 1, false, nullptr, p.term[i][j], st, 0, false,
                               nullptr, -1, Cs[j], Cs[i]));
          fa.in[fa.n_in] = p.term[i][j], fa.shift[fa.n_in] = j - i;
        } else {              // chain of (i - j) stride-2 3x3 convs; all but the last keep C_j channels and ReLU
          const op16* src = xj;
          int r = R[j];
          for (int k = 0; k < i - j; ++k) {
            const bool last = (k == i - j - 1);
            op16* dst = last ? p.term[i][j] : p.chain[k & 1];
            POEM_TRY(launch_conv(src, N, r, r, Cw[j], mod.fuse[i][j][k], last ? Cw[i] : Cw[j], 3, 2, !last, nullptr, dst, st, 0, false,
                                 nullptr, -1, Cs[j], last ? Cs[i] : Cs[j]));
            src = dst;
            r >>= 1;
          }
          fa.in[fa.n_in] = p.term[i][j], fa.shift[fa.n_in] = 0;
        }
        ++fa.n_in;
      }
      op16* dst = p.x[i][(cur[i] + 1) % 3];
      const size_t total16 = (size_t)N * R[i] * R[i] * Cs[i] / 16;
      prof_begin(st);
      fuse_sum_relu_kernel<<<(unsigned)((total16 + 255) / 256), 256, 0, st>>>(fa, dst, R[i], R[i], Cs[i], total16);
      LAUNCH_CHECK("fuse_sum_relu_kernel");
    }
    for (int i = 0; i < nb; ++i) cur[i] = (cur[i] + 1) % 3;
  }
  return POEM_OK;
}

static int hr_export(const HrPlan& p, const int* cur, const int* ch, const int* R, int N, float* const* out,
                     cudaStream_t st) {
  for (int i = 0; i < 4; ++i) {
    const int cs = pad16(ch[i]);
    dim3 grid((R[i] * R[i] + 31) / 32, (cs + 31) / 32, N), block(32, 8);
    prof_begin(st);
    nhwc_op16_to_nchw_f32_kernel<<<grid, block, 0, st>>>(p.x[i][cur[i]], out[i], ch[i], cs, R[i] * R[i]);
    LAUNCH_CHECK("nhwc_op16_to_nchw_f32_kernel");
  }
  return POEM_OK;
}

extern "C" int poem_hrnet_stage4_forward(const PoemHRStage4* w, int n_images, int base_res, const float* const* in,
                                         float* const* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!w || !in || !out || !workspace) return fail(POEM_E_NULL, "hrnet_stage4: null pointer");
  if (w->n_modules < 1 || w->n_modules > POEM_HR_MAX_MODULES) return fail(POEM_E_BADDIM, "n_modules=%d", w->n_modules);
  if (base_res != 64 && base_res != 32 && base_res != 128) return fail(POEM_E_BADDIM, "base_res=%d", base_res);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(POEM_E_ALIGN, "workspace must be 1024-byte aligned");
  const int N = n_images, R0 = base_res;
  const int* ch = w->channels;
  cudaStream_t st = (cudaStream_t)stream;
  HrPlan p;
  const size_t need = hr_plan(N, R0, ch, reinterpret_cast<uint8_t*>(workspace), &p);
  if (need > workspace_bytes) return fail(POEM_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, need);
  int Cp[4], R[4];
  for (int i = 0; i < 4; ++i) {
    Cp[i] = pad64(ch[i]);
    R[i] = R0 >> i;
    if (!in[i] || !out[i]) return fail(POEM_E_NULL, "hrnet_stage4: branch %d pointer missing", i);
    const int cs = pad16(ch[i]);
    dim3 grid((R[i] * R[i] + 31) / 32, (cs + 31) / 32, N), block(32, 8);
    prof_begin(st);
    nchw_f32_to_nhwc_op16_kernel<<<grid, block, 0, st>>>(in[i], p.x[i][0], ch[i], cs, R[i] * R[i]);
    LAUNCH_CHECK("nchw_f32_to_nhwc_op16_kernel");
  }
  int cur[4] = {0, 0, 0, 0};   // index of the buffer holding the branch's current activation
  POEM_TRY(run_hr_modules(w->modules, w->n_modules, 4, N, R, Cp, ch, p, cur, st));
  return hr_export(p, cur, ch, R, N, out, st);
}

// ------------------------------------------------------------------------------------------------
// whole HRNet-W40 backbone (hrnet.py:385-420): stem, layer1 (4 Bottlenecks), transitions, stages 2-4
// ------------------------------------------------------------------------------------------------
struct HrNetPlan {
  op16 *s1, *s2;        // stem outputs: (N,R/2,R/2,64), (N,R/4,R/4,64)
  op16 *l256[3], *l64[2];   // layer1 activations at R/4: 256 and 64 channels
  HrPlan hr;
};
static size_t hrnet_plan(int N, int img_res, const int* ch, uint8_t* base, HrNetPlan* p) {
  Bump b{base, 0};
  const size_t r2 = (size_t)(img_res / 2) * (img_res / 2), r4 = (size_t)(img_res / 4) * (img_res / 4);
  p->s1 = b.take<op16>((size_t)N * r2 * 64);
  p->s2 = b.take<op16>((size_t)N * r4 * 64);
  for (int k = 0; k < 3; ++k) p->l256[k] = b.take<op16>((size_t)N * r4 * 256);
  for (int k = 0; k < 2; ++k) p->l64[k] = b.take<op16>((size_t)N * r4 * 64);
  const size_t off = (b.off + 1023) & ~size_t(1023);
  const size_t hr = hr_plan(N, img_res / 4, ch, base ? base + off : nullptr, &p->hr);
  return off + hr;
}
extern "C" size_t poem_hrnet_workspace_bytes(const PoemHRNet* w, int n_images, int img_res) {
  if (!w || n_images < 1 || img_res < 64 || img_res % 32) return 0;
  HrNetPlan p;
  return hrnet_plan(n_images, img_res, w->channels, nullptr, &p);
}

// stem .. stage 4 on the planned workspace; leaves branch b's map in p.hr.x[b][cur[b]] (NHWC op16, padded channels)
static int hrnet_run(const PoemHRNet* w, int N, int img_res, const float* images, const HrNetPlan& p, int* cur,
                     cudaStream_t st) {
  const int* ch = w->channels;
  const int R2 = img_res / 2, R4 = img_res / 4;
  // stem: conv1 3->64 s2 (direct, fp32 weights) ; conv2 64->64 s2
  {
    const size_t pixels = (size_t)N * R2 * R2;
    prof_begin(st);
    stem_conv1_kernel<<<(unsigned)((pixels + 127) / 128), 128, 0, st>>>(images, w->stem1_w, w->stem1_b, p.s1, N, img_res,
                                                                      img_res);
    LAUNCH_CHECK("stem_conv1_kernel");
  }
  POEM_TRY(launch_conv(p.s1, N, R2, R2, 64, w->stem2, 64, 3, 2, true, nullptr, p.s2, st));
  // layer1: Bottleneck x4 (hrnet.py:70-104, 254-257)
  const op16* x = p.s2;
  int x_ch = 64;
  for (int k = 0; k < 4; ++k) {
    const PoemBottleneck& bt = w->layer1[k];
    POEM_TRY(launch_conv(x, N, R4, R4, x_ch, bt.c1, 64, 1, 1, true, nullptr, p.l64[0], st));
    POEM_TRY(launch_conv(p.l64[0], N, R4, R4, 64, bt.c2, 64, 3, 1, true, nullptr, p.l64[1], st));
    const op16* res = x;
    op16* y = p.l256[k % 2];
    if (bt.ds.w) {
      POEM_TRY(launch_conv(x, N, R4, R4, x_ch, bt.ds, 256, 1, 1, false, nullptr, p.l256[2], st));
      res = p.l256[2];
    } else if (x_ch != 256) {
      return fail(POEM_E_BADDIM, "hrnet: layer1 block %d needs a downsample", k);
    }
    POEM_TRY(launch_conv(p.l64[1], N, R4, R4, 64, bt.c3, 256, 1, 1, true, res, y, st));
    x = y;
    x_ch = 256;
  }
  int Cp[4], R[4];
  for (int i = 0; i < 4; ++i) {
    Cp[i] = pad64(ch[i]);
    R[i] = R4 >> i;
    cur[i] = 0;
  }
  // transition1 (hrnet.py:318-342): 3x3 256->40 ; 3x3 s2 256->80
  POEM_TRY(launch_conv(x, N, R4, R4, 256, w->trans1[0], Cp[0], 3, 1, true, nullptr, p.hr.x[0][0], st, ch[0], false, nullptr, 256, 256,
                       pad16(ch[0])));
  POEM_TRY(launch_conv(x, N, R4, R4, 256, w->trans1[1], Cp[1], 3, 2, true, nullptr, p.hr.x[1][0], st, 0, false, nullptr, -1, 256,
                       pad16(ch[1])));
  POEM_TRY(run_hr_modules(w->stage2, 1, 2, N, R, Cp, ch, p.hr, cur, st));
  // transition2: new branch from the lowest-resolution output, 3x3 s2 80->160
  POEM_TRY(launch_conv(p.hr.x[1][cur[1]], N, R[1], R[1], Cp[1], w->trans2, Cp[2], 3, 2, true, nullptr, p.hr.x[2][0], st, 0, false,
                       nullptr, -1, pad16(ch[1]), pad16(ch[2])));
  POEM_TRY(run_hr_modules(w->stage3, 4, 3, N, R, Cp, ch, p.hr, cur, st));
  // transition3: 3x3 s2 160->320
  POEM_TRY(launch_conv(p.hr.x[2][cur[2]], N, R[2], R[2], Cp[2], w->trans3, Cp[3], 3, 2, true, nullptr, p.hr.x[3][0], st, 0, false,
                       nullptr, -1, pad16(ch[2]), pad16(ch[3])));
  return run_hr_modules(w->stage4, 3, 4, N, R, Cp, ch, p.hr, cur, st);
}

static int hrnet_check(const PoemHRNet* w, int img_res, const void* images, const void* workspace) {
  if (!w || !images || !workspace) return fail(POEM_E_NULL, "hrnet: null pointer");
  if (img_res != 256) return fail(POEM_E_BADDIM, "hrnet: img_res=%d (256 supported)", img_res);
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(POEM_E_ALIGN, "workspace must be 1024-byte aligned");
  if (!w->stem1_w || !w->stem1_b) return fail(POEM_E_NULL, "hrnet: stem weights missing");
  return POEM_OK;
}

extern "C" int poem_hrnet_forward(const PoemHRNet* w, int n_images, int img_res, const float* images,
                                  float* const* out, void* workspace, size_t workspace_bytes, void* stream) {
  POEM_TRY(hrnet_check(w, img_res, images, workspace));
  if (!out) return fail(POEM_E_NULL, "hrnet: null pointer");
  const int N = n_images;
  cudaStream_t st = (cudaStream_t)stream;
  HrNetPlan p;
  const size_t need = hrnet_plan(N, img_res, w->channels, reinterpret_cast<uint8_t*>(workspace), &p);
  if (need > workspace_bytes) return fail(POEM_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, need);
  int R[4], cur[4];
  for (int i = 0; i < 4; ++i) {
    R[i] = (img_res / 4) >> i;
    if (!out[i]) return fail(POEM_E_NULL, "hrnet: output %d missing", i);
  }
  POEM_TRY(hrnet_run(w, N, img_res, images, p, cur, st));
  return hr_export(p.hr, cur, w->channels, R, N, out, st);
}

// ------------------------------------------------------------------------------------------------
// images -> mlvl_feat: backbone + feat_decode (reference lib/models/POEM.py:255-265 `extract_img_feat` + `feat_decode`,
// HRNet branch :189-203): three stride-2 ConvBlocks chained down the pyramid with the backbone maps added after the
// ReLU, bilinear x2 upsampling of the 8x8 map, 1x1 convolution 320 -> 160
// ------------------------------------------------------------------------------------------------
struct FeatPlan {
  HrNetPlan net;
  op16* d[3];     // pyramid sums at R/8, R/16, R/32
  float* f8;               // feat_in output at R/32, fp32 NHWC (padded channels)
  op16* cat[3];   // uv_decode: cat(upsampled, skip) at R/16, R/8, R/4
  op16* u[3];     // uv_decode: ConvBlock outputs at R/16, R/8, R/4
};
static size_t feat_plan(int N, int img_res, const int* ch, int out_ch, bool with_uv, uint8_t* base, FeatPlan* p) {
  const size_t net = hrnet_plan(N, img_res, ch, base, &p->net);
  Bump b{base ? base + ((net + 1023) & ~size_t(1023)) : nullptr, 0};
  for (int i = 0; i < 3; ++i) {
    const size_t r = (size_t)(img_res / 8) >> i;
    p->d[i] = b.take<op16>((size_t)N * r * r * pad16(ch[i + 1]));
  }
  const size_t r8 = (size_t)img_res / 32;
  p->f8 = b.take<float>((size_t)N * r8 * r8 * pad64(out_ch));
  for (int i = 0; i < 3; ++i) {
    p->cat[i] = p->u[i] = nullptr;
    if (!with_uv) continue;
    const size_t r = (size_t)(img_res / 16) << i;    // R/16, R/8, R/4
    p->cat[i] = b.take<op16>((size_t)N * r * r * pad16(ch[3 - i] + ch[2 - i]));
    p->u[i] = b.take<op16>((size_t)N * r * r * pad16(ch[2 - i]));
  }
  return ((net + 1023) & ~size_t(1023)) + b.off;
}
extern "C" size_t poem_image_features_workspace_bytes(const PoemHRNet* w, const PoemFeatDecode* fd, const PoemUVDecode* uv,
                                                      int n_images, int img_res) {
  if (!w || !fd || n_images < 1 || img_res < 64 || img_res % 32) return 0;
  FeatPlan p;
  return feat_plan(n_images, img_res, w->channels, fd->out_channels, uv != nullptr, nullptr, &p);
}

extern "C" int poem_image_features(const PoemHRNet* w, const PoemFeatDecode* fd, const PoemUVDecode* uv, int n_images,
                                   int img_res, const float* images, float* mlvl_feat, float* uv_px, float* heatmap,
                                   float* const* maps, void* workspace, size_t workspace_bytes, void* stream) {
  POEM_TRY(hrnet_check(w, img_res, images, workspace));
  if (!fd || !mlvl_feat) return fail(POEM_E_NULL, "image_features: null pointer");
  if (fd->out_channels < 1 || fd->out_channels > 256) return fail(POEM_E_BADDIM, "out_channels=%d", fd->out_channels);
  if (uv && (!uv_px || !uv->out_w || !uv->out_b)) return fail(POEM_E_NULL, "image_features: uv pointers missing");
  if (uv && (uv->n_joints < 1 || uv->n_joints > HEAT_MAX_J)) return fail(POEM_E_BADDIM, "n_joints=%d", uv->n_joints);
  const int N = n_images;
  const int* ch = w->channels;
  cudaStream_t st = (cudaStream_t)stream;
  FeatPlan p;
  const size_t need = feat_plan(N, img_res, ch, fd->out_channels, uv != nullptr, reinterpret_cast<uint8_t*>(workspace), &p);
  if (need > workspace_bytes) return fail(POEM_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, need);
  int Cp[4], R[4], cur[4];
  for (int i = 0; i < 4; ++i) {
    Cp[i] = pad64(ch[i]);
    R[i] = (img_res / 4) >> i;
    if (uv && ch[i] % 8) return fail(POEM_E_BADDIM, "uv_decode needs channel counts that are multiples of 8");
  }
  POEM_TRY(hrnet_run(w, N, img_res, images, p.net, cur, st));
  if (maps) {
    for (int i = 0; i < 4; ++i)
      if (!maps[i]) return fail(POEM_E_NULL, "image_features: map %d missing", i);
    POEM_TRY(hr_export(p.net.hr, cur, ch, R, N, maps, st));
  }
  // ---- feat_decode: x = f0 ; x = relu(bn(conv3x3 s2(x))) + f_{i+1}
  const op16* x = p.net.hr.x[0][cur[0]];
  for (int i = 0; i < 3; ++i) {
    POEM_TRY(launch_conv(x, N, R[i], R[i], Cp[i], fd->delayer[i], Cp[i + 1], 3, 2, true, p.net.hr.x[i + 1][cur[i + 1]],
                         p.d[i], st, 0, /*relu_before_res=*/true, nullptr, -1, pad16(ch[i]), pad16(ch[i + 1])));
    x = p.d[i];
  }
  // feat_in is a 1x1 convolution without norm / activation: it commutes with the bilinear upsampling (whose weights
  // sum to one, so the bias passes through), so it runs on the 8x8 map (4x fewer MACs) and the upsampling kernel
  // interpolates its fp32 output straight into the NCHW tensor the head consumes
  const int Co_p = pad64(fd->out_channels);
  POEM_TRY(launch_conv(x, N, R[3], R[3], Cp[3], fd->feat_in, Co_p, 1, 1, false, nullptr, nullptr, st, 0, false, p.f8, -1,
                       pad16(ch[3]), Co_p));
  {
    const int Ro = 2 * R[3];
    const size_t total = (size_t)N * fd->out_channels * Ro * Ro;
    prof_begin(st);
    upsample2x_nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.f8, mlvl_feat, N, R[3], R[3], Co_p,
                                                                                  fd->out_channels);
    LAUNCH_CHECK("upsample2x_nhwc_to_nchw_kernel");
  }
  if (!uv) return POEM_OK;
  // ---- uv_decode + heatmap_stage (POEM.py:205-229): x = f3; x = ConvBlock_i(cat(up2(x), f_{2-i})); max-pool; 1x1 +
  // sigmoid; soft-argmax
  const op16* h = p.net.hr.x[3][cur[3]];
  int h_cs = pad16(ch[3]), h_c = ch[3];
  for (int i = 0; i < 3; ++i) {
    const int lo = 2 - i;                      // skip branch, resolution R[lo]
    const int cat_cp = pad64(h_c + ch[lo]), cat_cs = pad16(h_c + ch[lo]);
    const size_t total8 = (size_t)N * R[lo] * R[lo] * cat_cs / 8;
    prof_begin(st);
    upsample2x_concat_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, st>>>(
        h, p.net.hr.x[lo][cur[lo]], p.cat[i], R[lo] / 2, R[lo] / 2, h_cs, h_c, pad16(ch[lo]), ch[lo], cat_cs, total8);
    LAUNCH_CHECK("upsample2x_concat_kernel");
    POEM_TRY(launch_conv(p.cat[i], N, R[lo], R[lo], cat_cp, uv->delayer[i], Cp[lo], 3, 1, true, nullptr, p.u[i], st, ch[lo], false,
                         nullptr, h_c + ch[lo], cat_cs, pad16(ch[lo])));
    h = p.u[i];
    h_cs = pad16(ch[lo]);
    h_c = ch[lo];
  }
  {
    const int J = uv->n_joints, Rp = R[0] / 2;
    const size_t smem = (size_t)(J * ch[0] + J + 8 * J * 3) * sizeof(float);
    prof_begin(st);
    heatmap_uv_kernel<<<N, 256, smem, st>>>(h, uv->out_w, uv->out_b, uv_px, heatmap, Rp, pad16(ch[0]), ch[0], J, (float)img_res,
                                           (float)img_res);
    LAUNCH_CHECK("heatmap_uv_kernel");
  }
  return POEM_OK;
}

extern "C" int poem_pa_metrics(const float* gt, const float* pred, int batch, int n_points, float* out, float* aligned,
                               void* stream) {
  if (!gt || !pred || !out) return fail(POEM_E_NULL, "pa_metrics: null pointer");
  if (batch < 1 || n_points < 3) return fail(POEM_E_BADDIM, "pa_metrics: batch=%d points=%d", batch, n_points);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  pa_metrics_kernel<<<batch, PA_THREADS, 0, st>>>(gt, pred, n_points, out, aligned);
  LAUNCH_CHECK("pa_metrics_kernel");
  return POEM_OK;
}

extern "C" int poem_triangulate_dlt(const float* uv_px, const float* cam_intr, const float* cam_extr,
                                    const int32_t* view_counts, int batch, int n_joints, float* ref_joints, void* stream) {
  if (!uv_px || !cam_intr || !cam_extr || !view_counts || !ref_joints) return fail(POEM_E_NULL, "triangulate: null pointer");
  if (batch < 1 || n_joints < 1 || n_joints > 32) return fail(POEM_E_BADDIM, "triangulate: batch=%d joints=%d", batch, n_joints);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  dlt_triangulate_kernel<<<batch, 32, 0, st>>>(uv_px, cam_intr, cam_extr, view_counts, ref_joints, n_joints);
  LAUNCH_CHECK("dlt_triangulate_kernel");
  return POEM_OK;
}

extern "C" int poem_linear(const poem_op16* A, int lda, const poem_op16* W, int ldw, const float* bias, int M, int N,
                           int K, int act, const float* residual, int ld_res, float* out_f32, int ld_f32,
                           poem_op16* out_op16, int ld_op16, void* stream) {
  GemmEpilogue e = epi_default(N);
  e.bias = bias;
  e.act = act;
  if (residual) {
    e.res_mode = RES_F32;
    e.res_f32 = residual;
    e.res_ld = ld_res;
  }
  e.out_f32 = out_f32;
  e.ld_f32 = ld_f32;
  e.out_op16 = reinterpret_cast<op16*>(out_op16);
  e.ld_op16 = ld_op16;
  return launch_gemm(reinterpret_cast<const op16*>(A), lda, reinterpret_cast<const op16*>(W), ldw, M,
                     N, K, e, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// MHA launch
// ------------------------------------------------------------------------------------------------
template <int HD>
static int launch_mha_hd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, op16* ctx,
                         int ld_ctx, int B, int Lq, int Lk, int n_heads, int q_col0, int k_col0, int v_col0,
                         cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(mha_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  MhaCfg<HD>::kSmemBytes));
    configured = true;
  }
  dim3 grid((unsigned)(((Lq + MHA_BQ - 1) / MHA_BQ) * n_heads * B));   // full tiles first, partial tiles last
  const float scale_log2e = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
  prof_begin(st);
  launch_pdl(mha_fwd_tc_kernel<HD>, dim3(grid), dim3(MHA_THREADS), (size_t)(MhaCfg<HD>::kSmemBytes), st, tq, tk, tv, ctx, ld_ctx, Lq, Lk, q_col0,
                                                                           k_col0, v_col0, scale_log2e, n_heads);
  LAUNCH_CHECK("mha_fwd_tc_kernel");
  return POEM_OK;
}

// Q [B*Lq, ldq] (columns q_col0..), K [B*Lk, ldk] (columns k_col0..), V [B*Lk, ldv] (columns v_col0..); all row-major
static int launch_mha(const op16* Q, int ldq, int q_col0, const op16* K, int ldk, int k_col0,
                      const op16* V, int ldv, int v_col0, op16* ctx, int ld_ctx, int B, int Lq,
                      int Lk, int D, int n_heads, cudaStream_t st) {
  if (D % n_heads) return fail(POEM_E_BADDIM, "mha: D %% heads != 0");
  const int hd = D / n_heads;
  if (Lk % MHA_BKEY) return fail(POEM_E_BADDIM, "mha: Lk=%d must be a multiple of %d", Lk, MHA_BKEY);
  if (ld_ctx % 8) return fail(POEM_E_ALIGN, "mha: ld_ctx must be a multiple of 8");
  const uint32_t boxc = hd < 64 ? (uint32_t)hd : 64u;
  CUtensorMap tq, tk, tv;
  POEM_TRY(make_tmap_op16(&tq, Q, (uint64_t)B * Lq, (uint64_t)ldq, (uint64_t)ldq, boxc, MHA_BQ));
  POEM_TRY(make_tmap_op16(&tk, K, (uint64_t)B * Lk, (uint64_t)ldk, (uint64_t)ldk, boxc, MHA_BKEY));
  POEM_TRY(make_tmap_op16(&tv, V, (uint64_t)B * Lk, (uint64_t)ldv, (uint64_t)ldv, boxc, MHA_BKEY));
  switch (hd) {
    case 32: return launch_mha_hd<32>(tq, tk, tv, ctx, ld_ctx, B, Lq, Lk, n_heads, q_col0, k_col0, v_col0, st);
    case 64: return launch_mha_hd<64>(tq, tk, tv, ctx, ld_ctx, B, Lq, Lk, n_heads, q_col0, k_col0, v_col0, st);
    case 128: return launch_mha_hd<128>(tq, tk, tv, ctx, ld_ctx, B, Lq, Lk, n_heads, q_col0, k_col0, v_col0, st);
    default: return fail(POEM_E_BADDIM, "mha: head dim %d unsupported (32, 64, 128)", hd);
  }
}

extern "C" int poem_mha(const poem_op16* Q, int ldq, const poem_op16* K, int ldk, const poem_op16* V, int ldv,
                        poem_op16* ctx, int ld_ctx, int B, int Lq, int Lk, int D, int n_heads, void* stream) {
  if (!Q || !K || !V || !ctx) return fail(POEM_E_NULL, "mha: null pointer");
  return launch_mha(reinterpret_cast<const op16*>(Q), ldq, 0, reinterpret_cast<const op16*>(K), ldk,
                    0, reinterpret_cast<const op16*>(V), ldv, 0, reinterpret_cast<op16*>(ctx), ld_ctx,
                    B, Lq, Lk, D, n_heads, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// small stages
// ------------------------------------------------------------------------------------------------
static int launch_knn(const float* q, const float* r, int* idx, int B, int Lq, int Lr, cudaStream_t st) {
  if (Lr < 32) return fail(POEM_E_BADDIM, "knn: need at least 32 reference points");
  const int total = B * Lq;
  const int threads = 256;
  const int blocks = (total * 32 + threads - 1) / threads;
  prof_begin(st);
  launch_pdl(knn32_kernel, dim3(blocks), dim3(threads), (size_t)(0), st, q, r, idx, Lq, Lr, total);
  LAUNCH_CHECK("knn32_kernel");
  return POEM_OK;
}
static int launch_knn_bps(const float* q, const float* r, const int* perm, const float* boxes, int* idx, int B, int Lq,
                          int Lr, cudaStream_t st) {
  if (Lr % 32 || Lr > 4096 || Lr < 32) return fail(POEM_E_BADDIM, "knn_bps: Lr=%d must be a multiple of 32, <= 4096", Lr);
  const int total = B * Lq;
  const int threads = 256;
  prof_begin(st);
  launch_pdl(knn32_bps_kernel, dim3((total * 32 + threads - 1) / threads), dim3(threads), (size_t)(0), st, q, r, perm, boxes, idx, Lq, Lr, total);
  LAUNCH_CHECK("knn32_bps_kernel");
  return POEM_OK;
}
extern "C" int poem_knn32_bps(const float* query_xyz, const float* ref_xyz, const int32_t* perm, const float* boxes,
                              int32_t* idx, int B, int Lq, int Lr, void* stream) {
  if (!query_xyz || !ref_xyz || !perm || !boxes || !idx) return fail(POEM_E_NULL, "knn_bps: null pointer");
  return launch_knn_bps(query_xyz, ref_xyz, perm, boxes, idx, B, Lq, Lr, (cudaStream_t)stream);
}
extern "C" int poem_knn32(const float* query_xyz, const float* ref_xyz, int32_t* idx, int B, int Lq, int Lr,
                          void* stream) {
  if (!query_xyz || !ref_xyz || !idx) return fail(POEM_E_NULL, "knn: null pointer");
  return launch_knn(query_xyz, ref_xyz, idx, B, Lq, Lr, (cudaStream_t)stream);
}

static int launch_layernorm(const float* x, const float* g, const float* b, float* y32, op16* y16, int rows,
                            int D, cudaStream_t st) {
  if (D % 32 || D > 1024) return fail(POEM_E_BADDIM, "layernorm: D=%d", D);
  const int threads = 256;
  prof_begin(st);
  launch_pdl(layernorm_kernel, dim3((rows * 32 + threads - 1) / threads), dim3(threads), (size_t)(0), st, x, g, b, y32, y16, rows, D, 1e-12f);
  LAUNCH_CHECK("layernorm_kernel");
  return POEM_OK;
}
extern "C" int poem_layernorm(const float* x, const float* gamma, const float* beta, float* y_f32, poem_op16* y_op16,
                              int rows, int D, void* stream) {
  if (!x || !gamma || !beta) return fail(POEM_E_NULL, "layernorm: null pointer");
  return launch_layernorm(x, gamma, beta, y_f32, reinterpret_cast<op16*>(y_op16), rows, D,
                          (cudaStream_t)stream);
}

// per-image / per-sample index tables derived from view_counts
struct ViewTables {
  int *img_sample, *img_view, *img_posrow, *sample_rowbase, *sample_views, *tile_start;
  int n_merge_tiles;   // row tiles of the fused sampler/merge kernel (tile_start[B])
};
static size_t view_tables_ints(int B, int NV) { return (size_t)3 * NV + 3 * B + 1; }
static int upload_view_tables(const int32_t* host_views, int B, int NV, int P, int max_views, int* dev, ViewTables* vt,
                              cudaStream_t st) {
  std::vector<int> h(view_tables_ints(B, NV));
  int* img_sample = h.data();
  int* img_view = img_sample + NV;
  int* img_posrow = img_view + NV;
  int* rowbase = img_posrow + NV;
  int* views = rowbase + B;
  int* tile_start = views + B;
  int img = 0, tiles = 0;
  long long rb = 0;
  for (int b = 0; b < B; ++b) {
    const int n = host_views[b];
    if (n < 1 || n > max_views) return fail(POEM_E_BADDIM, "view count %d of sample %d outside [1,%d]", n, b, max_views);
    rowbase[b] = (int)rb;
    views[b] = n;
    tile_start[b] = tiles;
    tiles += merge_tiles_of(n, P);
    for (int v = 0; v < n; ++v, ++img) {
      if (img >= NV) return fail(POEM_E_BADDIM, "sum(view_counts) exceeds n_images=%d", NV);
      img_sample[img] = b;
      img_view[img] = v;
      img_posrow[img] = n * (n - 1) / 2 + v;
    }
    rb += (long long)n * P;
    if (rb > 0x7fffffffLL) return fail(POEM_E_BADDIM, "too many merge rows");
  }
  if (img != NV) return fail(POEM_E_BADDIM, "sum(view_counts)=%d != n_images=%d", img, NV);
  tile_start[B] = tiles;
  vt->n_merge_tiles = tiles;
  if (B <= VIEW_PARAM_MAX) {
    // the view counts travel as kernel parameters and the tables are rebuilt on the device: no host-to-device copy, so
    // the whole forward can be captured into a CUDA graph (a memcpy node would keep a pointer to this stack frame)
    ViewCountsParam vp;
    for (int b = 0; b < B; ++b) vp.n[b] = host_views[b];
    prof_begin(st);
    launch_pdl(view_tables_kernel, dim3(1), dim3(256), (size_t)(0), st, vp, B, NV, P, dev);
    LAUNCH_CHECK("view_tables_kernel");
  } else {
    // a memcpy node would keep a pointer into this stack frame: refuse to be captured into a CUDA graph
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap != cudaStreamCaptureStatusNone)
      return fail(POEM_E_BADDIM, "batch %d > %d cannot be captured into a CUDA graph (view tables are copied from the host)", B,
                  VIEW_PARAM_MAX);
    CUDA_TRY(cudaMemcpyAsync(dev, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // h dies with this frame
  }
  vt->img_sample = dev;
  vt->img_view = dev + NV;
  vt->img_posrow = dev + 2 * NV;
  vt->sample_rowbase = dev + 3 * NV;
  vt->sample_views = dev + 3 * NV + B;
  vt->tile_start = dev + 3 * NV + 2 * B;
  return POEM_OK;
}

static int launch_project_sample(const float* xmap, const float* intr, const float* extr, const float* bps,
                                 const float* centre, const ViewTables& vt, float* proj, int NV, int D, int P, int fh,
                                 int fw, float img_w, float img_h, op16* X, cudaStream_t st) {
  if (P != SAMPLE_THREADS * 8) return fail(POEM_E_BADDIM, "sampler is specialised for P=4096 (got %d)", P);
  if (D % SAMPLE_CH || P % D) return fail(POEM_E_BADDIM, "sampler: D=%d must divide P and be a multiple of 32", D);
  prof_begin(st);
  launch_pdl(camera_prep_kernel, dim3((NV + 63) / 64), dim3(64), (size_t)(0), st, intr, extr, proj, NV);
  LAUNCH_CHECK("camera_prep_kernel");
  const size_t smem = (size_t)SAMPLE_PITCH * fh * fw * sizeof(float);
  if (smem > 48 * 1024) return fail(POEM_E_BADDIM, "feature map %dx%d too large for the sampler", fh, fw);
  dim3 grid(D / SAMPLE_CH, NV);
  prof_begin(st);
  project_sample_kernel<<<grid, SAMPLE_THREADS, smem, st>>>(xmap, proj, bps, centre, vt.img_sample, vt.img_view,
                                                           vt.sample_rowbase, X, D, P, fh, fw, 1.0f / img_w,
                                                           1.0f / img_h);
  LAUNCH_CHECK("project_sample_kernel");
  return POEM_OK;
}

extern "C" int poem_project_sample(const float* xmap, const float* cam_intr, const float* cam_extr, const float* bps,
                                   const float* centre, const int32_t* host_view_counts, int B, int n_images, int D,
                                   int P, int fh, int fw, float img_w, float img_h, poem_op16* X, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!xmap || !cam_intr || !cam_extr || !bps || !centre || !host_view_counts || !X || !workspace)
    return fail(POEM_E_NULL, "project_sample: null pointer");
  Bump bump{reinterpret_cast<uint8_t*>(workspace), 0};
  int* tab = bump.take<int>(view_tables_ints(B, n_images));
  float* proj = bump.take<float>((size_t)n_images * 24);
  if (bump.off > workspace_bytes) return fail(POEM_E_WORKSPACE, "project_sample: workspace %zu < %zu", workspace_bytes, bump.off);
  ViewTables vt;
  POEM_TRY(upload_view_tables(host_view_counts, B, n_images, P, 64, tab, &vt, (cudaStream_t)stream));
  return launch_project_sample(xmap, cam_intr, cam_extr, bps, centre, vt, proj, n_images, D, P, fh, fw, img_w, img_h,
                               reinterpret_cast<op16*>(X), (cudaStream_t)stream);
}

// Bilinear taps of every (image, BPS point): the gather indices and weights of rows a4 / a5 on their own
extern "C" int poem_sample_taps(const float* cam_intr, const float* cam_extr, const float* bps, const float* centre,
                                const int32_t* host_view_counts, int B, int n_images, int P, int fh, int fw, float img_w,
                                float img_h, int32_t* tap_pixels, float* tap_weights, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!cam_intr || !cam_extr || !bps || !centre || !host_view_counts || !tap_pixels || !tap_weights || !workspace)
    return fail(POEM_E_NULL, "sample_taps: null pointer");
  if (P != SM_P || fh * fw != SM_F) return fail(POEM_E_BADDIM, "sample_taps: P=%d, %dx%d (4096 points, 256 pixels)", P, fh, fw);
  cudaStream_t st = (cudaStream_t)stream;
  Bump bump{reinterpret_cast<uint8_t*>(workspace), 0};
  int* tab = bump.take<int>(view_tables_ints(B, n_images));
  float* proj = bump.take<float>((size_t)n_images * 24);
  uint32_t* taps = bump.take<uint32_t>((size_t)n_images * (P / 4) * SM_TAP_WORDS);
  if (bump.off > workspace_bytes) return fail(POEM_E_WORKSPACE, "sample_taps: workspace %zu < %zu", workspace_bytes, bump.off);
  ViewTables vt;
  POEM_TRY(upload_view_tables(host_view_counts, B, n_images, P, 64, tab, &vt, st));
  prof_begin(st);
  launch_pdl(camera_prep_kernel, dim3((n_images + 63) / 64), dim3(64), (size_t)0, st, cam_intr, cam_extr, proj, n_images);
  LAUNCH_CHECK("camera_prep_kernel");
  prof_begin(st);
  launch_pdl(sample_taps_kernel, dim3((unsigned)(((size_t)n_images * P + 255) / 256)), dim3(256), (size_t)0, st, proj, bps, centre,
             vt.img_sample, taps, n_images, fh, fw, 1.0f / img_w, 1.0f / img_h, 1);    // pixel pitch 1: offsets = pixel indices
  LAUNCH_CHECK("sample_taps_kernel");
  prof_begin(st);
  unpack_taps_kernel<<<(unsigned)(((size_t)n_images * P + 255) / 256), 256, 0, st>>>(taps, tap_pixels, tap_weights, n_images);
  LAUNCH_CHECK("unpack_taps_kernel");
  return POEM_OK;
}

// ------------------------------------------------------------------------------------------------
// vector attention
// ------------------------------------------------------------------------------------------------
extern "C" size_t poem_vector_attention_workspace_bytes(int B, int Lq, int D) {
  return 3 * (((size_t)B * Lq * 32 * D * 2 + 1023) & ~size_t(1023)) + 4096;
}

// Un-fused composition of the same folded math: token tensors (B*Lq*32, D) live in HBM between the D x D GEMMs.
//   t0: h -> logits ; t1: pos ; t2: gamma1_pre -> relu(gamma1_pre + qt_i - kt_j)
static int launch_vecattn(const PoemVecAttn* w, const op16* q, int ldq, const op16* ktab, int ldk,
                          const op16* vtab, int ldv, const float* q_xyz, const float* ref_xyz, const int* idx,
                          const int* anchor_idx, const float* anchor_xyz, int B, int Lq, int Lr, int D,
                          op16* res, op16* t0, op16* t1, op16* t2,
                          cudaStream_t st) {
  if (!w->wd1 || !w->bd1 || !w->delta2.w || !w->gamma1_delta2.w || !w->gamma2.w)
    return fail(POEM_E_NULL, "vector_attention: weight pointer missing");
  const size_t n_query = (size_t)B * Lq;
  const size_t T = n_query * 32;
  if (T > 0x7fffffffULL) return fail(POEM_E_BADDIM, "vector_attention: too many tokens");
  prof_begin(st);
  va_hdelta_kernel<<<(unsigned)((T + 7) / 8), 256, 0, st>>>(q_xyz, ref_xyz, idx, anchor_xyz, w->wd1, w->bd1, t0, Lq, Lr,
                                                          D, T);
  LAUNCH_CHECK("va_hdelta_kernel");
  auto lin = [&](const op16* A, const PoemLinear& l, bool bias, op16* out) {
    TagScope ts("va_token");
    GemmEpilogue e = epi_default(D);
    e.bias = bias ? l.b : nullptr;
    e.out_op16 = out;
    e.ld_op16 = D;
    return launch_gemm(A, D, reinterpret_cast<const op16*>(l.w), D, (int)T, D, D, e, st);
  };
  POEM_TRY(lin(t0, w->delta2, true, t1));          // pos
  POEM_TRY(lin(t0, w->gamma1_delta2, false, t2));  // (W_g1 W_d2) h
  {
    const size_t total = T * (D / 8);
    prof_begin(st);
    va_gmix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(q, ldq, ktab, ldk, idx, anchor_idx, t2, Lq, Lr, D, T);
    LAUNCH_CHECK("va_gmix_kernel");
  }
  POEM_TRY(lin(t2, w->gamma2, true, t0));          // attention logits
  {
    const size_t total = n_query * D;
    prof_begin(st);
    va_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(t0, t1, vtab, ldv, idx, anchor_idx, res, Lq, Lr, D,
                                                                     1.0f / sqrtf((float)D), n_query);
    LAUNCH_CHECK("va_reduce_kernel");
  }
  return POEM_OK;
}

template <int D>
static int launch_sample_merge(const PoemWeights* w, const SmParams& sp, cudaStream_t st) {
  using Cfg = SmCfg<D>;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(sample_merge_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  CUtensorMap t0a, t0b;
  POEM_TRY(make_tmap_op16(&t0a, w->merge0a.w, D, D, D, 64, 128));
  POEM_TRY(make_tmap_op16(&t0b, w->merge0b.w, D / 2, D, D, 64, (uint32_t)(D / 2 < 128 ? D / 2 : 128)));
  const int grid = sp.n_tiles < num_sms() ? sp.n_tiles : num_sms();
  prof_begin(st);
  launch_pdl(sample_merge_kernel<D>, dim3(grid), dim3(Cfg::THREADS), (size_t)(Cfg::SMEM_BYTES), st, t0a, t0b, sp);
  LAUNCH_CHECK("sample_merge_kernel");
  return POEM_OK;
}

// Fused kernel (vecattn.cuh): the three D x D GEMMs, both MLP epilogues, the neighbour gathers and the per-channel
// softmax/reduction in one launch; nothing token-sized touches HBM.
static bool g_force_unfused = false;
extern "C" void poem_debug_force_unfused(int on) { g_force_unfused = (on != 0); }

template <int D>
static int launch_vecattn_fused(const PoemVecAttn* w, const VaParams& prm, cudaStream_t st) {
  using Cfg = VaCfg<D>;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(va_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  CUtensorMap t1, t2, t3;
  POEM_TRY(make_tmap_op16(&t1, w->delta2.w, D, D, D, 64, 128));
  POEM_TRY(make_tmap_op16(&t2, w->gamma1_delta2.w, D, D, D, 64, 128));
  POEM_TRY(make_tmap_op16(&t3, w->gamma2.w, D, D, D, 64, 128));
  const int tiles = (prm.n_query + Cfg::QT - 1) / Cfg::QT;
  const int slots = num_sms() * Cfg::CTAS_PER_SM;
  const int grid = tiles < slots ? tiles : slots;
  prof_begin(st);
  launch_pdl(va_fused_kernel<D>, dim3(grid), dim3(Cfg::THREADS), (size_t)(Cfg::SMEM_BYTES), st, t1, t2, t3, prm);
  LAUNCH_CHECK("va_fused_kernel");
  return POEM_OK;
}

static int launch_vector_attention(const PoemVecAttn* w, const op16* q, int ldq, const op16* ktab,
                                   int ldk, const op16* vtab, int ldv, const float* q_xyz,
                                   const float* ref_xyz, const int* idx, const int* anchor_idx,
                                   const float* anchor_xyz, int B, int Lq, int Lr, int D, op16* res,
                                   op16* t0, op16* t1, op16* t2, cudaStream_t st) {
  if ((idx == nullptr) == (anchor_idx == nullptr)) return fail(POEM_E_NULL, "vector_attention: give idx XOR anchors");
  if (anchor_idx && !anchor_xyz) return fail(POEM_E_NULL, "vector_attention: anchor_xyz missing");
  if (D % 32) return fail(POEM_E_BADDIM, "vector_attention: D=%d", D);
  // the fused kernel gathers kt / v as 4-byte words (two channels): even leading dimension, 4-byte aligned tables
  if (!g_force_unfused && (D == 128 || D == 256 || D == 512) && ldk == ldv && (ldk % 2) == 0 &&
      ((reinterpret_cast<uintptr_t>(ktab) | reinterpret_cast<uintptr_t>(vtab)) & 3) == 0 &&
      (long long)B * Lr * ldk < 0x7fffffffLL) {
    if (!w->wd1 || !w->bd1 || !w->delta2.w || !w->delta2.b || !w->gamma1_delta2.w || !w->gamma2.w)
      return fail(POEM_E_NULL, "vector_attention: weight pointer missing");
    VaParams prm;
    prm.q = q, prm.ktab = ktab, prm.vtab = vtab;
    prm.ldq = ldq, prm.ldk = ldk, prm.ldv = ldv;
    prm.q_xyz = q_xyz, prm.ref_xyz = ref_xyz, prm.idx = idx, prm.anchor_idx = anchor_idx, prm.anchor_xyz = anchor_xyz;
    prm.wd1 = w->wd1, prm.bd1 = w->bd1, prm.bd2 = w->delta2.b;
    prm.res = res;
    prm.Lq = Lq, prm.Lr = Lr, prm.n_query = B * Lq;
    prm.softmax_scale_log2e = 1.4426950408889634f / sqrtf((float)D);
    if (D == 128) return launch_vecattn_fused<128>(w, prm, st);
    if (D == 256) return launch_vecattn_fused<256>(w, prm, st);
    return launch_vecattn_fused<512>(w, prm, st);
  }
  return launch_vecattn(w, q, ldq, ktab, ldk, vtab, ldv, q_xyz, ref_xyz, idx, anchor_idx, anchor_xyz, B, Lq, Lr, D,
                        res, t0, t1, t2, st);
}

extern "C" int poem_vector_attention(const PoemVecAttn* w, const poem_op16* q, int ldq, const poem_op16* ktab, int ldk,
                                     const poem_op16* vtab, int ldv, const float* q_xyz, const float* ref_xyz,
                                     const int32_t* idx, const int32_t* anchor_idx, const float* anchor_xyz, int B,
                                     int Lq, int Lr, int D, poem_op16* res, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (!w || !q || !ktab || !vtab || !q_xyz || !res || !workspace) return fail(POEM_E_NULL, "vector_attention: null");
  if (workspace_bytes < poem_vector_attention_workspace_bytes(B, Lq, D))
    return fail(POEM_E_WORKSPACE, "vector_attention: workspace too small");
  Bump bump{reinterpret_cast<uint8_t*>(workspace), 0};
  const size_t n = (size_t)B * Lq * 32 * D;
  op16* t0 = bump.take<op16>(n);
  op16* t1 = bump.take<op16>(n);
  op16* t2 = bump.take<op16>(n);
  return launch_vector_attention(w, reinterpret_cast<const op16*>(q), ldq,
                                 reinterpret_cast<const op16*>(ktab), ldk,
                                 reinterpret_cast<const op16*>(vtab), ldv, q_xyz, ref_xyz, idx, anchor_idx,
                                 anchor_xyz, B, Lq, Lr, D, reinterpret_cast<op16*>(res), t0, t1, t2,
                                 (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// whole path
// ------------------------------------------------------------------------------------------------
struct BlockPlan {   // decoder blocks (PtEmbedTRv4)
  float *pt_xyz, *pt_xyz_sorted, *xyz;  // xyz: (NB+1) buffers of B*Q*3; pt_xyz_sorted: BPS in k-d chunk order
  op16 *ptf, *KK;   // KK: (B*P, 6D) = K1 | K2 | kt_cross | v_cross | V1 | V2
  float *qf32, *qe32, *tmp32, *a1_32, *a2_32, *f1_32, *f2_32;
  op16 *qf16, *qe16, *qp16, *ctx16, *a1_16, *a2_16, *qkv16, *res16, *f1_16, *qc16, *f2_16, *r1_16, *ffn16;
  int *idx_self, *idx_cross;
  op16 *t0, *t1, *t2;
};
struct HeadPlan {    // everything in front of the blocks
  int* tables;
  float *proj, *centre, *xmap;
  op16 *featT, *X, *H1, *Mm, *S, *H2;
  float* sigma;   // per-token power-of-two scale of S (merge_reduce_kernel / sample_merge_kernel)
  uint32_t* taps; // bilinear taps of every (image, BPS point) (sample_taps_kernel)
  op16* Q1;       // (B*P, D) token-first rows of X (fused path)
};

static int check_dims(const PoemDims* d) {
  if (!d) return fail(POEM_E_NULL, "dims is NULL");
  const int D = d->embed_dims;
  if (D == 1024)   // POEM-huge (config/release/train_huge.yaml:188-190: D = 1024, 4 heads -> head dim 256)
    return fail(POEM_E_BADDIM, "embed_dims=1024 (POEM-huge) is not built: the attention kernel has no head-dim-256 instantiation and "
                               "the fused vector attention no D=1024 one");
  if (!(D == 128 || D == 256 || D == 512)) return fail(POEM_E_BADDIM, "embed_dims=%d unsupported (128, 256, 512)", D);
  if (d->n_heads <= 0 || D % d->n_heads) return fail(POEM_E_BADDIM, "n_heads=%d", d->n_heads);
  const int hd = D / d->n_heads;
  if (!(hd == 32 || hd == 64 || hd == 128)) return fail(POEM_E_BADDIM, "head dim %d unsupported", hd);
  if (d->n_sample != 4096) return fail(POEM_E_BADDIM, "n_sample=%d (kernels are specialised for 4096)", d->n_sample);
  if (d->n_neighbor != 32) return fail(POEM_E_BADDIM, "n_neighbor=%d (kernels are specialised for 32)", d->n_neighbor);
  if (d->feat_h * d->feat_w != 256) return fail(POEM_E_BADDIM, "feature map must have 256 pixels");
  if (d->n_blocks < 1 || d->n_blocks > POEM_MAX_BLOCKS) return fail(POEM_E_BADDIM, "n_blocks=%d", d->n_blocks);
  if (d->in_channels % 8) return fail(POEM_E_BADDIM, "in_channels=%d must be a multiple of 8", d->in_channels);
  if (d->n_query < 32 || d->max_views < 1) return fail(POEM_E_BADDIM, "n_query / max_views");
  return POEM_OK;
}

static void plan_blocks(const PoemDims* d, int B, Bump& b, BlockPlan* p) {
  const size_t D = d->embed_dims, P = d->n_sample, Q = d->n_query;
  const size_t BP = (size_t)B * P, BQ = (size_t)B * Q, T = BQ * 32;
  p->pt_xyz = b.take<float>(BP * 3);
  p->pt_xyz_sorted = b.take<float>(BP * 3);
  p->xyz = b.take<float>((size_t)(d->n_blocks + 1) * BQ * 3);
  p->ptf = b.take<op16>(BP * D);
  p->KK = b.take<op16>(BP * 6 * D);
  p->qf32 = b.take<float>(BQ * D);
  p->qe32 = b.take<float>(BQ * D);
  p->tmp32 = b.take<float>(BQ * D);
  p->a1_32 = b.take<float>(BQ * D);
  p->a2_32 = b.take<float>(BQ * D);
  p->f1_32 = b.take<float>(BQ * D);
  p->f2_32 = b.take<float>(BQ * D);
  p->qf16 = b.take<op16>(BQ * D);
  p->qe16 = b.take<op16>(BQ * D);
  p->qp16 = b.take<op16>(BQ * D);
  p->ctx16 = b.take<op16>(BQ * D);
  p->a1_16 = b.take<op16>(BQ * D);
  p->a2_16 = b.take<op16>(BQ * D);
  p->qkv16 = b.take<op16>(BQ * 3 * D);
  p->res16 = b.take<op16>(BQ * D);
  p->f1_16 = b.take<op16>(BQ * D);
  p->qc16 = b.take<op16>(BQ * D);
  p->f2_16 = b.take<op16>(BQ * D);
  p->r1_16 = b.take<op16>(BQ * D);
  p->ffn16 = b.take<op16>(BQ * 4 * D);
  p->idx_self = b.take<int>(T);
  p->idx_cross = b.take<int>(T);
  p->t0 = b.take<op16>(T * D);
  p->t1 = b.take<op16>(T * D);
  p->t2 = b.take<op16>(T * D);
}

// The view of a BlockPlan that starts at sample b0: every buffer is row-major over the samples of the batch
// (xyz: per block a slab of the whole batch, so only the start moves; run_blocks strides it by the whole batch).
static BlockPlan slice_plan(const BlockPlan& p, const PoemDims* d, int b0) {
  const size_t D = d->embed_dims, P = d->n_sample, Q = d->n_query;
  const size_t oP = (size_t)b0 * P, oQ = (size_t)b0 * Q, oT = oQ * 32;
  BlockPlan s = p;
  s.pt_xyz += oP * 3, s.pt_xyz_sorted += oP * 3, s.xyz += oQ * 3;
  s.ptf += oP * D, s.KK += oP * 6 * D;
  s.qf32 += oQ * D, s.qe32 += oQ * D, s.tmp32 += oQ * D, s.a1_32 += oQ * D, s.a2_32 += oQ * D, s.f1_32 += oQ * D, s.f2_32 += oQ * D;
  s.qf16 += oQ * D, s.qe16 += oQ * D, s.qp16 += oQ * D, s.ctx16 += oQ * D, s.a1_16 += oQ * D, s.a2_16 += oQ * D;
  s.qkv16 += oQ * 3 * D, s.res16 += oQ * D, s.f1_16 += oQ * D, s.qc16 += oQ * D, s.f2_16 += oQ * D, s.r1_16 += oQ * D;
  s.ffn16 += oQ * 4 * D;
  s.idx_self += oT, s.idx_cross += oT;
  s.t0 += oT * D, s.t1 += oT * D, s.t2 += oT * D;
  return s;
}

static void plan_head(const PoemDims* d, int B, int NV, Bump& b, HeadPlan* p) {
  const size_t D = d->embed_dims, C = d->in_channels, P = d->n_sample, F = 256;
  const size_t R = (size_t)NV * P, BP = (size_t)B * P;
  p->tables = b.take<int>(view_tables_ints(B, NV));
  p->proj = b.take<float>((size_t)NV * 24);
  p->centre = b.take<float>((size_t)B * 3);
  p->featT = b.take<op16>((size_t)NV * F * C);
  p->xmap = b.take<float>((size_t)NV * D * F);
  p->X = b.take<op16>(R * D);
  p->H1 = b.take<op16>(R * D);
  p->Mm = b.take<op16>(R * D / 2);
  p->S = b.take<op16>(BP * D / 2);
  p->H2 = b.take<op16>(BP * D / 2);
  p->sigma = b.take<float>(BP);
  p->taps = b.take<uint32_t>((size_t)NV * (P / 4) * SM_TAP_WORDS);
  p->Q1 = b.take<op16>(BP * D);
}

extern "C" size_t poem_workspace_bytes(const PoemDims* dims, int batch, int n_images) {
  if (check_dims(dims) != POEM_OK || batch < 1 || n_images < batch) return 0;
  Bump b{nullptr, 0};
  HeadPlan hp;
  BlockPlan bp;
  plan_head(dims, batch, n_images, b, &hp);
  plan_blocks(dims, batch, b, &bp);
  return b.off + 1024;
}
extern "C" size_t poem_transformer_workspace_bytes(const PoemDims* dims, int batch) {
  if (check_dims(dims) != POEM_OK || batch < 1) return 0;
  Bump b{nullptr, 0};
  BlockPlan bp;
  plan_blocks(dims, batch, b, &bp);
  return b.off + 1024;
}

constexpr int kHandCentreJoint = 9;   // reference_joints[:, 9, :] (ptEmb_head.py:873)

static inline const op16* W16(const PoemLinear& l) { return reinterpret_cast<const op16*>(l.w); }

static int linear(const char* tag, const op16* A, int lda, const PoemLinear& l, int M, int N, int K, int act,
                  const float* res32, float* o32, op16* o16, cudaStream_t st) {
  TagScope ts(tag);
  if (!l.w) return fail(POEM_E_NULL, "weight pointer missing");
  GemmEpilogue e = epi_default(N);
  e.bias = l.b;
  e.act = act;
  if (res32) {
    e.res_mode = RES_F32;
    e.res_f32 = res32;
    e.res_ld = N;
  }
  e.out_f32 = o32;
  e.ld_f32 = N;
  e.out_op16 = o16;
  e.ld_op16 = N;
  return launch_gemm(A, lda, W16(l), K, M, N, K, e, st);
}

// Side stream for the 32-NN searches: they only depend on the coordinates regressed by the previous block, are
// latency-bound and tiny in registers/smem, so they run concurrently with the next block's attention GEMMs instead of
// in front of its vector attention.  Fork/join with events keeps the call asynchronous on the caller's stream.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork[POEM_MAX_BLOCKS], join[POEM_MAX_BLOCKS];
  bool ok = false;
};
constexpr int kMaxDevices = 32;
// device a pointer lives on (-1 when it is not device memory); the streams / events below are created per device
static int device_of(const void* p) {
  cudaPointerAttributes a;
  if (p == nullptr || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}
// makes the device of the call's workspace current for the duration of an entry point (the caller's current device is
// restored on return), so streams / events are created on, and kernels launched to, the device that owns the buffers
struct DeviceGuard {
  int prev = -1, dev = -1;
  explicit DeviceGuard(const void* device_ptr) {
    dev = device_of(device_ptr);
    cudaGetDevice(&prev);
    if (dev >= 0 && dev != prev) cudaSetDevice(dev);
    else if (dev < 0) dev = prev;
  }
  ~DeviceGuard() {
    if (prev >= 0 && dev != prev) cudaSetDevice(prev);
  }
};
static int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}
constexpr int kMaxParts = 4;    // the decoder blocks run on up to this many batch slices concurrently
constexpr int kSideSlots = 2 * kMaxParts;   // slot k: 32-NN searches of slice k; slot kMaxParts + k: main stream of slice k (k >= 1)
static SideStream& side_stream(int slot = 0) {
  static thread_local SideStream per_dev[kMaxDevices][kSideSlots];
  SideStream& s = per_dev[current_device()][slot];
  if (!s.ok) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess) {
      s.ok = true;
      for (int i = 0; i < POEM_MAX_BLOCKS; ++i)
        s.ok = s.ok && cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming) == cudaSuccess;
    }
  }
  return s;
}

// Two chained Linear layers of the query stream in one kernel (qchain.cuh):
//   C1 = A·W1^T + b1 (+ res32) [-> LayerNorm] -> out1_f32 / out1_h16 ;  out2 = act2(fp16(C1)·W2^T + b2)
template <int D>
static int launch_chain2_d(const char* tag, const op16* A, const PoemLinear& l1, const PoemLinear& l2, Chain2Params prm,
                           cudaStream_t st) {
  using Cfg = QcCfg<D>;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(chain2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  if (!A || !l1.w || !l2.w || !prm.out2) return fail(POEM_E_NULL, "chain2 %s: null operand", tag);
  if (prm.N2 % Cfg::NC || prm.N2 > 4 * D || prm.ld2 % 16) return fail(POEM_E_BADDIM, "chain2 %s: N2=%d ld2=%d", tag, prm.N2, prm.ld2);
  CUtensorMap ta, t1, t2;
  POEM_TRY(make_tmap_op16(&ta, A, (uint64_t)prm.M, D, D, 64, 128));
  POEM_TRY(make_tmap_op16(&t1, l1.w, D, D, D, 64, Cfg::NC));
  POEM_TRY(make_tmap_op16(&t2, l2.w, (uint64_t)prm.N2, D, D, 64, Cfg::NC));
  prm.b1 = l1.b;
  prm.b2 = l2.b;
  const int tiles = (prm.M + 127) / 128;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  TagScope ts(tag);
  prof_begin(st);
  launch_pdl(chain2_kernel<D>, dim3(grid), dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, st, ta, t1, t2, prm);
  LAUNCH_CHECK("chain2_kernel");
  return POEM_OK;
}
static int launch_chain2(int D, const char* tag, const op16* A, const PoemLinear& l1, const PoemLinear& l2,
                         const Chain2Params& prm, cudaStream_t st) {
  return D == 128 ? launch_chain2_d<128>(tag, A, l1, l2, prm, st) : launch_chain2_d<256>(tag, A, l1, l2, prm, st);
}
static Chain2Params chain2_params(int M, int D) {
  Chain2Params c;
  memset(&c, 0, sizeof(c));
  c.M = M;
  c.N2 = D;
  c.ld2 = D;
  return c;
}

static thread_local op16* g_ptf_export = nullptr;     // test hook: merged BPS features (B*P, D) of the next whole-path call
static thread_local size_t g_ptf_capacity = 0;
extern "C" void poem_debug_export_pt_feats(poem_op16* device_buf, size_t capacity) {
  g_ptf_export = reinterpret_cast<op16*>(device_buf);
  g_ptf_capacity = device_buf ? capacity : 0;
}
static thread_local int32_t* g_nbr_export = nullptr;
static thread_local size_t g_nbr_capacity = 0;
extern "C" void poem_debug_export_neighbours(int32_t* device_buf, size_t capacity) {
  g_nbr_export = device_buf;
  g_nbr_capacity = device_buf ? capacity : 0;
}

static int launch_block_knn(const PoemWeights* w, int B, int Q, int P, const BlockPlan& p, const float* xyz,
                            bool pt_is_bps, cudaStream_t st) {
  POEM_TRY(launch_knn(xyz, xyz, p.idx_self, B, Q, Q, st));
  if (pt_is_bps && w->bps_perm && w->bps_chunk_box)
    return launch_knn_bps(xyz, p.pt_xyz_sorted, w->bps_perm, w->bps_chunk_box, p.idx_cross, B, Q, P, st);
  return launch_knn(xyz, p.pt_xyz, p.idx_cross, B, Q, P, st);
}

// a8-a13: the NB decoder blocks. Expects p.ptf (op16 BPS features), p.pt_xyz, p.xyz[0], p.qf32/p.qf16 filled.
// coords_out[i] = nan_to_num(xyz_i) * radius + centre when centre != NULL, else the raw normalised xyz_i.
// `B` samples starting at sample `b0` of a batch of `B_total` (plan, centre, coords_out, out_feats already point at
// sample b0; per-block slabs of xyz / coords are B_total samples apart); `half` selects the 32-NN side stream.
static int run_blocks(const PoemDims* dims, const PoemWeights* w, int B, const BlockPlan& p, const float* centre,
                      float* coords_out, float* out_feats, bool pt_is_bps, cudaStream_t st, int B_total = 0, int b0 = 0,
                      int half = 0) {
  const int D = dims->embed_dims, P = dims->n_sample, Q = dims->n_query, NB = dims->n_blocks;
  const int BQ = B * Q, BP = B * P;
  if (B_total <= 0) B_total = B;
  const size_t slab = (size_t)B_total * Q * 3;      // one block's coordinates of the whole batch
  SideStream& side = side_stream(half);
  const bool knn_on_side = side.ok && !g_prof_on;   // per-launch event timing assumes one stream
  for (int i = 0; i < NB; ++i) {
    const PoemBlock& k = w->blocks[i];
    float* xyz_in = p.xyz + (size_t)i * slab;
    float* xyz_out = p.xyz + (size_t)(i + 1) * slab;
    // BPS-token projections in one GEMM: K1 | K2 | kt_cross | v_cross | V1 | V2, all row-major
    {
      TagScope ts("pt_proj");
      POEM_TRY(linear("pt_proj", p.ptf, D, k.pt_proj, BP, 6 * D, D, ACT_NONE, nullptr, nullptr, p.KK, st));
    }
    // Query-stream segments: two chained layers per kernel for D <= 256 (qchain.cuh), separate GEMMs + LayerNorm otherwise
    const bool qchain = !g_force_unfused && (D == 128 || D == 256);
    if (qchain) {
      Chain2Params c = chain2_params(BQ, D);             // embedding -> attn.self.query
      c.out1_f32 = p.qe32;
      c.out2 = p.qp16;
      POEM_TRY(launch_chain2(D, "q_embed+mha_q", p.qf16, k.embedding, k.q1, c, st));
    } else {
      POEM_TRY(linear("q_embed", p.qf16, D, k.embedding, BQ, D, D, ACT_NONE, nullptr, p.qe32, p.qe16, st));
      POEM_TRY(linear("mha_q", p.qe16, D, k.q1, BQ, D, D, ACT_NONE, nullptr, nullptr, p.qp16, st));
    }
    // MHA 1
    POEM_TRY(launch_mha(p.qp16, D, 0, p.KK, 6 * D, 0, p.KK, 6 * D, 4 * D, p.ctx16, D, B, Q, P, D, dims->n_heads, st));
    if (qchain) {
      Chain2Params c = chain2_params(BQ, D);             // attn.output.dense + residual + LayerNorm -> cross_attn.self.query
      c.res32 = p.qe32, c.ln = 1, c.ln_g = k.ln1_g, c.ln_b = k.ln1_b;
      c.out1_f32 = p.a1_32;
      c.out2 = p.qp16;
      POEM_TRY(launch_chain2(D, "mha_out+ln+mha_q", p.ctx16, k.o1, k.q2, c, st));
    } else {
      POEM_TRY(linear("mha_out", p.ctx16, D, k.o1, BQ, D, D, ACT_NONE, p.qe32, p.tmp32, nullptr, st));
      POEM_TRY(launch_layernorm(p.tmp32, k.ln1_g, k.ln1_b, p.a1_32, p.a1_16, BQ, D, st));
      POEM_TRY(linear("mha_q", p.a1_16, D, k.q2, BQ, D, D, ACT_NONE, nullptr, nullptr, p.qp16, st));
    }
    // MHA 2
    POEM_TRY(launch_mha(p.qp16, D, 0, p.KK, 6 * D, D, p.KK, 6 * D, 5 * D, p.ctx16, D, B, Q, P, D, dims->n_heads, st));
    if (qchain) {
      Chain2Params c = chain2_params(BQ, D);             // cross_attn.output.dense + residual + LayerNorm -> self-attention q | k | v
      c.res32 = p.a1_32, c.ln = 1, c.ln_g = k.ln2_g, c.ln_b = k.ln2_b;
      c.out1_f32 = p.a2_32;
      c.N2 = 3 * D, c.ld2 = 3 * D, c.out2 = p.qkv16;
      POEM_TRY(launch_chain2(D, "mha_out+ln+va_self_qkv", p.ctx16, k.o2, k.self_qkv, c, st));
    } else {
      POEM_TRY(linear("mha_out", p.ctx16, D, k.o2, BQ, D, D, ACT_NONE, p.a1_32, p.tmp32, nullptr, st));
      POEM_TRY(launch_layernorm(p.tmp32, k.ln2_g, k.ln2_b, p.a2_32, p.a2_16, BQ, D, st));
      POEM_TRY(linear("va_self_qkv", p.a2_16, D, k.self_qkv, BQ, 3 * D, D, ACT_NONE, nullptr, nullptr, p.qkv16, st));
    }
    // vector self-attention
    const bool anchors = (i == 0);
    if (!anchors) {   // neighbour indices of this block: computed on the side stream since the previous block ended
      if (knn_on_side) CUDA_TRY(cudaStreamWaitEvent(st, side.join[i], 0));
      else POEM_TRY(launch_block_knn(w, B, Q, P, p, xyz_in, pt_is_bps, st));
      if (g_nbr_export != nullptr) {   // test hook: the index sets this block is about to use
        const size_t T = (size_t)BQ * 32, T_all = (size_t)B_total * Q * 32;
        if ((size_t)(NB - 1) * 2 * T_all > g_nbr_capacity) return fail(POEM_E_WORKSPACE, "neighbour export buffer too small");
        int32_t* dst = g_nbr_export + (size_t)(i - 1) * 2 * T_all + (size_t)b0 * Q * 32;
        CUDA_TRY(cudaMemcpyAsync(dst, p.idx_self, T * 4, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(dst + T_all, p.idx_cross, T * 4, cudaMemcpyDeviceToDevice, st));
      }
    }
    POEM_TRY(launch_vector_attention(&k.self_attn, p.qkv16, 3 * D, p.qkv16 + D, 3 * D, p.qkv16 + 2 * D, 3 * D, xyz_in,
                                     xyz_in, anchors ? nullptr : p.idx_self, anchors ? w->anchor_idx : nullptr,
                                     anchors ? w->anchor_xyz : nullptr, B, Q, Q, D, p.res16, p.t0, p.t1, p.t2, st));
    if (qchain) {
      Chain2Params c = chain2_params(BQ, D);             // query_self_attn.fc2 + residual -> query_cross_attn.w_qs (folded)
      c.res32 = p.a2_32;
      c.out1_f32 = p.f1_32;
      c.out2 = p.qc16;
      POEM_TRY(launch_chain2(D, "va_fc2+va_cross_q", p.res16, k.self_attn.fc2, k.cross_q, c, st));
    } else {
      POEM_TRY(linear("va_fc2", p.res16, D, k.self_attn.fc2, BQ, D, D, ACT_NONE, p.a2_32, p.f1_32, p.f1_16, st));
      // vector cross-attention (queries <- BPS tokens)
      POEM_TRY(linear("va_cross_q", p.f1_16, D, k.cross_q, BQ, D, D, ACT_NONE, nullptr, nullptr, p.qc16, st));
    }
    POEM_TRY(launch_vector_attention(&k.cross_attn, p.qc16, D, p.KK + 2 * D, 6 * D, p.KK + 3 * D, 6 * D, xyz_in,
                                     p.pt_xyz, anchors ? nullptr : p.idx_cross, anchors ? w->anchor_idx : nullptr,
                                     anchors ? w->anchor_xyz : nullptr, B, Q, P, D, p.res16, p.t0, p.t1, p.t2, st));
    if (qchain) {
      Chain2Params c = chain2_params(BQ, D);             // query_cross_attn.fc2 + residual -> reg_branch.0 + ReLU
      c.res32 = p.f1_32;
      c.out1_f32 = p.f2_32, c.out1_h16 = p.f2_16;
      c.act2 = ACT_RELU, c.out2 = p.r1_16;
      POEM_TRY(launch_chain2(D, "va_fc2+reg1", p.res16, k.cross_attn.fc2, k.reg1, c, st));
    } else {
      POEM_TRY(linear("va_fc2", p.res16, D, k.cross_attn.fc2, BQ, D, D, ACT_NONE, p.f1_32, p.f2_32, p.f2_16, st));
      // coordinate regression
      POEM_TRY(linear("reg1", p.f2_16, D, k.reg1, BQ, D, D, ACT_RELU, nullptr, nullptr, p.r1_16, st));
    }
    {
      if (!k.reg2_w || !k.reg2_b) return fail(POEM_E_NULL, "block %d: reg_branch.2 missing", i);
      const int threads = 256;
      prof_begin(st);
      launch_pdl(reg_out_kernel, dim3(((size_t)BQ * 32 + threads - 1) / threads), dim3(threads), (size_t)(0), st, 
          p.r1_16, k.reg2_w, k.reg2_b, xyz_in, xyz_out, coords_out + (size_t)i * slab, centre, dims->radius, Q, D,
          BQ);
      LAUNCH_CHECK("reg_out_kernel");
    }
    if (knn_on_side && i + 1 < NB) {   // fork: 32-NN of block i+1 on the side stream, overlapping FFN / MHA of the main one
      CUDA_TRY(cudaEventRecord(side.fork[i + 1], st));
      CUDA_TRY(cudaStreamWaitEvent(side.stream, side.fork[i + 1], 0));
      POEM_TRY(launch_block_knn(w, B, Q, P, p, xyz_out, pt_is_bps, side.stream));
      CUDA_TRY(cudaEventRecord(side.join[i + 1], side.stream));
    }
    // feed-forward (its output only feeds the next block)
    const bool last = (i == NB - 1);
    if (!last || dims->run_last_ffn) {
      POEM_TRY(linear("ffn1", p.f2_16, D, k.ffn1, BQ, 4 * D, D, ACT_GELU, nullptr, nullptr, p.ffn16, st));
      POEM_TRY(linear("ffn2", p.ffn16, 4 * D, k.ffn2, BQ, D, 4 * D, ACT_NONE, p.f2_32, p.tmp32, nullptr, st));
      float* dst32 = (last && out_feats) ? out_feats : p.qf32;
      POEM_TRY(launch_layernorm(p.tmp32, k.ln3_g, k.ln3_b, dst32, p.qf16, BQ, D, st));
    }
  }
  return POEM_OK;
}

extern "C" int poem_transformer_forward(const PoemDims* dims, const PoemWeights* w, int B, const float* query_xyz,
                                        const float* query_feat, const float* pt_xyz, const float* pt_feats,
                                        float* out_xyz, float* out_feats, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  POEM_TRY(check_dims(dims));
  if (!w || !query_xyz || !query_feat || !pt_xyz || !pt_feats || !out_xyz || !workspace)
    return fail(POEM_E_NULL, "transformer_forward: null pointer");
  if (B < 1) return fail(POEM_E_BADDIM, "batch=%d", B);
  if (out_feats && !dims->run_last_ffn) return fail(POEM_E_BADDIM, "out_feats needs dims->run_last_ffn");
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(POEM_E_ALIGN, "workspace must be 1024-byte aligned");
  DeviceGuard guard(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  Bump b{reinterpret_cast<uint8_t*>(workspace), 0};
  BlockPlan p;
  plan_blocks(dims, B, b, &p);
  if (b.off + 1024 > workspace_bytes) return fail(POEM_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, b.off + 1024);
  const size_t D = dims->embed_dims, BP = (size_t)B * dims->n_sample, BQ = (size_t)B * dims->n_query;
  CUDA_TRY(cudaMemcpyAsync(p.pt_xyz, pt_xyz, BP * 3 * 4, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(p.xyz, query_xyz, BQ * 3 * 4, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(p.qf32, query_feat, BQ * D * 4, cudaMemcpyDeviceToDevice, st));
  prof_begin(st);
  launch_pdl(f32_to_op16_kernel, dim3((unsigned)((BQ * D + 255) / 256)), dim3(256), (size_t)(0), st, query_feat, p.qf16, BQ * D);
  LAUNCH_CHECK("f32_to_op16_kernel");
  prof_begin(st);
  launch_pdl(f32_to_op16_kernel, dim3((unsigned)((BP * D + 255) / 256)), dim3(256), (size_t)(0), st, pt_feats, p.ptf, BP * D);
  LAUNCH_CHECK("f32_to_op16_kernel");
  return run_blocks(dims, w, B, p, nullptr, out_xyz, out_feats, /*pt_is_bps=*/false, st);
}

// a16: flat_verts + mano_linear + rot6d -> axis-angle + MANO forward; `flat` is B*D floats of scratch.
static int launch_parametric_tail(const PoemDims* dims, const PoemManoTail* m, int B, const float* feats,
                                  const float* ref_joints, float* coords, float* pose, float* shape, float* flat,
                                  cudaStream_t st) {
  if (!m->flat_w || !m->flat_b || !m->lin_w || !m->lin_b || !m->v_template || !m->shapedirs || !m->posedirs ||
      !m->j_regressor || !m->skin_weights)
    return fail(POEM_E_NULL, "parametric tail: null weight / MANO parameter");
  if (dims->n_query != 21 + kManoVerts) return fail(POEM_E_BADDIM, "parametric tail needs n_query = 799");
  if (dims->center_idx < 0 || dims->center_idx >= 21) return fail(POEM_E_BADDIM, "center_idx=%d", dims->center_idx);
  const int rows = B * dims->embed_dims;
  prof_begin(st);
  launch_pdl(flat_verts_kernel, dim3((unsigned)(((size_t)rows * 32 + 255) / 256)), dim3(256), (size_t)(0), st, feats, m->flat_w, m->flat_b, flat,
                                                                                 dims->n_query, rows);
  LAUNCH_CHECK("flat_verts_kernel");
  ManoTailArgs a;
  a.lin_w = m->lin_w, a.lin_b = m->lin_b;
  a.v_template = m->v_template, a.shapedirs = m->shapedirs, a.posedirs = m->posedirs;
  a.j_regressor = m->j_regressor, a.skin_weights = m->skin_weights;
  a.flat = flat, a.ref_joints = ref_joints;
  a.coords = coords, a.pose_out = pose, a.shape_out = shape;
  a.D = dims->embed_dims, a.center_idx = dims->center_idx;
  prof_begin(st);
  launch_pdl(mano_tail_kernel, dim3(B), dim3(kManoThreads), (size_t)(0), st, a);
  LAUNCH_CHECK("mano_tail_kernel");
  return POEM_OK;
}

extern "C" size_t poem_parametric_tail_workspace_bytes(const PoemDims* dims, int batch) {
  if (check_dims(dims) != POEM_OK || batch < 1) return 0;
  return (size_t)batch * dims->embed_dims * sizeof(float) + 1024;
}

extern "C" int poem_parametric_tail(const PoemDims* dims, const PoemManoTail* mano, int batch, const float* query_feats,
                                    const float* reference_joints, float* coords, float* pred_pose, float* pred_shape,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  POEM_TRY(check_dims(dims));
  if (!mano || !query_feats || !coords || !pred_pose || !pred_shape || !workspace)
    return fail(POEM_E_NULL, "parametric_tail: null pointer");
  if (batch < 1) return fail(POEM_E_BADDIM, "batch=%d", batch);
  if (workspace_bytes < (size_t)batch * dims->embed_dims * sizeof(float))
    return fail(POEM_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes,
                (size_t)batch * dims->embed_dims * sizeof(float));
  return launch_parametric_tail(dims, mano, batch, query_feats, reference_joints, coords, pred_pose, pred_shape,
                                reinterpret_cast<float*>(workspace), (cudaStream_t)stream);
}

static int head_forward_impl(const PoemDims* dims, const PoemWeights* w, const PoemInputs* in, float* out_coords,
                             float* out_feats, const PoemManoTail* mano, float* pred_pose, float* pred_shape,
                             void* workspace, size_t workspace_bytes, void* stream);

extern "C" int poem_head_forward(const PoemDims* dims, const PoemWeights* w, const PoemInputs* in, float* out_coords,
                                 float* out_feats, void* workspace, size_t workspace_bytes, void* stream) {
  return head_forward_impl(dims, w, in, out_coords, out_feats, nullptr, nullptr, nullptr, workspace, workspace_bytes,
                           stream);
}

extern "C" int poem_head_forward_parametric(const PoemDims* dims, const PoemWeights* w, const PoemManoTail* mano,
                                            const PoemInputs* in, float* out_coords, float* pred_pose,
                                            float* pred_shape, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dims || !mano || !pred_pose || !pred_shape) return fail(POEM_E_NULL, "head_forward_parametric: null pointer");
  PoemDims d = *dims;
  d.run_last_ffn = 1;   // the tail reads the last block's feed-forward output
  return head_forward_impl(&d, w, in, out_coords, nullptr, mano, pred_pose, pred_shape, workspace, workspace_bytes,
                           stream);
}

static int head_forward_impl(const PoemDims* dims, const PoemWeights* w, const PoemInputs* in, float* out_coords,
                             float* out_feats, const PoemManoTail* mano, float* pred_pose, float* pred_shape,
                             void* workspace, size_t workspace_bytes, void* stream) {
  POEM_TRY(check_dims(dims));
  if (!w || !in || !out_coords || !workspace) return fail(POEM_E_NULL, "head_forward: null pointer");
  if (!in->view_counts || !in->mlvl_feat || !in->cam_intr || !in->cam_extr || !in->reference_joints)
    return fail(POEM_E_NULL, "head_forward: null input");
  if (!w->pos_table || !w->query_embed || !w->bps || !w->anchor_xyz || !w->anchor_idx || !w->template_xyz)
    return fail(POEM_E_NULL, "head_forward: null constant table");
  const int B = in->batch, NV = in->n_images;
  if (B < 1 || NV < B) return fail(POEM_E_BADDIM, "batch=%d n_images=%d", B, NV);
  if (out_feats && !dims->run_last_ffn) return fail(POEM_E_BADDIM, "out_feats needs dims->run_last_ffn");
  const int D = dims->embed_dims, C = dims->in_channels, P = dims->n_sample, Q = dims->n_query;
  const int F = 256, H = D / 2;
  const int BQ = B * Q, BP = B * P;
  const long long R = (long long)NV * P;
  if (R > 0x7fffffffLL || (long long)BQ * 32 > 0x7fffffffLL) return fail(POEM_E_BADDIM, "problem too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(POEM_E_ALIGN, "workspace must be 1024-byte aligned");
  DeviceGuard guard(workspace);
  Bump bump{reinterpret_cast<uint8_t*>(workspace), 0};
  HeadPlan h;
  BlockPlan p;
  plan_head(dims, B, NV, bump, &h);
  plan_blocks(dims, B, bump, &p);
  if (bump.off + 1024 > workspace_bytes)
    return fail(POEM_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, bump.off + 1024);

  ViewTables vt;
  POEM_TRY(upload_view_tables(in->view_counts, B, NV, P, dims->max_views, h.tables, &vt, st));

  // ---- a2: x = input_proj(feat) + positional term, written channel-planar (NV, D, 256) fp32
  {
    dim3 grid((F + 31) / 32, (C + 31) / 32, NV), block(32, 8);
    prof_begin(st);
    launch_pdl(nchw_to_rows_op16_kernel, dim3(grid), dim3(block), (size_t)(0), st, in->mlvl_feat, h.featT, C, F);
    LAUNCH_CHECK("nchw_to_rows_op16_kernel");
    GemmEpilogue e = epi_default(D);
    e.bias = w->input_proj.b;
    e.res_mode = RES_POSADD;
    e.res_f32 = w->pos_table;
    e.res_ld = D;
    e.row_tab = vt.img_posrow;
    e.trans_from = 0;
    e.t_rows = F;
    e.t_group_stride = (long long)D * F;
    e.out_t_f32 = h.xmap;
    TagScope ts("input_proj");
    POEM_TRY(launch_gemm(h.featT, C, W16(w->input_proj), C, NV * F, D, C, e, st));
  }
  // ---- a3/a7: centre, normalised point sets
  prof_begin(st);
  // the hand centre is ALWAYS joint 9 (ptEmb_head.py:873); dims->center_idx only roots the MANO layer (a16)
  launch_pdl(gather_centre_kernel, dim3((B * 3 + 127) / 128), dim3(128), (size_t)(0), st, in->reference_joints, h.centre, kHandCentreJoint, B);
  LAUNCH_CHECK("gather_centre_kernel");
  {
    const int total = B * (P + Q) * 3;
    prof_begin(st);
    launch_pdl(normalise_points_kernel, dim3((total + 255) / 256), dim3(256), (size_t)(0), st, w->bps, w->template_xyz, h.centre, p.pt_xyz, p.xyz, P,
                                                                Q, dims->radius, B, w->bps_chunk_box ? w->bps_perm : nullptr,
                                                                p.pt_xyz_sorted);
    LAUNCH_CHECK("normalise_points_kernel");
  }
  // ---- a4/a5/a6: projection + bilinear sampling + merge MLP0 + cross-view reduce.  Fused kernel (sample_merge.cuh)
  //      for D <= 256: nothing row-sized (X, H1, Mm) touches HBM; else (D = 512, or the test hook) the four-kernel chain.
  const bool fused_merge = !g_force_unfused && (D == 128 || D == 256) && P == SM_P && dims->max_views <= 16;
  if (fused_merge) {
    if (!w->merge0a.w || !w->merge0a.b || !w->merge0b.w || !w->merge0b.b) return fail(POEM_E_NULL, "merge_net_feature.0 missing");
    prof_begin(st);
    launch_pdl(camera_prep_kernel, dim3((NV + 63) / 64), dim3(64), (size_t)(0), st, in->cam_intr, in->cam_extr, h.proj, NV);
    LAUNCH_CHECK("camera_prep_kernel");
    prof_begin(st);
    launch_pdl(sample_taps_kernel, dim3((unsigned)(((size_t)NV * P + 255) / 256)), dim3(256), (size_t)(0), st, h.proj, w->bps, h.centre, vt.img_sample, h.taps, NV,
                                                                               dims->feat_h, dims->feat_w, 1.0f / in->inp_img_w,
                                                                               1.0f / in->inp_img_h,
                                                                               (D == 128 ? SmCfg<128>::SLOTS : SmCfg<256>::SLOTS) * 4);
    LAUNCH_CHECK("sample_taps_kernel");
    SmParams sp;
    sp.xmap = h.xmap, sp.taps = h.taps, sp.tile_start = vt.tile_start, sp.sample_views = vt.sample_views;
    sp.sample_rowbase = vt.sample_rowbase, sp.b0a = w->merge0a.b, sp.b0b = w->merge0b.b;
    sp.q1 = h.Q1, sp.s = h.S, sp.sigma = h.sigma, sp.n_samples = B, sp.n_tiles = vt.n_merge_tiles;
    POEM_TRY(D == 128 ? launch_sample_merge<128>(w, sp, st) : launch_sample_merge<256>(w, sp, st));
  } else {
  POEM_TRY(launch_project_sample(h.xmap, in->cam_intr, in->cam_extr, w->bps, h.centre, vt, h.proj, NV, D, P,
                                 dims->feat_h, dims->feat_w, in->inp_img_w, in->inp_img_h, h.X, st));
  POEM_TRY(linear("merge0a", h.X, D, w->merge0a, (int)R, D, D, ACT_RELU, nullptr, nullptr, h.H1, st));
  POEM_TRY(linear("merge0b", h.H1, D, w->merge0b, (int)R, H, D, ACT_NONE, nullptr, nullptr, h.Mm, st));
  {
    const int threads = 256;
    prof_begin(st);
    const unsigned blocks = (unsigned)(((size_t)BP * 32 + threads - 1) / threads);
    switch (H) {
      case 64: merge_reduce_kernel<2><<<blocks, threads, 0, st>>>(h.Mm, vt.sample_rowbase, vt.sample_views, h.S, h.sigma, P, BP); break;
      case 128: merge_reduce_kernel<4><<<blocks, threads, 0, st>>>(h.Mm, vt.sample_rowbase, vt.sample_views, h.S, h.sigma, P, BP); break;
      case 256: merge_reduce_kernel<8><<<blocks, threads, 0, st>>>(h.Mm, vt.sample_rowbase, vt.sample_views, h.S, h.sigma, P, BP); break;
      default: merge_reduce_kernel<16><<<blocks, threads, 0, st>>>(h.Mm, vt.sample_rowbase, vt.sample_views, h.S, h.sigma, P, BP); break;
    }
    LAUNCH_CHECK("merge_reduce_kernel");
  }
  }
  {   // MLP1 hidden layer on S / sigma: relu(W s / sigma + b / sigma) = H2 / sigma
    if (!w->merge1a.w) return fail(POEM_E_NULL, "merge_net_feature.1.0 missing");
    GemmEpilogue e = epi_default(H);
    e.bias = w->merge1a.b;
    e.act = ACT_RELU;
    e.row_sigma = h.sigma;
    e.sigma_mode = 1;
    e.out_op16 = h.H2;
    e.ld_op16 = H;
    TagScope ts("merge1a");
    POEM_TRY(launch_gemm(h.S, H, W16(w->merge1a), H, BP, H, H, e, st));
  }
  {
    if (!w->merge1b.w) return fail(POEM_E_NULL, "merge_net_feature.1.2 missing");
    GemmEpilogue e = epi_default(D);
    e.bias = w->merge1b.b;
    e.res_mode = RES_MERGE;
    e.res_op16 = fused_merge ? h.Q1 : h.X;
    e.res_ld = D;
    e.row_tab = fused_merge ? nullptr : vt.sample_rowbase;   // fused path: the token-first rows are already gathered
    e.row_cnt = vt.sample_views;
    e.rows_per_group = P;
    e.row_sigma = h.sigma;
    e.sigma_mode = 2;
    e.out_op16 = p.ptf;
    e.ld_op16 = D;
    TagScope ts("merge1b");
    POEM_TRY(launch_gemm(h.H2, H, W16(w->merge1b), H, BP, D, H, e, st));
  }
  if (g_ptf_export != nullptr) {   // test hook: stage boundary a6 (merged BPS features, the `pt_feats` of ptEmb_head.py:926)
    if ((size_t)BP * D > g_ptf_capacity) return fail(POEM_E_WORKSPACE, "pt_feats export buffer too small");
    CUDA_TRY(cudaMemcpyAsync(g_ptf_export, p.ptf, (size_t)BP * D * sizeof(op16), cudaMemcpyDeviceToDevice, st));
  }
  // ---- a7: query features
  {
    const size_t total = (size_t)BQ * D;
    prof_begin(st);
    launch_pdl(broadcast_queries_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), st, w->query_embed, p.qf32, p.qf16, Q * D,
                                                                             total);
    LAUNCH_CHECK("broadcast_queries_kernel");
  }
  // ---- a8-a15.  The blocks of two (POEM_SPLIT=n: up to four) slices of the batch run on their own streams: every kernel of the path is a
  // persistent grid with a tail (the attention kernel's last wave, the last partial round of tiles of the fused
  // kernels, the 1.35 tiles per CTA of the query-stream kernels), and samples are independent, so the other half's
  // kernels fill the SMs a tail leaves idle.  POEM_SPLIT=1 (or a profiled run) keeps everything on one stream.
  int parts = g_prof_on ? 1 : split_parts();
  while (parts > 1 && B / parts < 4) --parts;       // at least 4 samples per slice
  for (int k = 0; k < parts && parts > 1; ++k)
    if (!side_stream(k).ok || !side_stream(kMaxParts + k).ok) parts = 1;
  if (parts <= 1) {
    POEM_TRY(run_blocks(dims, w, B, p, h.centre, out_coords, out_feats, /*pt_is_bps=*/true, st));
  } else {
    SideStream& fork = side_stream(kMaxParts);       // its events fork / join the slices' streams
    CUDA_TRY(cudaEventRecord(fork.fork[0], st));
    int b0 = 0;
    for (int k = 0; k < parts; ++k) {
      const int Bk = B / parts + (k < B % parts ? 1 : 0);
      cudaStream_t sk = (k == 0) ? st : side_stream(kMaxParts + k).stream;
      if (k > 0) CUDA_TRY(cudaStreamWaitEvent(sk, fork.fork[0], 0));
      const BlockPlan pk = slice_plan(p, dims, b0);
      POEM_TRY(run_blocks(dims, w, Bk, pk, h.centre + (size_t)b0 * 3, out_coords + (size_t)b0 * Q * 3,
                          out_feats ? out_feats + (size_t)b0 * Q * D : nullptr, true, sk, B, b0, k));
      if (k > 0) {
        CUDA_TRY(cudaEventRecord(side_stream(kMaxParts + k).join[0], sk));
        CUDA_TRY(cudaStreamWaitEvent(st, side_stream(kMaxParts + k).join[0], 0));
      }
      b0 += Bk;
    }
  }
  // ---- a16: the last block's joints / vertices are replaced by the MANO output (tmp32 is free after the last LayerNorm)
  if (mano)
    POEM_TRY(launch_parametric_tail(dims, mano, B, out_feats ? out_feats : p.qf32, in->reference_joints,
                                    out_coords + (size_t)(dims->n_blocks - 1) * BQ * 3, pred_pose, pred_shape, p.tmp32,
                                    st));
  return POEM_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer variant
// ------------------------------------------------------------------------------------------------
static size_t align1k(size_t x) { return (x + 1023) & ~size_t(1023); }
static size_t staging_slot_bytes(const PoemDims* d, int B, int NV) {
  return align1k((size_t)NV * d->in_channels * 256 * 4) + align1k((size_t)NV * 9 * 4) + align1k((size_t)NV * 16 * 4) +
         align1k((size_t)B * 63 * 4) + align1k((size_t)d->n_blocks * B * d->n_query * 3 * 4) +
         align1k((size_t)B * 58 * 4) /* pred_pose | pred_shape of a parametric head */ + 1024;
}
// two staging slots: the host->device copy of call i + 1 overlaps the kernels of call i
extern "C" size_t poem_staging_bytes(const PoemDims* d, int B, int NV) {
  if (check_dims(d) != POEM_OK) return 0;
  return 2 * staging_slot_bytes(d, B, NV) + 2048;
}

// copy stream + events of the host-buffer entry point (per host thread and device)
struct HostPipe {
  cudaStream_t copy = nullptr;
  cudaEvent_t h2d_done[2], slot_free[2];
  int next = 0;
  bool ok = false;
  const void* staging = nullptr;   // staging buffer of the previous call: a different one means the slots moved
  size_t staging_bytes = 0;
};
static HostPipe& host_pipe() {
  static thread_local HostPipe per_dev[kMaxDevices];
  HostPipe& p = per_dev[current_device()];
  if (!p.ok && cudaStreamCreateWithFlags(&p.copy, cudaStreamNonBlocking) == cudaSuccess) {
    p.ok = true;
    for (int i = 0; i < 2; ++i)
      p.ok = p.ok && cudaEventCreateWithFlags(&p.h2d_done[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&p.slot_free[i], cudaEventDisableTiming) == cudaSuccess;
  }
  return p;
}

static int head_forward_host_impl(const PoemDims* dims, const PoemWeights* w, const PoemManoTail* mano,
                                  const PoemInputs* hin, float* host_out, float* host_pose, float* host_shape,
                                  void* staging, size_t staging_bytes, void* workspace, size_t workspace_bytes,
                                  void* stream);

extern "C" int poem_head_forward_host(const PoemDims* dims, const PoemWeights* w, const PoemInputs* hin,
                                      float* host_out, void* staging, size_t staging_bytes, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  return head_forward_host_impl(dims, w, nullptr, hin, host_out, nullptr, nullptr, staging, staging_bytes, workspace,
                                workspace_bytes, stream);
}

extern "C" int poem_head_forward_parametric_host(const PoemDims* dims, const PoemWeights* w, const PoemManoTail* mano,
                                                 const PoemInputs* hin, float* host_out, float* host_pose,
                                                 float* host_shape, void* staging, size_t staging_bytes,
                                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!mano || !host_pose || !host_shape) return fail(POEM_E_NULL, "head_forward_parametric_host: null pointer");
  return head_forward_host_impl(dims, w, mano, hin, host_out, host_pose, host_shape, staging, staging_bytes, workspace,
                                workspace_bytes, stream);
}

static int head_forward_host_impl(const PoemDims* dims, const PoemWeights* w, const PoemManoTail* mano,
                                  const PoemInputs* hin, float* host_out, float* host_pose, float* host_shape,
                                  void* staging, size_t staging_bytes, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  POEM_TRY(check_dims(dims));
  if (!hin || !host_out || !workspace || !staging) return fail(POEM_E_NULL, "head_forward_host: null pointer");
  if (reinterpret_cast<uintptr_t>(staging) & 1023) return fail(POEM_E_ALIGN, "staging must be 1024-byte aligned");
  const int B = hin->batch, NV = hin->n_images;
  const size_t slot_bytes = staging_slot_bytes(dims, B, NV);
  // the two slots sit at fixed offsets 0 and staging_bytes / 2, whatever the shape of this call: consecutive calls
  // with different (batch, views) — the last partial batch of an epoch, ragged view counts — must not overlap the
  // slot the previous call is still reading
  const size_t slot_stride = (staging_bytes / 2) & ~size_t(1023);
  if (slot_stride < slot_bytes) return fail(POEM_E_WORKSPACE, "staging %zu < required %zu", staging_bytes, 2 * slot_bytes + 2048);
  DeviceGuard guard(workspace);
  if (workspace_bytes < poem_workspace_bytes(dims, B, NV))
    return fail(POEM_E_WORKSPACE, "workspace %zu < required %zu", workspace_bytes, poem_workspace_bytes(dims, B, NV));
  cudaStream_t st = (cudaStream_t)stream;
  // Inputs travel on a copy stream into one of two staging slots, so the transfer of the next call overlaps the
  // kernels of this one; a slot is reused only after the call that read it has finished (slot_free).  A capturing
  // stream (CUDA graph) or a failed stream creation falls back to copying on the compute stream.
  HostPipe& hp = host_pipe();
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  const bool piped = hp.ok && cap == cudaStreamCaptureStatusNone && !g_prof_on;
  const int slot = piped ? hp.next : 0;
  if (piped) hp.next ^= 1;
  cudaStream_t cs = piped ? hp.copy : st;
  if (piped && (hp.staging != staging || hp.staging_bytes != staging_bytes)) {
    // another staging buffer (or size) than the previous call's: its slot layout is unrelated, wait for both slots
    for (int k = 0; k < 2; ++k) CUDA_TRY(cudaStreamWaitEvent(cs, hp.slot_free[k], 0));
    hp.staging = staging;
    hp.staging_bytes = staging_bytes;
  }
  Bump b{reinterpret_cast<uint8_t*>(staging) + (size_t)slot * slot_stride, 0};
  const size_t n_feat = (size_t)NV * dims->in_channels * 256;
  float* d_feat = b.take<float>(n_feat);
  float* d_intr = b.take<float>((size_t)NV * 9);
  float* d_extr = b.take<float>((size_t)NV * 16);
  float* d_ref = b.take<float>((size_t)B * 63);
  const size_t n_out = (size_t)dims->n_blocks * B * dims->n_query * 3;
  float* d_out = b.take<float>(n_out);
  float* d_pose = b.take<float>((size_t)B * 58);   // pred_pose (B,48) then pred_shape (B,10)
  float* d_shape = d_pose + (size_t)B * 48;
  if (piped) CUDA_TRY(cudaStreamWaitEvent(cs, hp.slot_free[slot], 0));   // no-op until the slot has been used once
  CUDA_TRY(cudaMemcpyAsync(d_feat, hin->mlvl_feat, n_feat * 4, cudaMemcpyHostToDevice, cs));
  CUDA_TRY(cudaMemcpyAsync(d_intr, hin->cam_intr, (size_t)NV * 9 * 4, cudaMemcpyHostToDevice, cs));
  CUDA_TRY(cudaMemcpyAsync(d_extr, hin->cam_extr, (size_t)NV * 16 * 4, cudaMemcpyHostToDevice, cs));
  CUDA_TRY(cudaMemcpyAsync(d_ref, hin->reference_joints, (size_t)B * 63 * 4, cudaMemcpyHostToDevice, cs));
  if (piped) {
    CUDA_TRY(cudaEventRecord(hp.h2d_done[slot], cs));
    CUDA_TRY(cudaStreamWaitEvent(st, hp.h2d_done[slot], 0));
  }
  PoemInputs din = *hin;
  din.mlvl_feat = d_feat;
  din.cam_intr = d_intr;
  din.cam_extr = d_extr;
  din.reference_joints = d_ref;
  if (mano) {
    POEM_TRY(poem_head_forward_parametric(dims, w, mano, &din, d_out, d_pose, d_shape, workspace, workspace_bytes, stream));
    CUDA_TRY(cudaMemcpyAsync(host_pose, d_pose, (size_t)B * 48 * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(host_shape, d_shape, (size_t)B * 10 * 4, cudaMemcpyDeviceToHost, st));
  } else {
    POEM_TRY(poem_head_forward(dims, w, &din, d_out, nullptr, workspace, workspace_bytes, stream));
  }
  CUDA_TRY(cudaMemcpyAsync(host_out, d_out, n_out * 4, cudaMemcpyDeviceToHost, st));
  if (piped) CUDA_TRY(cudaEventRecord(hp.slot_free[slot], st));
  return POEM_OK;
}
