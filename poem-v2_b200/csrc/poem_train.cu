// libpoem_train.so: launchers of the training-path primitives declared in include/poem_train.h.
// Kernels: tgemm.cuh (TF32 tcgen05 GEMM, all operand major-ness combinations, batch + split-K) and train_simt.cuh.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/poem_train.h"
#include "tgemm.cuh"
#include "train_simt.cuh"
#include "mano_bwd.cuh"

using namespace poem;

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define TR_CHECK(name)                                                                              \
  do {                                                                                              \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                             \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess) return fail(POEM_TR_E_CUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

extern "C" int poem_tr_abi_version(void) { return POEM_TR_ABI_VERSION; }
extern "C" const char* poem_tr_last_error(void) { return g_err; }
extern "C" long long poem_tr_kernel_launches(void) { return g_launches.load(); }

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
static int agg_cap() {   // resident-block cap per SM of the per-query vector-attention kernels (POEM_TR_AGG_CAP, experiments)
  static int c = 0;
  if (c == 0) {
    const char* e = getenv("POEM_TR_AGG_CAP");
    c = e ? atoi(e) : 32;
    if (c < 1) c = 1;
  }
  return c;
}
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64) ? dev : 0;
}
static inline int grid_for(long long n, int block = 256, int per_sm = 8) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------ tensor maps
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

// fp32 operand of the GEMM as a rank-4 tensor map (inner, rows, batch1, batch2).
//   K-major : stored [rows x K]  -> inner = K,    outer = rows, box (32, box_rows)
//   MN-major: stored [K x rows]  -> inner = rows, outer = K,    box (32, 32)
static int make_tmap_f32(CUtensorMap* tm, const float* base, int mn_major, long long rows, long long K, long long ld,
                         long long s1, long long s2, int nb1, int nb2, int box_rows, int* bc1, int* bc2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(POEM_TR_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  *bc1 = (s1 == 0 || nb1 == 1), *bc2 = (s2 == 0 || nb2 == 1);
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3) || (!*bc1 && (s1 & 3)) || (!*bc2 && (s2 & 3)))
    return fail(POEM_TR_E_ALIGN, "GEMM operand needs a 16-byte aligned base and pitches in multiples of 4 floats (base=%p ld=%lld s1=%lld s2=%lld)",
                (const void*)base, ld, s1, s2);
  const long long inner = mn_major ? rows : K, outer = mn_major ? K : rows;
  if (inner <= 0 || outer <= 0 || ld < inner) return fail(POEM_TR_E_BADARG, "GEMM operand extents (inner=%lld outer=%lld ld=%lld)", inner, outer, ld);
  cuuint64_t gdim[4] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)(*bc1 ? 1 : nb1), (cuuint64_t)(*bc2 ? 1 : nb2)};
  const cuuint64_t plane = (cuuint64_t)ld * 4ull * (cuuint64_t)outer;
  cuuint64_t gstride[3] = {(cuuint64_t)ld * 4ull, *bc1 ? plane : (cuuint64_t)s1 * 4ull, *bc2 ? plane : (cuuint64_t)s2 * 4ull};
  cuuint32_t box[4] = {32u, (cuuint32_t)(mn_major ? 32 : box_rows), 1u, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(POEM_TR_E_CUDA, "cuTensorMapEncodeTiled (fp32 rank 4) failed (%d)", (int)r);
  return POEM_TR_OK;
}

template <int BN, bool A_MN, bool B_MN, int STAGES>
static int launch_tgemm(const CUtensorMap& ta, const CUtensorMap& tb, const TgParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = TgCfg<BN, STAGES>;
  static bool attr_done[64] = {false};       // function attributes are per device
  const int dev = current_device();
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tgemm_kernel<BN, A_MN, B_MN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return fail(POEM_TR_E_CUDA, "tgemm smem attribute: %s", cudaGetErrorString(e));
    attr_done[dev] = true;
  }
  tgemm_kernel<BN, A_MN, B_MN, STAGES><<<grid, TG_THREADS, Cfg::kSmemBytes, st>>>(ta, tb, p);
  TR_CHECK("tgemm");
  return POEM_TR_OK;
}
template <int BN, int STAGES>
static int launch_tgemm_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const TgParams& p, dim3 grid,
                              cudaStream_t st) {
  if (a_mn) return b_mn ? launch_tgemm<BN, true, true, STAGES>(ta, tb, p, grid, st) : launch_tgemm<BN, true, false, STAGES>(ta, tb, p, grid, st);
  return b_mn ? launch_tgemm<BN, false, true, STAGES>(ta, tb, p, grid, st) : launch_tgemm<BN, false, false, STAGES>(ta, tb, p, grid, st);
}

extern "C" int poem_tr_gemm(const float* A, int a_mn, long long lda, long long a_s1, long long a_s2, const float* B,
                            int b_mn, long long ldb, long long b_s1, long long b_s2, float* C, long long ldc,
                            long long c_s1, long long c_s2, int M, int N, int K, int nb1, int nb2, float alpha,
                            const float* bias, int bias_on_m, int accumulate, int relu, const float* relu_mask, long long ld_mask,
                            int round_ops, int round_out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || nb1 < 1 || nb2 < 1) return fail(POEM_TR_E_BADARG, "poem_tr_gemm: bad arguments");
  const int BN = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  CUtensorMap ta, tb;
  TgParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_f32(&ta, A, a_mn, M, K, lda, a_s1, a_s2, nb1, nb2, TG_BM, &p.a_bc1, &p.a_bc2);
  if (rc) return rc;
  rc = make_tmap_f32(&tb, B, b_mn, N, K, ldb, b_s1, b_s2, nb1, nb2, BN, &p.b_bc1, &p.b_bc2);
  if (rc) return rc;
  p.M = M, p.N = N, p.K = K, p.nb1 = nb1, p.nb2 = nb2;
  p.alpha = alpha, p.bias = bias, p.bias_on_m = bias_on_m, p.relu = relu;
  p.C = C, p.ldc = ldc, p.c_s1 = c_s1, p.c_s2 = c_s2;
  const int tiles = ((M + TG_BM - 1) / TG_BM) * ((N + BN - 1) / BN) * nb1 * nb2;
  const bool batch_sum = (nb1 > 1 && c_s1 == 0) || (nb2 > 1 && c_s2 == 0);
  int splits = 1;
  const int kblocks = (K + TG_BK - 1) / TG_BK;
  if (tiles < num_sms() && kblocks >= 32) {      // tall reduction (wgrad): spread K over the idle SMs
    splits = (2 * num_sms() + tiles - 1) / tiles;
    if (splits > kblocks / 8) splits = kblocks / 8;
    if (splits < 1) splits = 1;
  }
  if (round_out || relu || relu_mask) splits = 1;   // these epilogues need the complete sum in one CTA
  const int kb_per_split = (kblocks + splits - 1) / splits;
  splits = (kblocks + kb_per_split - 1) / kb_per_split;
  p.splits = splits, p.k_per_split = kb_per_split * TG_BK;
  p.mode = (splits > 1 || batch_sum) ? TG_ATOMIC : (accumulate ? TG_ADD : TG_STORE);
  p.relu_mask = relu_mask, p.ld_mask = ld_mask, p.round_ops = round_ops & 3, p.round_out = round_out;
  if (round_out && p.mode != TG_STORE) return fail(POEM_TR_E_BADARG, "poem_tr_gemm: round_out needs a plain store");
  if (relu_mask && (nb1 * nb2 != 1 || p.mode != TG_STORE)) return fail(POEM_TR_E_BADARG, "poem_tr_gemm: relu_mask needs an unbatched plain store");
  if (relu && p.mode != TG_STORE) return fail(POEM_TR_E_BADARG, "poem_tr_gemm: relu needs a plain store (no split / accumulate)");
  if (p.mode == TG_ATOMIC && !accumulate) {      // atomic partial sums need a zeroed destination
    const int e1 = c_s1 == 0 ? 1 : nb1, e2 = c_s2 == 0 ? 1 : nb2;
    for (int i2 = 0; i2 < e2; ++i2)
      for (int i1 = 0; i1 < e1; ++i1) {
        cudaError_t e = cudaMemset2DAsync(C + i1 * c_s1 + i2 * c_s2, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
        if (e != cudaSuccess) return fail(POEM_TR_E_CUDA, "poem_tr_gemm memset: %s", cudaGetErrorString(e));
      }
  }
  dim3 grid((N + BN - 1) / BN, (M + TG_BM - 1) / TG_BM, nb1 * nb2 * splits);
  if (grid.y > 65535u || grid.z > 65535u) return fail(POEM_TR_E_BADARG, "poem_tr_gemm: grid too large (%u, %u)", grid.y, grid.z);
  // two-stage ring / three CTAs per SM: K <= 64 (the ring is the whole K loop), and the tall streaming GEMMs (many tiles,
  // no split): they are bound by bytes in flight and by the epilogue's store bursts, a third resident CTA buys more of both
  // (per-edge GEMM 0.47 -> 0.44 ms); the split-K wgrads keep three stages (0.26 vs 0.29 ms).
  const bool shallow = kb_per_split <= 2 || (splits == 1 && !a_mn && tiles >= 4 * num_sms());      // K <= 64: a two-stage ring is the whole K loop; three CTAs per SM
  switch (BN) {
    case 32: return shallow ? launch_tgemm_major<32, 2>(a_mn, b_mn, ta, tb, p, grid, st) : launch_tgemm_major<32, 3>(a_mn, b_mn, ta, tb, p, grid, st);
    case 64: return shallow ? launch_tgemm_major<64, 2>(a_mn, b_mn, ta, tb, p, grid, st) : launch_tgemm_major<64, 3>(a_mn, b_mn, ta, tb, p, grid, st);
    default: return shallow ? launch_tgemm_major<128, 2>(a_mn, b_mn, ta, tb, p, grid, st) : launch_tgemm_major<128, 3>(a_mn, b_mn, ta, tb, p, grid, st);
  }
}

// ------------------------------------------------------------------------------------------------ SIMT launchers
#define ST static_cast<cudaStream_t>(stream)

extern "C" int poem_tr_relu(float* y, long long n, void* stream) {
  tr_relu_kernel<<<grid_for(n), 256, 0, ST>>>(y, n);
  TR_CHECK("relu");
  return 0;
}
extern "C" int poem_tr_relu_bwd(float* dy, const float* y, long long n, void* stream) {
  tr_relu_bwd_kernel<<<grid_for(n), 256, 0, ST>>>(dy, y, n);
  TR_CHECK("relu_bwd");
  return 0;
}
extern "C" int poem_tr_gelu(const float* x, float* y, long long n, void* stream) {
  tr_gelu_kernel<<<grid_for(n), 256, 0, ST>>>(x, y, n);
  TR_CHECK("gelu");
  return 0;
}
extern "C" int poem_tr_gelu_bwd(float* dy, const float* x, long long n, void* stream) {
  tr_gelu_bwd_kernel<<<grid_for(n), 256, 0, ST>>>(dy, x, n);
  TR_CHECK("gelu_bwd");
  return 0;
}
extern "C" int poem_tr_axpy(float* y, const float* x, float a, long long n, void* stream) {
  tr_axpy_kernel<<<grid_for(n), 256, 0, ST>>>(y, x, a, n);
  TR_CHECK("axpy");
  return 0;
}
extern "C" int poem_tr_affine_rows(const float* x, const float* off, float a, float* out, long long rows,
                                   int rows_per_group, int n_groups, int cols, void* stream) {
  tr_affine_rows_kernel<<<grid_for(rows * cols), 256, 0, ST>>>(x, off, a, out, rows, rows_per_group, n_groups, cols);
  TR_CHECK("affine_rows");
  return 0;
}
extern "C" int poem_tr_colsum(const float* dy, long long ld, long long M, int N, float* out, void* stream) {
  long long slabs = (M + 255) / 256;
  if (slabs > 4 * num_sms()) slabs = 4 * num_sms();
  if (slabs < 1) slabs = 1;
  dim3 grid((N + 31) / 32, (unsigned)slabs), block(32, 8);
  tr_colsum_kernel<<<grid, block, 0, ST>>>(dy, ld, M, N, out);
  TR_CHECK("colsum");
  return 0;
}
extern "C" int poem_tr_rowsum_groups(const float* x, long long rows, int cols, int group, float* out, void* stream) {
  tr_rowsum_groups_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST>>>(x, rows, cols, group, out);
  TR_CHECK("rowsum_groups");
  return 0;
}
extern "C" int poem_tr_sum_batch(const float* x, int B, long long n, float* out, void* stream) {
  tr_sum_batch_kernel<<<grid_for(n), 256, 0, ST>>>(x, B, n, out);
  TR_CHECK("sum_batch");
  return 0;
}
extern "C" int poem_tr_bcast_batch(const float* x, int B, long long n, float* out, void* stream) {
  tr_bcast_batch_kernel<<<grid_for(n * B), 256, 0, ST>>>(x, B, n, out);
  TR_CHECK("bcast_batch");
  return 0;
}

extern "C" int poem_tr_layernorm(const float* x, const float* res, const float* gamma, const float* beta, float eps,
                                 float* y, float* xhat, float* rstd, long long M, int D, void* stream) {
  if (D > 1024) return fail(POEM_TR_E_BADARG, "layernorm: D = %d > 1024", D);
  const int wpb = 8;
  const unsigned grid = (unsigned)((M + wpb - 1) / wpb);
  if (D <= 128) tr_layernorm_fwd_kernel<4><<<grid, wpb * 32, 0, ST>>>(x, res, gamma, beta, eps, y, xhat, rstd, M, D);
  else if (D <= 256) tr_layernorm_fwd_kernel<8><<<grid, wpb * 32, 0, ST>>>(x, res, gamma, beta, eps, y, xhat, rstd, M, D);
  else if (D <= 512) tr_layernorm_fwd_kernel<16><<<grid, wpb * 32, 0, ST>>>(x, res, gamma, beta, eps, y, xhat, rstd, M, D);
  else tr_layernorm_fwd_kernel<32><<<grid, wpb * 32, 0, ST>>>(x, res, gamma, beta, eps, y, xhat, rstd, M, D);
  TR_CHECK("layernorm");
  return 0;
}
extern "C" int poem_tr_layernorm_bwd(const float* dy, const float* xhat, const float* rstd, const float* gamma, float* dx,
                                     float* dgamma, float* dbeta, long long M, int D, void* stream) {
  if (D > 1024) return fail(POEM_TR_E_BADARG, "layernorm_bwd: D = %d > 1024", D);
  const int wpb = 8;
  long long g = (M + wpb - 1) / wpb;
  if (g > 2 * num_sms()) g = 2 * num_sms();
  const size_t sh = (size_t)2 * D * 4;
  if (D <= 128) tr_layernorm_bwd_kernel<4><<<(unsigned)g, wpb * 32, sh, ST>>>(dy, xhat, rstd, gamma, dx, dgamma, dbeta, M, D);
  else if (D <= 256) tr_layernorm_bwd_kernel<8><<<(unsigned)g, wpb * 32, sh, ST>>>(dy, xhat, rstd, gamma, dx, dgamma, dbeta, M, D);
  else if (D <= 512) tr_layernorm_bwd_kernel<16><<<(unsigned)g, wpb * 32, sh, ST>>>(dy, xhat, rstd, gamma, dx, dgamma, dbeta, M, D);
  else tr_layernorm_bwd_kernel<32><<<(unsigned)g, wpb * 32, sh, ST>>>(dy, xhat, rstd, gamma, dx, dgamma, dbeta, M, D);
  TR_CHECK("layernorm_bwd");
  return 0;
}

extern "C" int poem_tr_dropout(const float* x, float* y, long long n, float p, const unsigned long long* seed,
                               unsigned long long site, void* stream) {
  if (!(p >= 0.f && p < 1.f) || !seed) return fail(POEM_TR_E_BADARG, "dropout: p in [0, 1) and a device seed are needed");
  tr_dropout_kernel<<<grid_for(n), 256, 0, ST>>>(x, y, n, p, seed, site);
  TR_CHECK("dropout");
  return 0;
}
extern "C" int poem_tr_softmax_rows(float* S, long long rows, int L, float scale, float* P_dropped, float p_drop,
                                    const unsigned long long* seed, unsigned long long site, void* stream) {
  if (rows > 0x7fffffffLL) return fail(POEM_TR_E_BADARG, "softmax_rows: too many rows");
  if (P_dropped && (!seed || !(p_drop >= 0.f && p_drop < 1.f))) return fail(POEM_TR_E_BADARG, "softmax_rows: dropout needs a seed and p in [0, 1)");
  const bool al = ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(P_dropped)) & 15) == 0;
  if (al && L == 4096) tr_softmax_rows_kernel<4><<<(unsigned)rows, 256, 0, ST>>>(S, L, scale, P_dropped, p_drop, seed, site);
  else if (al && L == 1024) tr_softmax_rows_kernel<1><<<(unsigned)rows, 256, 0, ST>>>(S, L, scale, P_dropped, p_drop, seed, site);
  else tr_softmax_rows_kernel<0><<<(unsigned)rows, 256, 0, ST>>>(S, L, scale, P_dropped, p_drop, seed, site);
  TR_CHECK("softmax_rows");
  return 0;
}
extern "C" int poem_tr_softmax_rows_bwd(const float* P, float* dP, long long rows, int L, float scale, float p_drop,
                                        const unsigned long long* seed, unsigned long long site, void* stream) {
  if (rows > 0x7fffffffLL) return fail(POEM_TR_E_BADARG, "softmax_rows_bwd: too many rows");
  const bool al = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(dP)) & 15) == 0;
  if (al && L == 4096) tr_softmax_rows_bwd_kernel<4><<<(unsigned)rows, 256, 0, ST>>>(P, dP, L, scale, p_drop, seed, site);
  else if (al && L == 1024) tr_softmax_rows_bwd_kernel<1><<<(unsigned)rows, 256, 0, ST>>>(P, dP, L, scale, p_drop, seed, site);
  else tr_softmax_rows_bwd_kernel<0><<<(unsigned)rows, 256, 0, ST>>>(P, dP, L, scale, p_drop, seed, site);
  TR_CHECK("softmax_rows_bwd");
  return 0;
}

extern "C" int poem_tr_va_make_idx(const int32_t* local_idx, const int32_t* anchor_idx, int B, int Q, int R, int32_t* gidx,
                                   void* stream) {
  tr_va_make_idx_kernel<<<grid_for((long long)B * Q * TR_NBR), 256, 0, ST>>>(local_idx, anchor_idx, B, Q, R, gidx);
  TR_CHECK("va_make_idx");
  return 0;
}
extern "C" int poem_tr_va_rel(const float* q_xyz, const float* ref_xyz, const float* anchor_xyz, const int32_t* gidx,
                              long long E, float* rel, void* stream) {
  tr_va_rel_kernel<<<grid_for(E), 256, 0, ST>>>(q_xyz, ref_xyz, anchor_xyz, gidx, E, rel);
  TR_CHECK("va_rel");
  return 0;
}
extern "C" int poem_tr_lin3_relu(const float* rel, const float* W, const float* b, float* h, long long E, int D, void* stream) {
  if (D % 4 || D > 1024) return fail(POEM_TR_E_BADARG, "lin3_relu: D = %d", D);
  {
    dim3 block(D / 4, 256 / (D / 4) > 0 ? 256 / (D / 4) : 1);
    long long g = (E + block.y - 1) / block.y;
    if (g > 16LL * num_sms()) g = 16LL * num_sms();
    tr_lin3_relu_kernel<<<(unsigned)g, block, 0, ST>>>(rel, W, b, h, E, D);
  }
  TR_CHECK("lin3_relu");
  return 0;
}
extern "C" int poem_tr_lin3_bwd(const float* dh, const float* rel, const float* W, float* dW, float* db, float* drel,
                                long long E, int D, void* stream) {
  if (D > 1024) return fail(POEM_TR_E_BADARG, "lin3_bwd: D = %d > 1024", D);
  const int wpb = 8;
  long long blocks = (E + wpb - 1) / wpb;
  if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
  if (blocks < 1) blocks = 1;
  const size_t shb = (size_t)4 * D * 4;
  if (D <= 128) tr_lin3_bwd_kernel<4><<<(unsigned)blocks, wpb * 32, shb, ST>>>(dh, rel, W, dW, db, drel, E, D);
  else if (D <= 256) tr_lin3_bwd_kernel<8><<<(unsigned)blocks, wpb * 32, shb, ST>>>(dh, rel, W, dW, db, drel, E, D);
  else if (D <= 512) tr_lin3_bwd_kernel<16><<<(unsigned)blocks, wpb * 32, shb, ST>>>(dh, rel, W, dW, db, drel, E, D);
  else tr_lin3_bwd_kernel<32><<<(unsigned)blocks, wpb * 32, shb, ST>>>(dh, rel, W, dW, db, drel, E, D);
  TR_CHECK("lin3_bwd");
  return 0;
}
extern "C" int poem_tr_va_gather_t(const float* q, const float* ktab, const int32_t* gidx, const float* pos, float* t,
                                   long long E, int D, void* stream) {
  if (D % 4 || D > 1024) return fail(POEM_TR_E_BADARG, "va_gather_t: D = %d", D);
  {
    dim3 block(D / 4, 256 / (D / 4) > 0 ? 256 / (D / 4) : 1);
    long long g = (E + block.y - 1) / block.y;
    if (g > 16LL * num_sms()) g = 16LL * num_sms();
    tr_va_gather_t_kernel<<<(unsigned)g, block, 0, ST>>>(q, ktab, gidx, pos, t, E, D);
  }
  TR_CHECK("va_gather_t");
  return 0;
}
extern "C" int poem_tr_va_softmax_agg(float* a_w, const float* vtab, const float* pos, const int32_t* gidx, float scale,
                                      float* res, long long NQ, int D, void* stream) {
  tr_va_softmax_agg_kernel<<<(unsigned)(NQ < (long long)agg_cap() * num_sms() ? NQ : (long long)agg_cap() * num_sms()), (D < 256 ? D : 256), 0, ST>>>(a_w, vtab, pos, gidx, scale, res, NQ, D);
  TR_CHECK("va_softmax_agg");
  return 0;
}
extern "C" int poem_tr_va_softmax_agg_bwd(const float* dres, float* w_da, const float* vtab, const float* pos,
                                          const int32_t* gidx, float scale, float* dvp, long long NQ, int D, void* stream) {
  tr_va_softmax_agg_bwd_kernel<<<(unsigned)(NQ < (long long)agg_cap() * num_sms() ? NQ : (long long)agg_cap() * num_sms()), (D < 256 ? D : 256), 0, ST>>>(dres, w_da, vtab, pos, gidx, scale, dvp, NQ, D);
  TR_CHECK("va_softmax_agg_bwd");
  return 0;
}
extern "C" int poem_tr_va_scatter(float* dt_dpos, const float* dvp, const int32_t* gidx, float* dq, float* dktab,
                                  float* dvtab, long long NQ, int D, void* stream) {
  tr_va_scatter_kernel<<<(unsigned)(NQ < (long long)agg_cap() * num_sms() ? NQ : (long long)agg_cap() * num_sms()), (D < 256 ? D : 256), 0, ST>>>(dt_dpos, dvp, gidx, dq, dktab, dvtab, NQ, D);
  TR_CHECK("va_scatter");
  return 0;
}
extern "C" int poem_tr_va_drel_scatter(const float* drel, const int32_t* gidx, float* dxyz_q, float* dxyz_ref, long long NQ,
                                       void* stream) {
  tr_va_drel_scatter_kernel<<<grid_for(NQ * 3), 256, 0, ST>>>(drel, gidx, dxyz_q, dxyz_ref, NQ);
  TR_CHECK("va_drel_scatter");
  return 0;
}

extern "C" int poem_tr_lin_n3(const float* x, const float* W, const float* b, const float* base, float* y, long long M,
                              int D, void* stream) {
  const int wpb = 8;
  tr_lin_n3_kernel<<<(unsigned)((M + wpb - 1) / wpb), wpb * 32, 0, ST>>>(x, W, b, base, y, M, D);
  TR_CHECK("lin_n3");
  return 0;
}
extern "C" int poem_tr_lin_n3_bwd(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db,
                                  long long M, int D, int x_is_relu, void* stream) {
  long long g = M < 4 * num_sms() ? M : 4 * num_sms();
  if (g < 1) g = 1;
  tr_lin_n3_bwd_kernel<<<(unsigned)g, 256, 0, ST>>>(dy, x, W, dx, dW, db, M, D, x_is_relu);
  TR_CHECK("lin_n3_bwd");
  return 0;
}

extern "C" int poem_tr_project(const float* bps, const float* centre, const float* cam_intr, const float* cam_extr,
                               const int32_t* img_sample, int NV, int P, float inp_w, float inp_h, float* grid, void* stream) {
  tr_project_kernel<<<grid_for((long long)NV * P), 256, 0, ST>>>(bps, centre, cam_intr, cam_extr, img_sample, NV, P, inp_w, inp_h, grid);
  TR_CHECK("project");
  return 0;
}
extern "C" int poem_tr_sample(const float* planes, const float* grid, float* S, int NV, int D, int P, int hw, void* stream) {
  tr_sample_fwd_kernel<<<grid_for((long long)NV * P, 128, 16), 128, 0, ST>>>(planes, grid, S, NV, D, P, hw);
  TR_CHECK("sample");
  return 0;
}
extern "C" int poem_tr_sample_bwd(const float* dS, const float* grid_, float* dplanes, int NV, int D, int P, int hw, void* stream) {
  const size_t shb = (size_t)kSbCh * hw * hw * 4;
  static bool attr_done[64] = {false};       // function attributes are per device
  const int dev = current_device();
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tr_sample_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail(POEM_TR_E_CUDA, "sample_bwd smem attribute: %s", cudaGetErrorString(e));
    attr_done[dev] = true;
  }
  if (shb > 200 * 1024) return fail(POEM_TR_E_BADARG, "sample_bwd: feature map %d x %d too large", hw, hw);
  dim3 grid(NV, (D + kSbCh - 1) / kSbCh);
  tr_sample_bwd_kernel<<<grid, 512, shb, ST>>>(dS, grid_, dplanes, NV, D, P, hw);
  TR_CHECK("sample_bwd");
  return 0;
}

extern "C" int poem_tr_merge_agg(const float* m, const int32_t* row0, const int32_t* nviews, int B, int P, int Dm,
                                 float* agg, void* stream) {
  if (Dm > 256) return fail(POEM_TR_E_BADARG, "merge_agg: Dm = %d > 256", Dm);
  const long long total = (long long)B * P;
  tr_merge_agg_kernel<<<(unsigned)((total + 7) / 8), 256, 0, ST>>>(m, row0, nviews, P, Dm, agg, total);
  TR_CHECK("merge_agg");
  return 0;
}
extern "C" int poem_tr_merge_agg_bwd(const float* dagg, const float* m, const int32_t* row0, const int32_t* nviews, int B,
                                     int P, int Dm, float* dm, void* stream) {
  if (Dm > 256) return fail(POEM_TR_E_BADARG, "merge_agg_bwd: Dm = %d > 256", Dm);
  const long long total = (long long)B * P;
  tr_merge_agg_bwd_kernel<<<(unsigned)((total + 7) / 8), 256, 0, ST>>>(dagg, m, row0, nviews, P, Dm, dm, total);
  TR_CHECK("merge_agg_bwd");
  return 0;
}
extern "C" int poem_tr_merge_out(const float* X, const float* y, const int32_t* row0, const int32_t* nviews, int B, int P,
                                 int D, float* out, void* stream) {
  const long long total = (long long)B * P * D;
  tr_merge_out_kernel<<<grid_for(total), 256, 0, ST>>>(X, y, row0, nviews, P, D, out, total);
  TR_CHECK("merge_out");
  return 0;
}
extern "C" int poem_tr_merge_out_bwd(const float* dout, const int32_t* row0, const int32_t* nviews, int B, int P, int D,
                                     float* dX, float* dy, void* stream) {
  const long long total = (long long)B * P * D;
  tr_merge_out_bwd_kernel<<<grid_for(total), 256, 0, ST>>>(dout, row0, nviews, P, D, dX, dy, total);
  TR_CHECK("merge_out_bwd");
  return 0;
}

extern "C" int poem_tr_sumsq(const float* g, long long n, float* sumsq, void* stream) {
  tr_sumsq_kernel<<<grid_for(n, 256, 2), 256, 0, ST>>>(g, n, sumsq);
  TR_CHECK("sumsq");
  return 0;
}
extern "C" int poem_tr_clip_scale(float* g, long long n, const float* sumsq, float max_norm, void* stream) {
  tr_clip_scale_kernel<<<grid_for(n), 256, 0, ST>>>(g, n, sumsq, max_norm);
  TR_CHECK("clip_scale");
  return 0;
}

extern "C" int poem_tr_seg_sumsq(const float* g, const long long* off, const long long* len, int n_seg, float* sumsq, void* stream) {
  tr_seg_sumsq_kernel<<<n_seg, 256, 0, ST>>>(g, off, len, sumsq);
  TR_CHECK("seg_sumsq");
  return 0;
}
extern "C" int poem_tr_seg_clip(float* g, const long long* off, const long long* len, int n_seg, const float* sumsq,
                                float max_norm, void* stream) {
  tr_seg_clip_kernel<<<n_seg, 256, 0, ST>>>(g, off, len, sumsq, max_norm);
  TR_CHECK("seg_clip");
  return 0;
}
extern "C" int poem_tr_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, int step, void* stream) {
  if (step < 1) return fail(POEM_TR_E_BADARG, "adam: step counts from 1");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  tr_adam_kernel<<<grid_for(n), 256, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2);
  TR_CHECK("adam");
  return 0;
}
extern "C" int poem_tr_coord_loss(const float* coords, const float* gt_joints, const float* gt_verts, int n_blocks, int B,
                                  int n_joints, int n_verts, float w_joints, float w_verts, float* loss, float* dcoords,
                                  void* stream) {
  tr_coord_loss_kernel<<<grid_for((long long)n_blocks * B * (n_joints + n_verts) * 3), 256, 0, ST>>>(
      coords, gt_joints, gt_verts, n_blocks, B, n_joints, n_verts, w_joints, w_verts, loss, dcoords);
  TR_CHECK("coord_loss");
  return 0;
}

extern "C" int poem_tr_round_tf32(const float* x, float* y, long long n, void* stream) {
  tr_round_tf32_kernel<<<grid_for(n), 256, 0, ST>>>(x, y, n);
  TR_CHECK("round_tf32");
  return 0;
}

// ------------------------------------------------------------------------------------------------ parametric (MANO) tail
extern "C" int poem_tr_mano_tail(const float* feats, const float* flat_w, const float* flat_b, const float* lin_w,
                                 const float* lin_b, const float* v_template, const float* shapedirs, const float* posedirs,
                                 const float* j_regressor, const float* skin_weights, const float* ref_joints, int center_idx,
                                 int B, int Q, int D, float* flat, float* coords, float* pose, float* shape, void* stream) {
  if (Q != 21 + kManoVerts) return fail(POEM_TR_E_BADARG, "mano_tail: Q = %d, expected 799", Q);
  const int rows = B * D;
  flat_verts_kernel<<<(rows * 32 + 255) / 256, 256, 0, ST>>>(feats, flat_w, flat_b, flat, Q, rows);
  TR_CHECK("flat_verts");
  ManoTailArgs a;
  a.lin_w = lin_w, a.lin_b = lin_b, a.v_template = v_template, a.shapedirs = shapedirs, a.posedirs = posedirs;
  a.j_regressor = j_regressor, a.skin_weights = skin_weights, a.flat = flat, a.ref_joints = ref_joints;
  a.coords = coords, a.pose_out = pose, a.shape_out = shape, a.D = D, a.center_idx = center_idx;
  mano_tail_kernel<<<B, kManoThreads, 0, ST>>>(a);
  TR_CHECK("mano_tail");
  return 0;
}
extern "C" int poem_tr_mano_tail_bwd(const float* feats, const float* flat_w, const float* lin_w, const float* lin_b,
                                     const float* v_template, const float* shapedirs, const float* posedirs,
                                     const float* j_regressor, const float* skin_weights, int center_idx, int B, int Q, int D,
                                     const float* flat, const float* dcoords, const float* dpose, const float* dshape,
                                     float* dflat, float* dfeats, float* dflat_w, float* dflat_b, float* dlin_w, float* dlin_b,
                                     void* stream) {
  if (Q != 21 + kManoVerts) return fail(POEM_TR_E_BADARG, "mano_tail_bwd: Q = %d, expected 799", Q);
  ManoTailBwdArgs a;
  a.lin_w = lin_w, a.lin_b = lin_b, a.v_template = v_template, a.shapedirs = shapedirs, a.posedirs = posedirs;
  a.j_regressor = j_regressor, a.skin_weights = skin_weights, a.flat = flat, a.dcoords = dcoords, a.dpose = dpose;
  a.dshape = dshape, a.dflat = dflat, a.dlin_w = dlin_w, a.dlin_b = dlin_b, a.D = D, a.center_idx = center_idx;
  mano_tail_bwd_kernel<<<B, kManoThreads, 0, ST>>>(a);
  TR_CHECK("mano_tail_bwd");
  const int rows = B * D;
  int slabs = (rows + 63) / 64;
  if (slabs > 4 * num_sms()) slabs = 4 * num_sms();
  dim3 grid((Q + 31) / 32, slabs), block(32, 8);
  tr_flat_verts_bwd_kernel<<<grid, block, 0, ST>>>(dflat, feats, flat_w, dfeats, dflat_w, dflat_b, Q, rows);
  TR_CHECK("flat_verts_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------ compute_loss (head terms)
extern "C" int poem_tr_compute_loss(const float* coords_last, const float* gt_joints, const float* gt_verts,
                                    const float* j_regressor, const float* cam_intr, const float* cam_extr,
                                    const int32_t* img_sample, const float* target_joints_2d, int B, int NV, float img_scale,
                                    float w_joints, float w_verts, float w_joints_2d, float w_verts_2d, const float* pred_pose,
                                    const float* gt_pose, const float* pred_shape, const float* gt_shape, float w_pose,
                                    float w_shape, float* losses, float* dcoords_last, float* dpose, float* dshape, void* stream) {
  cudaError_t e = cudaMemsetAsync(losses, 0, 8 * sizeof(float), ST);
  if (e == cudaSuccess) e = cudaMemsetAsync(dcoords_last, 0, (size_t)B * kLossQ * 3 * sizeof(float), ST);
  if (e != cudaSuccess) return fail(POEM_TR_E_CUDA, "compute_loss memset: %s", cudaGetErrorString(e));
  TrLossArgs a;
  a.coords = coords_last, a.gt_joints = gt_joints, a.gt_verts = gt_verts, a.j_regressor = j_regressor;
  a.cam_intr = cam_intr, a.cam_extr = cam_extr, a.target_2d = target_joints_2d, a.img_sample = img_sample;
  a.B = B, a.NV = NV, a.img_scale = img_scale, a.w_j = w_joints, a.w_v = w_verts, a.w_j2d = w_joints_2d, a.w_v2d = w_verts_2d;
  a.losses = losses, a.dcoords = dcoords_last;
  tr_loss_3d_kernel<<<B, 256, 0, ST>>>(a);
  TR_CHECK("loss_3d");
  if (w_joints_2d != 0.f || w_verts_2d != 0.f) {
    if (!cam_intr || !cam_extr || !img_sample || !target_joints_2d) return fail(POEM_TR_E_BADARG, "compute_loss: the 2-D terms need cameras and targets");
    tr_loss_2d_kernel<<<grid_for((long long)NV * kLossQ), 256, 0, ST>>>(a);
    TR_CHECK("loss_2d");
  }
  if (pred_pose && gt_pose) {
    tr_loss_mse_kernel<<<1, 256, 0, ST>>>(pred_pose, gt_pose, B * 48, w_pose, losses, 5, dpose);
    TR_CHECK("loss_pose");
  }
  if (pred_shape && gt_shape) {
    tr_loss_mse_kernel<<<1, 256, 0, ST>>>(pred_shape, gt_shape, B * 10, w_shape, losses, 6, dshape);
    TR_CHECK("loss_shape");
  }
  return 0;
}

#if POEM_TG_TRACE
extern "C" int poem_tr_debug_trace(unsigned long long* host_out, int n) {   // experiments: scripts/tgemm_trace.py
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, g_tg_trace, (size_t)n * 8, 0, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
#endif
