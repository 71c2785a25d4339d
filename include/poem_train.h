/* C-ABI of the training-path primitives (libpoem_train.so, poem-v2_b200/csrc/poem_train.cu).
 *
 * SURVEY.md §8 row f3 ("training path"): what the reference gets from torch.autograd over
 * lib/models/heads/ptEmb_head.py:825-964, lib/models/bricks/pt_metro_transformer.py:34-200 and
 * lib/models/bricks/point_transformers.py:70-156 inside `scripts/train_ddp.py:96-116` (loss.backward(), clip_gradient,
 * optimizer.step()).  The library holds the device kernels only — fp32 tensors in HBM, TF32 tcgen05 GEMMs, SIMT kernels
 * for everything else; the forward/backward schedule of the head is host code (poem-v2_b200/train.py), the way the
 * reference's schedule is Python.  Every pointer is a DEVICE pointer, every tensor dense row-major fp32 unless stated;
 * `stream` is a cudaStream_t (NULL = default stream).  Functions return 0 or a negative POEM_TR_E_* code and leave a
 * message for poem_tr_last_error().  Gradient outputs documented "+=" ACCUMULATE into the caller's buffer.
 */
#ifndef POEM_TRAIN_H_
#define POEM_TRAIN_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POEM_TR_ABI_VERSION 1
#define POEM_TR_OK 0
#define POEM_TR_E_BADARG (-1)
#define POEM_TR_E_ALIGN (-2)
#define POEM_TR_E_CUDA (-3)

int poem_tr_abi_version(void);
const char* poem_tr_last_error(void);
long long poem_tr_kernel_launches(void);

/* C[b2,b1] (+)= alpha * op(A[b2,b1]) . op(B[b2,b1])^T (+ bias)         TF32 tensor cores, fp32 accumulate
 *   a_mn == 0: A stored [M x K] (row pitch lda);  a_mn == 1: A stored [K x M] (row pitch lda)   — same for B with N.
 *   batch: nb1 x nb2 problems; element strides (a_s1, a_s2), (b_s1, b_s2), (c_s1, c_s2); a stride of 0 shares the
 *   operand along that axis; c stride 0 with extent > 1 sums the batch into one C (atomic accumulation).
 *   bias: NULL, [N] (bias_on_m == 0) or [M] (bias_on_m == 1).  accumulate != 0: C += instead of C =.
 *   relu != 0: C = max(., 0) after the bias (plain stores only).  relu_mask != NULL ([M x N], pitch ld_mask): C = 0 where
 *   relu_mask <= 0 — the dgrad GEMM that feeds a ReLU's backward writes the masked gradient directly.
 *   round_ops: bit 0 / bit 1 = round operand A / B to TF32 (nearest) before the tensor core reads it (forward GEMMs: the
 *   1e-3 bound needs it); 0 = the tensor core truncates the fp32 operands (gradient GEMMs).
 *   round_out != 0: C is stored rounded to TF32 — for tensors that are only ever GEMM operands again (their consumer then
 *   passes round_ops without that operand's bit and skips the shared-memory pass).
 *   pitches must be multiples of 4 elements and bases 16-byte aligned (TMA).  This one primitive is the forward, dgrad and
 *   wgrad of every nn.Linear / 1x1 conv of the path and the five GEMMs of the attention core and its backward. */
int poem_tr_gemm(const float* A, int a_mn, long long lda, long long a_s1, long long a_s2, const float* B, int b_mn,
                 long long ldb, long long b_s1, long long b_s2, float* C, long long ldc, long long c_s1, long long c_s2,
                 int M, int N, int K, int nb1, int nb2, float alpha, const float* bias, int bias_on_m, int accumulate,
                 int relu, const float* relu_mask, long long ld_mask, int round_ops, int round_out, void* stream);

/* elementwise */
int poem_tr_relu(float* y, long long n, void* stream);
int poem_tr_relu_bwd(float* dy, const float* y, long long n, void* stream);          /* dy *= (y > 0) */
int poem_tr_gelu(const float* x, float* y, long long n, void* stream);               /* exact erf GELU */
int poem_tr_gelu_bwd(float* dy, const float* x, long long n, void* stream);
int poem_tr_round_tf32(const float* x, float* y, long long n, void* stream);        /* y = x rounded to TF32 (nearest) */
int poem_tr_axpy(float* y, const float* x, float a, long long n, void* stream);      /* y += a x */
int poem_tr_affine_rows(const float* x, const float* off, float a, float* out, long long rows, int rows_per_group,
                        int n_groups, int cols, void* stream);                       /* out = a x + off[group] */
int poem_tr_colsum(const float* dy, long long ld, long long M, int N, float* out, void* stream);        /* out[n] += */
int poem_tr_rowsum_groups(const float* x, long long rows, int cols, int group, float* out, void* stream);  /* out[row % group] += */
int poem_tr_sum_batch(const float* x, int B, long long n, float* out, void* stream);                     /* out += sum_b */
int poem_tr_bcast_batch(const float* x, int B, long long n, float* out, void* stream);

/* LayerNorm of (x + res) (res may be NULL), eps as given; saves xhat [M x D] and rstd [M] for the backward */
int poem_tr_layernorm(const float* x, const float* res, const float* gamma, const float* beta, float eps, float* y,
                      float* xhat, float* rstd, long long M, int D, void* stream);
int poem_tr_layernorm_bwd(const float* dy, const float* xhat, const float* rstd, const float* gamma, float* dx,
                          float* dgamma, float* dbeta, long long M, int D, void* stream);   /* dgamma, dbeta += */

/* Dropout without stored masks: element i of site `site` is kept iff mix(seed[0], site, i) >= p * 2^32 (seed: DEVICE
 * scalar the caller bumps every step, so captured graphs draw fresh masks).  y = keep ? x / (1 - p) : 0; the same call on
 * a gradient is the backward.  Reference: nn.Dropout(hidden_dropout_prob) at pt_metro_transformer.py:117,185-186 and in
 * the HF BertSelfOutput / BertOutput / BertSelfAttention the layers are built from (config/release: DROPOUT 0.1). */
int poem_tr_dropout(const float* x, float* y, long long n, float p, const unsigned long long* seed, unsigned long long site,
                    void* stream);

/* softmax over rows of length L.  P_dropped == NULL: S <- P = softmax(S * scale), stored TF32-rounded.  P_dropped != NULL
 * (attention-probability dropout): S <- P in full precision (the backward needs the un-dropped P), P_dropped <-
 * keep ? P / (1 - p_drop) : 0, TF32-rounded (the operand of P.V and P^T.dO).
 * backward: dS = P (d - sum P d) * scale over dP, d = dP, or with seed != NULL d = keep ? dP / (1 - p_drop) : 0 */
int poem_tr_softmax_rows(float* S, long long rows, int L, float scale, float* P_dropped, float p_drop,
                         const unsigned long long* seed, unsigned long long site, void* stream);
int poem_tr_softmax_rows_bwd(const float* P, float* dP, long long rows, int L, float scale, float p_drop,
                             const unsigned long long* seed, unsigned long long site, void* stream);

/* vector attention (32 neighbours per query); edge e = query * 32 + slot */
int poem_tr_va_make_idx(const int32_t* local_idx, const int32_t* anchor_idx, int B, int Q, int R, int32_t* gidx, void* stream);
int poem_tr_va_rel(const float* q_xyz, const float* ref_xyz, const float* anchor_xyz, const int32_t* gidx, long long E,
                   float* rel, void* stream);
int poem_tr_lin3_relu(const float* rel, const float* W, const float* b, float* h, long long E, int D, void* stream);
int poem_tr_lin3_bwd(const float* dh, const float* rel, const float* W, float* dW, float* db, float* drel /* or NULL */,
                     long long E, int D, void* stream);
int poem_tr_va_gather_t(const float* q, const float* ktab, const int32_t* gidx, const float* pos, float* t, long long E,
                        int D, void* stream);
int poem_tr_va_softmax_agg(float* a_w, const float* vtab, const float* pos, const int32_t* gidx, float scale, float* res,
                           long long NQ, int D, void* stream);
int poem_tr_va_softmax_agg_bwd(const float* dres, float* w_da, const float* vtab, const float* pos, const int32_t* gidx,
                               float scale, float* dvp, long long NQ, int D, void* stream);
int poem_tr_va_scatter(float* dt_dpos, const float* dvp, const int32_t* gidx, float* dq, float* dktab, float* dvtab,
                       long long NQ, int D, void* stream);
int poem_tr_va_drel_scatter(const float* drel, const int32_t* gidx, float* dxyz_q, float* dxyz_ref /* or NULL */,
                            long long NQ, void* stream);

/* reg_branch.2 : Linear(D, 3) (+ base coordinates) */
int poem_tr_lin_n3(const float* x, const float* W, const float* b, const float* base, float* y, long long M, int D, void* stream);
int poem_tr_lin_n3_bwd(const float* dy, const float* x, const float* W, float* dx, float* dW, float* db, long long M,
                       int D, int x_is_relu /* dx = 0 where x <= 0 */, void* stream);

/* camera projection of the BPS points + bilinear sampler (planes NCHW, hw x hw) and its scatter backward */
int poem_tr_project(const float* bps, const float* centre, const float* cam_intr, const float* cam_extr,
                    const int32_t* img_sample, int NV, int P, float inp_w, float inp_h, float* grid, void* stream);
int poem_tr_sample(const float* planes, const float* grid, float* S, int NV, int D, int P, int hw, void* stream);
int poem_tr_sample_bwd(const float* dS, const float* grid, float* dplanes, int NV, int D, int P, int hw, void* stream);

/* cross-view merge (raw `.view` regroup: rows row0[b] + p * n[b] + v) */
int poem_tr_merge_agg(const float* m, const int32_t* row0, const int32_t* nviews, int B, int P, int Dm, float* agg, void* stream);
int poem_tr_merge_agg_bwd(const float* dagg, const float* m, const int32_t* row0, const int32_t* nviews, int B, int P,
                          int Dm, float* dm, void* stream);
int poem_tr_merge_out(const float* X, const float* y, const int32_t* row0, const int32_t* nviews, int B, int P, int D,
                      float* out, void* stream);
int poem_tr_merge_out_bwd(const float* dout, const int32_t* row0, const int32_t* nviews, int B, int P, int D, float* dX,
                          float* dy, void* stream);

/* gradient clipping as lib/utils/net_utils.py:122-132 applies it (clip_grad_norm_ on every parameter tensor by itself):
 * sumsq[0] += |g|^2 ; g *= min(1, max_norm / (sqrt(sumsq[0]) + 1e-6)) */
int poem_tr_sumsq(const float* g, long long n, float* sumsq, void* stream);
int poem_tr_clip_scale(float* g, long long n, const float* sumsq, float max_norm, void* stream);

/* the same on a flat gradient buffer (segment s = [off[s], off[s] + len[s]), device arrays): one launch per pass */
int poem_tr_seg_sumsq(const float* g, const long long* off, const long long* len, int n_seg, float* sumsq, void* stream);
int poem_tr_seg_clip(float* g, const long long* off, const long long* len, int n_seg, const float* sumsq, float max_norm,
                     void* stream);
/* torch.optim.Adam step on flat buffers (lib/utils/net_utils.py:57-63): L2 weight decay added to the gradient,
 * bias-corrected first / second moments; `step` counts from 1 */
int poem_tr_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step, void* stream);
/* 3-D terms of compute_loss on the last block (lib/models/POEM.py:398-412): loss[0] += w_joints * MSE(joints) +
 * w_verts * L1(verts); dcoords [n_blocks, B, n_joints + n_verts, 3] = d loss / d all_coords_preds */
int poem_tr_coord_loss(const float* coords, const float* gt_joints, const float* gt_verts, int n_blocks, int B,
                       int n_joints, int n_verts, float w_joints, float w_verts, float* loss, float* dcoords, void* stream);

/* The head's terms of `PtEmbedMultiviewStereoV2.compute_loss` (lib/models/POEM.py:363-466, release loss types: joints l2,
 * vertices l1, parameters l2) on the LAST block's prediction coords_last [B, 799, 3], with d loss_recon / d coords_last:
 * losses[8] = {loss_3d_joints_from_mesh, loss_3d_joints, loss_3d_verts, loss_2d_joints, loss_2d_verts, loss_pose,
 * loss_shape, loss_recon}.  j_regressor [16, 778] (MANO), cameras / target_joints_2d per image (cam_extr: camera -> master,
 * inverted inside as the reference does), img_sample [NV] = sample of each image, img_scale = sqrt(W^2 + H^2).
 * pred_pose / gt_pose [B, 48] and pred_shape / gt_shape [B, 10] (GT of each sample's first view) or NULL.
 * The heat-map term of the reference's total loss belongs to the image half and is not included. */
int poem_tr_compute_loss(const float* coords_last, const float* gt_joints, const float* gt_verts, const float* j_regressor,
                         const float* cam_intr, const float* cam_extr, const int32_t* img_sample, const float* target_joints_2d,
                         int B, int NV, float img_scale, float w_joints, float w_verts, float w_joints_2d, float w_verts_2d,
                         const float* pred_pose, const float* gt_pose, const float* pred_shape, const float* gt_shape,
                         float w_pose, float w_shape, float* losses, float* dcoords_last, float* dpose, float* dshape,
                         void* stream);

/* Parametric (medium_MANO) tail of the last block, pt_metro_transformer.py:139-151 (forward as in poem_parametric_tail of
 * poem_b200.h, on raw fp32 parameter pointers) and its backward.  MANO constants with the blend axis first: v_template
 * [778*3], shapedirs [10][778*3], posedirs [135][778*3], j_regressor [16][778], skin_weights [778][16].
 * forward: feats [B, Q, D] (re-interpreted as (B*D, Q) rows like the reference) -> flat [B*D] (kept for the backward),
 * coords [B, Q, 3] (metric, root-centred on center_idx, + the sample's hand centre ref_joints[:, 9]), pose [B, 48], shape [B, 10].
 * backward: dcoords (+ dpose / dshape or NULL) -> dfeats [B, Q, D] written, dflat [B*D] scratch; dflat_w [Q], dflat_b [1],
 * dlin_w [106, D], dlin_b [106] += . */
int poem_tr_mano_tail(const float* feats, const float* flat_w, const float* flat_b, const float* lin_w, const float* lin_b,
                      const float* v_template, const float* shapedirs, const float* posedirs, const float* j_regressor,
                      const float* skin_weights, const float* ref_joints, int center_idx, int B, int Q, int D, float* flat,
                      float* coords, float* pose, float* shape, void* stream);
int poem_tr_mano_tail_bwd(const float* feats, const float* flat_w, const float* lin_w, const float* lin_b,
                          const float* v_template, const float* shapedirs, const float* posedirs, const float* j_regressor,
                          const float* skin_weights, int center_idx, int B, int Q, int D, const float* flat,
                          const float* dcoords, const float* dpose, const float* dshape, float* dflat, float* dfeats,
                          float* dflat_w, float* dflat_b, float* dlin_w, float* dlin_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif
