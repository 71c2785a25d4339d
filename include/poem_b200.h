/* poem_b200.h — C ABI of the B200-native POEM-v2 point-embedded transformer decoder.
 *
 * The reference has no native code and no FFI: its plug-in boundary for this path is a Python class
 * resolved through a registry (`@HEAD.register_module() class POEM_Generalized_Head`,
 * /root/reference/lib/models/heads/ptEmb_head.py:683-684; `@TRANSFORMER.register_module() class PtEmbedTRv4`,
 * lib/models/layers/ptEmb_transformer.py:303-304; resolved by lib/utils/builder.py:9-47).  The entry points
 * below are what a reference-side `nn.Module` binds through ctypes; each cites the
 * reference call it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative POEM_E_* code; poem_last_error() gives a message
 *   - all device buffers are caller-allocated; the library allocates nothing and keeps no mutable global
 *     state besides the last-error string (thread local)
 *   - launches are asynchronous on `stream` (a cudaStream_t passed as void*)
 *   - "op16" buffers hold the library's 16-bit tensor-core operand format: IEEE fp16 (binary16; uint16_t storage).  Every
 *     fp32 -> op16 conversion inside the library saturates (no inf / NaN from a finite value); weights handed to the
 *     library must already be finite fp16 (the Python packer raises when |w| > 65504)
 */
#ifndef POEM_B200_H
#define POEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POEM_ABI_VERSION 2

#define POEM_OK 0
#define POEM_E_BADDIM (-1)      /* unsupported / inconsistent dimensions */
#define POEM_E_NULL (-2)        /* required pointer is NULL */
#define POEM_E_WORKSPACE (-3)   /* workspace too small */
#define POEM_E_CUDA (-4)        /* CUDA runtime / driver error */
#define POEM_E_ALIGN (-5)       /* pointer or leading dimension not 16-byte aligned */

typedef uint16_t poem_op16;

/* Dimensions read from cfg.MODEL.HEAD (ptEmb_head.py:57-76,686-695; ptEmb_transformer.py:312-324). */
typedef struct PoemDims {
  int32_t embed_dims;   /* D: 128 | 256 | 512 (POEM-small / -medium / -large); 1024 (POEM-huge, head dim 256) is rejected:
                         * poem_workspace_bytes returns 0 and every entry point POEM_E_BADDIM */
  int32_t in_channels;  /* C: 160 */
  int32_t n_sample;     /* P: 4096 */
  int32_t n_query;      /* Q: 799 */
  int32_t n_blocks;     /* NB: 3 */
  int32_t n_heads;      /* h: 4 */
  int32_t n_neighbor;   /* K: 32 */
  int32_t feat_h;       /* 16 */
  int32_t feat_w;       /* 16 */
  int32_t center_idx;   /* 9 */
  float radius;         /* 0.1 */
  int32_t max_views;    /* rows of the positional table cover view counts 1..max_views */
  int32_t run_last_ffn; /* 1 if the last block's FFN output is needed (parametric tail) */
} PoemDims;

/* One Linear layer in kernel layout: weight op16 [out, in] row-major (K-major), bias fp32 [out] or NULL. */
typedef struct PoemLinear {
  const poem_op16* w;
  const float* b;
} PoemLinear;

/* Vector-attention (Point-Transformer) layer, reference lib/models/bricks/point_transformers.py:47-156.
 * fc_gamma.0 is linear, so it is distributed over (q_i - k_j + pos_ij) at pack time:
 *   fc_gamma.0(q_i - k_j + pos_ij) = gamma1_delta2 · h_ij + qt_i - kt_j,   h_ij = relu(fc_delta.0(xyz_i - nbr_j))
 *   gamma1_delta2 = W_g1 · W_d2;   qt_i = W_g1 q_i + W_g1 b_d2 + b_g1;   kt_j = W_g1 k_j
 * qt / kt are produced by the query / key projections (their weights are pre-multiplied by W_g1). */
typedef struct PoemVecAttn {
  const float* wd1;          /* fc_delta.0 weight fp32 [D,3] */
  const float* bd1;          /* fc_delta.0 bias   fp32 [D]   */
  PoemLinear delta2;         /* fc_delta.2 */
  PoemLinear gamma1_delta2;  /* W_g1 · W_d2 (bias unused: folded into qt) */
  PoemLinear gamma2;         /* fc_gamma.2 */
  PoemLinear fc2;            /* fc2 */
} PoemVecAttn;

/* One point_METRO_block (pt_metro_transformer.py:94-200), weights folded at pack time:
 *   pt_proj : [6D, D]  rows = K1 | K2 | kt_cross | v_cross | V1 | V2   each composed with `embedding`
 *             (and with query_cross_attn.fc1 for kt, v; kt additionally with fc_gamma.0); bias folded likewise
 *   self_qkv: [3D, D]  rows = qt | kt | v of query_self_attn (w_qs·fc1, w_ks·fc1 pre-multiplied by fc_gamma.0)
 *   cross_q : [D, D]   qt of query_cross_attn (fc_gamma.0 · w_qs, bias W_g1 b_d2 + b_g1)                  */
typedef struct PoemBlock {
  PoemLinear embedding;     /* embedding (applied to the query stream) */
  PoemLinear pt_proj;
  PoemLinear q1, o1;        /* encoder.attn.self.query, encoder.attn.output.dense */
  const float *ln1_g, *ln1_b;
  PoemLinear q2, o2;        /* encoder.cross_attn.* */
  const float *ln2_g, *ln2_b;
  PoemLinear self_qkv;
  PoemVecAttn self_attn;
  PoemLinear cross_q;       /* qt projection of query_cross_attn */
  PoemVecAttn cross_attn;
  PoemLinear reg1;          /* reg_branch.0 */
  const float* reg2_w;      /* reg_branch.2 weight fp32 [3,D] */
  const float* reg2_b;      /* fp32 [3] */
  PoemLinear ffn1, ffn2;    /* encoder.intermediate.dense, encoder.output.dense */
  const float *ln3_g, *ln3_b;
} PoemBlock;

#define POEM_MAX_BLOCKS 8

typedef struct PoemWeights {
  PoemLinear input_proj;        /* [D, C] (1x1 conv, ptEmb_head.py:94) */
  const float* pos_table;       /* fp32 [sum_{N=1..max_views} N, feat_h*feat_w, D]:
                                   adapt_pos3d(sine3d(N))[n] + adapt_pos3d.bias, rows ordered N=1:(n=0), N=2:(n=0,1), ... */
  PoemLinear merge0a, merge0b;  /* merge_net_feature.0.{0,2} */
  PoemLinear merge1a, merge1b;  /* merge_net_feature.1.{0,2} */
  const float* query_embed;     /* query_feat_embedding.weight fp32 [Q, D] */
  const float* bps;             /* fp32 [P,3]  (assets/bps.npy) */
  const float* anchor_xyz;      /* fp32 [32,3] (assets/anchor.npy) */
  const int32_t* anchor_idx;    /* int32 [32]  (assets/anchor_idx.npy) */
  const float* template_xyz;    /* fp32 [Q,3]  MANO zero-pose template, joints then vertices */
  const int32_t* bps_perm;      /* optional int32 [P]: BPS indices in k-d chunk order of bps / radius (32-NN pruning) */
  const float* bps_chunk_box;   /* optional fp32 [P/32, 6]: (min xyz, max xyz) of each 32-point chunk of that order,
                                   in normalised units, grown by 1e-4; NULL -> brute-force 32-NN */
  PoemBlock blocks[POEM_MAX_BLOCKS];
} PoemWeights;

/* Per-call inputs of POEM_Generalized_Head.forward (ptEmb_head.py:825). */
typedef struct PoemInputs {
  int32_t batch;                /* B */
  int32_t n_images;             /* sum of views */
  const int32_t* view_counts;   /* HOST pointer, int32 [B]  (img_metas["cam_view_num"]) */
  const float* mlvl_feat;       /* fp32 [n_images, C, feat_h, feat_w] */
  const float* cam_intr;        /* fp32 [n_images, 3, 3] */
  const float* cam_extr;        /* fp32 [n_images, 4, 4]  camera -> master */
  const float* reference_joints;/* fp32 [B, 21, 3] metres, master frame */
  float inp_img_w, inp_img_h;   /* img_metas["inp_img_shape"] */
} PoemInputs;

int poem_abi_version(void);
const char* poem_last_error(void);

/* Instrumentation (bench.py): number of kernels launched by this library so far; optional per-launch CUDA-event
 * timing of every kernel on its launching stream, summarised as a JSON object keyed by "<kernel>:<stage>". */
long long poem_kernel_launches(void);
void poem_profile_enable(int on);
size_t poem_profile_summary(char* buf, size_t cap);
/* Test hook: route poem_vector_attention through the un-fused composition (token tensors in HBM) so the fused
 * kernel can be checked against it on the device.  Not used by the product path. */
void poem_debug_force_unfused(int on);
/* Test hook: 0 = 3x3 stride-1 C->C convolutions use the halo-reuse kernel on the live channels (default); 1 = halo
 * kernel on all padded channels; 2 = every convolution takes the generic implicit-GEMM path (so the variants can be
 * compared on the device). */
void poem_debug_conv_mode(int mode);
/* Test hook: while `device_buf` is non-NULL, every whole-path / transformer call of this host thread copies the
 * 32-NN index sets it used in blocks 1..NB-1 into it (device int32, layout [block - 1][0 = self, 1 = cross][B*Q*32];
 * a call that needs more than `capacity` int32 fails with POEM_E_WORKSPACE).  32-NN selection is discontinuous in the
 * regressed coordinates, so the parity tests compare against the oracle run on the SAME sets, and check the sets
 * separately (bit-exact given the coordinates).  NULL disables.  Not used by the product path. */
void poem_debug_export_neighbours(int32_t* device_buf, size_t capacity);
/* Test hook: while non-NULL, every whole-path call of this host thread copies the merged BPS features (stage boundary a6:
 * `pt_feats` of ptEmb_head.py:926, op16 [B*P, D]) into the buffer (capacity in elements).  NULL disables. */
void poem_debug_export_pt_feats(poem_op16* device_buf, size_t capacity);

/* Bytes of device workspace poem_head_forward needs for (batch, n_images). */
size_t poem_workspace_bytes(const PoemDims* dims, int batch, int n_images);

/* Whole decoder path, device buffers.  Replaces POEM_Generalized_Head.forward + PtEmbedTRv4.forward
 * (ptEmb_head.py:825-964, ptEmb_transformer.py:371-376).
 *   out_coords : fp32 [NB, B, Q, 3] metres (`all_coords_preds`)
 *   out_feats  : optional fp32 [B, Q, D] output of the last block's FFN (needs dims->run_last_ffn) or NULL */
int poem_head_forward(const PoemDims* dims, const PoemWeights* w, const PoemInputs* in, float* out_coords,
                      float* out_feats, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Parametric output (config/release/train_medium_MANO.yaml: TRANSFORMER.PARAMETRIC_OUTPUT), SURVEY §8a row a16.
 * Replaces point_METRO_block.get_parametric_output (pt_metro_transformer.py:139-151) of the LAST block, rot6d_to_aa
 * (utils/transform.py:448-466) and the manotorch ManoLayer forward it calls (axis-angle, no PCA, flat hand mean,
 * centred on joint dims->center_idx).  All tensors fp32 on the device; the MANO model parameters are the layer's
 * `th_*` buffers re-laid so that the blend loops read consecutive vertices:
 *   shapedirs [10, 778*3]  (= th_shapedirs (778,3,10) with the coefficient axis first),
 *   posedirs  [135, 778*3] (= th_posedirs (778,3,135) likewise). */
typedef struct PoemManoTail {
  const float* flat_w;        /* flat_verts.weight [Q] */
  const float* flat_b;        /* flat_verts.bias   [1] */
  const float* lin_w;         /* mano_linear.weight [106, D] */
  const float* lin_b;         /* mano_linear.bias   [106] */
  const float* v_template;    /* [778, 3] */
  const float* shapedirs;     /* [10, 778*3] */
  const float* posedirs;      /* [135, 778*3] */
  const float* j_regressor;   /* [16, 778] */
  const float* skin_weights;  /* [778, 16] */
} PoemManoTail;

/* Stage-level: query_feats [B,Q,D] (output of the last block, re-interpreted as (B*D, Q) rows like the reference) ->
 *   coords     [B,Q,3]: 21 MANO joints then 778 vertices, centred on joint center_idx, nan_to_num'd, plus
 *                       reference_joints[b, center_idx] when `reference_joints` ([B,21,3]) is not NULL;
 *   pred_pose  [B,48] axis-angle, pred_shape [B,10].
 * workspace: poem_parametric_tail_workspace_bytes(dims, batch) bytes. */
size_t poem_parametric_tail_workspace_bytes(const PoemDims* dims, int batch);
int poem_parametric_tail(const PoemDims* dims, const PoemManoTail* mano, int batch, const float* query_feats,
                         const float* reference_joints, float* coords, float* pred_pose, float* pred_shape,
                         void* workspace, size_t workspace_bytes, void* stream);

/* poem_head_forward for a PARAMETRIC_OUTPUT head (ptEmb_head.py:950-963): blocks 0..NB-2 give xyz*radius + centre, the
 * last block's joints/vertices come from the MANO tail (+ centre, not scaled).  Same workspace as poem_head_forward.
 * out_coords [NB,B,Q,3]; pred_pose [B,48] (`pred_pose` reshaped (B,16,3)); pred_shape [B,10]. */
int poem_head_forward_parametric(const PoemDims* dims, const PoemWeights* w, const PoemManoTail* mano,
                                 const PoemInputs* in, float* out_coords, float* pred_pose, float* pred_shape,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* The same with HOST buffers (see poem_head_forward_host): host_pose [B,48], host_shape [B,10] are copied back with
 * the coordinates. */
int poem_head_forward_parametric_host(const PoemDims* dims, const PoemWeights* w, const PoemManoTail* mano,
                                      const PoemInputs* host_in, float* host_out, float* host_pose, float* host_shape,
                                      void* staging, size_t staging_bytes, void* workspace, size_t workspace_bytes,
                                      void* stream);

/* Decoder blocks only.  Replaces PtEmbedTRv4.forward(query_xyz, query_feat, pt_xyz, pt_feats)
 * (ptEmb_transformer.py:371-376): all inputs fp32 device buffers in normalised (radius) units,
 *   query_xyz [B,Q,3], query_feat [B,Q,D], pt_xyz [B,P,3], pt_feats [B,P,D];
 *   out_xyz fp32 [NB,B,Q,3] (normalised, stacked per block); out_feats optional fp32 [B,Q,D]. */
size_t poem_transformer_workspace_bytes(const PoemDims* dims, int batch);
int poem_transformer_forward(const PoemDims* dims, const PoemWeights* w, int batch, const float* query_xyz,
                             const float* query_feat, const float* pt_xyz, const float* pt_feats, float* out_xyz,
                             float* out_feats, void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST input/output buffers (pinned memory recommended): copies the inputs into `staging`, runs,
 * copies `all_coords_preds` back.  Weights, staging and workspace stay on the device.  `staging` (poem_staging_bytes(),
 * 1024-byte aligned) is private to this entry point: it has two slots and the inputs travel on an internal copy stream,
 * so the host->device transfer of call i + 1 overlaps the kernels of call i; nothing else may use it while calls are
 * in flight.  `workspace` (poem_workspace_bytes()) is only touched on `stream` and may be shared with other calls on
 * that stream.  The host buffers of a call must stay valid until `stream` has reached the end of that call. */
size_t poem_staging_bytes(const PoemDims* dims, int batch, int n_images);
int poem_head_forward_host(const PoemDims* dims, const PoemWeights* w, const PoemInputs* host_in, float* host_out,
                           void* staging, size_t staging_bytes, void* workspace, size_t workspace_bytes, void* stream);

/* ---- HRNet-W40 stage 4 (reference lib/models/backbones/hrnet.py:272-277: 3 HighResolutionModules of 4 branches x
 * 4 BasicBlocks + all-to-all fuse layers).  Convolution weights have eval-mode BatchNorm folded in and are stored as
 * op16 [Cout_p, k*k*Cin_p] (K ordered (ky, kx, c); channel counts padded to multiples of 64 with zeros), bias fp32
 * [Cout_p].  Activations are converted NCHW fp32 <-> NHWC op16 inside the call. */
#define POEM_HR_MAX_MODULES 4
typedef struct PoemHRModule {
  PoemLinear branch[4][4][2];   /* [branch][block][conv1 | conv2] */
  PoemLinear fuse[4][4][3];     /* [i][j][k]: j > i: k = 0 is the 1x1 conv; j < i: k = 0..i-j-1 stride-2 3x3 chain */
} PoemHRModule;
typedef struct PoemHRStage4 {
  int32_t n_modules;            /* 3 */
  int32_t channels[4];          /* 40, 80, 160, 320 */
  PoemHRModule modules[POEM_HR_MAX_MODULES];
} PoemHRStage4;
size_t poem_hrnet_stage4_workspace_bytes(const PoemHRStage4* w, int n_images, int base_res);
/* in[b] / out[b]: fp32 NCHW (n_images, channels[b], base_res >> b, base_res >> b), device pointers.
 * Replaces `y_list = self.stage4(x_list)` (hrnet.py:417). */
int poem_hrnet_stage4_forward(const PoemHRStage4* w, int n_images, int base_res, const float* const* in,
                              float* const* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- whole HRNet-W40 backbone (reference lib/models/backbones/hrnet.py:240-420): stem (two
 * stride-2 3x3 convs), layer1 (4 Bottlenecks), transitions 1-3, stage 2 (1 module, 2 branches), stage 3 (4 modules,
 * 3 branches), stage 4 (3 modules, 4 branches).  Same weight conventions as stage 4; the first convolution (3 input
 * channels) keeps fp32 weights [64, 27] with k = (ky*3 + kx)*3 + c. */
typedef struct PoemBottleneck {
  PoemLinear c1, c2, c3;   /* 1x1 in->64, 3x3 64->64, 1x1 64->256 (BN folded) */
  PoemLinear ds;           /* 1x1 in->256 downsample of the identity path, w == NULL when absent */
} PoemBottleneck;
typedef struct PoemHRNet {
  const float* stem1_w;
  const float* stem1_b;
  PoemLinear stem2;
  PoemBottleneck layer1[4];
  PoemLinear trans1[2];    /* 3x3 256->40 ; 3x3 stride-2 256->80 */
  PoemLinear trans2;       /* 3x3 stride-2 80->160 */
  PoemLinear trans3;       /* 3x3 stride-2 160->320 */
  int32_t channels[4];
  PoemHRModule stage2[1];
  PoemHRModule stage3[4];
  PoemHRModule stage4[3];
} PoemHRNet;
size_t poem_hrnet_workspace_bytes(const PoemHRNet* w, int n_images, int img_res);
/* images: fp32 NCHW (n_images, 3, 256, 256); out[b]: fp32 NCHW (n_images, channels[b], 64 >> b, 64 >> b).
 * Replaces `HighResolutionNet.forward` (hrnet.py:385-420) as called at lib/models/POEM.py:262. */
int poem_hrnet_forward(const PoemHRNet* w, int n_images, int img_res, const float* images, float* const* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- images -> mlvl_feat: backbone + `feat_decode` (reference lib/models/POEM.py:189-203, 255-265, HRNet branch):
 * x = f0; x = ConvBlock_i(x) + f_{i+1} for the three stride-2 ConvBlocks (3x3 conv with bias + BN + ReLU, folded);
 * bilinear x2 upsampling of the 8x8 map; 1x1 convolution 320 -> out_channels (bias, no norm). */
typedef struct PoemFeatDecode {
  PoemLinear delayer[3];   /* 3x3 stride-2: 40->80, 80->160, 160->320 (padded to 64; BN and conv bias folded) */
  PoemLinear feat_in;      /* 1x1 320 -> out_channels (padded to 64) */
  int32_t out_channels;    /* 160 */
} PoemFeatDecode;
/* `uv_decode` + `heatmap_stage` (POEM.py:205-229, HRNet branch): x = f3; x = ConvBlock_i(cat(bilinear x2(x), f_{2-i}))
 * for i = 0..2 (3x3 conv + bias + BN + ReLU: 480->160, 240->80, 120->40); 2x2 max-pool; 1x1 conv 40 -> n_joints +
 * sigmoid; pdf normalisation and integral soft-argmax (lib/models/integal_pose.py:196-220) scaled to pixels. */
typedef struct PoemUVDecode {
  PoemLinear delayer[3];   /* input channels ordered (upsampled, skip) as torch.cat does, then zero-padded to 64 */
  const float* out_w;      /* fp32 [n_joints, 40] */
  const float* out_b;      /* fp32 [n_joints] */
  int32_t n_joints;        /* 21 */
} PoemUVDecode;
size_t poem_image_features_workspace_bytes(const PoemHRNet* w, const PoemFeatDecode* fd, const PoemUVDecode* uv,
                                           int n_images, int img_res);
/* images fp32 NCHW (n_images, 3, 256, 256) -> mlvl_feat fp32 NCHW (n_images, out_channels, 16, 16), the tensor
 * poem_head_forward takes.  uv != NULL additionally runs the heatmap branch: uv_px fp32 (n_images, n_joints, 2) in
 * pixels (`pred_joints_uv`, POEM.py:331) and, if heatmap != NULL, the sigmoid maps fp32 (n_images, n_joints, 32, 32).
 * maps: optional four fp32 NCHW backbone outputs (NULL to skip the export). */
int poem_image_features(const PoemHRNet* w, const PoemFeatDecode* fd, const PoemUVDecode* uv, int n_images, int img_res,
                        const float* images, float* mlvl_feat, float* uv_px, float* heatmap, float* const* maps,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Batched DLT triangulation (lib/utils/triangulation.py:5-45 in the per-sample loop of POEM.py:284-299): uv_px
 * (n_images, n_joints, 2), cam_intr (n_images, 3, 3), cam_extr (n_images, 4, 4) camera-to-master, view_counts int32
 * [batch] -> ref_joints fp32 (batch, n_joints, 3).  All device pointers. */
int poem_triangulate_dlt(const float* uv_px, const float* cam_intr, const float* cam_extr, const int32_t* view_counts,
                         int batch, int n_joints, float* ref_joints, void* stream);

/* Evaluation metrics on the device (reference lib/metrics/pa_eval.py:41-124 `PAEval.feed` /
 * `align_w_scale`, lib/metrics/mean_epe.py:23-33): per sample the Procrustes-aligned and the raw mean point distance.
 * gt, pred: fp32 (batch, n_points, 3); out: fp32 (batch, 2) = [aligned, raw]; aligned: optional fp32 (batch, n_points, 3)
 * aligned prediction.  Replaces the per-sample host loop around scipy.linalg.orthogonal_procrustes. */
int poem_pa_metrics(const float* gt, const float* pred, int batch, int n_points, float* out, float* aligned,
                    void* stream);

/* One convolution of the stage (building block of the call above): NHWC op16 in/out, channels padded to 64,
 * w op16 [Cout_p, k*k*Cin_p], b fp32 [Cout_p], ksize 1|3 (padding k/2), stride 1|2, optional ReLU and NHWC residual.
 * c_live_in / c_live_out > 0 promise that only the first c_live_in input / c_live_out output channels are non-zero
 * (the rest of the padded tensors, weights and bias is zero padding), which lets the 3x3 stride-1 kernel skip the
 * padding; 0 = no promise.
 * Replaces nn.Conv2d + nn.BatchNorm2d(eval) (+ReLU, + identity) of hrnet.py:38-67,177-207. */
int poem_conv_nhwc(const poem_op16* in, int n_images, int H, int W, int Cin_p, const poem_op16* w, const float* b,
                   int Cout_p, int ksize, int stride, int relu, const poem_op16* res, poem_op16* out, int c_live_in,
                   int c_live_out, void* stream);

/* ---- stage-level entry points (unit-testable building blocks; same kernels the whole path uses) ---- */

/* C = act(A·W^T + bias) (+ residual); A op16 [M,K] (lda), W op16 [N,K] (ldw); outputs optional.
 * act: 0 none, 1 relu, 2 gelu(erf).  Replaces nn.Linear / 1x1 nn.Conv2d call sites (cuBLAS/cuDNN). */
int poem_linear(const poem_op16* A, int lda, const poem_op16* W, int ldw, const float* bias, int M, int N, int K,
                int act, const float* residual, int ld_res, float* out_f32, int ld_f32, poem_op16* out_op16,
                int ld_op16, void* stream);

/* softmax(Q K^T / sqrt(hd)) V per head, no mask.  Q op16 [B*Lq, ldq], K op16 [B*Lk, ldk], V op16 [B*Lk, ldv] (all
 * row-major, head h in columns [h*hd, (h+1)*hd)), ctx op16 [B*Lq, ld_ctx].  Lk % 128 == 0.
 * Replaces HF BertSelfAttention's matmul-softmax-matmul (pt_metro_transformer.py:57-72). */
int poem_mha(const poem_op16* Q, int ldq, const poem_op16* K, int ldk, const poem_op16* V, int ldv, poem_op16* ctx,
             int ld_ctx, int B, int Lq, int Lk, int D, int n_heads, void* stream);

/* idx int32 [B, Lq, 32]: 32 nearest reference points, ascending squared-L2, lower index wins ties.
 * Replaces pytorch3d.ops.knn_points(K=32) (point_transformers.py:83,134). */
int poem_knn32(const float* query_xyz, const float* ref_xyz, int32_t* idx, int B, int Lq, int Lr, void* stream);

/* Same result as poem_knn32 for a reference set that is (up to rounding) the fixed BPS: perm / boxes as in PoemWeights;
 * ref_xyz_sorted[b][k] = ref_xyz[b][perm[k]] (chunk order). */
int poem_knn32_bps(const float* query_xyz, const float* ref_xyz_sorted, const int32_t* perm, const float* boxes,
                   int32_t* idx, int B, int Lq, int Lr, void* stream);

/* Camera projection + bilinear sampling in the reference's reinterpreted (token, view, channel) row order.
 * xmap fp32 [n_images, D, fh*fw]; X op16 [sum_views*P, D].  Replaces collation.py:48-65 + F.grid_sample +
 * the raw .view regroup (ptEmb_head.py:874-915). */
int poem_project_sample(const float* xmap, const float* cam_intr, const float* cam_extr, const float* bps,
                        const float* centre, const int32_t* host_view_counts, int B, int n_images, int D, int P,
                        int fh, int fw, float img_w, float img_h, poem_op16* X, void* workspace,
                        size_t workspace_bytes, void* stream);

/* The gather side of rows a4 / a5 on its own: for every (image, BPS point) the four bilinear taps of
 * F.grid_sample(align_corners=False, zero padding) after generate_grid_sample_proj (collation.py:48-65): pixel index
 * y * fw + x of the taps nw, ne, sw, se (0 with weight 0 when the tap lies outside the map) and their weights.
 * tap_pixels int32 [n_images, P, 4], tap_weights fp32 [n_images, P, 4]; workspace >= 128 KB + 96 bytes per (image, point)/4.
 * This is the table the fused sampler / merge kernel gathers with. */
int poem_sample_taps(const float* cam_intr, const float* cam_extr, const float* bps, const float* centre,
                     const int32_t* host_view_counts, int B, int n_images, int P, int fh, int fw, float img_w, float img_h,
                     int32_t* tap_pixels, float* tap_weights, void* workspace, size_t workspace_bytes, void* stream);

/* Vector attention core: res[b,i,:] = sum_j softmax_j(gamma(q_i - k_j + pos_ij)/sqrt(D)) * (v_j + pos_ij),
 * pos_ij = delta(xyz_i - nbr_xyz_j), in the folded form documented at PoemVecAttn:
 * q = qt op16 [B*Lq, ldq]; ktab = kt, vtab = v op16 [B*Lr, ldk/ldv]; idx int32 [B*Lq*32]
 * (or NULL with anchors: anchor_idx int32[32], anchor_xyz fp32[32,3]); res op16 [B*Lq, D].
 * Replaces point_transformers.py:86-94 / 139-150. */
int poem_vector_attention(const PoemVecAttn* w, const poem_op16* q, int ldq, const poem_op16* ktab, int ldk,
                          const poem_op16* vtab, int ldv, const float* q_xyz, const float* ref_xyz,
                          const int32_t* idx, const int32_t* anchor_idx, const float* anchor_xyz, int B, int Lq,
                          int Lr, int D, poem_op16* res, void* workspace, size_t workspace_bytes, void* stream);
size_t poem_vector_attention_workspace_bytes(int B, int Lq, int D);

/* y = LayerNorm(x) over the last dim, eps 1e-12 (HF BertSelfOutput/BertOutput). */
int poem_layernorm(const float* x, const float* gamma, const float* beta, float* y_f32, poem_op16* y_op16, int rows,
                   int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POEM_B200_H */
