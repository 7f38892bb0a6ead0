/*
 * airpose_b200 -- C ABI of the B200-native AirPose hot path (copenet_twoview forward).
 *
 * The reference has no FFI: its boundary is the Python object protocol of
 * `model_copenet.getcopenet()` / `smplx.SMPLX`.  This header is
 * what the Python shim in airpose_b200/ binds with ctypes; every entry point names the
 * reference call it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is available
 *     from airpose_last_error() (thread-local).  Nothing throws.
 *   - all tensor arguments are DEVICE pointers owned by the caller unless the name ends
 *     in `_host`; the library allocates nothing persistent except the opaque handles
 *     (which own packed weights and activation workspaces).
 *   - every launch is asynchronous on the `stream` argument (a cudaStream_t passed as
 *     void*); one handle per GPU / rank; handles are thread-compatible, not thread-safe.
 *   - fp32 tensors are contiguous row-major exactly as the reference's torch tensors.
 */
#ifndef AIRPOSE_B200_H_
#define AIRPOSE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIRPOSE_B200_ABI_VERSION 1

const char* airpose_last_error(void);
int airpose_abi_version(void);

/* ------------------------------------------------------------------------------------
 * SMPL-X body model  (copenet/src/copenet/smplx/smplx/body_models.py:648-994, lbs.py)
 * ---------------------------------------------------------------------------------- */
typedef struct airpose_smplx airpose_smplx_t;

/* Host-side description of the buffers SMPL.__init__/SMPLX.__init__ register
 * (body_models.py:205-296,727-730).  All pointers are HOST pointers, read during create. */
typedef struct {
  int32_t num_verts;        /* V = 10475 */
  int32_t num_joints;       /* J = 55 */
  int32_t num_shape;        /* shapedirs columns (20: 10 betas + 10 expression) */
  int32_t num_pose_basis;   /* (J-1)*9 = 486 */
  int32_t num_faces;
  int32_t num_landmarks;    /* 51 static face landmarks */
  int32_t num_extra;        /* 21 vertex-picked joints (vertex_joint_selector.py:38-68) */
  const float*   v_template;      /* [V,3] */
  const float*   shapedirs;       /* [V,3,num_shape] */
  const float*   posedirs;        /* [num_pose_basis, V*3]  (already transposed, body_models.py:284-288) */
  const float*   J_regressor;     /* [J,V] dense */
  const int64_t* parents;         /* [J], parents[0] = -1 */
  const float*   lbs_weights;     /* [V,J] */
  const int64_t* faces;           /* [F,3] */
  const int64_t* lmk_faces_idx;   /* [L] */
  const float*   lmk_bary_coords; /* [L,3] */
  const int64_t* extra_joint_idx; /* [E] */
} airpose_smplx_model_host;

int airpose_smplx_create(airpose_smplx_t** out, const airpose_smplx_model_host* model, int device);
int airpose_smplx_destroy(airpose_smplx_t* h);
/* max non-zeros per row found in lbs_weights (the skinning kernel iterates exactly these) */
int airpose_smplx_skin_nnz(const airpose_smplx_t* h);

/* Arguments of one SMPL-X forward.  Replaces, in one call:
 *   SMPLX.forward(pose2rot=False)      body_models.py:820-994  (lbs.py:135-222, :96-132, :316-370)
 *   VertexJointSelector.forward        vertex_joint_selector.py:73-77
 *   transform_smpl                     copenet/src/copenet/utils/utils.py:237-256   (if root_R/root_t given)
 *   perspective_projection             copenet/src/copenet/utils/geometry.py:63-91  (if out_j2d given)
 * Rotation inputs are row-major 3x3 matrices.  A NULL rotation segment means identity
 * (what the reference obtains from its zero Parameters via batch_rodrigues, :878-925).
 * `*_stride` are in floats between consecutive meshes (so a [B,55,3,3] full_pose can be
 * passed as three segments of one buffer with stride 495). */
typedef struct {
  int32_t batch;
  int32_t num_betas;             /* columns of `betas` actually supplied (10, or 20 with expression) */
  const float* betas;            /* [B,num_betas] */
  int32_t betas_stride;
  const float* global_orient;    /* joint 0        [B,1,3,3] or NULL */
  int32_t global_orient_stride;
  const float* body_pose;        /* joints 1..21   [B,21,3,3] or NULL */
  int32_t body_pose_stride;
  const float* tail_pose;        /* joints 22..54  [B,33,3,3] (jaw, eyes, hands) or NULL */
  int32_t tail_pose_stride;
  const float* transl;           /* [B,3] or NULL  (body_models.py:980-982) */
  const float* root_R;           /* [B,3,3] or NULL: camera-frame rotation about the origin */
  int32_t root_R_stride;
  const float* root_t;           /* [B,3] or NULL */
  int32_t root_t_stride;
  float focal_x, focal_y;        /* constants.py:7 */
  const float* center;           /* [B,2] principal point, or NULL */
  int32_t center_stride;
  float* out_vertices;           /* [B,V,3]   ModelOutput.vertices */
  float* out_joints;             /* [B,127,3] ModelOutput.joints (55 + 21 + 51) */
  float* out_vertices_cam;       /* [B,V,3]   or NULL */
  float* out_joints_cam;         /* [B,127,3] or NULL */
  float* out_joints_2d;          /* [B,127,2] or NULL */
  const float* proj_t;           /* [B,3] or NULL: `translation` of perspective_projection (geometry.py:63-91), added
                                    to the camera-frame joints for the projection only (hmr.py:149-158) */
  int32_t proj_t_stride;
} airpose_smplx_fwd_args;

int airpose_smplx_fwd(airpose_smplx_t* h, const airpose_smplx_fwd_args* args, void* stream);

/* Backward of airpose_smplx_fwd for the training step and for bundle adjustment: what torch autograd derives for
 * SMPLX.forward(pose2rot=False) + transform_smpl + perspective_projection (copenet_twoview.py:281-317;
 * copenet_real_data/scripts/bundle_adj.py:301-401).  Call pattern of the hot path: betas, 21 body rotations, optional
 * global orientation (NULL = identity), joints 22..54 identity, no translation.  Upstream gradients (any subset):
 * d/d vertices, d/d joints (canonical), d/d joints_cam and d/d joints_2d (these two need the forward's `joints` output
 * and the same root_R / root_t / focal).  All outputs are fp32; the rotation gradients are with respect to the nine
 * matrix entries, as autograd gives them.  Deterministic apart from the shared-memory accumulation of dL/dA inside a
 * 128-vertex tile (fp32 atomics, ~1e-7 relative). */
typedef struct {
  int32_t batch;
  int32_t num_betas;
  const float* betas;          int32_t betas_stride;          /* [B,num_betas] */
  const float* global_orient;  int32_t global_orient_stride;  /* [B,1,3,3] or NULL */
  const float* body_pose;      int32_t body_pose_stride;      /* [B,21,3,3] or NULL */
  const float* joints;                                        /* [B,127,3] forward output, or NULL */
  const float* root_R;         int32_t root_R_stride;         /* [B,3,3] or NULL */
  const float* root_t;         int32_t root_t_stride;         /* [B,3] or NULL */
  float focal_x, focal_y;
  const float* grad_vertices;                                 /* [B,V,3] or NULL */
  const float* grad_joints;                                   /* [B,127,3] or NULL */
  const float* grad_joints_cam;                               /* [B,127,3] or NULL */
  const float* grad_joints_2d;                                /* [B,127,2] or NULL */
  float* grad_betas;                                          /* [B,num_betas] */
  float* grad_body_pose;                                      /* [B,21,3,3] or NULL */
  float* grad_global_orient;                                  /* [B,3,3] or NULL */
  float* grad_root_R;                                         /* [B,3,3] or NULL */
  float* grad_root_t;                                         /* [B,3] or NULL */
} airpose_smplx_bwd_args;
int airpose_smplx_bwd(airpose_smplx_t* h, const airpose_smplx_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------
 * geometry helpers
 * ---------------------------------------------------------------------------------- */
/* rot6d_to_rotmat (copenet/src/copenet/utils/geometry.py:47-61): x [n,6] -> R [n,3,3].
 * `x_row_stride` floats between rows of 6 (lets callers pass pred_pose[:,3:] in place). */
int airpose_rot6d_to_rotmat(const float* x, int64_t n, float* R, void* stream);
/* strided form: `groups` rows of `per_group` consecutive 6-vectors, rows `row_stride` floats apart */
int airpose_rot6d_to_rotmat_strided(const float* x, int64_t groups, int32_t per_group, int64_t row_stride,
                                    float* R, void* stream);
/* backward of the above: grad_R [groups*per_group,3,3] -> grad_x, written with the same strided layout as x
 * (rows `grad_x_row_stride` floats apart), e.g. straight into d(loss)/d(pred_pose)[:, 3:]. */
int airpose_rot6d_to_rotmat_bwd_strided(const float* x, int64_t groups, int32_t per_group, int64_t row_stride,
                                        const float* grad_R, float* grad_x, int64_t grad_x_row_stride, void* stream);
/* SMPL joint -> 14 OpenPose joints (copenet_real_data/scripts/bundle_adj.py:48,116):
 * out[b,i,:] = joints[b,map[i],:], a bit-exact gather. map=NULL selects the reference map. */
int airpose_j14_gather(const float* joints, int32_t batch, int32_t num_joints, const int32_t* map_host,
                       float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Test-mode outputs and metrics (copenet/src/copenet/copenet_twoview.py:323-326,541-559,583-586)
 * The two conversions restate torchgeometry 0.1.2 (requirements.txt), which is absent offline: PARITY UNPINNED, checked by
 * known-answer identities only (csrc/testmode.cu).
 * ---------------------------------------------------------------------------------- */
/* tgm.rotation_matrix_to_angle_axis: n row-major rotation matrices, matrix i at R + i*mat_stride, its rows row_stride floats
 * apart ((9, 3) for [n,3,3]; (12, 4) for the reference's [n,3,4] input, whose 4th column is ignored) -> out [n,3]. */
int airpose_rotmat_to_angle_axis(const float* R, int64_t n, int32_t mat_stride, int32_t row_stride, float* out, void* stream);
/* tgm.angle_axis_to_rotation_matrix (the 3x3 block of the 4x4 it returns): aa [n,3] -> R [n,3,3]. */
int airpose_angle_axis_to_rotmat(const float* aa, int64_t n, float* R, void* stream);
/* mean over items and over the first points_used of each item's points_per_item 3-vectors of ||a - b||_2 -> out[0]:
 * MPJPE with (127, 22) (copenet_twoview.py:583-586), the mean position error with (1, 1) (:541-551).  Deterministic. */
int airpose_mean_distance(const float* a, const float* b, int64_t items, int32_t points_per_item, int32_t points_used, float* out,
                          void* stream);

/* ------------------------------------------------------------------------------------
 * Input preprocessing (the step right before the path)
 * ---------------------------------------------------------------------------------- */
/* The drone server's stage-0 conversion (catkin_ws/.../airpose_server/server.py:93-98): u8 BGR [n,size,size,3] (device) ->
 * fp32 RGB NCHW, x * (1/255) then (x - mean[c]) / std[c] as three separately rounded fp32 operations (bit-exact with the
 * reference's torch ops).  mean3 / std3 are HOST pointers to three floats each. */
int airpose_preprocess_bgr8(const uint8_t* bgr_hwc, int32_t n_images, int32_t size, const float* mean3, const float* std3,
                            float* out_nchw, void* stream);
/* The dataset's per-camera preprocessing (copenet/src/copenet/dsets/aerialpeople.py:125-141,174; utils/utils.py:214-235):
 * frame[y0:y1, x0:x1] of a u8 BGR frame -> RGB / 255 -> cv2.resize(INTER_LINEAR) to a longer side of `size` -> zero
 * letterbox to size x size -> Normalize(mean, std) -> fp32 NCHW.  frames_bgr: n frames of frame_h x frame_w x 3 bytes on the
 * device, frame_stride_bytes apart (0: every crop is taken from the same frame);
 * rects_dev: int32 [n,4] = (y0, y1, x0, x1) on the device.  The bilinear arithmetic follows cv2's CV_64F path (double work
 * type, float coefficients) so the result agrees with the reference to the final float32 rounding. */
int airpose_preprocess_crop_resize(const uint8_t* frames_bgr, int64_t frame_stride_bytes, int32_t frame_h, int32_t frame_w,
                                   const int32_t* rects_dev, int32_t n_images, int32_t size, const float* mean3, const float* std3,
                                   float* out_nchw, void* stream);

/* ------------------------------------------------------------------------------------
 * ResNet-50 trunk + IEF regressor  (copenet/src/copenet/models/model_copenet.py)
 * ---------------------------------------------------------------------------------- */
typedef struct airpose_net airpose_net_t;

/* One conv + its eval-mode BatchNorm, in the reference's parameter layout (DEVICE pointers,
 * fp32).  Order of the 53 entries = forward order, see airpose_b200/synthetic.py conv_specs(). */
typedef struct {
  const float* weight;        /* [Cout,Cin,k,k]  nn.Conv2d.weight */
  const float* bn_weight;     /* [Cout] gamma */
  const float* bn_bias;       /* [Cout] beta */
  const float* bn_mean;       /* [Cout] running_mean */
  const float* bn_var;        /* [Cout] running_var */
} airpose_conv_params;

typedef struct {
  airpose_conv_params conv[53];
  const float* fc1_w;  const float* fc1_b;          /* [1024,2332], [1024] */
  const float* fc2_w;  const float* fc2_b;          /* [1024,1024], [1024] */
  const float* decpose_w;  const float* decpose_b;  /* [135,1024], [135] */
  const float* decshape_w; const float* decshape_b; /* [10,1024], [10] */
  const float* init_pose;                           /* [144] */
  const float* init_shape;                          /* [10] */
  float bn_eps;                                     /* 1e-5 */
} airpose_net_params;

/* max_images bounds the workspace: images per trunk call (two views of B pairs = 2B). */
int airpose_net_create(airpose_net_t** out, int max_images, int device);
int airpose_net_destroy(airpose_net_t* h);
/* (Re)pack weights: bf16 K-major conv matrices, folded BN scale/shift, the collapsed regressor
 * matrix G = Wdec*W2*W1 (formed in fp64; ief.cu). */
int airpose_net_load(airpose_net_t* h, const airpose_net_params* p, void* stream);

/* copenet.forward_feat_ext (model_copenet.py:161-176), eval mode:
 * x [n,3,224,224] fp32 NCHW -> feat [n,2048] fp32.  bf16 operands, fp32 accumulate. */
int airpose_backbone_fwd(airpose_net_t* h, const float* x_nchw, int n_images, float* out_feat, void* stream);
/* The two trunk passes of copenet.forward (model_copenet.py:140-141) in one call, without concatenating
 * the views: x0, x1 [B,3,224,224] -> out_feat [2B,2048] (rows [0,B) = view 0, [B,2B) = view 1).
 * Eval-mode BatchNorm makes images independent, so both views share every launch. */
int airpose_backbone_fwd_pair(airpose_net_t* h, const float* x0_nchw, const float* x1_nchw, int n_pairs,
                              float* out_feat, void* stream);

/* copenet.forward_feat_ext with the module in train() mode (Lightning's training_step; also the frozen trunk of
 * copenet_real's train_reg_only): every BatchNorm normalises with the statistics of this batch of n images (biased
 * variance) and updates running_mean / running_var in place (momentum 0.1, unbiased variance), torch semantics.  One
 * call = one view, 2 <= n <= the handle's chunk size.  Conv weights are the ones packed by airpose_net_load; the
 * BatchNorm affine parameters and running statistics are read / written live (DEVICE pointers, forward order as in
 * airpose_net_params.conv).  saved_stats (optional, airpose_bn_saved_stats_floats() floats) receives per layer
 * [mean(C) | invstd(C)] for a backward pass. */
typedef struct {
  const float* bn_weight[53];
  const float* bn_bias[53];
  float* running_mean[53];     /* may be NULL per layer: no running-statistics update */
  float* running_var[53];
  float momentum;              /* 0.1 */
  float eps;                   /* 1e-5 */
  float* saved_stats;
  int32_t tape;                /* 0 / 1: keep every layer's raw and normalised output of this call in the handle's tape
                                  (one per view) for airpose_backbone_bwd_train; -1: forward only */
} airpose_bn_train_params;
int64_t airpose_bn_saved_stats_floats(void);
int airpose_backbone_fwd_train(airpose_net_t* h, const float* x_nchw, int n_images, const airpose_bn_train_params* bn,
                               float* out_feat, void* stream);
/* Backward of airpose_backbone_fwd_train for one view (tape 0 / 1): d loss / d features [n,2048] -> gradients of every
 * conv weight (reference layout [Cout,Cin,kh,kw]) and BatchNorm weight / bias, fp32.  accumulate = 0 overwrites the
 * output buffers, 1 adds (the second view of a pair: the two views share the weights).  x_nchw: the images of that
 * forward call (the stem's weight gradient needs them).  Any n (the reference trains with 30 pairs per rank).
 * bf16 tensor-core GEMMs throughout: data gradients as implicit-GEMM convolutions of dz with the flipped, transposed
 * weights (stride-2 layers through a zero-dilated dz), weight gradients as [Cout, K] = dz^T . im2col(x) with both operands
 * transposed to K-major and the long pixel contraction split stream-K over all SMs; BatchNorm backward as two HBM-bound
 * passes per layer. */
typedef struct {
  float* g_weight[53];
  float* g_bn_weight[53];
  float* g_bn_bias[53];
  int32_t accumulate;
  /* Optional (NULL = off): called on the calling thread, once, when everything that writes the gradients of layer3 and layer4
   * (convs 24..52 and their BatchNorms: 95 % of the trunk's parameters) has been enqueued on the stream -- the point where a
   * data-parallel caller can start all-reducing that part (DDP's bucketed overlap, copenet_twoview.py:376-390 under Lightning's
   * DDP) while layer2, layer1 and the stem are still being differentiated. */
  void (*upper_done)(void* user);
  void* user;
} airpose_trunk_grads;
int airpose_backbone_bwd_train(airpose_net_t* h, const float* x_nchw, int n_images, int tape, const airpose_bn_train_params* bn,
                               const float* g_feat, const airpose_trunk_grads* grads, const float* const* conv_weights_f32,
                               void* stream);
/* Both views of a batch of n pairs through ONE set of launches (2n <= chunk = 64 images): every conv GEMM, pooling and, in
 * the backward, every data- / weight-gradient GEMM runs once over the 2n images (the weight gradient contracts over both
 * views' pixels, so there is no second accumulation pass), while BatchNorm runs per view on its half of the rows with its own
 * batch statistics -- the reference calls forward_feat_ext once per view (model_copenet.py:140-141) -- view 0 first, so the
 * running statistics receive the same two updates in the same order.  out_feat / g_feat: [2n,2048], rows [0,n) = view 0.
 * The tape (bn->tape = 0 / 1) then holds both views; the matching backward is airpose_backbone_bwd_train_pair. */
int airpose_backbone_fwd_train_pair(airpose_net_t* h, const float* x0_nchw, const float* x1_nchw, int n_pairs,
                                    const airpose_bn_train_params* bn, float* out_feat, void* stream);
int airpose_backbone_bwd_train_pair(airpose_net_t* h, const float* x0_nchw, const float* x1_nchw, int n_pairs, int tape,
                                    const airpose_bn_train_params* bn, const float* g_feat, const airpose_trunk_grads* grads,
                                    const float* const* conv_weights, void* stream);    /* conv_weights_f32: the 53 live fp32 conv weights [Cout,Cin,kh,kw] (DEVICE) */
/* Building blocks of the trunk backward, exported for the per-layer parity tests (same code paths as above).
 * conv_bwd: weight gradient (fp32 [Cout,Cin,k,k], overwritten or accumulated) and, when out_dx is given, data gradient
 * (+ add) of trunk conv `conv_idx` for n images: dz bf16 [n,Ho,Wo,Cout], x_in bf16 [n,Hin,Win,Cin].
 * bn_bwd: BatchNorm(+ReLU) backward on [M,C] bf16: dy, y (post-activation output; NULL = no ReLU), z (conv output),
 * stats [mean(C) | invstd(C)], gamma -> dz, dpre (optional), dgamma, dbeta. */
int airpose_debug_conv_bwd(airpose_net_t* h, int conv_idx, int n_images, const void* dz, const void* x_in, const float* w_f32,
                           const void* add, void* out_dx, float* out_gw, int accumulate, void* stream);
int airpose_debug_bn_bwd(airpose_net_t* h, int64_t M, int C, const void* dy, const void* y, const void* z, const float* stats,
                         const float* gamma, void* out_dz, void* out_dpre, float* g_gamma, float* g_beta, int accumulate,
                         void* stream);
/* copies one bf16 tensor of a training tape (which: 0 = raw conv output z, 1 = BN(+residual)(+ReLU) output y, 2 = max-pooled
 * stem output) into a caller buffer of exactly that many elements (tests). */
int airpose_debug_tape_get(airpose_net_t* h, int tape, int conv_idx, int which, void* dst, int64_t dst_elems, void* stream);
/* Re-packs only the conv weights / folded BatchNorm of airpose_net_load (what a full training step invalidates). */
int airpose_net_load_trunk(airpose_net_t* h, const airpose_net_params* p, void* stream);
/* Re-forms only the collapsed regressor matrix of airpose_net_load (the conv weights stay packed): what changes
 * between the steps of a regressor-only training run. */
int airpose_net_load_regressor(airpose_net_t* h, const airpose_net_params* p, void* stream);

/* The regressor half of copenet.forward (model_copenet.py:118-159,178-204), eval mode: with dropout
 * inactive the three Linears have no nonlinearity between them, so the pass is evaluated as one
 * affine map per iteration (fp32 FMA; differs from the reference chain by summation order only).
 * xf0/xf1 [B,2048]; bb0/bb1 [B,3]; pos0/pos1 [B,3] (already scaled, copenet_twoview.py:199-203);
 * init_theta0/1 [B,>=132] or NULL (module init_pose); init_shape0/1 [B,10] or NULL.
 * Outputs pred_pose [B,135], pred_betas [B,10] per view. */
typedef struct {
  int32_t batch;
  int32_t iters;
  const float* xf0; const float* xf1;
  const float* bb0; const float* bb1;
  const float* pos0; const float* pos1;
  const float* init_theta0; const float* init_theta1; int32_t init_theta_stride;
  const float* init_shape0; const float* init_shape1; int32_t init_shape_stride;
  float* out_pose0; float* out_betas0; float* out_pose1; float* out_betas1;
} airpose_ief_args;
int airpose_ief_fwd(airpose_net_t* h, const airpose_ief_args* a, void* stream);

/* Training-mode regressor (dropout active, so no collapse): forward that saves its activations, and the backward pass
 * autograd would run through copenet.forward's regressor loop (model_copenet.py:118-159,178-204).  Parameters are read
 * live from the module's fp32 tensors (DEVICE pointers).  mask1/mask2: multiplicative dropout masks
 * [iters][2 views][B][1024] (0 or 1/(1-p)) drawn by the caller, or both NULL for eval semantics.  `saved` must hold
 * airpose_ief_train_saved_floats(B, iters) floats and live from fwd to bwd; `workspace` holds
 * airpose_ief_train_workspace_floats(B) floats.  bwd overwrites every g_* buffer (parameter gradients are summed over
 * the iterations and views in a fixed order); g_xf0/g_xf1 (d loss / d trunk features, the entry point of the trunk
 * backward) are optional.  This is the trainable part of copenet_real's `train_reg_only` mode
 * (copenet_real/src/copenet_real/copenet_twoview.py:357-372). */
typedef struct {
  int32_t batch, iters;
  const float* xf0; const float* xf1;
  const float* bb0; const float* bb1;
  const float* pos0; const float* pos1;
  const float* init_theta0; const float* init_theta1; int32_t init_theta_stride;
  const float* init_shape0; const float* init_shape1; int32_t init_shape_stride;
  const float* fc1_w; const float* fc1_b; const float* fc2_w; const float* fc2_b;
  const float* decpose_w; const float* decpose_b; const float* decshape_w; const float* decshape_b;
  const float* init_pose; const float* init_shape;
  const float* mask1; const float* mask2;
  float* saved; float* workspace;
  float* out_pose0; float* out_betas0; float* out_pose1; float* out_betas1;               /* fwd */
  const float* g_pose0; const float* g_betas0; const float* g_pose1; const float* g_betas1; /* bwd: upstream */
  float* g_fc1_w; float* g_fc1_b; float* g_fc2_w; float* g_fc2_b;
  float* g_decpose_w; float* g_decpose_b; float* g_decshape_w; float* g_decshape_b;
  float* g_xf0; float* g_xf1;
} airpose_ief_train_args;
int64_t airpose_ief_train_saved_floats(int32_t batch, int32_t iters);
int64_t airpose_ief_train_workspace_floats(int32_t batch);
int airpose_ief_train_fwd(const airpose_ief_train_args* a, void* stream);
int airpose_ief_train_bwd(const airpose_ief_train_args* a, void* stream);

/* ------------------------------------------------------------------------------------
 * hmr baseline: single view, same trunk  (copenet/src/copenet/models/model_hmr.py:48-172; BASELINE config 1)
 * ---------------------------------------------------------------------------------- */
typedef struct {
  airpose_conv_params conv[53];
  const float* fc1_w;  const float* fc1_b;          /* [1024,2193], [1024]   2193 = 2048 + 132 + 10 + 3 (:66) */
  const float* fc2_w;  const float* fc2_b;          /* [1024,1024], [1024] */
  const float* decpose_w;  const float* decpose_b;  /* [132,1024], [132] */
  const float* decshape_w; const float* decshape_b; /* [10,1024], [10] */
  const float* deccam_w;   const float* deccam_b;   /* [3,1024], [3] */
  const float* init_pose;                           /* [144], the first 132 are used (:117) */
  const float* init_shape;                          /* [10] */
  const float* init_cam;                            /* [3] */
  float bn_eps;
} airpose_hmr_params;
/* Loads the trunk and the hmr regressor into a handle made by airpose_net_create (the trunk entry points
 * airpose_backbone_fwd* then serve model_hmr.forward_feat_ext :143-158). */
int airpose_hmr_load(airpose_net_t* h, const airpose_hmr_params* p, void* stream);

/* The regressor loop of model_hmr.copenet.forward (:112-141, forward_reg :160-172), eval mode, collapsed to one
 * affine map per iteration like airpose_ief_fwd.  xf [B,2048]; init_theta [B,>=132] / init_shape [B,10] /
 * init_cam [B,3] or NULL (module buffers).  Outputs pred_pose [B,132] (6D, before rot6d_to_rotmat :140),
 * pred_betas [B,10], pred_cam [B,3]. */
typedef struct {
  int32_t batch;
  int32_t iters;
  const float* xf;
  const float* init_theta; int32_t init_theta_stride;
  const float* init_shape; int32_t init_shape_stride;
  const float* init_cam;   int32_t init_cam_stride;
  float* out_pose; float* out_betas; float* out_cam;
} airpose_hmr_ief_args;
int airpose_hmr_ief_fwd(airpose_net_t* h, const airpose_hmr_ief_args* a, void* stream);

/* ------------------------------------------------------------------------------------
 * training-step loss  (copenet/src/copenet/copenet_twoview.py:83-161, get_loss)
 * ---------------------------------------------------------------------------------- */
/* All seven MSE terms, their weighted sum x60 and (optionally) the gradient of the total loss with
 * respect to every prediction, in one pass.  Shapes follow the reference tensors; `trans*` may be a
 * strided view (pred_pose[:, :3], stride 135).  `out` receives 8 floats in the order of the reference's
 * `losses` dict: loss, loss_regr_trans, loss_keypoints, loss_keypoints_3d, loss_regr_shape, loss_rootrot,
 * loss_regr_pose, loss_regul_betas -- one device buffer instead of eight `.item()` syncs.
 * Gradient buffers: all NULL (forward only) or all given; same shapes as the predictions (g_trans* dense [B,3]). */
typedef struct {
  int32_t batch, num_verts, num_joints;                 /* B, 10475, 127 */
  const float* trans0; const float* trans1; int32_t trans_stride;        /* pred_smpltrans  [B,3] */
  const float* rotmat0; const float* rotmat1;           /* pred_rotmat        [B,22,3,3] */
  const float* betas0; const float* betas1;             /* pred_betas         [B,10] */
  const float* verts0; const float* verts1;             /* pred_output_cam.vertices [B,V,3] (canonical) */
  const float* joints0; const float* joints1;           /* pred_output_cam.joints   [B,127,3] */
  const float* j2d0; const float* j2d1;                 /* pred_joints_2d_cam [B,127,2] */
  const float* gt_pose_rotmat;                          /* smplpose_rotmat    [B,21,3,3] */
  const float* gt_trans0; const float* gt_trans1;       /* smpltrans_rel{0,1} [B,3] */
  const float* gt_orient0; const float* gt_orient1;     /* smplorient_rel{0,1} [B,1,3,3] */
  const float* gt_verts;                                /* smpl_vertices      [B,1,V,3] */
  const float* gt_joints;                               /* smpl_joints        [B,1,127,3] */
  const float* gt_j2d0; const float* gt_j2d1;           /* smpl_joints_2d{0,1} [B,1,127,2] */
  float w_shape, w_kp2d, w_kp3d, w_limbs3d, w_limbstheta, w_trans, w_rootrot, w_pose, w_beta;   /* :655-677 */
  float* out;                                           /* [8] */
  float* g_verts0; float* g_verts1; float* g_joints0; float* g_joints1; float* g_j2d0; float* g_j2d1;
  float* g_rotmat0; float* g_rotmat1; float* g_betas0; float* g_betas1; float* g_trans0; float* g_trans1;
} airpose_twoview_loss_args;
int airpose_twoview_loss(const airpose_twoview_loss_args* a, void* stream);

/* copenet_real's get_loss without the VPoser prior (copenet_real/src/copenet_real/copenet_twoview.py:99-160; SURVEY.md 8(f)
 * rank 4): confidence-weighted 2D keypoint loss over the first 22 joints with the limb weights, cross-view pose and beta
 * consistency, beta regulariser, exp(-t_z)^2 depth barrier, x60; `vposer_term` (loss_regul_vposer, computed by the caller if it
 * has the VPoser model, else 0) enters the total with `w_vposer`.  gt_j2d* = smpl_joints_2d{0,1}[:, 0]: [B, gt_joints, 3] =
 * (x, y, confidence).  `out` receives 5 floats in the order of the reference's `losses` dict: loss, loss_regul_vposer,
 * loss_regr_pose, loss_keypoints, loss_regul_betas.  Gradient buffers: all NULL or all given (g_trans* dense [B,3]). */
typedef struct {
  int32_t batch, num_joints, gt_joints;                 /* B, 127, joints per ground-truth row (>= 22) */
  const float* trans0; const float* trans1; int32_t trans_stride;        /* pred_smpltrans  [B,3] */
  const float* rotmat0; const float* rotmat1;           /* pred_rotmat        [B,22,3,3] */
  const float* betas0; const float* betas1;             /* pred_betas         [B,10] */
  const float* j2d0; const float* j2d1;                 /* pred_joints_2d_cam [B,127,2] */
  const float* gt_j2d0; const float* gt_j2d1;           /* [B, gt_joints, 3] */
  float w_kp2d, w_limbs2d, w_beta, w_pose, w_vposer, vposer_term;
  float* out;                                           /* [5] */
  float* g_j2d0; float* g_j2d1; float* g_rotmat0; float* g_rotmat1; float* g_betas0; float* g_betas1; float* g_trans0; float* g_trans1;
} airpose_real_loss_args;
int airpose_real_loss(const airpose_real_loss_args* a, void* stream);

/* torch.optim.Adam(..., weight_decay=0, amsgrad=True) (copenet_twoview.py:416-425) over a FLAT fp32 buffer:
 * one launch per step for all 27.1 M parameters.  `step` is the 1-based step count (bias corrections are
 * formed on the host in fp64 like torch's python scalars).  max_exp_avg_sq = NULL gives plain Adam.
 * grad_scale (0 = 1) multiplies the gradient on the fly (e.g. 1/world_size after a sum all-reduce). */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq; float* max_exp_avg_sq;
  int64_t n;
  float lr, beta1, beta2, eps;
  int32_t step;
  float grad_scale;
} airpose_adam_args;
int airpose_adam_step(const airpose_adam_args* a, void* stream);

/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t airpose_launch_count(void);

/* ------------------------------------------------------------------------------------
 * low-level building blocks, exported for the per-layer parity tests
 * ---------------------------------------------------------------------------------- */
/* D[M,N] = A[M,K] * B[N,K]^T on tcgen05 (bf16 operands, fp32 accumulate in TMEM).
 * Epilogue: v = acc*scale[n] + shift[n] (+ residual[m,n] bf16) ; relu ; store bf16 and/or fp32.
 * lda/ldb/ldd in elements.  scale may be NULL (=1). */
typedef struct {
  const void* A; int64_t lda;      /* bf16 [M,K] */
  const void* B; int64_t ldb;      /* bf16 [N,K] */
  int32_t M, N, K;
  const float* scale; const float* shift;
  const void* residual; int64_t ldr; /* bf16 [M,N] or NULL */
  int32_t relu;
  void*  out_bf16; int64_t ldd;     /* bf16 [M,N] or NULL */
  float* out_f32;  int64_t ldf;     /* fp32 [M,N] or NULL */
  /* a_t != 0: A is given transposed, bf16 [K,M] with row pitch lda (M contiguous); b_t likewise for B as [K,N].  The operand is
   * then read through MN-major tcgen05 descriptors -- no transpose pass.  This is the weight-gradient form: both operands of
   * dW = dZ^T . X contract over the pixels, the OUTER dimension of NHWC tensors (what torch autograd computes for
   * copenet_twoview.py:378-386).  Needs out_bf16 with N % 64 == 0, and M % 8 == 0 (a_t) / N % 8 == 0 (b_t). */
  int32_t a_t, b_t;
} airpose_gemm_args;
int airpose_gemm_bf16(const airpose_gemm_args* g, void* stream);

/* Implicit-GEMM convolution on NHWC bf16 through TMA im2col (3x3 / strided 1x1 convs).
 * x [n,H,W,Cin] bf16, w [Cout, kh*kw*Cin] bf16 (tap-major, channel-minor), out [n,Ho,Wo,Cout]. */
typedef struct {
  const void* x; int32_t n, H, W, Cin;
  const void* w; int32_t Cout, ksize, stride, pad;
  const float* scale; const float* shift;
  const void* residual;             /* bf16 [n,Ho,Wo,Cout] or NULL */
  int32_t relu;
  void* out;                        /* bf16 [n,Ho,Wo,Cout] */
} airpose_conv_args;
int airpose_conv_bf16(const airpose_conv_args* c, void* stream);

/* Fused tail of a stride-1 bottleneck (Bottleneck.forward, model_copenet.py:33-46): conv2 3x3 pad 1 + bn2 + ReLU ->
 * conv3 1x1 + bn3 + residual + ReLU in ONE launch (halo slab in shared memory, nine row-shifted tcgen05 windows).
 * t1 [n,H,W,Cm] bf16 (conv1's output), w2 [Cm, 9*Cm] bf16 (tap-major, channel-minor), w3 [4*Cm, Cm] bf16,
 * residual / out [n,H,W,4*Cm] bf16.  Supported: Cm = 64, 2*(W+2) <= 128 (the 56x56 stage). */
typedef struct {
  const void* t1; int32_t n, H, W, Cm;
  const void* w2; const float* scale2; const float* shift2;
  const void* w3; const float* scale3; const float* shift3;
  const void* residual;
  void* out;
} airpose_bneck_tail_args;
int airpose_bneck_tail_bf16(const airpose_bneck_tail_args* a, void* stream);

/* Stem only (conv 7x7 s2 + BN + ReLU + MaxPool 3x3 s2, model_copenet.py:163-166):
 * x [n,3,224,224] fp32 NCHW -> out [n,56,56,64] bf16 NHWC.  n <= the handle's chunk size. */
int airpose_backbone_stem(airpose_net_t* h, const float* x_nchw, int n_images, void* out_nhwc_bf16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIRPOSE_B200_H_ */
