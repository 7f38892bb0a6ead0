// The per-GPU network handle shared by trunk.cu (ResNet-50) and ief.cu (regressor).
#pragma once
#include <cuda_bf16.h>

#include <map>
#include <utility>
#include <vector>

#include "common.cuh"
#include "gemm.cuh"

namespace airpose {

struct ConvSpec { int cout, cin, k, stride, pad; };

struct PlanOp { int kind; int idx; };  // kind 0: gemms[idx] (one conv), 1: tails[idx] (fused conv2 + conv3, bneck.cu), 2: slabs[idx] (conv3x3.cu)
struct TrunkPlan {
  std::vector<GemmLaunch> gemms;       // [stem,] then per block conv1, conv2, [down], conv3 (conv2/conv3 absent when fused)
  std::vector<TailLaunch> tails;
  std::vector<SlabLaunch> slabs;       // 3x3 convs of the 128-channel stage (conv3x3.cu)
  StemLaunch stem;                     // stage-A plans: the fused stem (when enabled)
  std::vector<PlanOp> ops;             // launch order (the stem GEMM, gemms[0] of a stage-A plan, is not listed)
  const __nv_bfloat16* final_act = nullptr;
};

constexpr int kFeat = 2048;            // trunk feature width (model_copenet.py:173-174)

// Collapsed regressor (ief.cu): G = Wdec * W2 * W1, stored transposed and padded for coalesced reads.
struct IefState {
  float* GxT = nullptr;        // [2048][160]  columns of G that multiply the image feature
  float* GuT = nullptr;        // [284][160]   columns of G that multiply the iterated state
  float* g = nullptr;          // [160]        Wdec (W2 b1 + b2) + bdec
  double* T = nullptr;         // [145][1024]  Wdec * W2 (load-time scratch, fp64)
  float* wdec = nullptr;       // [145][1024] the decoders, concatenated at load
  float* bdec = nullptr;       // [145]
  float* init_pose = nullptr;  // [144]
  float* init_shape = nullptr; // [10]
  float* init_cam = nullptr;   // [3]   (hmr only)
  float* partial = nullptr;    // [kIefKSlices][rows][160] split-K partials of GxT . xf
  int partial_rows = 0;
};

}  // namespace airpose

struct airpose_net {
  int device = 0;
  int max_images = 0;
  int chunk = 0;
  bool loaded = false;
  std::vector<airpose::ConvSpec> specs;
  std::vector<__nv_bfloat16*> wq;       // packed conv weights
  __nv_bfloat16* wq_stem_pairs = nullptr;   // conv1 in the column-pair layout of the fused stem (stem.cu)
  std::vector<float*> scale, shift;     // folded BN
  airpose::IefState ief;                // two-view regressor (model_copenet.py)
  airpose::IefState ief_hmr;            // single-view hmr regressor (model_hmr.py)
  bool hmr_loaded = false;
  // workspaces
  // stage A (stem, layer1, layer2) works on `chunk` images at a time; consecutive chunks alternate between two
  // buffer sets and two streams so that one chunk's kernel tails overlap the other's ramp-up (trunk.cu)
  static constexpr int kSets = 2;
  __nv_bfloat16* colS[kSets] = {nullptr, nullptr};         // packed stem operand
  __nv_bfloat16* stem_outS[kSets] = {nullptr, nullptr};
  __nv_bfloat16* actS[kSets][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int sets = 1;
  __nv_bfloat16* actB[4] = {nullptr, nullptr, nullptr, nullptr};   // stage B (layer3, layer4): `group` images
  // split mode (trunk.cu): every chunk runs stage B itself, on its own stream, in its own buffers (set 1: actB1, `chunk` images)
  __nv_bfloat16* actB1[4] = {nullptr, nullptr, nullptr, nullptr};
  bool split_b = false;
  int group = 0;
  // training-mode forward (airpose_backbone_fwd_train): raw conv output, BatchNorm partial sums, scale/shift of the layer
  __nv_bfloat16* ztrain = nullptr;
  float* bn_part = nullptr;
  float* bn_scale = nullptr;
  float* bn_shift = nullptr;
  std::vector<int64_t> bn_save_off;
  // tape of a training-mode forward (one per view): what the trunk backward needs
  struct Tape {
    int n = 0, cap = 0;                 // images on the tape (both views together for a two-view tape)
    int views = 1;                      // 2: images [0, n/2) are view 0, [n/2, n) view 1 (airpose_backbone_fwd_train_pair)
    std::vector<__nv_bfloat16*> z, y;   // per conv: raw output, and BN(+residual)+ReLU output
    __nv_bfloat16* pooled = nullptr;    // max-pooled stem output = input of layer1
    uint8_t* pool_idx = nullptr;        // window position of each pooled element's (first) maximum, for the backward
    float* stats = nullptr;             // per conv [mean(C) | invstd(C)]
    float* stats1 = nullptr;            // the same for view 1 of a two-view tape
  } tape[2];
  // scratch of the backward pass
  __nv_bfloat16* bw[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // gradient ping-pong / dz / dpre / dilated
  __nv_bfloat16* bw_t0 = nullptr;       // transposed dz   [Cout][M]
  __nv_bfloat16* bw_t1 = nullptr;       // transposed im2col(x) [K][M]
  __nv_bfloat16* bw_w = nullptr;        // packed dgrad weights / wgrad output (per-layer debug entry point)
  __nv_bfloat16* bw_wd = nullptr;       // dgrad operands of ALL convs, packed by one launch per backward pass (bw_wd_off[i])
  __nv_bfloat16* bw_wg = nullptr;       // wgrad GEMM outputs of ALL convs, unpacked by one launch at the end of the pass
  size_t bw_w_off[64] = {0};            // element offset of conv i inside bw_wd / bw_wg
  bool bw_batched = false;              // inside a whole backward pass: use the per-layer slots above
  bool bw_reduced[64] = {false};        // conv i's weight gradient was written by wgrad_reduce_kernel in this pass (no unpack)
  float* bw_coef = nullptr;             // [3][2048] BatchNorm backward coefficients
  int bw_cap = 0;
  std::map<std::pair<int, int>, airpose::TrunkPlan> plansA;        // (images, 2 * first image inside the group + buffer set)
  std::map<int, airpose::TrunkPlan> plansB;                        // images
};

namespace airpose {
int ief_create(airpose_net* h);
void ief_destroy(airpose_net* h);
int ief_load(airpose_net* h, const airpose_net_params* p, cudaStream_t st);
int ief_load_hmr(airpose_net* h, const airpose_hmr_params* p, cudaStream_t st);
}  // namespace airpose
