// Shared declarations of the SMPL-X kernels (smplx.cu: pose / joints / generic vertex kernel,
// smplx_tc.cu: tensor-core vertex kernel of the hot path).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <vector>

#include "common.cuh"

namespace airpose {

constexpr int kMaxJoints = 64;
constexpr int kMaxShape = 20;

struct SmplxDev {
  int V, J, NS, P, L, E, KW;
  const float* v_template;   // [V,3]
  const float* shapedirs;    // [V,3,NS]
  const float* posedirs;     // [P, V*3]
  const float* J_template;   // [J,3]
  const float* J_shapedirs;  // [J,3,NS]
  const int* parents;        // [J]
  const int* skin_idx;       // [KW,V]
  const float* skin_w;       // [KW,V]
  const int* lmk_vidx;       // [L,3]
  const float* lmk_bary;     // [L,3]
  const int* extra_idx;      // [E]
};

// ---- tensor-core path (hot path only: 21 body rotations given, joints 22.. identity, <= 10 betas)
constexpr int kTcBodyJoints = 22;              // joints whose A matrix is distinct on the hot path
constexpr int kTcKPose = 192;                  // 21*9 = 189 pose features padded to 3 k-blocks of 64
constexpr int kTcKShape = 32;                  // one 32-wide k-block: template + shape blend as split-fp16 products (below)
constexpr int kTcK = kTcKPose + kTcKShape;     // row length of P and F
// columns 192..223 -- vertex side (P):  T_hi  T_lo | S_hi[0..9] | S_hi[0..9] | S_lo[0..9]     (all x 2^10)
//                     mesh side (F_hi): 1     1    | b_hi[0..9] | b_lo[0..9] | b_hi[0..9]
// so that D = 2^10 (v_template + shapedirs . beta + posedirs . feat) to ~2^-22 relative (the S_lo b_lo term is dropped)
constexpr int kTcMeshTile = 32;                // meshes per MMA tile (N)
constexpr int kTcSub = 8;                      // meshes per record sub-batch (one bulk copy, one epilogue warp's share of a tile)
constexpr int kTcRecFloats = 280;              // per-mesh record: A[22][12] | camR[9] camt[3] | transl[3] pad
constexpr int kTcRecCam = 264, kTcRecTransl = 276;
constexpr int kTcMaxKW = 8;
constexpr int kMlVertTile = 32;                // vertices per tile of the mesh-lane kernel (smplx_ml.cu)
constexpr int kMlMinBatch = 256;               // batches from here on run the mesh-lane kernel (128 meshes per CTA)
constexpr float kTcPScale = 1024.f;            // posedirs are stored as fp16(P * 2^10): keeps small entries normal

struct SmplxTc {
  bool ok = false;             // model supports the tensor-core path
  int KW = 0;                  // max non-zeros per vertex after folding joints 22.. onto their body ancestor
  int vtiles = 0;              // ceil(V / 128)
  __half* P = nullptr;         // [3][vtiles*128][224] fp16, coordinate-major, K-major rows
  CUtensorMap tmP;             // 64-column boxes (128-byte swizzle): the pose k-blocks
  CUtensorMap tmPx;            // 32-column boxes (64-byte swizzle): the template / shape k-block
  const int* sk_off = nullptr; // [KW][V] byte offset of the joint's A inside a record (j*48)
  const float* sk_w = nullptr; // [KW][V]
  const int* sk_cnt = nullptr; // [V]
  // mesh-lane kernel (smplx_ml.cu)
  bool ml_ok = false;
  const uint32_t* ml_vtab = nullptr;   // [ml_vtiles * 32][8]: 4 slot ids (float offset j * 12 in a record) + 4 weights per vertex
  int ml_vtiles = 0;                   // ceil(V / 32)
  CUtensorMap tmP32, tmPx32;           // 32-row boxes of P (pose k-blocks / shape k-block)
};

struct TcCall {
  int B, nb, has_transl;
  const float* rec;            // [Bpad][280]: pairs of meshes element-interleaved (smplx_tc.cu) or plain rows (smplx_ml.cu)
  const __half* fh;            // [B][224]
  const __half* fl;            // [B][224]
  float* out;                  // [B,V,3]
  float* out_cam;              // [B,V,3] or null
};

int smplx_tc_create(const airpose_smplx_model_host* mh, const SmplxDev& d, SmplxTc* tc, std::vector<void*>* owned);
int smplx_tc_forward(const SmplxDev& d, const SmplxTc& tc, const TcCall& c, cudaStream_t stream);
int smplx_ml_create(const airpose_smplx_model_host* mh, const SmplxDev& d, SmplxTc* tc, std::vector<void*>* owned);
int smplx_ml_forward(const SmplxDev& d, const SmplxTc& tc, const TcCall& c, cudaStream_t stream);

}  // namespace airpose
