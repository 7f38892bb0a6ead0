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
constexpr int kTcK = 192;                      // 21*9 = 189 pose features padded to 3 k-blocks of 64
constexpr int kTcMeshTile = 32;                // meshes per MMA tile (N)
constexpr int kTcSub = 16;                     // meshes per record sub-batch (one bulk copy)
constexpr int kTcRecFloats = 292;              // per-mesh record: A[22][12] | betas[10]+pad2 | camR[9] camt[3] | transl[3] pad
constexpr int kTcRecBetas = 264, kTcRecCam = 276, kTcRecTransl = 288;
constexpr int kTcMaxKW = 8;
constexpr float kTcPScale = 1024.f;            // posedirs are stored as fp16(P * 2^10): keeps small entries normal

struct SmplxTc {
  bool ok = false;             // model supports the tensor-core path
  int KW = 0;                  // max non-zeros per vertex after folding joints 22.. onto their body ancestor
  int vtiles = 0;              // ceil(V / 128)
  __half* P = nullptr;         // [3][vtiles*128][192] fp16, coordinate-major, K-major rows
  CUtensorMap tmP;
  const int* sk_off = nullptr; // [KW][V] byte offset of the joint's A inside a record (j*48)
  const float* sk_w = nullptr; // [KW][V]
  const int* sk_cnt = nullptr; // [V]
};

struct TcCall {
  int B, nb, has_transl;
  const float* rec;            // [Bpad][292]
  const __half* fh;            // [B][192]
  const __half* fl;            // [B][192]
  float* out;                  // [B,V,3]
  float* out_cam;              // [B,V,3] or null
};

int smplx_tc_create(const airpose_smplx_model_host* mh, const SmplxDev& d, SmplxTc* tc, std::vector<void*>* owned);
int smplx_tc_forward(const SmplxDev& d, const SmplxTc& tc, const TcCall& c, cudaStream_t stream);

}  // namespace airpose
