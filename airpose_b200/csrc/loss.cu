// copenet_twoview.get_loss on the device, plus its gradient with respect to every prediction.
//
// Replaces /root/reference/copenet/src/copenet/copenet_twoview.py:83-161: seven mean-squared-error
// terms (2D keypoints, limb-weighted 3D keypoints, vertices incl. the cross-view consistency term,
// translation, root rotation, limb-weighted pose rotation matrices, beta regulariser), their weighted
// sum, x60 -- and the eight `.item()` host syncs of the reference become one 8-float result buffer.
//
// The vertex term dominates (three [B,10475,3] tensors): one pass, HBM-bound, float4 loads; when the
// gradient buffers are given the same pass writes dL/dvertices for both views (the loss is a sum of
// squares, so its gradient is elementwise).  Reductions are deterministic: per-thread fp32 partial ->
// fixed-order warp/CTA tree -> fp64 partial per CTA -> one finishing CTA sums them in order.
#include "common.cuh"

namespace airpose {

namespace {

constexpr int kTerms = 7;              // trans, kp2d, kp3d, shape, rootrot, pose, betas (sums of squares)
constexpr int kLossThreads = 256;
constexpr int kLossGrid = 148 * 4;

struct LossK {
  airpose_twoview_loss_args a;
  double* partial;                     // [kLossGrid][kTerms]
};

__device__ __forceinline__ float joint_w3d(int j, float w) {     // copenet_twoview.py:114-115
  if (j == 4 || j == 5 || j == 18 || j == 19) return w;
  if (j == 7 || j == 8 || j == 20 || j == 21) return w * w;
  return 1.f;
}
__device__ __forceinline__ float joint_wtheta(int j, float w) {  // :133-134 (index into the 21 body rotations)
  if (j == 3 || j == 4 || j == 17 || j == 18) return w;
  if (j == 6 || j == 7 || j == 19 || j == 20) return w * w;
  return 1.f;
}

// three-way squared error of one element: (p0-g)^2 + (p1-g)^2 + (p0-p1)^2 and its gradients
__device__ __forceinline__ float tri(float p0, float p1, float g, float& d0, float& d1) {
  const float a = p0 - g, b = p1 - g, c = p0 - p1;
  d0 = a + c;                           // d/dp0 of the sum, without the factor 2
  d1 = b - c;
  return a * a + b * b + c * c;
}

__global__ void __launch_bounds__(kLossThreads) loss_partial_kernel(LossK k) {
  const airpose_twoview_loss_args& a = k.a;
  const int B = a.batch, V = a.num_verts, J = a.num_joints;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  float acc[kTerms];
#pragma unroll
  for (int i = 0; i < kTerms; ++i) acc[i] = 0.f;
  const bool grads = a.g_verts0 != nullptr;
  const float c60 = 60.f;

  // ---- vertices (term 3): mean over B*V*3                                                    :118-120
  {
    const int64_t n = (int64_t)B * V * 3, n4 = n / 4;
    const float gs = grads ? c60 * a.w_shape * 2.f / (float)n : 0.f;
    const float4* p0 = reinterpret_cast<const float4*>(a.verts0);
    const float4* p1 = reinterpret_cast<const float4*>(a.verts1);
    const float4* pg = reinterpret_cast<const float4*>(a.gt_verts);
    float4* g0 = reinterpret_cast<float4*>(a.g_verts0);
    float4* g1 = reinterpret_cast<float4*>(a.g_verts1);
    for (int64_t i = tid; i < n4; i += nth) {
      const float4 x = __ldg(p0 + i), y = __ldg(p1 + i), g = __ldg(pg + i);
      float4 dx, dy;
      acc[3] += tri(x.x, y.x, g.x, dx.x, dy.x) + tri(x.y, y.y, g.y, dx.y, dy.y) + tri(x.z, y.z, g.z, dx.z, dy.z) +
                tri(x.w, y.w, g.w, dx.w, dy.w);
      if (grads) {
        g0[i] = make_float4(gs * dx.x, gs * dx.y, gs * dx.z, gs * dx.w);
        g1[i] = make_float4(gs * dy.x, gs * dy.y, gs * dy.z, gs * dy.w);
      }
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += nth) {
      float dx, dy;
      acc[3] += tri(a.verts0[i], a.verts1[i], a.gt_verts[i], dx, dy);
      if (grads) { a.g_verts0[i] = gs * dx; a.g_verts1[i] = gs * dy; }
    }
  }
  // ---- 3D keypoints (term 2): first 22 joints, limb weights, mean over B*22*3                 :110-116
  {
    const int64_t n = (int64_t)B * 22 * 3;
    const float gs = c60 * a.w_kp3d * 2.f / (float)n;
    for (int64_t i = tid; i < (int64_t)B * J * 3; i += nth) {
      const int b = (int)(i / (J * 3)), r = (int)(i % (J * 3)), j = r / 3;
      float d0 = 0.f, d1 = 0.f;
      if (j < 22) {
        const float w = joint_w3d(j, a.w_limbs3d);
        acc[2] += w * tri(a.joints0[i], a.joints1[i], a.gt_joints[i], d0, d1);
        d0 *= w * gs; d1 *= w * gs;
      }
      if (a.g_joints0) { a.g_joints0[i] = d0; a.g_joints1[i] = d1; }
      (void)b;
    }
  }
  // ---- 2D keypoints (term 1): first 22 joints of each view, mean over B*22*2                  :107-108
  {
    const int64_t n = (int64_t)B * 22 * 2;
    const float gs = c60 * a.w_kp2d * 2.f / (float)n;
    for (int64_t i = tid; i < (int64_t)B * J * 2; i += nth) {
      const int j = (int)(i % (J * 2)) / 2;
      float d0 = 0.f, d1 = 0.f;
      if (j < 22) {
        d0 = a.j2d0[i] - a.gt_j2d0[i];
        d1 = a.j2d1[i] - a.gt_j2d1[i];
        acc[1] += d0 * d0 + d1 * d1;
        d0 *= gs; d1 *= gs;
      }
      if (a.g_j2d0) { a.g_j2d0[i] = d0; a.g_j2d1[i] = d1; }
    }
  }
  // ---- rotation matrices: root (term 4, :125-126) and the 21 body joints (term 5, :128-135)
  {
    const float gs_root = c60 * a.w_rootrot * 2.f / (float)((int64_t)B * 9);
    const float gs_pose = c60 * a.w_pose * 2.f / (float)((int64_t)B * 21 * 9);
    for (int64_t i = tid; i < (int64_t)B * 22 * 9; i += nth) {
      const int b = (int)(i / 198), r = (int)(i % 198), j = r / 9, e = r % 9;
      float d0, d1;
      if (j == 0) {
        d0 = a.rotmat0[i] - a.gt_orient0[(int64_t)b * 9 + e];
        d1 = a.rotmat1[i] - a.gt_orient1[(int64_t)b * 9 + e];
        acc[4] += d0 * d0 + d1 * d1;
        d0 *= gs_root; d1 *= gs_root;
      } else {
        const float w = joint_wtheta(j - 1, a.w_limbstheta);
        acc[5] += w * tri(a.rotmat0[i], a.rotmat1[i], a.gt_pose_rotmat[(int64_t)b * 189 + (j - 1) * 9 + e], d0, d1);
        d0 *= w * gs_pose; d1 *= w * gs_pose;
      }
      if (a.g_rotmat0) { a.g_rotmat0[i] = d0; a.g_rotmat1[i] = d1; }
    }
  }
  // ---- translation (term 0, :122-123) and betas (term 6, :137-139)
  {
    const float gs_t = c60 * a.w_trans * 2.f / (float)((int64_t)B * 3);
    for (int64_t i = tid; i < (int64_t)B * 3; i += nth) {
      const int b = (int)(i / 3), e = (int)(i % 3);
      const float d0 = a.trans0[(int64_t)b * a.trans_stride + e] - a.gt_trans0[i];
      const float d1 = a.trans1[(int64_t)b * a.trans_stride + e] - a.gt_trans1[i];
      acc[0] += d0 * d0 + d1 * d1;
      if (a.g_trans0) { a.g_trans0[i] = gs_t * d0; a.g_trans1[i] = gs_t * d1; }
    }
    const float gs_b = c60 * a.w_beta * 2.f / (float)((int64_t)B * 10);
    for (int64_t i = tid; i < (int64_t)B * 10; i += nth) {
      const float b0 = a.betas0[i], b1 = a.betas1[i], c = b0 - b1;
      acc[6] += b0 * b0 + b1 * b1 + c * c;
      if (a.g_betas0) { a.g_betas0[i] = gs_b * (b0 + c); a.g_betas1[i] = gs_b * (b1 - c); }
    }
  }

  // fixed-order reduction: warp tree, then warp 0 over the 8 warp sums, fp64 partial per CTA
  __shared__ float red[kLossThreads / 32][kTerms];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < kTerms; ++t) {
    float v = acc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < kTerms) {
    double s = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) s += (double)red[w][threadIdx.x];
    k.partial[(size_t)blockIdx.x * kTerms + threadIdx.x] = s;
  }
}

__global__ void loss_final_kernel(LossK k, int nparts) {
  const airpose_twoview_loss_args& a = k.a;
  __shared__ double tot[kTerms];
  if (threadIdx.x < kTerms) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += k.partial[(size_t)i * kTerms + threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double B = a.batch;
    const double l_trans = tot[0] / (B * 3), l_kp = tot[1] / (B * 22 * 2), l_kp3d = tot[2] / (B * 22 * 3);
    const double l_shape = tot[3] / (B * a.num_verts * 3), l_root = tot[4] / (B * 9), l_pose = tot[5] / (B * 21 * 9);
    const double l_beta = tot[6] / (B * 10);
    const double loss = 60.0 * (a.w_trans * l_trans + a.w_kp2d * l_kp + a.w_kp3d * l_kp3d + a.w_shape * l_shape +
                                a.w_rootrot * l_root + a.w_pose * l_pose + a.w_beta * l_beta);
    // order of the reference's `losses` dict (:152-159)
    a.out[0] = (float)loss; a.out[1] = (float)l_trans; a.out[2] = (float)l_kp; a.out[3] = (float)l_kp3d;
    a.out[4] = (float)l_shape; a.out[5] = (float)l_root; a.out[6] = (float)l_pose; a.out[7] = (float)l_beta;
  }
}

double* g_partial[16] = {nullptr};

// ------------------------------------------------------------------------------ copenet_real's get_loss (VPoser-free part)
// /root/reference/copenet_real/src/copenet_real/copenet_twoview.py:99-160: confidence-weighted 2D keypoints over the first 22
// joints with the limb weights (:115-121), cross-view pose consistency (:137), beta regulariser + cross-view beta consistency
// (:139-141), depth barrier exp(-t_z)^2 (:148-149), x60 (:151).  A few thousand elements: ONE CTA, every term a strided loop
// with fp64 per-thread partials reduced in a fixed order (deterministic); gradients are elementwise and written in the same loops.
constexpr int kRealThreads = 256;
constexpr int kRealTerms = 4;          // keypoints, pose, betas, depth barrier

__global__ void __launch_bounds__(kRealThreads) real_loss_kernel(airpose_real_loss_args a) {
  __shared__ double red[kRealTerms][kRealThreads];
  const int B = a.batch, J = a.num_joints, JG = a.gt_joints, t = threadIdx.x;
  const bool grads = a.g_j2d0 != nullptr;
  double acc[kRealTerms] = {0.0, 0.0, 0.0, 0.0};
  const float wl = a.w_limbs2d;
  // ---- 2D keypoints: mean over B * 22 * 2 of [(p0 - g0)^2 c0 + (p1 - g1)^2 c1] * limb weight             :115-121
  {
    const int n = B * 22 * 2;
    const float gs = 60.f * a.w_kp2d * 2.f / (float)n;
    if (grads)
      for (int i = t; i < B * J * 2; i += kRealThreads) { a.g_j2d0[i] = 0.f; a.g_j2d1[i] = 0.f; }
    __syncthreads();
    for (int i = t; i < n; i += kRealThreads) {
      const int b = i / 44, j = (i % 44) >> 1, c = i & 1;
      const float lw = (j == 4 || j == 5 || j == 18 || j == 19) ? wl : ((j == 7 || j == 8 || j == 20 || j == 21) ? wl * wl : 1.f);
      const int ip = (b * J + j) * 2 + c, ig = (b * JG + j) * 3;
      const float d0 = a.j2d0[ip] - a.gt_j2d0[ig + c], d1 = a.j2d1[ip] - a.gt_j2d1[ig + c];
      const float c0 = a.gt_j2d0[ig + 2], c1 = a.gt_j2d1[ig + 2];
      acc[0] += (double)((d0 * d0 * c0 + d1 * d1 * c1) * lw);
      if (grads) { a.g_j2d0[ip] = gs * lw * c0 * d0; a.g_j2d1[ip] = gs * lw * c1 * d1; }
    }
  }
  // ---- cross-view pose consistency: mean over B * 21 * 9 of (R0 - R1)^2 on the body rotations              :137
  {
    const int n = B * 21 * 9;
    const float gs = 60.f * a.w_pose * 2.f / (float)n;
    for (int i = t; i < B * 22 * 9; i += kRealThreads) {
      const int j = (i / 9) % 22;
      float g = 0.f;
      if (j > 0) {
        const float d = a.rotmat0[i] - a.rotmat1[i];
        acc[1] += (double)(d * d);
        g = gs * d;
      }
      if (grads) { a.g_rotmat0[i] = g; a.g_rotmat1[i] = -g; }
    }
  }
  // ---- betas: mean(b0^2) + mean(b1^2) + mean((b0 - b1)^2)                                                   :139-141
  {
    const int n = B * 10;
    const float gs = 60.f * a.w_beta * 2.f / (float)n;
    for (int i = t; i < n; i += kRealThreads) {
      const float b0 = a.betas0[i], b1 = a.betas1[i], d = b0 - b1;
      acc[2] += (double)(b0 * b0 + b1 * b1 + d * d);
      if (grads) { a.g_betas0[i] = gs * (b0 + d); a.g_betas1[i] = gs * (b1 - d); }
    }
  }
  // ---- depth barrier: mean_B exp(-t_z)^2 per view                                                            :148-149
  for (int b = t; b < B; b += kRealThreads) {
    const float e0 = expf(-a.trans0[(size_t)b * a.trans_stride + 2]), e1 = expf(-a.trans1[(size_t)b * a.trans_stride + 2]);
    acc[3] += (double)(e0 * e0 + e1 * e1);
    if (grads) {
      const float gs = 60.f * -2.f / (float)B;
      a.g_trans0[b * 3] = 0.f; a.g_trans0[b * 3 + 1] = 0.f; a.g_trans0[b * 3 + 2] = gs * e0 * e0;
      a.g_trans1[b * 3] = 0.f; a.g_trans1[b * 3 + 1] = 0.f; a.g_trans1[b * 3 + 2] = gs * e1 * e1;
    }
  }
#pragma unroll
  for (int k = 0; k < kRealTerms; ++k) red[k][t] = acc[k];
  __syncthreads();
  for (int s = kRealThreads / 2; s > 0; s >>= 1) {          // fixed-order tree
    if (t < s)
#pragma unroll
      for (int k = 0; k < kRealTerms; ++k) red[k][t] += red[k][t + s];
    __syncthreads();
  }
  if (t == 0) {
    const double l_kp = red[0][0] / (B * 22 * 2), l_pose = red[1][0] / (B * 21 * 9), l_beta = red[2][0] / (B * 10);
    const double l_depth = red[3][0] / B;
    const double loss = 60.0 * (a.w_kp2d * l_kp + a.w_beta * l_beta + a.w_vposer * a.vposer_term + a.w_pose * l_pose + l_depth);
    // order of the reference's `losses` dict (:153-157)
    a.out[0] = (float)loss; a.out[1] = a.vposer_term; a.out[2] = (float)l_pose; a.out[3] = (float)l_kp; a.out[4] = (float)l_beta;
  }
}

}  // namespace
}  // namespace airpose

using namespace airpose;

extern "C" int airpose_twoview_loss(const airpose_twoview_loss_args* a, void* stream_) {
  AP_REQUIRE(a, "airpose_twoview_loss: null argument");
  AP_REQUIRE(a->batch > 0 && a->num_verts > 0 && a->num_joints >= 22, "airpose_twoview_loss: bad sizes");
  AP_REQUIRE(a->trans0 && a->trans1 && a->rotmat0 && a->rotmat1 && a->betas0 && a->betas1 && a->verts0 && a->verts1 &&
             a->joints0 && a->joints1 && a->j2d0 && a->j2d1, "airpose_twoview_loss: null prediction");
  AP_REQUIRE(a->gt_pose_rotmat && a->gt_trans0 && a->gt_trans1 && a->gt_orient0 && a->gt_orient1 && a->gt_verts &&
             a->gt_joints && a->gt_j2d0 && a->gt_j2d1 && a->out, "airpose_twoview_loss: null ground truth / output");
  const bool any_g = a->g_verts0 || a->g_verts1 || a->g_joints0 || a->g_joints1 || a->g_j2d0 || a->g_j2d1 || a->g_rotmat0 ||
                     a->g_rotmat1 || a->g_betas0 || a->g_betas1 || a->g_trans0 || a->g_trans1;
  const bool all_g = a->g_verts0 && a->g_verts1 && a->g_joints0 && a->g_joints1 && a->g_j2d0 && a->g_j2d1 && a->g_rotmat0 &&
                     a->g_rotmat1 && a->g_betas0 && a->g_betas1 && a->g_trans0 && a->g_trans1;
  AP_REQUIRE(!any_g || all_g, "airpose_twoview_loss: gradient buffers must be given all together or not at all");
  const uintptr_t al = (uintptr_t)a->verts0 | (uintptr_t)a->verts1 | (uintptr_t)a->gt_verts | (uintptr_t)a->g_verts0 |
                       (uintptr_t)a->g_verts1;
  AP_REQUIRE((al & 15) == 0, "airpose_twoview_loss: vertex tensors must be 16-byte aligned");
  int dev = 0;
  AP_CHECK_CUDA(cudaGetDevice(&dev));
  AP_REQUIRE(dev >= 0 && dev < 16, "airpose_twoview_loss: device index %d out of range", dev);
  if (!g_partial[dev]) AP_CHECK_CUDA(cudaMalloc((void**)&g_partial[dev], (size_t)kLossGrid * kTerms * sizeof(double)));
  LossK k;
  k.a = *a;
  k.partial = g_partial[dev];
  cudaStream_t st = (cudaStream_t)stream_;
  loss_partial_kernel<<<kLossGrid, kLossThreads, 0, st>>>(k);
  AP_LAUNCH_CHECK();
  loss_final_kernel<<<1, 32, 0, st>>>(k, kLossGrid);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_real_loss(const airpose_real_loss_args* a, void* stream_) {
  AP_REQUIRE(a, "airpose_real_loss: null argument");
  AP_REQUIRE(a->batch > 0 && a->num_joints >= 22 && a->gt_joints >= 22, "airpose_real_loss: bad sizes (B=%d J=%d gt J=%d)", a->batch,
             a->num_joints, a->gt_joints);
  AP_REQUIRE(a->trans0 && a->trans1 && a->rotmat0 && a->rotmat1 && a->betas0 && a->betas1 && a->j2d0 && a->j2d1 && a->gt_j2d0 &&
             a->gt_j2d1 && a->out, "airpose_real_loss: null argument");
  const bool any_g = a->g_j2d0 || a->g_j2d1 || a->g_rotmat0 || a->g_rotmat1 || a->g_betas0 || a->g_betas1 || a->g_trans0 || a->g_trans1;
  const bool all_g = a->g_j2d0 && a->g_j2d1 && a->g_rotmat0 && a->g_rotmat1 && a->g_betas0 && a->g_betas1 && a->g_trans0 && a->g_trans1;
  AP_REQUIRE(!any_g || all_g, "airpose_real_loss: gradient buffers must be given all together or not at all");
  real_loss_kernel<<<1, kRealThreads, 0, (cudaStream_t)stream_>>>(*a);
  AP_LAUNCH_CHECK();
  return 0;
}
