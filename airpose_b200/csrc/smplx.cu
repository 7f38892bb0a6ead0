// SMPL-X forward for the AirPose hot path: blend shapes, kinematic chain, linear blend
// skinning, extra joints + landmarks, camera transform and 2D projection.
//
// Replaces (paths relative to /root/reference/copenet/src/copenet):
//   smplx/smplx/body_models.py:820-994  SMPLX.forward(pose2rot=False)
//   smplx/smplx/lbs.py:135-222,245-266,225-242,316-370,96-132
//   smplx/smplx/vertex_joint_selector.py:73-77
//   utils/utils.py:237-256 transform_smpl, utils/geometry.py:63-91 perspective_projection
//
// Three kernels per call (DESIGN.md "SMPL-X path"):
//   1. smplx_pose_kernel    one CTA per mesh: joints from the pre-contracted regressor
//                           (J = J_template + J_shapedirs.beta, exact by linearity), the
//                           kinematic chain, A_j = G_j - [0 | G_j j_rest], pose feature.
//   2. smplx_vertex_kernel  CTA = 128 vertices x MB meshes: pose-corrective offsets as a
//                           register-tiled fp32 contraction, shape blend, sparse skinning
//                           from smem-resident A, optional camera transform.
//   3. smplx_joints_kernel  one CTA per mesh: 55 + 21 + 51 joints, camera transform,
//                           perspective projection.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"
#include "smplx.cuh"

namespace airpose {

constexpr int kVertsPerCta = 128;
constexpr int kMeshTile = 16;

struct PoseArgs {
  int B, nb, n_active;       // n_active = joints 1..n_active feed the pose feature
  const float* betas; int betas_stride;
  const float* seg[3]; int seg_stride[3];   // joint 0 | joints 1..21 | joints 22..J-1
  float* A;                  // [B,J,12]  (generic vertex kernel) or null
  float* Jt;                 // [B,J,3]
  float* feat;               // [B,PF]    (generic vertex kernel) or null
  // tensor-core vertex kernel (smplx_tc.cu): per-mesh record + split-fp16 pose feature
  float* rec;                // [Bpad, kTcRecFloats] or null
  int rec_plain;             // 1: one plain row per mesh (smplx_ml.cu); 0: pairs of meshes element-interleaved (smplx_tc.cu)
  __half* fh; __half* fl;    // [B, kTcK]
  const float* transl;
  const float* root_R; int root_R_stride;
  const float* root_t; int root_t_stride;
};

__device__ __forceinline__ void load_rot(const PoseArgs& a, int b, int j, float R[9]) {
  const float* src = nullptr;
  if (j == 0) {
    if (a.seg[0]) src = a.seg[0] + (size_t)b * a.seg_stride[0];
  } else if (j < 22) {
    if (a.seg[1]) src = a.seg[1] + (size_t)b * a.seg_stride[1] + (j - 1) * 9;
  } else {
    if (a.seg[2]) src = a.seg[2] + (size_t)b * a.seg_stride[2] + (j - 22) * 9;
  }
#pragma unroll
  for (int e = 0; e < 9; ++e) R[e] = src ? __ldg(src + e) : ((e % 4 == 0) ? 1.f : 0.f);
}

// lbs.py:316-370 batch_rigid_transform, one mesh per CTA, one joint per thread.
__global__ void __launch_bounds__(kMaxJoints) smplx_pose_kernel(SmplxDev m, PoseArgs a) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  __shared__ float Ts[kMaxJoints][12];
  __shared__ float Js[kMaxJoints][3];
  __shared__ float beta_s[kMaxShape];
  __shared__ int par_s[kMaxJoints];
  const int b = blockIdx.x, j = threadIdx.x;
  if (j < kMaxShape) beta_s[j] = (j < a.nb) ? __ldg(a.betas + (size_t)b * a.betas_stride + j) : 0.f;
  if (j < m.J) par_s[j] = m.parents[j];
  __syncthreads();
  float R[9];
  if (j < m.J) {
    load_rot(a, b, j, R);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float acc = __ldg(m.J_template + j * 3 + k);
      const float* sd = m.J_shapedirs + ((size_t)j * 3 + k) * m.NS;
      for (int l = 0; l < a.nb; ++l) acc = fmaf(__ldg(sd + l), beta_s[l], acc);
      Js[j][k] = acc;
    }
  }
  __syncthreads();
  if (j < m.J) {
    const int p = par_s[j];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Ts[j][r * 4 + 0] = R[r * 3 + 0];
      Ts[j][r * 4 + 1] = R[r * 3 + 1];
      Ts[j][r * 4 + 2] = R[r * 3 + 2];
      Ts[j][r * 4 + 3] = Js[j][r] - (p >= 0 ? Js[p][r] : 0.f);   // rel_joints (:343-345)
    }
  }
  __syncthreads();
  if (j >= m.J) return;
  // G_j = T_root ... T_parent T_j, accumulated from the leaf upwards.
  float G[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) G[e] = Ts[j][e];
  for (int p = par_s[j]; p >= 0; p = par_s[p]) {
    float N[12];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float t0 = Ts[p][r * 4 + 0], t1 = Ts[p][r * 4 + 1], t2 = Ts[p][r * 4 + 2];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v = t0 * G[0 * 4 + c];
        v = fmaf(t1, G[1 * 4 + c], v);
        v = fmaf(t2, G[2 * 4 + c], v);
        if (c == 3) v += Ts[p][r * 4 + 3];
        N[r * 4 + c] = v;
      }
    }
#pragma unroll
    for (int e = 0; e < 12; ++e) G[e] = N[e];
  }
  float* Jt = a.Jt + ((size_t)b * m.J + j) * 3;
  Jt[0] = G[3]; Jt[1] = G[7]; Jt[2] = G[11];                    // posed joints (:360)
  float Aj[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {                                  // rel_transforms (:365-368)
    const float gj = G[r * 4 + 0] * Js[j][0] + G[r * 4 + 1] * Js[j][1] + G[r * 4 + 2] * Js[j][2];
    Aj[r * 4 + 0] = G[r * 4 + 0];
    Aj[r * 4 + 1] = G[r * 4 + 1];
    Aj[r * 4 + 2] = G[r * 4 + 2];
    Aj[r * 4 + 3] = G[r * 4 + 3] - gj;
  }
  if (a.A) {
    float* A = a.A + ((size_t)b * m.J + j) * 12;
#pragma unroll
    for (int e = 0; e < 12; ++e) A[e] = Aj[e];
  }
  if (a.feat && j >= 1 && j <= a.n_active) {                     // pose_feature (lbs.py:197)
    float* f = a.feat + (size_t)b * (a.n_active * 9) + (j - 1) * 9;
#pragma unroll
    for (int e = 0; e < 9; ++e) f[e] = R[e] - ((e % 4 == 0) ? 1.f : 0.f);
  }
  if (a.rec) {
    // records are stored in PAIRS of meshes, element-interleaved (field f of mesh b at pair_base + 2 f + (b & 1)), so
    // that the vertex kernel reads {mesh 2p, mesh 2p+1} of a field with one 64-bit load and does its fp32 math with
    // packed FFMA2 (two meshes per instruction)
    const int rs = a.rec_plain ? 1 : 2;                          // element stride inside the record
    float* rec = a.rec_plain ? a.rec + (size_t)b * kTcRecFloats : a.rec + (size_t)(b >> 1) * 2 * kTcRecFloats + (b & 1);
    if (j < kTcBodyJoints) {
#pragma unroll
      for (int e = 0; e < 12; ++e) rec[rs * (j * 12 + e)] = Aj[e];
    }
    if (j >= 1 && j < kTcBodyJoints) {                           // split-fp16 pose feature: f = hi + lo
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        const float f = R[e] - ((e % 4 == 0) ? 1.f : 0.f);
        const __half hi = __float2half_rn(f);
        a.fh[(size_t)b * kTcK + (j - 1) * 9 + e] = hi;
        a.fl[(size_t)b * kTcK + (j - 1) * 9 + e] = __float2half_rn(f - __half2float(hi));
      }
    }
    if (j < kTcKPose - 189) {                                    // zero the K padding of the pose blocks
      a.fh[(size_t)b * kTcK + 189 + j] = __float2half_rn(0.f);
      a.fl[(size_t)b * kTcK + 189 + j] = __float2half_rn(0.f);
    }
    if (j < kTcKShape) {                                         // template / shape k-block (smplx.cuh): 1 1 | b_hi | b_lo | b_hi
      float val = 1.f;
      if (j >= 2) {
        const int l = (j - 2) % 10;
        const float be = l < a.nb ? beta_s[l] : 0.f;
        const float hi = __half2float(__float2half_rn(be));
        val = (j >= 12 && j < 22) ? be - hi : hi;
      }
      a.fh[(size_t)b * kTcK + kTcKPose + j] = __float2half_rn(val);
      a.fl[(size_t)b * kTcK + kTcKPose + j] = __float2half_rn(0.f);
    }
    if (j >= 12 && j < 21) rec[rs * (kTcRecCam + (j - 12))] = a.root_R ? __ldg(a.root_R + (size_t)b * a.root_R_stride + (j - 12))
                                                                        : (((j - 12) % 4 == 0) ? 1.f : 0.f);
    if (j >= 21 && j < 24) rec[rs * (kTcRecCam + 9 + (j - 21))] = a.root_t ? __ldg(a.root_t + (size_t)b * a.root_t_stride + (j - 21)) : 0.f;
    if (j >= 24 && j < 28) rec[rs * (kTcRecTransl + (j - 24))] = (a.transl && j < 27) ? __ldg(a.transl + (size_t)b * 3 + (j - 24)) : 0.f;
  }
}

struct VertexArgs {
  int B, nb, PF;
  const float* betas; int betas_stride;
  const float* A;            // [B,J,12]
  const float* feat;         // [B,PF]
  const float* transl;       // [B,3] or null
  const float* root_R; int root_R_stride;
  const float* root_t; int root_t_stride;
  float* out;                // [B,V,3]
  float* out_cam;            // [B,V,3] or null
};

template <int MB>
__global__ void __launch_bounds__(kVertsPerCta) smplx_vertex_kernel(SmplxDev m, VertexArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* f_s = smem;                                   // [PF][MB]
  float* A_s = f_s + (size_t)a.PF * MB;                // [MB][J*12]
  float* beta_s = A_s + (size_t)MB * m.J * 12;         // [MB][kMaxShape]
  float* cam_s = beta_s + MB * kMaxShape;              // [MB][16]: R(9) t(3) transl(3)
  const int tid = threadIdx.x;
  const int v = blockIdx.x * kVertsPerCta + tid;
  const int mesh0 = blockIdx.y * MB;
  const int nmesh = min(MB, a.B - mesh0);
  const bool valid = v < m.V;
  const int vc = valid ? v : m.V - 1;

  for (int i = tid; i < a.PF * MB; i += kVertsPerCta) {
    const int p = i / MB, b = i % MB;
    f_s[i] = (b < nmesh) ? __ldg(a.feat + (size_t)(mesh0 + b) * a.PF + p) : 0.f;
  }
  const int A_per_mesh = m.J * 12;
  for (int i = tid; i < MB * A_per_mesh; i += kVertsPerCta) {
    const int b = i / A_per_mesh;
    A_s[i] = (b < nmesh) ? __ldg(a.A + (size_t)(mesh0 + b) * A_per_mesh + (i - b * A_per_mesh)) : 0.f;
  }
  for (int i = tid; i < MB * kMaxShape; i += kVertsPerCta) {
    const int b = i / kMaxShape, l = i % kMaxShape;
    beta_s[i] = (b < nmesh && l < a.nb) ? __ldg(a.betas + (size_t)(mesh0 + b) * a.betas_stride + l) : 0.f;
  }
  for (int i = tid; i < MB * 16; i += kVertsPerCta) {
    const int b = i / 16, e = i % 16;
    float val = 0.f;
    if (b < nmesh) {
      if (e < 9) val = a.root_R ? __ldg(a.root_R + (size_t)(mesh0 + b) * a.root_R_stride + e) : ((e % 4 == 0) ? 1.f : 0.f);
      else if (e < 12) val = a.root_t ? __ldg(a.root_t + (size_t)(mesh0 + b) * a.root_t_stride + (e - 9)) : 0.f;
      else if (e < 15) val = a.transl ? __ldg(a.transl + (size_t)(mesh0 + b) * 3 + (e - 12)) : 0.f;
    }
    cam_s[i] = val;
  }
  __syncthreads();

  // pose-corrective offsets: acc[b][k] = sum_p feat[b][p] * posedirs[p][3v+k]   (lbs.py:200-201)
  float acc[MB][3];
#pragma unroll
  for (int b = 0; b < MB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0.f;
  const float* Pv = m.posedirs + (size_t)vc * 3;
  const size_t prow = (size_t)m.V * 3;
#pragma unroll 2
  for (int p = 0; p < a.PF; ++p) {
    const float p0 = __ldg(Pv + p * prow), p1 = __ldg(Pv + p * prow + 1), p2 = __ldg(Pv + p * prow + 2);
    const float4* fr = reinterpret_cast<const float4*>(f_s + (size_t)p * MB);
#pragma unroll
    for (int q = 0; q < MB / 4; ++q) {
      const float4 f4 = fr[q];
      const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[q * 4 + r][0] = fmaf(fv[r], p0, acc[q * 4 + r][0]);
        acc[q * 4 + r][1] = fmaf(fv[r], p1, acc[q * 4 + r][1]);
        acc[q * 4 + r][2] = fmaf(fv[r], p2, acc[q * 4 + r][2]);
      }
    }
  }

  const float vt0 = __ldg(m.v_template + vc * 3), vt1 = __ldg(m.v_template + vc * 3 + 1),
              vt2 = __ldg(m.v_template + vc * 3 + 2);
  const float* Sv = m.shapedirs + (size_t)vc * 3 * m.NS;
#pragma unroll
  for (int b = 0; b < MB; ++b) {
    if (b >= nmesh) break;
    // v_shaped + pose offsets (lbs.py:179,203)
    float x = vt0, y = vt1, z = vt2;
    for (int l = 0; l < a.nb; ++l) {
      const float be = beta_s[b * kMaxShape + l];
      x = fmaf(__ldg(Sv + l), be, x);
      y = fmaf(__ldg(Sv + m.NS + l), be, y);
      z = fmaf(__ldg(Sv + 2 * m.NS + l), be, z);
    }
    x += acc[b][0]; y += acc[b][1]; z += acc[b][2];
    // T = sum_j w_vj A_j over the non-zero weights (lbs.py:209-213)
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    for (int k = 0; k < m.KW; ++k) {
      const float w = __ldg(m.skin_w + (size_t)k * m.V + vc);
      const int jn = __ldg(m.skin_idx + (size_t)k * m.V + vc);
      const float4* Aj = reinterpret_cast<const float4*>(A_s + (size_t)b * A_per_mesh + jn * 12);
      const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
      T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]); T[3] = fmaf(w, r0.w, T[3]);
      T[4] = fmaf(w, r1.x, T[4]); T[5] = fmaf(w, r1.y, T[5]); T[6] = fmaf(w, r1.z, T[6]); T[7] = fmaf(w, r1.w, T[7]);
      T[8] = fmaf(w, r2.x, T[8]); T[9] = fmaf(w, r2.y, T[9]); T[10] = fmaf(w, r2.z, T[10]); T[11] = fmaf(w, r2.w, T[11]);
    }
    const float* cs = cam_s + b * 16;
    // v = T [v_posed; 1]  (lbs.py:215-220), then += transl (body_models.py:980-982)
    const float ox = fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[3]))) + cs[12];
    const float oy = fmaf(T[4], x, fmaf(T[5], y, fmaf(T[6], z, T[7]))) + cs[13];
    const float oz = fmaf(T[8], x, fmaf(T[9], y, fmaf(T[10], z, T[11]))) + cs[14];
    if (valid) {
      float* o = a.out + ((size_t)(mesh0 + b) * m.V + v) * 3;
      o[0] = ox; o[1] = oy; o[2] = oz;
      if (a.out_cam) {   // transform_smpl (utils.py:237-239): R v + t about the origin
        float* oc = a.out_cam + ((size_t)(mesh0 + b) * m.V + v) * 3;
        oc[0] = fmaf(cs[0], ox, fmaf(cs[1], oy, cs[2] * oz)) + cs[9];
        oc[1] = fmaf(cs[3], ox, fmaf(cs[4], oy, cs[5] * oz)) + cs[10];
        oc[2] = fmaf(cs[6], ox, fmaf(cs[7], oy, cs[8] * oz)) + cs[11];
      }
    }
  }
}

struct JointArgs {
  int B;
  const float* verts;        // [B,V,3] (transl already applied)
  const float* Jt;           // [B,J,3]
  const float* transl;
  const float* root_R; int root_R_stride;
  const float* root_t; int root_t_stride;
  float fx, fy;
  const float* center; int center_stride;
  const float* proj_t; int proj_t_stride;
  float* joints; float* joints_cam; float* j2d;
};

__global__ void __launch_bounds__(128) smplx_joints_kernel(SmplxDev m, JointArgs a) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int b = blockIdx.x;
  const int nj = m.J + m.E + m.L;
  const float* vb = a.verts + (size_t)b * m.V * 3;
  for (int i = threadIdx.x; i < nj; i += blockDim.x) {
    float x, y, z;
    if (i < m.J) {
      const float* s = a.Jt + ((size_t)b * m.J + i) * 3;
      x = s[0]; y = s[1]; z = s[2];
      if (a.transl) { x += a.transl[b * 3]; y += a.transl[b * 3 + 1]; z += a.transl[b * 3 + 2]; }
    } else if (i < m.J + m.E) {                       // vertex_joint_selector.py:73-77 (pure gather)
      const float* s = vb + (size_t)m.extra_idx[i - m.J] * 3;
      x = s[0]; y = s[1]; z = s[2];
    } else {                                          // vertices2landmarks (lbs.py:96-132)
      const int l = i - m.J - m.E;
      x = y = z = 0.f;
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const float w = m.lmk_bary[l * 3 + f];
        const float* s = vb + (size_t)m.lmk_vidx[l * 3 + f] * 3;
        x = fmaf(s[0], w, x); y = fmaf(s[1], w, y); z = fmaf(s[2], w, z);
      }
    }
    float* o = a.joints + ((size_t)b * nj + i) * 3;
    o[0] = x; o[1] = y; o[2] = z;
    if (a.joints_cam || a.j2d) {
      float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
      if (a.root_R) for (int e = 0; e < 9; ++e) R[e] = a.root_R[(size_t)b * a.root_R_stride + e];
      if (a.root_t) for (int e = 0; e < 3; ++e) t[e] = a.root_t[(size_t)b * a.root_t_stride + e];
      const float cx = fmaf(R[0], x, fmaf(R[1], y, R[2] * z)) + t[0];
      const float cy = fmaf(R[3], x, fmaf(R[4], y, R[5] * z)) + t[1];
      const float cz = fmaf(R[6], x, fmaf(R[7], y, R[8] * z)) + t[2];
      if (a.joints_cam) {
        float* oc = a.joints_cam + ((size_t)b * nj + i) * 3;
        oc[0] = cx; oc[1] = cy; oc[2] = cz;
      }
      if (a.j2d) {   // perspective_projection (geometry.py:63-91) with rotation = I, translation = proj_t (or 0)
        const float px = a.center ? a.center[(size_t)b * a.center_stride] : 0.f;
        const float py = a.center ? a.center[(size_t)b * a.center_stride + 1] : 0.f;
        float qx = cx, qy = cy, qz = cz;
        if (a.proj_t) {
          const float* pt = a.proj_t + (size_t)b * a.proj_t_stride;
          qx += pt[0]; qy += pt[1]; qz += pt[2];
        }
        float* o2 = a.j2d + ((size_t)b * nj + i) * 2;
        o2[0] = fmaf(a.fx, qx / qz, px);
        o2[1] = fmaf(a.fy, qy / qz, py);
      }
    }
  }
}

__global__ void rot6d_kernel(const float* __restrict__ x, int64_t groups, int per_group, int64_t row_stride,
                             float* __restrict__ R) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * per_group) return;
  const float* s = x + (i / per_group) * row_stride + (i % per_group) * 6;
  // reshape(-1,3,2): a1 = elements (0,2,4), a2 = (1,3,5)   (geometry.py:55-57)
  const float a1x = s[0], a1y = s[2], a1z = s[4], a2x = s[1], a2y = s[3], a2z = s[5];
  const float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);   // F.normalize eps
  const float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  const float d = b1x * a2x + b1y * a2y + b1z * a2z;
  const float ux = a2x - d * b1x, uy = a2y - d * b1y, uz = a2z - d * b1z;
  const float n2 = fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-12f);
  const float b2x = ux / n2, b2y = uy / n2, b2z = uz / n2;
  const float b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
  float* o = R + i * 9;          // columns (b1,b2,b3)  (geometry.py:61)
  o[0] = b1x; o[1] = b2x; o[2] = b3x;
  o[3] = b1y; o[4] = b2y; o[5] = b3y;
  o[6] = b1z; o[7] = b2z; o[8] = b3z;
}

// Backward of rot6d_to_rotmat (geometry.py:47-61): gR [n,3,3] -> gx (6 per rotation, same strided layout as x).
__global__ void rot6d_bwd_kernel(const float* __restrict__ x, int64_t groups, int per_group, int64_t row_stride,
                                 const float* __restrict__ gR, float* __restrict__ gx, int64_t gx_row_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * per_group) return;
  const float* s = x + (i / per_group) * row_stride + (i % per_group) * 6;
  const float a1[3] = {s[0], s[2], s[4]}, a2[3] = {s[1], s[3], s[5]};
  const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
  const float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
  const float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  const float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
  const float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
  const float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
  const float* g = gR + i * 9;
  float gb1[3] = {g[0], g[3], g[6]}, gb2[3] = {g[1], g[4], g[7]};
  const float gb3[3] = {g[2], g[5], g[8]};
  // b3 = b1 x b2:  gb1 += b2 x gb3,  gb2 += gb3 x b1
  gb1[0] += b2[1] * gb3[2] - b2[2] * gb3[1]; gb1[1] += b2[2] * gb3[0] - b2[0] * gb3[2]; gb1[2] += b2[0] * gb3[1] - b2[1] * gb3[0];
  gb2[0] += gb3[1] * b1[2] - gb3[2] * b1[1]; gb2[1] += gb3[2] * b1[0] - gb3[0] * b1[2]; gb2[2] += gb3[0] * b1[1] - gb3[1] * b1[0];
  // b2 = u / |u|
  const float p2 = b2[0] * gb2[0] + b2[1] * gb2[1] + b2[2] * gb2[2];
  const float gu[3] = {(gb2[0] - b2[0] * p2) / n2, (gb2[1] - b2[1] * p2) / n2, (gb2[2] - b2[2] * p2) / n2};
  // u = a2 - (b1.a2) b1
  const float q = gu[0] * b1[0] + gu[1] * b1[1] + gu[2] * b1[2];
  const float ga2[3] = {gu[0] - q * b1[0], gu[1] - q * b1[1], gu[2] - q * b1[2]};
#pragma unroll
  for (int k = 0; k < 3; ++k) gb1[k] -= q * a2[k] + d * gu[k];
  // b1 = a1 / |a1|
  const float p1 = b1[0] * gb1[0] + b1[1] * gb1[1] + b1[2] * gb1[2];
  float* o = gx + (i / per_group) * gx_row_stride + (i % per_group) * 6;
  o[0] = (gb1[0] - b1[0] * p1) / n1; o[2] = (gb1[1] - b1[1] * p1) / n1; o[4] = (gb1[2] - b1[2] * p1) / n1;
  o[1] = ga2[0]; o[3] = ga2[1]; o[5] = ga2[2];
}

__global__ void j14_gather_kernel(const float* __restrict__ joints, int B, int nj, const int* __restrict__ map,
                                  int nmap, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nmap * 3) return;
  const int c = i % 3, k = (i / 3) % nmap, b = i / (3 * nmap);
  out[i] = joints[((size_t)b * nj + map[k]) * 3 + c];
}

}  // namespace airpose

#include "smplx_bwd.inl"

using namespace airpose;

struct airpose_smplx {
  int device = 0;
  SmplxDev d{};
  SmplxTc tc;
  std::vector<void*> owned;
  float* ws = nullptr;
  size_t ws_floats = 0;
  // backward (smplx_bwd.inl)
  float* Pt = nullptr;         // [3V][ldq]: posedirs (P columns) then shapedirs (NS columns), vertex-major
  int ldq = 0;
  int* xoff = nullptr;         // [V+1] CSR: which extra joints / landmarks gather each vertex ...
  int* xj = nullptr;           // ... their row in the joints output ...
  float* xw = nullptr;         // ... and the weight (1 for the vertex-picked joints, barycentric for landmarks)
  int* seg_off = nullptr;      // per 128-vertex tile: skinning weights regrouped by joint (smplx_vertex_bwd_kernel) ...
  int* seg = nullptr;          // ... {joint, first entry, end entry, 0} ...
  int* ent = nullptr;          // ... {vertex in tile, weight bits}
  float* bws = nullptr;        // backward workspace
  size_t bws_floats = 0;
};

extern "C" int airpose_smplx_create(airpose_smplx_t** out, const airpose_smplx_model_host* mh, int device) {
  AP_REQUIRE(out && mh, "airpose_smplx_create: null argument");
  const int V = mh->num_verts, J = mh->num_joints, NS = mh->num_shape, P = mh->num_pose_basis;
  AP_REQUIRE(J >= 1 && J <= kMaxJoints, "airpose_smplx_create: num_joints %d out of range [1,%d]", J, kMaxJoints);
  AP_REQUIRE(NS >= 0 && NS <= kMaxShape, "airpose_smplx_create: num_shape %d > %d", NS, kMaxShape);
  AP_REQUIRE(P == (J - 1) * 9, "airpose_smplx_create: num_pose_basis %d != (J-1)*9", P);
  AP_REQUIRE(mh->parents[0] < 0, "airpose_smplx_create: parents[0] must be -1");
  for (int j = 1; j < J; ++j)
    AP_REQUIRE(mh->parents[j] >= 0 && mh->parents[j] < j, "airpose_smplx_create: parents[%d]=%lld is not < %d", j,
               (long long)mh->parents[j], j);
  AP_CHECK_CUDA(cudaSetDevice(device));
  auto* h = new airpose_smplx();
  h->device = device;
  SmplxDev& d = h->d;
  d.V = V; d.J = J; d.NS = NS; d.P = P; d.L = mh->num_landmarks; d.E = mh->num_extra;

  // Joint regressor contracted with the template and the shape basis (exact by linearity of
  // lbs.py:179-183): J = J_regressor (v_template + shapedirs beta).
  std::vector<float> Jt((size_t)J * 3), Js((size_t)J * 3 * NS);
  for (int j = 0; j < J; ++j) {
    double t[3] = {0, 0, 0};
    std::vector<double> s((size_t)3 * NS, 0.0);
    for (int v = 0; v < V; ++v) {
      const double w = mh->J_regressor[(size_t)j * V + v];
      if (w == 0.0) continue;
      for (int k = 0; k < 3; ++k) {
        t[k] += w * mh->v_template[(size_t)v * 3 + k];
        for (int l = 0; l < NS; ++l) s[(size_t)k * NS + l] += w * mh->shapedirs[((size_t)v * 3 + k) * NS + l];
      }
    }
    for (int k = 0; k < 3; ++k) {
      Jt[j * 3 + k] = (float)t[k];
      for (int l = 0; l < NS; ++l) Js[((size_t)j * 3 + k) * NS + l] = (float)s[(size_t)k * NS + l];
    }
  }
  // Skinning weights in ELL form: exactly the non-zeros of lbs_weights, widest row decides KW.
  int KW = 1;
  for (int v = 0; v < V; ++v) {
    int n = 0;
    for (int j = 0; j < J; ++j) n += mh->lbs_weights[(size_t)v * J + j] != 0.f;
    KW = std::max(KW, n);
  }
  d.KW = KW;
  std::vector<int> sidx((size_t)KW * V, 0);
  std::vector<float> sw((size_t)KW * V, 0.f);
  for (int v = 0; v < V; ++v) {
    int n = 0;
    for (int j = 0; j < J; ++j) {
      const float w = mh->lbs_weights[(size_t)v * J + j];
      if (w != 0.f) { sidx[(size_t)n * V + v] = j; sw[(size_t)n * V + v] = w; ++n; }
    }
  }
  std::vector<int> par(J), lv((size_t)d.L * 3), ex(d.E);
  for (int j = 0; j < J; ++j) par[j] = (int)mh->parents[j];
  for (int l = 0; l < d.L; ++l) {
    const int64_t f = mh->lmk_faces_idx[l];
    AP_REQUIRE(f >= 0 && f < mh->num_faces, "airpose_smplx_create: lmk_faces_idx[%d] out of range", l);
    for (int k = 0; k < 3; ++k) {
      const int64_t vi = mh->faces[f * 3 + k];
      AP_REQUIRE(vi >= 0 && vi < V, "airpose_smplx_create: face vertex id out of range");
      lv[l * 3 + k] = (int)vi;
    }
  }
  for (int e = 0; e < d.E; ++e) {
    AP_REQUIRE(mh->extra_joint_idx[e] >= 0 && mh->extra_joint_idx[e] < V, "airpose_smplx_create: extra joint id out of range");
    ex[e] = (int)mh->extra_joint_idx[e];
  }

  float* fp; int* ip;
#define UP_F(field, src, n) do { if (device_upload(&fp, (const float*)(src), (size_t)(n))) return 1; d.field = fp; h->owned.push_back(fp); } while (0)
#define UP_I(field, src, n) do { if (device_upload(&ip, (const int*)(src), (size_t)(n))) return 1; d.field = ip; h->owned.push_back(ip); } while (0)
  UP_F(v_template, mh->v_template, (size_t)V * 3);
  UP_F(shapedirs, mh->shapedirs, (size_t)V * 3 * NS);
  UP_F(posedirs, mh->posedirs, (size_t)P * V * 3);
  UP_F(J_template, Jt.data(), Jt.size());
  UP_F(J_shapedirs, Js.data(), std::max<size_t>(Js.size(), 1));
  UP_I(parents, par.data(), par.size());
  UP_I(skin_idx, sidx.data(), sidx.size());
  UP_F(skin_w, sw.data(), sw.size());
  UP_I(lmk_vidx, lv.data(), std::max<size_t>(lv.size(), 1));
  UP_F(lmk_bary, mh->lmk_bary_coords, std::max<size_t>((size_t)d.L * 3, 1));
  UP_I(extra_idx, ex.data(), std::max<size_t>(ex.size(), 1));
#undef UP_F
#undef UP_I
  if (smplx_tc_create(mh, d, &h->tc, &h->owned)) return 1;
  if (smplx_ml_create(mh, d, &h->tc, &h->owned)) return 1;
  {  // backward constants: [posedirs | shapedirs] vertex-major, and the vertex -> gathered-joint lists
    h->ldq = (P + NS + 31) / 32 * 32;
    std::vector<float> Pt((size_t)V * 3 * h->ldq, 0.f);
    for (int p = 0; p < P; ++p) {
      const float* src = mh->posedirs + (size_t)p * V * 3;
      for (size_t r = 0; r < (size_t)V * 3; ++r) Pt[r * h->ldq + p] = src[r];
    }
    for (size_t r = 0; r < (size_t)V * 3; ++r)
      for (int l = 0; l < NS; ++l) Pt[r * h->ldq + P + l] = mh->shapedirs[r * NS + l];
    if (device_upload(&h->Pt, Pt.data(), Pt.size())) return 1;
    h->owned.push_back(h->Pt);
    std::vector<std::vector<std::pair<int, float>>> lists(V);
    for (int e = 0; e < d.E; ++e) lists[ex[e]].push_back({J + e, 1.f});
    for (int l = 0; l < d.L; ++l)
      for (int k = 0; k < 3; ++k) lists[lv[l * 3 + k]].push_back({J + d.E + l, mh->lmk_bary_coords[l * 3 + k]});
    std::vector<int> xoff(V + 1, 0), xj;
    std::vector<float> xw;
    for (int v = 0; v < V; ++v) {
      for (auto& e : lists[v]) { xj.push_back(e.first); xw.push_back(e.second); }
      xoff[v + 1] = (int)xj.size();
    }
    if (xj.empty()) { xj.push_back(0); xw.push_back(0.f); }
    if (device_upload(&h->xoff, xoff.data(), xoff.size())) return 1;
    h->owned.push_back(h->xoff);
    if (device_upload(&h->xj, xj.data(), xj.size())) return 1;
    h->owned.push_back(h->xj);
    if (device_upload(&h->xw, xw.data(), xw.size())) return 1;
    h->owned.push_back(h->xw);
    // skinning weights regrouped per vertex tile by joint: dL/dA_j of a tile is a fixed-order sum over these lists
    const int vtiles = ceil_div(V, kVertsPerCta);
    std::vector<int> seg_off(vtiles + 1, 0), seg, ent;
    for (int t = 0; t < vtiles; ++t) {
      const int v0 = t * kVertsPerCta, v1 = std::min(V, v0 + kVertsPerCta);
      for (int j = 0; j < J; ++j) {
        const int first = (int)ent.size() / 2;
        for (int v = v0; v < v1; ++v) {
          const float w = mh->lbs_weights[(size_t)v * J + j];
          if (w == 0.f) continue;
          int bits; std::memcpy(&bits, &w, sizeof(bits));
          ent.push_back(v - v0); ent.push_back(bits);
        }
        const int end = (int)ent.size() / 2;
        if (end > first) { seg.push_back(j); seg.push_back(first); seg.push_back(end); seg.push_back(0); }
      }
      seg_off[t + 1] = (int)seg.size() / 4;
    }
    if (device_upload(&h->seg_off, seg_off.data(), seg_off.size())) return 1;
    h->owned.push_back(h->seg_off);
    if (device_upload(&h->seg, seg.data(), seg.size())) return 1;
    h->owned.push_back(h->seg);
    if (device_upload(&h->ent, ent.data(), ent.size())) return 1;
    h->owned.push_back(h->ent);
  }
  *out = h;
  return 0;
}

extern "C" int airpose_smplx_destroy(airpose_smplx_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (void* p : h->owned) cudaFree(p);
  cudaFree(h->ws);
  cudaFree(h->bws);
  delete h;
  return 0;
}

extern "C" int airpose_smplx_skin_nnz(const airpose_smplx_t* h) { return h ? h->d.KW : -1; }

extern "C" int airpose_smplx_fwd(airpose_smplx_t* h, const airpose_smplx_fwd_args* g, void* stream_) {
  AP_REQUIRE(h && g, "airpose_smplx_fwd: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const SmplxDev& d = h->d;
  const int B = g->batch;
  AP_REQUIRE(B >= 0, "airpose_smplx_fwd: negative batch");
  if (B == 0) return 0;
  AP_REQUIRE(g->betas && g->out_vertices && g->out_joints, "airpose_smplx_fwd: betas/out_vertices/out_joints are required");
  AP_REQUIRE(g->num_betas >= 0 && g->num_betas <= d.NS, "airpose_smplx_fwd: num_betas %d > shapedirs columns %d", g->num_betas, d.NS);
  AP_REQUIRE(!g->tail_pose || d.J > 22, "airpose_smplx_fwd: tail_pose given but the model has %d joints", d.J);
  AP_REQUIRE(d.J == 55 || (!g->body_pose && !g->tail_pose) || d.J >= 22, "airpose_smplx_fwd: unsupported joint count %d", d.J);
  const int n_active = g->tail_pose ? d.J - 1 : (g->body_pose ? std::min(21, d.J - 1) : 0);
  const int PF = n_active * 9;

  // tensor-core vertex kernel for the hot-path call pattern (smplx_tc.cu), generic fp32 kernel otherwise
  const bool use_tc = h->tc.ok && g->body_pose && !g->tail_pose && g->num_betas <= 10 && d.J == 55;
  const int Bpad = ceil_div(B, kTcMeshTile) * kTcMeshTile;
  const size_t n_jt = ((size_t)B * d.J * 3 + 3) & ~size_t(3);    // keeps the record block 16-byte aligned
  const size_t need = use_tc ? n_jt + (size_t)Bpad * kTcRecFloats + (size_t)B * kTcK      // Jt | records | fh+fl (2 x B x 192 halves)
                             : n_jt + (size_t)B * d.J * 12 + (size_t)B * std::max(PF, 1);   // Jt | A | feat
  if (need > h->ws_floats) {
    AP_CHECK_CUDA(cudaStreamSynchronize(stream));
    cudaFree(h->ws);
    h->ws = nullptr; h->ws_floats = 0;
    AP_CHECK_CUDA(cudaMalloc((void**)&h->ws, need * sizeof(float)));
    h->ws_floats = need;
  }
  float* Jt = h->ws;
  float* A = use_tc ? nullptr : Jt + n_jt;
  float* feat = use_tc ? nullptr : A + (size_t)B * d.J * 12;
  float* rec = use_tc ? Jt + n_jt : nullptr;
  __half* fh = use_tc ? reinterpret_cast<__half*>(rec + (size_t)Bpad * kTcRecFloats) : nullptr;
  __half* fl = use_tc ? fh + (size_t)B * kTcK : nullptr;

  PoseArgs pa{};
  pa.B = B; pa.nb = g->num_betas; pa.n_active = n_active;
  pa.betas = g->betas; pa.betas_stride = g->betas_stride;
  pa.seg[0] = g->global_orient; pa.seg_stride[0] = g->global_orient_stride;
  pa.seg[1] = g->body_pose; pa.seg_stride[1] = g->body_pose_stride;
  pa.seg[2] = g->tail_pose; pa.seg_stride[2] = g->tail_pose_stride;
  pa.A = A; pa.Jt = Jt; pa.feat = feat;
  const bool use_ml = use_tc && h->tc.ml_ok && B >= kMlMinBatch && ((int64_t)B + 128) * d.V * 3 < (int64_t)1 << 31;   // 32-bit element offsets
  pa.rec = rec; pa.fh = fh; pa.fl = fl; pa.rec_plain = use_ml;
  pa.transl = g->transl;
  pa.root_R = g->root_R; pa.root_R_stride = g->root_R_stride;
  pa.root_t = g->root_t; pa.root_t_stride = g->root_t_stride;
  AP_CHECK_CUDA(launch_chain(smplx_pose_kernel, dim3(B), dim3(kMaxJoints), stream, d, pa));

  if (use_tc) {
    TcCall tcall{};
    tcall.B = B; tcall.nb = g->num_betas; tcall.has_transl = g->transl != nullptr;
    tcall.rec = rec; tcall.fh = fh; tcall.fl = fl;
    tcall.out = g->out_vertices; tcall.out_cam = g->out_vertices_cam;
    if (use_ml ? smplx_ml_forward(d, h->tc, tcall, stream) : smplx_tc_forward(d, h->tc, tcall, stream)) return 1;
  } else {
    VertexArgs va{};
    va.B = B; va.nb = g->num_betas; va.PF = PF;
    va.betas = g->betas; va.betas_stride = g->betas_stride;
    va.A = A; va.feat = feat; va.transl = g->transl;
    va.root_R = g->root_R; va.root_R_stride = g->root_R_stride;
    va.root_t = g->root_t; va.root_t_stride = g->root_t_stride;
    va.out = g->out_vertices; va.out_cam = g->out_vertices_cam;
    constexpr int MB = kMeshTile;
    const size_t smem = ((size_t)PF * MB + (size_t)MB * d.J * 12 + MB * kMaxShape + MB * 16) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      AP_CHECK_CUDA(cudaFuncSetAttribute(smplx_vertex_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_set = true;
    }
    AP_REQUIRE(smem <= 160 * 1024, "airpose_smplx_fwd: shared memory %zu too large", smem);
    dim3 grid(ceil_div(d.V, kVertsPerCta), ceil_div(B, MB));
    smplx_vertex_kernel<MB><<<grid, kVertsPerCta, smem, stream>>>(d, va);
    AP_LAUNCH_CHECK();
  }

  JointArgs ja{};
  ja.B = B; ja.verts = g->out_vertices; ja.Jt = Jt; ja.transl = g->transl;
  ja.root_R = g->root_R; ja.root_R_stride = g->root_R_stride;
  ja.root_t = g->root_t; ja.root_t_stride = g->root_t_stride;
  ja.fx = g->focal_x; ja.fy = g->focal_y; ja.center = g->center; ja.center_stride = g->center_stride;
  ja.proj_t = g->proj_t; ja.proj_t_stride = g->proj_t_stride;
  ja.joints = g->out_joints; ja.joints_cam = g->out_joints_cam; ja.j2d = g->out_joints_2d;
  AP_CHECK_CUDA(launch_chain(smplx_joints_kernel, dim3(B), dim3(128), stream, d, ja));
  return 0;
}

extern "C" int airpose_rot6d_to_rotmat_strided(const float* x, int64_t groups, int32_t per_group, int64_t row_stride,
                                               float* R, void* stream) {
  AP_REQUIRE(x && R && groups >= 0 && per_group > 0, "airpose_rot6d_to_rotmat: bad argument");
  const int64_t n = groups * per_group;
  if (n == 0) return 0;
  rot6d_kernel<<<(unsigned)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(x, groups, per_group, row_stride, R);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_rot6d_to_rotmat_bwd_strided(const float* x, int64_t groups, int32_t per_group, int64_t row_stride,
                                                   const float* grad_R, float* grad_x, int64_t grad_x_row_stride, void* stream) {
  AP_REQUIRE(x && grad_R && grad_x && groups >= 0 && per_group > 0, "airpose_rot6d_to_rotmat_bwd: bad argument");
  const int64_t n = groups * per_group;
  if (n == 0) return 0;
  rot6d_bwd_kernel<<<(unsigned)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(x, groups, per_group, row_stride, grad_R,
                                                                                    grad_x, grad_x_row_stride);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_rot6d_to_rotmat(const float* x, int64_t n, float* R, void* stream) {
  return airpose_rot6d_to_rotmat_strided(x, n, 1, 6, R, stream);
}

extern "C" int airpose_j14_gather(const float* joints, int32_t batch, int32_t num_joints, const int32_t* map_host,
                                  float* out, void* stream) {
  AP_REQUIRE(joints && out && batch >= 0 && num_joints > 0, "airpose_j14_gather: bad argument");
  static const int32_t ref_map[14] = {15, 12, 17, 19, 21, 16, 18, 20, 2, 5, 8, 1, 4, 7};
  const int32_t* mp = map_host ? map_host : ref_map;
  for (int i = 0; i < 14; ++i) AP_REQUIRE(mp[i] >= 0 && mp[i] < num_joints, "airpose_j14_gather: map[%d]=%d out of range", i, mp[i]);
  if (batch == 0) return 0;
  // the map is tiny: pass it through a per-call device buffer on the stream
  int* dmap = nullptr;
  AP_CHECK_CUDA(cudaMallocAsync((void**)&dmap, 14 * sizeof(int), (cudaStream_t)stream));
  AP_CHECK_CUDA(cudaMemcpyAsync(dmap, mp, 14 * sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  j14_gather_kernel<<<ceil_div(batch * 14 * 3, 128), 128, 0, (cudaStream_t)stream>>>(joints, batch, num_joints, dmap, 14, out);
  AP_LAUNCH_CHECK();
  AP_CHECK_CUDA(cudaFreeAsync(dmap, (cudaStream_t)stream));
  return 0;
}

extern "C" int airpose_smplx_bwd(airpose_smplx_t* h, const airpose_smplx_bwd_args* g, void* stream_) {
  AP_REQUIRE(h && g, "airpose_smplx_bwd: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const SmplxDev& d = h->d;
  const int B = g->batch;
  AP_REQUIRE(B >= 0, "airpose_smplx_bwd: negative batch");
  if (B == 0) return 0;
  AP_REQUIRE(g->betas && g->grad_betas, "airpose_smplx_bwd: betas / grad_betas are required");
  AP_REQUIRE(g->num_betas >= 1 && g->num_betas <= d.NS, "airpose_smplx_bwd: num_betas %d out of range", g->num_betas);
  AP_REQUIRE(g->grad_vertices || g->grad_joints || g->grad_joints_cam || g->grad_joints_2d, "airpose_smplx_bwd: no upstream gradient");
  AP_REQUIRE(!(g->grad_joints_cam || g->grad_joints_2d) || g->joints, "airpose_smplx_bwd: the forward joints are needed for camera-frame gradients");
  AP_REQUIRE(d.J >= 22, "airpose_smplx_bwd: unsupported joint count %d", d.J);
  const int n_active = g->body_pose ? 21 : 0;
  const int PF = n_active * 9, NQ = PF + g->num_betas;
  const int nj = d.J + d.E + d.L;
  const int vtiles = ceil_div(d.V, kVertsPerCta);
  auto al4 = [](size_t n) { return (n + 3) & ~size_t(3); };
  const size_t n_jt = al4((size_t)B * d.J * 3), n_A = al4((size_t)B * d.J * 12), n_feat = al4((size_t)B * std::max(PF, 1)),
               n_gj = al4((size_t)B * nj * 3), n_gA = al4((size_t)vtiles * B * d.J * 12), n_gq = al4((size_t)vtiles * B * NQ);
  const size_t need = n_jt + n_A + n_feat + n_gj + n_gA + n_gq;
  if (need > h->bws_floats) {
    AP_CHECK_CUDA(cudaStreamSynchronize(stream));
    cudaFree(h->bws);
    h->bws = nullptr; h->bws_floats = 0;
    AP_CHECK_CUDA(cudaMalloc((void**)&h->bws, need * sizeof(float)));
    h->bws_floats = need;
  }
  float* Jt = h->bws;
  float* A = Jt + n_jt;
  float* feat = A + n_A;
  float* gjt = feat + n_feat;
  float* gA_part = gjt + n_gj;
  float* gq_part = gA_part + n_gA;

  BwdJointArgs ja{};
  ja.B = B; ja.nj = nj; ja.joints = g->joints;
  ja.g_joints = g->grad_joints; ja.g_joints_cam = g->grad_joints_cam; ja.g_j2d = g->grad_joints_2d;
  ja.root_R = g->root_R; ja.root_R_stride = g->root_R_stride; ja.root_t = g->root_t; ja.root_t_stride = g->root_t_stride;
  ja.fx = g->focal_x; ja.fy = g->focal_y;
  ja.g_tot = gjt; ja.g_root_R = g->grad_root_R; ja.g_root_t = g->grad_root_t;
  AP_CHECK_CUDA(launch_chain(smplx_bwd_joints_kernel, dim3(B), dim3(128), stream, ja));

  PoseArgs pa{};
  pa.B = B; pa.nb = g->num_betas; pa.n_active = n_active;
  pa.betas = g->betas; pa.betas_stride = g->betas_stride;
  pa.seg[0] = g->global_orient; pa.seg_stride[0] = g->global_orient_stride;
  pa.seg[1] = g->body_pose; pa.seg_stride[1] = g->body_pose_stride;
  pa.A = A; pa.Jt = Jt; pa.feat = feat;
  AP_CHECK_CUDA(launch_chain(smplx_pose_kernel, dim3(B), dim3(kMaxJoints), stream, d, pa));

  BwdVertexArgs va{};
  va.B = B; va.nb = g->num_betas; va.PF = PF; va.NQ = NQ; va.ldq = h->ldq; va.nj = nj; va.P = d.P;
  va.betas = g->betas; va.betas_stride = g->betas_stride;
  va.A = A; va.feat = feat; va.g_verts = g->grad_vertices; va.g_jtot = gjt;
  va.Pt = h->Pt; va.xoff = h->xoff; va.xj = h->xj; va.xw = h->xw;
  va.seg_off = h->seg_off; va.seg = reinterpret_cast<const int4*>(h->seg); va.ent = reinterpret_cast<const int2*>(h->ent);
  va.gA_part = gA_part; va.gq_part = gq_part;
  {
    constexpr int MB = kBwdMB;
    const size_t smem = ((size_t)kVertsPerCta * (3 * MB + 4) + (size_t)PF * MB + (size_t)MB * d.J * 12 + MB * kMaxShape +
                         (size_t)kVertsPerCta * kBwdXStride) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      AP_CHECK_CUDA(cudaFuncSetAttribute(smplx_vertex_bwd_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_set = true;
    }
    AP_REQUIRE(smem <= 160 * 1024, "airpose_smplx_bwd: shared memory %zu too large", smem);
    dim3 grid(vtiles, ceil_div(B, MB));
    AP_CHECK_CUDA(launch_chain_smem(smplx_vertex_bwd_kernel<MB>, grid, dim3(kVertsPerCta), smem, stream, d, va));
  }

  BwdChainArgs ca{};
  ca.B = B; ca.nb = g->num_betas; ca.PF = PF; ca.NQ = NQ; ca.n_active = n_active; ca.vtiles = vtiles; ca.nj = nj;
  ca.betas = g->betas; ca.betas_stride = g->betas_stride;
  ca.seg[0] = g->global_orient; ca.seg_stride[0] = g->global_orient_stride;
  ca.seg[1] = g->body_pose; ca.seg_stride[1] = g->body_pose_stride;
  ca.seg[2] = nullptr; ca.seg_stride[2] = 0;
  ca.A = A; ca.gA_part = gA_part; ca.gq_part = gq_part; ca.g_jtot = gjt;
  ca.g_betas = g->grad_betas; ca.g_body_pose = g->grad_body_pose; ca.g_global_orient = g->grad_global_orient;
  AP_CHECK_CUDA(launch_chain(smplx_bwd_chain_kernel, dim3(B), dim3(kMaxJoints), stream, d, ca));
  return 0;
}
