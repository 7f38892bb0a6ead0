// SMPL-X backward (included at the end of smplx.cu: same translation unit as the forward kernels).
//
// Gradient of SMPLX.forward(pose2rot=False) + transform_smpl + perspective_projection
// (/root/reference/copenet/src/copenet/smplx/smplx/body_models.py:820-994, lbs.py:135-222,316-370,96-132;
// utils/utils.py:237-256; utils/geometry.py:63-91) with respect to betas, the 21 body rotations, the global
// orientation and the camera transform -- what torch autograd derives for the reference in the training step
// (copenet_twoview.py:281-317) and in AirPose+ bundle adjustment (copenet_real_data/scripts/bundle_adj.py:301-401).
//
//   K0  smplx_bwd_joints_kernel   per mesh: fold d/d(joints_2d), d/d(joints_cam) into d/d(joints) and d/d(root R, t)
//   K1  smplx_pose_kernel         forward recompute of A_j, posed joints, pose feature (as in the forward call)
//   K2  smplx_vertex_bwd_kernel   CTA = 128 vertices x MB meshes.  Phase 1 (thread = vertex): recompute v_posed and T_v,
//                                 g = dL/dv (+ the extra-joint / landmark terms that gather this vertex),
//                                 X = g (x) [v_posed;1] -> smem, dL/dA_j = sum_v w_vj X_v over host-built per-tile
//                                 (joint, vertex, weight) lists in a fixed order (no atomics), h = T_v.R^T g -> smem.
//                                 Phase 2 (thread = column q of [posedirs | shapedirs] stored vertex-major): dL/dq =
//                                 sum over the tile's vertices of Pt[3v+c][q] h[v][c] -- coalesced over q.
//                                 Per-(vertex tile, mesh) partials go to a workspace: no global atomics.
//   K3  smplx_bwd_chain_kernel    per mesh: fixed-order sum of the partials, reverse pass over the kinematic chain
//                                 (children before parents), joint regressor -> betas, pose feature -> rotations.
// All fp32 on the CUDA cores: the backward batch is the training batch (32 pairs per GPU in config 4), three orders
// of magnitude below the lbs() inference config the tensor-core forward kernel exists for.

namespace airpose {

constexpr int kBwdMB = 4;       // meshes per CTA: the kernel is bound by the latency of its global loads, so more and smaller CTAs win

struct BwdJointArgs {
  int B, nj;
  const float* joints;           // [B,nj,3] forward output (canonical)
  const float* g_joints;         // [B,nj,3] or null
  const float* g_joints_cam;     // [B,nj,3] or null
  const float* g_j2d;            // [B,nj,2] or null
  const float* root_R; int root_R_stride;
  const float* root_t; int root_t_stride;
  float fx, fy;
  float* g_tot;                  // [B,nj,3] total gradient w.r.t. the canonical joints
  float* g_root_R;               // [B,9] or null
  float* g_root_t;               // [B,3] or null
};

__global__ void __launch_bounds__(128) smplx_bwd_joints_kernel(BwdJointArgs a) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  __shared__ float red[128][12];
  const int b = blockIdx.x, t = threadIdx.x;
  float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tr[3] = {0, 0, 0};
  if (a.root_R) for (int e = 0; e < 9; ++e) R[e] = a.root_R[(size_t)b * a.root_R_stride + e];
  if (a.root_t) for (int e = 0; e < 3; ++e) tr[e] = a.root_t[(size_t)b * a.root_t_stride + e];
  float acc[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) acc[e] = 0.f;
  for (int i = t; i < a.nj; i += 128) {
    const size_t o = ((size_t)b * a.nj + i) * 3;
    float g[3] = {0, 0, 0};
    if (a.g_joints) { g[0] = a.g_joints[o]; g[1] = a.g_joints[o + 1]; g[2] = a.g_joints[o + 2]; }
    if (a.g_joints_cam || a.g_j2d) {
      const float x = a.joints[o], y = a.joints[o + 1], z = a.joints[o + 2];
      const float cx = fmaf(R[0], x, fmaf(R[1], y, R[2] * z)) + tr[0];
      const float cy = fmaf(R[3], x, fmaf(R[4], y, R[5] * z)) + tr[1];
      const float cz = fmaf(R[6], x, fmaf(R[7], y, R[8] * z)) + tr[2];
      float gc[3] = {0, 0, 0};
      if (a.g_joints_cam) { gc[0] = a.g_joints_cam[o]; gc[1] = a.g_joints_cam[o + 1]; gc[2] = a.g_joints_cam[o + 2]; }
      if (a.g_j2d) {                       // u = fx cx/cz + px, v = fy cy/cz + py  (geometry.py:84-91)
        const float gu = a.g_j2d[((size_t)b * a.nj + i) * 2], gv = a.g_j2d[((size_t)b * a.nj + i) * 2 + 1];
        const float iz = 1.f / cz;
        gc[0] += a.fx * gu * iz;
        gc[1] += a.fy * gv * iz;
        gc[2] -= (a.fx * gu * cx + a.fy * gv * cy) * iz * iz;
      }
      // c = R j + t  (utils.py:237-239)
      g[0] += R[0] * gc[0] + R[3] * gc[1] + R[6] * gc[2];
      g[1] += R[1] * gc[0] + R[4] * gc[1] + R[7] * gc[2];
      g[2] += R[2] * gc[0] + R[5] * gc[1] + R[8] * gc[2];
      const float jv[3] = {x, y, z};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[r * 3 + c] += gc[r] * jv[c];
        acc[9 + r] += gc[r];
      }
    }
    a.g_tot[o] = g[0]; a.g_tot[o + 1] = g[1]; a.g_tot[o + 2] = g[2];
  }
#pragma unroll
  for (int e = 0; e < 12; ++e) red[t][e] = acc[e];
  __syncthreads();
  if (t < 12) {
    float s = 0.f;
    for (int i = 0; i < 128; ++i) s += red[i][t];       // fixed order
    if (t < 9) { if (a.g_root_R) a.g_root_R[(size_t)b * 9 + t] = s; }
    else if (a.g_root_t) a.g_root_t[(size_t)b * 3 + (t - 9)] = s;
  }
}

struct BwdVertexArgs {
  int B, nb, PF, NQ, ldq, nj;
  const float* betas; int betas_stride;
  const float* A;            // [B,J,12]
  const float* feat;         // [B,PF]
  const float* g_verts;      // [B,V,3] or null
  const float* g_jtot;       // [B,nj,3] total joint gradient (rows >= J scatter onto vertices)
  const float* Pt;           // [3V][ldq]: columns [0,P) posedirs, [P,P+NS) shapedirs
  const int* xoff;           // [V+1] CSR of the joints that gather each vertex
  const int* xj;             // joint row (>= J) ...
  const float* xw;           // ... and its weight
  const int* seg_off;        // [vtiles+1] per vertex tile: its skinning segments (one per joint that has a non-zero weight there)
  const int4* seg;           // {joint, first entry, end entry, 0}
  const int2* ent;           // {vertex in tile, weight bits}, grouped by (tile, joint), vertices ascending
  float* gA_part;            // [vtiles][B][J*12]
  float* gq_part;            // [vtiles][B][NQ]     NQ = PF + nb
  int P;                     // column of the first shape direction in Pt
};

constexpr int kBwdXStride = 13;      // 12 floats of g (x) [v_posed; 1] per vertex, odd stride: conflict-free both ways
constexpr int kBwdSegGroups = kVertsPerCta / 12;

template <int MB>
__global__ void __launch_bounds__(kVertsPerCta) smplx_vertex_bwd_kernel(SmplxDev m, BwdVertexArgs a) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  constexpr int HS = 3 * MB + 4;     // H_s row: [c][mesh] of one vertex, padded (16-byte rows, 4-way conflicts on the writes only)
  extern __shared__ __align__(16) float smem[];
  const int A_per_mesh = m.J * 12;
  float* H_s = smem;                                    // [128][HS]
  float* f_s = H_s + (size_t)kVertsPerCta * HS;         // [PF][MB]
  float* A_s = f_s + (size_t)a.PF * MB;                 // [MB][J*12]
  float* beta_s = A_s + (size_t)MB * A_per_mesh;        // [MB][kMaxShape]
  float* X_s = beta_s + MB * kMaxShape;                 // [128][13]  (one mesh at a time)
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
  for (int i = tid; i < MB * A_per_mesh; i += kVertsPerCta) {
    const int b = i / A_per_mesh;
    A_s[i] = (b < nmesh) ? __ldg(a.A + (size_t)(mesh0 + b) * A_per_mesh + (i - b * A_per_mesh)) : 0.f;
  }
  for (int i = tid; i < MB * kMaxShape; i += kVertsPerCta) {
    const int b = i / kMaxShape, l = i % kMaxShape;
    beta_s[i] = (b < nmesh && l < a.nb) ? __ldg(a.betas + (size_t)(mesh0 + b) * a.betas_stride + l) : 0.f;
  }
  // dL/dA partials of this vertex tile: joints without a weight in the tile stay zero, the others are overwritten below
  for (int i = tid; i < nmesh * A_per_mesh; i += kVertsPerCta)
    a.gA_part[((size_t)blockIdx.x * a.B + mesh0) * A_per_mesh + i] = 0.f;
  __syncthreads();

  // ---- phase 1: thread = vertex.  Pose-corrective offsets exactly as in the forward kernel.
  float acc[MB][3];
#pragma unroll
  for (int b = 0; b < MB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0.f;
  const float* Pv = m.posedirs + (size_t)vc * 3;
  const size_t prow = (size_t)m.V * 3;
  // the loop is a chain of global-load latencies unless the loads of several rows are issued together: 8 rows (24 loads) per batch
  for (int pb = 0; pb < a.PF; pb += 8) {
    float pv[8][3];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int p = min(pb + k, a.PF - 1);
      const float ok = pb + k < a.PF ? 1.f : 0.f;       // straight-line code (no branch): ptxas keeps the 24 loads together
      pv[k][0] = ok * __ldg(Pv + p * prow); pv[k][1] = ok * __ldg(Pv + p * prow + 1); pv[k][2] = ok * __ldg(Pv + p * prow + 2);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4* fr = reinterpret_cast<const float4*>(f_s + (size_t)min(pb + k, a.PF - 1) * MB);
#pragma unroll
      for (int q = 0; q < MB / 4; ++q) {
        const float4 f4 = fr[q];
        const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[q * 4 + r][0] = fmaf(fv[r], pv[k][0], acc[q * 4 + r][0]);
          acc[q * 4 + r][1] = fmaf(fv[r], pv[k][1], acc[q * 4 + r][1]);
          acc[q * 4 + r][2] = fmaf(fv[r], pv[k][2], acc[q * 4 + r][2]);
        }
      }
    }
  }
  const float vt0 = __ldg(m.v_template + vc * 3), vt1 = __ldg(m.v_template + vc * 3 + 1),
              vt2 = __ldg(m.v_template + vc * 3 + 2);
  const float* Sv = m.shapedirs + (size_t)vc * 3 * m.NS;
  const int x0 = __ldg(a.xoff + vc), x1 = __ldg(a.xoff + vc + 1);
  const int seg0 = __ldg(a.seg_off + blockIdx.x), seg1 = __ldg(a.seg_off + blockIdx.x + 1);
  const int se = tid % 12, sg = tid / 12;               // reduction role: element of A_j, segment group
#pragma unroll
  for (int b = 0; b < MB; ++b) {
    float hx = 0.f, hy = 0.f, hz = 0.f;
    float X[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) X[e] = 0.f;
    if (b < nmesh && valid) {
      float x = vt0, y = vt1, z = vt2;
      for (int l = 0; l < a.nb; ++l) {
        const float be = beta_s[b * kMaxShape + l];
        x = fmaf(__ldg(Sv + l), be, x);
        y = fmaf(__ldg(Sv + m.NS + l), be, y);
        z = fmaf(__ldg(Sv + 2 * m.NS + l), be, z);
      }
      x += acc[b][0]; y += acc[b][1]; z += acc[b][2];
      // g = dL/dv: the vertex's own gradient plus every extra joint / landmark that gathers it
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (a.g_verts) {
        const float* gp = a.g_verts + ((size_t)(mesh0 + b) * m.V + v) * 3;
        g0 = __ldg(gp); g1 = __ldg(gp + 1); g2 = __ldg(gp + 2);
      }
      for (int e = x0; e < x1; ++e) {
        const float w = __ldg(a.xw + e);
        const float* gj = a.g_jtot + ((size_t)(mesh0 + b) * a.nj + __ldg(a.xj + e)) * 3;
        g0 = fmaf(w, __ldg(gj), g0); g1 = fmaf(w, __ldg(gj + 1), g1); g2 = fmaf(w, __ldg(gj + 2), g2);
      }
      // T = sum_k w_k A_k;  h = T.R^T g;  X = g (x) [v_posed; 1]  (dL/dA_k = sum_v w_vk X_v, reduced below)
      float T[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) T[e] = 0.f;
      for (int k = 0; k < m.KW; ++k) {
        const float w = __ldg(m.skin_w + (size_t)k * m.V + vc);
        if (w == 0.f) continue;
        const int jn = __ldg(m.skin_idx + (size_t)k * m.V + vc);
        const float* Aj = A_s + (size_t)b * A_per_mesh + jn * 12;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) T[r * 3 + c] = fmaf(w, Aj[r * 4 + c], T[r * 3 + c]);
      }
      const float gv[3] = {g0, g1, g2}, vp[4] = {x, y, z, 1.f};
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) X[r * 4 + c] = gv[r] * vp[c];
      hx = T[0] * g0 + T[3] * g1 + T[6] * g2;
      hy = T[1] * g0 + T[4] * g1 + T[7] * g2;
      hz = T[2] * g0 + T[5] * g1 + T[8] * g2;
    }
    H_s[tid * HS + 0 * MB + b] = hx;
    H_s[tid * HS + 1 * MB + b] = hy;
    H_s[tid * HS + 2 * MB + b] = hz;
    if (b < nmesh) {                                    // block-uniform
      if (b > 0) __syncthreads();                       // the previous mesh's reduction has read X_s
#pragma unroll
      for (int e = 0; e < 12; ++e) X_s[tid * kBwdXStride + e] = X[e];
      __syncthreads();
      // dL/dA_j[e] of this mesh over the tile: fixed order (entries ascend by vertex), no atomics
      if (sg < kBwdSegGroups) {
        for (int s = seg0 + sg; s < seg1; s += kBwdSegGroups) {
          const int4 sd = __ldg(a.seg + s);
          float r = 0.f;
          for (int i = sd.y; i < sd.z; ++i) {
            const int2 en = __ldg(a.ent + i);
            r = fmaf(__int_as_float(en.y), X_s[en.x * kBwdXStride + se], r);
          }
          a.gA_part[((size_t)blockIdx.x * a.B + mesh0 + b) * A_per_mesh + sd.x * 12 + se] = r;
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2: thread = columns q and q + 128 of Pt (pose feature rows, then shape directions)
  const int v_base = blockIdx.x * kVertsPerCta;
  const int nv = min(kVertsPerCta, m.V - v_base);
  for (int q0 = tid; q0 < a.NQ; q0 += 2 * kVertsPerCta) {
    const int q1 = q0 + kVertsPerCta;
    const bool two = q1 < a.NQ;
    const int col0 = q0 < a.PF ? q0 : a.P + (q0 - a.PF);
    const int col1 = two ? (q1 < a.PF ? q1 : a.P + (q1 - a.PF)) : col0;
    float s0[MB], s1[MB];
#pragma unroll
    for (int b = 0; b < MB; ++b) s0[b] = s1[b] = 0.f;
    const float* pt = a.Pt + (size_t)v_base * 3 * a.ldq;
    for (int vb = 0; vb < nv; vb += 4) {              // 4 vertices = 24 loads per batch, issued together
      float pa[4][3], pb[4][3];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int vv = min(vb + k, nv - 1);
        const float ok = vb + k < nv ? 1.f : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          pa[k][c] = ok * __ldg(pt + (size_t)(vv * 3 + c) * a.ldq + col0);
          pb[k][c] = ok * __ldg(pt + (size_t)(vv * 3 + c) * a.ldq + col1);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float4* hr = reinterpret_cast<const float4*>(H_s + min(vb + k, nv - 1) * HS + c * MB);
#pragma unroll
          for (int qd = 0; qd < MB / 4; ++qd) {
            const float4 h4 = hr[qd];
            const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              s0[qd * 4 + r] = fmaf(pa[k][c], hv[r], s0[qd * 4 + r]);
              s1[qd * 4 + r] = fmaf(pb[k][c], hv[r], s1[qd * 4 + r]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < MB; ++b)
      if (b < nmesh) {
        a.gq_part[((size_t)blockIdx.x * a.B + mesh0 + b) * a.NQ + q0] = s0[b];
        if (two) a.gq_part[((size_t)blockIdx.x * a.B + mesh0 + b) * a.NQ + q1] = s1[b];
      }
  }
}

struct BwdChainArgs {
  int B, nb, PF, NQ, n_active, vtiles, nj;
  const float* betas; int betas_stride;
  const float* seg[3]; int seg_stride[3];
  const float* A;            // [B,J,12]  (rotation part = global rotations Rg_j)
  const float* gA_part; const float* gq_part;
  const float* g_jtot;       // [B,nj,3]
  float* g_betas;            // [B,nb]
  float* g_body_pose;        // [B,21,9] or null
  float* g_global_orient;    // [B,9] or null
};

__global__ void __launch_bounds__(kMaxJoints) smplx_bwd_chain_kernel(SmplxDev m, BwdChainArgs a) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  __shared__ float gRg[kMaxJoints][9], gtg[kMaxJoints][3], gJr[kMaxJoints][3];
  __shared__ float Rg[kMaxJoints][9], Rl[kMaxJoints][9], Jr[kMaxJoints][3];
  __shared__ float gR[kMaxJoints][9];
  __shared__ float beta_s[kMaxShape];
  __shared__ int par_s[kMaxJoints];
  const int b = blockIdx.x, j = threadIdx.x;
  if (j < kMaxShape) beta_s[j] = (j < a.nb) ? __ldg(a.betas + (size_t)b * a.betas_stride + j) : 0.f;
  if (j < m.J) par_s[j] = m.parents[j];
  __syncthreads();
  if (j < m.J) {
    PoseArgs pa{};
    for (int s = 0; s < 3; ++s) { pa.seg[s] = a.seg[s]; pa.seg_stride[s] = a.seg_stride[s]; }
    float R[9];
    load_rot(pa, b, j, R);
#pragma unroll
    for (int e = 0; e < 9; ++e) Rl[j][e] = R[e];
#pragma unroll
    for (int k = 0; k < 3; ++k) {                       // rest joints (lbs.py:183 through the pre-contracted regressor)
      float acc = __ldg(m.J_template + j * 3 + k);
      const float* sd = m.J_shapedirs + ((size_t)j * 3 + k) * m.NS;
      for (int l = 0; l < a.nb; ++l) acc = fmaf(__ldg(sd + l), beta_s[l], acc);
      Jr[j][k] = acc;
    }
    // fixed-order sum of the per-vertex-tile partials of dL/dA_j
    float gA[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) gA[e] = 0.f;
    {
      // four tiles (twelve 16-byte loads) in flight, added in tile order: the sums are serial chains of L2 latencies otherwise
      const size_t tstride = (size_t)a.B * m.J * 12;
      const float* src0 = a.gA_part + ((size_t)b * m.J + j) * 12;
      int t = 0;
      for (; t + 4 <= a.vtiles; t += 4) {
        float4 v[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int e = 0; e < 3; ++e) v[k][e] = __ldg(reinterpret_cast<const float4*>(src0 + (size_t)(t + k) * tstride) + e);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int e = 0; e < 3; ++e) { gA[4 * e] += v[k][e].x; gA[4 * e + 1] += v[k][e].y; gA[4 * e + 2] += v[k][e].z; gA[4 * e + 3] += v[k][e].w; }
      }
      for (; t < a.vtiles; ++t) {
        const float* src = src0 + (size_t)t * tstride;
#pragma unroll
        for (int e = 0; e < 12; ++e) gA[e] += __ldg(src + e);
      }
    }
    const float* Aj = a.A + ((size_t)b * m.J + j) * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) Rg[j][r * 3 + c] = __ldg(Aj + r * 4 + c);
    // A_j = [Rg_j | tg_j - Rg_j Jr_j],  posed joint = tg_j   (lbs.py:360-368)
    const float gAt[3] = {gA[3], gA[7], gA[11]};
    const float* gJp = a.g_jtot + ((size_t)b * a.nj + j) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      gtg[j][r] = gAt[r] + __ldg(gJp + r);
#pragma unroll
      for (int c = 0; c < 3; ++c) gRg[j][r * 3 + c] = gA[r * 4 + c] - gAt[r] * Jr[j][c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      gJr[j][c] = -(Rg[j][0 * 3 + c] * gAt[0] + Rg[j][1 * 3 + c] * gAt[1] + Rg[j][2 * 3 + c] * gAt[2]);
  }
  __syncthreads();
  // reverse pass over the chain: children (larger index) before parents; 54 sequential steps, each spread over nine lanes of
  // warp 0 (lane = element (r, c) of the 3x3 updates; one thread doing all of it was 80 us of shared-memory latencies per mesh)
  if (j < 32) {
    const int l = j, r = l / 3, c = l - r * 3;
    for (int i = m.J - 1; i >= 1; --i) {
      const int p = par_s[i];
      if (l < 3) {                                       // tg_i = Rg_p (Jr_i - Jr_p) + tg_p : lane = component
        const float u = Rg[p][0 * 3 + l] * gtg[i][0] + Rg[p][1 * 3 + l] * gtg[i][1] + Rg[p][2 * 3 + l] * gtg[i][2];
        gJr[i][l] += u;
        gJr[p][l] -= u;
        gtg[p][l] += gtg[i][l];
      }
      if (l < 9) {                                       // Rg_i = Rg_p R_i
        const float d = Jr[i][c] - Jr[p][c];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          s1 += gRg[i][r * 3 + k] * Rl[i][c * 3 + k];      // gRg_p += gRg_i R_i^T
          s2 += Rg[p][k * 3 + r] * gRg[i][k * 3 + c];      // gR_i = Rg_p^T gRg_i
        }
        float t = gRg[p][r * 3 + c];
        t += gtg[i][r] * d;
        t += s1;
        gRg[p][r * 3 + c] = t;
        gR[i][r * 3 + c] = s2;
      }
      __syncwarp();
    }
    if (l < 9) gR[0][l] = gRg[0][l];                     // Rg_0 = R_0
    if (l < 3) gJr[0][l] += gtg[0][l];                   // tg_0 = Jr_0
  }
  __syncthreads();
  // pose feature (lbs.py:197): R_i - I for i = 1..n_active, and the shape columns, summed over the vertex tiles
  for (int q = j; q < a.NQ; q += kMaxJoints) {
    float s = 0.f;
    {
      const size_t tstride = (size_t)a.B * a.NQ;
      const float* src0 = a.gq_part + (size_t)b * a.NQ + q;
      int t = 0;
      for (; t + 8 <= a.vtiles; t += 8) {               // eight loads in flight, added in tile order
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(src0 + (size_t)(t + k) * tstride);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k];
      }
      for (; t < a.vtiles; ++t) s += __ldg(src0 + (size_t)t * tstride);
    }
    if (q < a.PF) {
      gR[1 + q / 9][q % 9] += s;                              // distinct (joint, element) per q: no race
    } else {
      const int l = q - a.PF;                                 // shape blend of the vertices (lbs.py:179)
      float g = s;
      for (int i = 0; i < m.J; ++i)                           // ... and of the rest joints
#pragma unroll
        for (int k = 0; k < 3; ++k) g = fmaf(__ldg(m.J_shapedirs + ((size_t)i * 3 + k) * m.NS + l), gJr[i][k], g);
      a.g_betas[(size_t)b * a.nb + l] = g;
    }
  }
  __syncthreads();
  if (j < m.J) {
    if (j == 0 && a.g_global_orient) {
#pragma unroll
      for (int e = 0; e < 9; ++e) a.g_global_orient[(size_t)b * 9 + e] = gR[0][e];
    }
    if (j >= 1 && j <= 21 && a.g_body_pose) {
#pragma unroll
      for (int e = 0; e < 9; ++e) a.g_body_pose[((size_t)b * 21 + (j - 1)) * 9 + e] = gR[j][e];
    }
  }
}

}  // namespace airpose
