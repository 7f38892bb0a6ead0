// Training-mode IEF regressor of copenet: forward with dropout and the backward pass.
//
// Replaces, for the training step, copenet.forward's regressor loop and forward_reg
// (/root/reference/copenet/src/copenet/models/model_copenet.py:118-159,178-204) as autograd differentiates them:
// per iteration and view   z = [xf | bb, position, orient, art_pose, shape | other view's art_pose, shape]   (:185,:192)
//                          h1 = drop1(fc1 z)   h2 = drop2(fc2 h1)   state += [decpose; decshape] h2        (:186-202)
// with weights shared by the two views and the iterations.  With dropout active the eval-mode collapse of ief.cu does
// not apply, so the three Linears run as fp32 GEMMs (both views stacked: 2B rows).  The dropout masks are INPUTS
// (multiplicative, 0 or 1/(1-p)): the caller draws them (torch.bernoulli on the device), which is also what makes the
// parity tests against autograd exact.  The whole regressor is 7.2 MFLOP per row and iteration -- tiny next to the trunk --
// so a plain shared-memory-tiled SGEMM on the CUDA cores is enough; parameter gradients are fp32, accumulated over the
// iterations in a fixed order (no atomics).
//
// This is also the whole trainable part of the reference's `train_reg_only` fine-tuning mode
// (copenet_real/src/copenet_real/copenet_twoview.py:357-372).
#include <algorithm>
#include "common.cuh"
#include "ptx.cuh"

namespace airpose {
namespace {

constexpr int kF = 2048, kU = 284, kZ = kF + kU, kH = 1024, kD = 145;   // feature, state part of z, fc1 in, hidden, decoded

// C[M,N] = alpha * sum_k A(m,k) B(k,n) + beta * C[M,N];  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
// 64 x 64 tiles, 4 x 4 outputs per thread (two 16-byte shared loads per 16 FMAs), k tiles of 16 with the next tile's global
// loads in registers while the current one is multiplied.  The products have 2B <= 128 rows: parallelism comes from split-K.
constexpr int kTM = 64, kTN = 64, kTK = 16, kTP = kTM + 4;      // row pitch 68 floats: 16-byte aligned, 2-way conflicts at worst
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                    const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                                    float* __restrict__ C, int64_t ldc, int M, int N, int K, float alpha, float beta,
                                                    int k_per_split, float* __restrict__ partial) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  __shared__ __align__(16) float As[2][kTK][kTP], Bs[2][kTK][kTP];
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;        // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // split-K (gridDim.z > 1): this CTA contracts k in [z * k_per_split, ...) and writes its partial tile to partial[z][M][N];
  // splitk_reduce_kernel adds the partials in a fixed order (no atomics: bitwise reproducible)
  const int k_begin = blockIdx.z * k_per_split;
  K = min(K, k_begin + k_per_split);
  // element e of a 64 x 16 tile handled by this thread: 4 per operand; the faster index follows whichever stride is 1
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = threadIdx.x + e * 256;
    if (sak == 1) { ak[e] = i % kTK; am[e] = i / kTK; } else { am[e] = i % kTM; ak[e] = i / kTM; }
    if (sbk == 1) { bk[e] = i % kTK; bn[e] = i / kTK; } else { bn[e] = i % kTN; bk[e] = i / kTN; }
  }
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gm = m0 + am[e], gka = k0 + ak[e], gn = n0 + bn[e], gkb = k0 + bk[e];
      ra[e] = (gm < M && gka < K) ? __ldg(A + gm * sam + gka * sak) : 0.f;
      rb[e] = (gn < N && gkb < K) ? __ldg(B + gkb * sbk + gn * sbn) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { As[buf][ak[e]][am[e]] = ra[e]; Bs[buf][bk[e]][bn[e]] = rb[e]; }
  };
  if (k_begin < K) {
    fetch(k_begin);
    stash(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < K; k0 += kTK) {
    const bool more = k0 + kTK < K;
    if (more) fetch(k0 + kTK);
#pragma unroll
    for (int k = 0; k < kTK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) {
        if (gridDim.z > 1) {
          partial[((size_t)blockIdx.z * M + gm) * N + gn] = acc[i][j];
        } else {
          float* c = C + gm * ldc + gn;
          *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
        }
      }
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, float* __restrict__ C, int64_t ldc, int M, int N, float beta) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  float acc = 0.f;
  for (int z = 0; z < splits; ++z) acc += partial[(size_t)z * M * N + i];
  float* c = C + (i / N) * ldc + (i % N);
  *c = acc + (beta != 0.f ? beta * *c : 0.f);
}

constexpr int64_t kSplitKFloats = (int64_t)4 << 20;     // partial-tile workspace at the end of the caller's workspace (16 MB)

int sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc, int M, int N,
          int K, float beta, cudaStream_t st, float* splitk_ws = nullptr) {
  dim3 grid(ceil_div(N, kTN), ceil_div(M, kTM));
  // the skinny products (2B rows against a 1024 x 2332 weight) have a few dozen tiles and a long serial k loop: cut k
  int splits = 1;
  if (splitk_ws) {
    const int tiles = (int)(grid.x * grid.y);
    splits = std::min(std::min(32, 592 / std::max(tiles, 1)), K / (4 * kTK));
    while (splits > 1 && (int64_t)splits * M * N > kSplitKFloats) --splits;
    splits = std::max(splits, 1);
  }
  int k_per = K;
  if (splits > 1) {
    k_per = ceil_div(ceil_div(K, splits), kTK) * kTK;
    splits = ceil_div(K, k_per);
    grid.z = splits;
  }
  AP_CHECK_CUDA(launch_chain(sgemm_kernel, grid, dim3(256), st, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, 1.f, beta, k_per, splitk_ws));
  if (splits > 1) {
    AP_CHECK_CUDA(launch_chain(splitk_reduce_kernel, dim3((unsigned)ceil_div64((int64_t)M * N, 256)), dim3(256), st, splitk_ws, splits, C, ldc, M, N, beta));
  }
  return 0;
}

// state rows: [2B][145] = pose(135) | shape(10); row m = view * B + b.  z rows: [2B][2332].
__global__ void ief_init_state_kernel(int B, const float* pos0, const float* pos1, const float* th0, const float* th1, int th_stride,
                                      const float* sh0, const float* sh1, int sh_stride, const float* init_pose,
                                      const float* init_shape, float* state) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * kD) return;
  const int m = i / kD, o = i % kD, v = m / B, b = m % B;
  float x;
  if (o < 3) x = (v ? pos1 : pos0)[b * 3 + o];
  else if (o < 135) { const float* th = v ? th1 : th0; x = th ? th[(size_t)b * th_stride + (o - 3)] : init_pose[o - 3]; }
  else { const float* sh = v ? sh1 : sh0; x = sh ? sh[(size_t)b * sh_stride + (o - 135)] : init_shape[o - 135]; }
  state[i] = x;
}

__global__ void ief_assemble_kernel(int B, const float* __restrict__ xf0, const float* __restrict__ xf1, const float* __restrict__ bb0,
                                    const float* __restrict__ bb1, const float* __restrict__ state, float* __restrict__ z) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)2 * B * kZ) return;
  const int m = (int)(i / kZ), k = (int)(i % kZ), v = m / B, b = m % B;
  float x;
  if (k < kF) x = (v ? xf1 : xf0)[(size_t)b * kF + k];
  else {
    const int u = k - kF;                                  // [bb(3) | pose_self(135) | shape_self(10) | art_other(126) | shape_other(10)]
    const float* self = state + (size_t)m * kD;
    const float* other = state + (size_t)((1 - v) * B + b) * kD;
    if (u < 3) x = (v ? bb1 : bb0)[b * 3 + u];
    else if (u < 148) x = self[u - 3];
    else if (u < 274) x = other[9 + (u - 148)];
    else x = other[135 + (u - 274)];
  }
  z[i] = x;
}

// h = (h + bias) * mask   (mask may be null = eval)
__global__ void bias_mask_kernel(float* __restrict__ h, const float* __restrict__ bias, const float* __restrict__ mask, int rows, int cols) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const float x = h[i] + bias[i % cols];
  h[i] = mask ? x * mask[i] : x;
}
__global__ void mul_mask_kernel(float* __restrict__ g, const float* __restrict__ mask, int n) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && mask) g[i] *= mask[i];
}
// state += d + bdec
__global__ void state_update_kernel(float* __restrict__ state, const float* __restrict__ d, const float* __restrict__ bdec, int rows) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * kD) state[i] += d[i] + bdec[i % kD];
}
// out[c] (+)= sum_r g[r][c], rows summed in order (deterministic)
__global__ void colsum_kernel(const float* __restrict__ g, int rows, int cols, float* __restrict__ out, float beta) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += g[(size_t)r * cols + c];
  out[c] = s + (beta != 0.f ? out[c] : 0.f);
}
// gradient of the assembled input back onto the previous iteration's states (the identity path state_new = state_old + d
// is already in g_state): g_state[self] += gz[u-part of self], g_state[other] += gz[cross part]
__global__ void scatter_gu_kernel(int B, const float* __restrict__ gz, float* __restrict__ g_state) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * kD) return;
  const int m = i / kD, o = i % kD, v = m / B, b = m % B;
  const float* gs = gz + (size_t)m * kZ + kF;                       // own row: bb(3) | pose_self | shape_self
  const float* go = gz + (size_t)((1 - v) * B + b) * kZ + kF;       // the other view's row reads our art_pose / shape
  float g = gs[3 + o];                                              // pose (o < 135) and shape (135..144) of self
  if (o >= 9 && o < 135) g += go[148 + (o - 9)];
  else if (o >= 135) g += go[274 + (o - 135)];
  g_state[i] += g;
}
__global__ void copy_gxf_kernel(int B, const float* __restrict__ gz, float* __restrict__ g_xf0, float* __restrict__ g_xf1, float beta) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)2 * B * kF) return;
  const int m = (int)(i / kF), k = (int)(i % kF), v = m / B, b = m % B;
  float* dst = (v ? g_xf1 : g_xf0) + (size_t)b * kF + k;
  *dst = gz[(size_t)m * kZ + k] + (beta != 0.f ? *dst : 0.f);
}

inline unsigned blocks(int64_t n) { return (unsigned)ceil_div64(n, 256); }

}  // namespace
}  // namespace airpose

using namespace airpose;

extern "C" int64_t airpose_ief_train_saved_floats(int32_t batch, int32_t iters) {
  return (int64_t)iters * 2 * batch * (kZ + kH + kH);
}
extern "C" int64_t airpose_ief_train_workspace_floats(int32_t batch) {
  // state, g_state, d / gd, gz, gh (x2), wdec, bdec, g_wdec, g_bdec
  return (int64_t)2 * batch * (kD + kD + kD + kZ + kH + kH) + (int64_t)kD * kH * 2 + 2 * kD + 64 + kSplitKFloats;
}

static int check_common(const airpose_ief_train_args* a, const char* who) {
  AP_REQUIRE(a, "%s: null argument", who);
  AP_REQUIRE(a->batch > 0 && a->iters >= 1, "%s: bad batch/iters", who);
  AP_REQUIRE(a->xf0 && a->xf1 && a->bb0 && a->bb1 && a->pos0 && a->pos1, "%s: null input", who);
  AP_REQUIRE(a->fc1_w && a->fc1_b && a->fc2_w && a->fc2_b && a->decpose_w && a->decpose_b && a->decshape_w && a->decshape_b &&
             a->init_pose && a->init_shape, "%s: null parameter", who);
  AP_REQUIRE((a->mask1 == nullptr) == (a->mask2 == nullptr), "%s: give both dropout masks or neither", who);
  AP_REQUIRE(a->saved && a->workspace, "%s: saved / workspace buffers are required", who);
  AP_REQUIRE((a->init_theta0 == nullptr) == (a->init_theta1 == nullptr) && (a->init_shape0 == nullptr) == (a->init_shape1 == nullptr),
             "%s: init_theta / init_shape must be given for both views or neither", who);
  return 0;
}

struct Ws {
  float *state, *g_state, *d, *gz, *gh1, *gh2, *wdec, *bdec, *g_wdec, *g_bdec, *splitk;
};
static Ws carve(float* w, int B) {
  Ws s;
  const size_t R = (size_t)2 * B;
  s.state = w; w += R * kD;
  s.g_state = w; w += R * kD;
  s.d = w; w += R * kD;
  s.gz = w; w += R * kZ;
  s.gh1 = w; w += R * kH;
  s.gh2 = w; w += R * kH;
  s.wdec = w; w += (size_t)kD * kH;
  s.g_wdec = w; w += (size_t)kD * kH;
  s.bdec = w; w += kD;
  s.g_bdec = w; w += kD;
  s.splitk = w + ((16 - ((uintptr_t)w / 4) % 16) % 16);      // 64-byte aligned; the size includes 64 floats of slack
  return s;
}
static int load_dec(const airpose_ief_train_args* a, const Ws& s, cudaStream_t st) {
  AP_CHECK_CUDA(cudaMemcpyAsync(s.wdec, a->decpose_w, (size_t)135 * kH * 4, cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(s.wdec + (size_t)135 * kH, a->decshape_w, (size_t)10 * kH * 4, cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(s.bdec, a->decpose_b, 135 * 4, cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(s.bdec + 135, a->decshape_b, 10 * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int airpose_ief_train_fwd(const airpose_ief_train_args* a, void* stream_) {
  if (check_common(a, "airpose_ief_train_fwd")) return 2;
  AP_REQUIRE(a->out_pose0 && a->out_pose1 && a->out_betas0 && a->out_betas1, "airpose_ief_train_fwd: null output");
  cudaStream_t st = (cudaStream_t)stream_;
  const int B = a->batch, R = 2 * B;
  const Ws s = carve(a->workspace, B);
  if (load_dec(a, s, st)) return 1;
  AP_CHECK_CUDA(launch_chain(ief_init_state_kernel, dim3(blocks((int64_t)R * kD)), dim3(256), st, B, a->pos0, a->pos1, a->init_theta0, a->init_theta1, a->init_theta_stride,
                                                                a->init_shape0, a->init_shape1, a->init_shape_stride, a->init_pose,
                                                                a->init_shape, s.state));
  const size_t per_it = (size_t)R * (kZ + kH + kH);
  for (int it = 0; it < a->iters; ++it) {
    float* z = a->saved + it * per_it;
    float* h1 = z + (size_t)R * kZ;
    float* h2 = h1 + (size_t)R * kH;
    const float* m1 = a->mask1 ? a->mask1 + (size_t)it * R * kH : nullptr;
    const float* m2 = a->mask2 ? a->mask2 + (size_t)it * R * kH : nullptr;
    AP_CHECK_CUDA(launch_chain(ief_assemble_kernel, dim3(blocks((int64_t)R * kZ)), dim3(256), st, B, a->xf0, a->xf1, a->bb0, a->bb1, s.state, z));
    if (sgemm(z, kZ, 1, a->fc1_w, 1, kZ, h1, kH, R, kH, kZ, 0.f, st, s.splitk)) return 1;                 // h1 = z W1^T
    AP_CHECK_CUDA(launch_chain(bias_mask_kernel, dim3(blocks((int64_t)R * kH)), dim3(256), st, h1, a->fc1_b, m1, R, kH));
    if (sgemm(h1, kH, 1, a->fc2_w, 1, kH, h2, kH, R, kH, kH, 0.f, st, s.splitk)) return 1;                // h2 = h1 W2^T
    AP_CHECK_CUDA(launch_chain(bias_mask_kernel, dim3(blocks((int64_t)R * kH)), dim3(256), st, h2, a->fc2_b, m2, R, kH));
    if (sgemm(h2, kH, 1, s.wdec, 1, kH, s.d, kD, R, kD, kH, 0.f, st, s.splitk)) return 1;                 // d = h2 Wdec^T
    AP_CHECK_CUDA(launch_chain(state_update_kernel, dim3(blocks((int64_t)R * kD)), dim3(256), st, s.state, s.d, s.bdec, R));
  }
  // outputs: pose [B,135], betas [B,10] per view
  for (int v = 0; v < 2; ++v) {
    float* pose = v ? a->out_pose1 : a->out_pose0;
    float* betas = v ? a->out_betas1 : a->out_betas0;
    AP_CHECK_CUDA(cudaMemcpy2DAsync(pose, 135 * 4, s.state + (size_t)v * B * kD, kD * 4, 135 * 4, B, cudaMemcpyDeviceToDevice, st));
    AP_CHECK_CUDA(cudaMemcpy2DAsync(betas, 10 * 4, s.state + (size_t)v * B * kD + 135, kD * 4, 10 * 4, B, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

extern "C" int airpose_ief_train_bwd(const airpose_ief_train_args* a, void* stream_) {
  if (check_common(a, "airpose_ief_train_bwd")) return 2;
  AP_REQUIRE(a->g_pose0 && a->g_pose1 && a->g_betas0 && a->g_betas1, "airpose_ief_train_bwd: null upstream gradient");
  AP_REQUIRE(a->g_fc1_w && a->g_fc1_b && a->g_fc2_w && a->g_fc2_b && a->g_decpose_w && a->g_decpose_b && a->g_decshape_w &&
             a->g_decshape_b, "airpose_ief_train_bwd: null parameter-gradient buffer");
  AP_REQUIRE((a->g_xf0 == nullptr) == (a->g_xf1 == nullptr), "airpose_ief_train_bwd: give both feature-gradient buffers or neither");
  cudaStream_t st = (cudaStream_t)stream_;
  const int B = a->batch, R = 2 * B;
  const Ws s = carve(a->workspace, B);
  if (load_dec(a, s, st)) return 1;
  // g_state = upstream gradient of the final state
  for (int v = 0; v < 2; ++v) {
    const float* gp = v ? a->g_pose1 : a->g_pose0;
    const float* gb = v ? a->g_betas1 : a->g_betas0;
    AP_CHECK_CUDA(cudaMemcpy2DAsync(s.g_state + (size_t)v * B * kD, kD * 4, gp, 135 * 4, 135 * 4, B, cudaMemcpyDeviceToDevice, st));
    AP_CHECK_CUDA(cudaMemcpy2DAsync(s.g_state + (size_t)v * B * kD + 135, kD * 4, gb, 10 * 4, 10 * 4, B, cudaMemcpyDeviceToDevice, st));
  }
  const size_t per_it = (size_t)R * (kZ + kH + kH);
  for (int it = a->iters - 1; it >= 0; --it) {
    const float beta = (it == a->iters - 1) ? 0.f : 1.f;          // first visit overwrites, later ones accumulate
    const float* z = a->saved + it * per_it;
    const float* h1 = z + (size_t)R * kZ;
    const float* h2 = h1 + (size_t)R * kH;
    const float* m1 = a->mask1 ? a->mask1 + (size_t)it * R * kH : nullptr;
    const float* m2 = a->mask2 ? a->mask2 + (size_t)it * R * kH : nullptr;
    const float* gd = s.g_state;                                   // state_new = state_old + d  =>  dL/dd = dL/dstate_new
    // decoders: g_wdec += gd^T h2, g_bdec += colsum(gd), gh2 = (gd Wdec) * m2
    if (sgemm(gd, 1, kD, h2, kH, 1, s.g_wdec, kH, kD, kH, R, beta, st, s.splitk)) return 1;
    AP_CHECK_CUDA(launch_chain(colsum_kernel, dim3(ceil_div(kD, 128)), dim3(128), st, gd, R, kD, s.g_bdec, beta));
    if (sgemm(gd, kD, 1, s.wdec, kH, 1, s.gh2, kH, R, kH, kD, 0.f, st, s.splitk)) return 1;
    AP_CHECK_CUDA(launch_chain(mul_mask_kernel, dim3(blocks((int64_t)R * kH)), dim3(256), st, s.gh2, m2, R * kH));
    // fc2
    if (sgemm(s.gh2, 1, kH, h1, kH, 1, a->g_fc2_w, kH, kH, kH, R, beta, st, s.splitk)) return 1;
    AP_CHECK_CUDA(launch_chain(colsum_kernel, dim3(ceil_div(kH, 128)), dim3(128), st, s.gh2, R, kH, a->g_fc2_b, beta));
    if (sgemm(s.gh2, kH, 1, a->fc2_w, kH, 1, s.gh1, kH, R, kH, kH, 0.f, st, s.splitk)) return 1;
    AP_CHECK_CUDA(launch_chain(mul_mask_kernel, dim3(blocks((int64_t)R * kH)), dim3(256), st, s.gh1, m1, R * kH));
    // fc1
    if (sgemm(s.gh1, 1, kH, z, kZ, 1, a->g_fc1_w, kZ, kH, kZ, R, beta, st, s.splitk)) return 1;
    AP_CHECK_CUDA(launch_chain(colsum_kernel, dim3(ceil_div(kH, 128)), dim3(128), st, s.gh1, R, kH, a->g_fc1_b, beta));
    if (sgemm(s.gh1, kH, 1, a->fc1_w, kZ, 1, s.gz, kZ, R, kZ, kH, 0.f, st, s.splitk)) return 1;
    if (a->g_xf0) {
      AP_CHECK_CUDA(launch_chain(copy_gxf_kernel, dim3(blocks((int64_t)R * kF)), dim3(256), st, B, s.gz, a->g_xf0, a->g_xf1, beta));
    }
    if (it > 0) {                                                  // the first iteration's state is the constant initialisation
      AP_CHECK_CUDA(launch_chain(scatter_gu_kernel, dim3(blocks((int64_t)R * kD)), dim3(256), st, B, s.gz, s.g_state));
    }
  }
  // split the concatenated decoder gradient
  AP_CHECK_CUDA(cudaMemcpyAsync(a->g_decpose_w, s.g_wdec, (size_t)135 * kH * 4, cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(a->g_decshape_w, s.g_wdec + (size_t)135 * kH, (size_t)10 * kH * 4, cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(a->g_decpose_b, s.g_bdec, 135 * 4, cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(a->g_decshape_b, s.g_bdec + 135, 10 * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}
