// IEF regressor of copenet, eval mode -- collapsed.
//
// Replaces the 3-iteration loop of copenet.forward and forward_reg
// (/root/reference/copenet/src/copenet/models/model_copenet.py:118-159,178-204).
//
// In eval mode (dropout = identity) there is NO nonlinearity between fc1, fc2 and the decoders
// (model_copenet.py:186-189,195-199), so one regressor pass is an affine map of its input
//     d = Wdec (W2 (W1 z + b1) + b2) + bdec = G z + g,   G = Wdec W2 W1  [145 x 2332]
// with z = [xf(2048) | u(284)], u = [bb, pos, orient, art_self, shape_self, art_other, shape_other]
// (:185,:192) and Wdec = [decpose; decshape].  G and g are formed once per weight load in fp64;
// a forward is then
//   1. base = Gx xf + g           iteration-invariant, 2B x 2048 x 145  (split-K over CTAs, fixed-order sum)
//   2. `iters` times: state += base + Gu u(state)      2B x 284 x 145, one CTA per frame pair
// i.e. 38 + 3*16 MMAC per 64 pairs instead of 840 MMAC in nine latency-bound GEMMs.  All fp32 FMA:
// the result differs from the reference's fp32 chain only by summation order (~1e-6 relative).
#include "common.cuh"
#include "ptx.cuh"
#include "net.cuh"

namespace airpose {

constexpr int kState = 284;            // fc1 input minus the image feature
constexpr int kHid = 1024;
constexpr int kDec = 145;              // 135 pose + 10 shape
constexpr int kDecPad = 160;
constexpr int kIefKSlices = 16;        // split-K of the 2048-wide feature product
constexpr int kIefKPer = kFeat / kIefKSlices;
constexpr int kIefRows = 8;            // rows of [xf0; xf1] per CTA in the base product

// ------------------------------------------------------------------------------ load-time kernels
// Both regressors decode 145 numbers: two-view [decpose 135 | decshape 10] (model_copenet.py:71-72),
// hmr [decpose 132 | decshape 10 | deccam 3] (model_hmr.py:69-71).  The decoder matrices are copied into
// one contiguous Wdec [145][1024] / bdec [145] at load so the fold kernels serve both.

// T[o][j] = sum_i Wdec[o][i] * W2[i][j]      (145 x 1024, fp64)
__global__ void ief_fold_t_kernel(const float* __restrict__ wdec, const float* __restrict__ fc2_w, double* __restrict__ T) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
  if (j >= kHid) return;
  const float* wd = wdec + (size_t)o * kHid;
  double acc = 0.0;
  for (int i = 0; i < kHid; ++i) acc += (double)__ldg(wd + i) * (double)__ldg(fc2_w + (size_t)i * kHid + j);
  T[(size_t)o * kHid + j] = acc;
}

// G[o][k] = sum_j T[o][j] * W1[j][k]  -> GxT[k][o] (k < 2048) / GuT[k-2048][o]
__global__ void ief_fold_g_kernel(const double* __restrict__ T, const float* __restrict__ fc1_w, int fc1_in,
                                  float* __restrict__ GxT, float* __restrict__ GuT) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
  if (k >= fc1_in) return;
  double acc = 0.0;
  for (int j = 0; j < kHid; ++j) acc += T[(size_t)o * kHid + j] * (double)__ldg(fc1_w + (size_t)j * fc1_in + k);
  if (k < kFeat) GxT[(size_t)k * kDecPad + o] = (float)acc;
  else GuT[(size_t)(k - kFeat) * kDecPad + o] = (float)acc;
}

// g[o] = T[o] . b1 + Wdec[o] . b2 + bdec[o]
__global__ void ief_fold_bias_kernel(const double* __restrict__ T, const float* __restrict__ fc1_b,
                                     const float* __restrict__ wdec, const float* __restrict__ fc2_b,
                                     const float* __restrict__ bdec, float* __restrict__ g) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= kDecPad) return;
  if (o >= kDec) { g[o] = 0.f; return; }
  const float* wd = wdec + (size_t)o * kHid;
  double acc = (double)bdec[o];
  for (int j = 0; j < kHid; ++j) acc += T[(size_t)o * kHid + j] * (double)fc1_b[j] + (double)wd[j] * (double)fc2_b[j];
  g[o] = (float)acc;
}

// ------------------------------------------------------------------------------ forward kernels
// partial[ks][m][o] = sum_{k in slice ks} GxT[k][o] * xf[m][k];  rows m in [0,2B): view m / B, pair m % B.
__global__ void __launch_bounds__(kDecPad) ief_base_kernel(int B, const float* __restrict__ xf0, const float* __restrict__ xf1,
                                                           const float* __restrict__ GxT, float* __restrict__ partial) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  __shared__ float xs[kIefRows][kIefKPer];
  const int m0 = blockIdx.x * kIefRows, ks = blockIdx.y, o = threadIdx.x, M = 2 * B;
  for (int i = threadIdx.x; i < kIefRows * kIefKPer; i += kDecPad) {
    const int r = i / kIefKPer, k = i % kIefKPer, m = m0 + r;
    float v = 0.f;
    if (m < M) v = __ldg((m < B ? xf0 + (size_t)m * kFeat : xf1 + (size_t)(m - B) * kFeat) + ks * kIefKPer + k);
    xs[r][k] = v;
  }
  __syncthreads();
  float acc[kIefRows];
#pragma unroll
  for (int r = 0; r < kIefRows; ++r) acc[r] = 0.f;
  const float* gp = GxT + (size_t)ks * kIefKPer * kDecPad + o;
#pragma unroll 16
  for (int k = 0; k < kIefKPer; ++k) {
    const float gv = __ldg(gp + (size_t)k * kDecPad);
#pragma unroll
    for (int r = 0; r < kIefRows; ++r) acc[r] = fmaf(gv, xs[r][k], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kIefRows; ++r)
    if (m0 + r < M) partial[((size_t)ks * M + m0 + r) * kDecPad + o] = acc[r];
}

// same product for a single feature tensor (hmr): partial[ks][m][o], rows m in [0,M)
__global__ void __launch_bounds__(kDecPad) ief_base_rows_kernel(int M, const float* __restrict__ xf, const float* __restrict__ GxT,
                                                                float* __restrict__ partial) {
  __shared__ float xs[kIefRows][kIefKPer];
  const int m0 = blockIdx.x * kIefRows, ks = blockIdx.y, o = threadIdx.x;
  for (int i = threadIdx.x; i < kIefRows * kIefKPer; i += kDecPad) {
    const int r = i / kIefKPer, k = i % kIefKPer, m = m0 + r;
    xs[r][k] = m < M ? __ldg(xf + (size_t)m * kFeat + ks * kIefKPer + k) : 0.f;
  }
  __syncthreads();
  float acc[kIefRows];
#pragma unroll
  for (int r = 0; r < kIefRows; ++r) acc[r] = 0.f;
  const float* gp = GxT + (size_t)ks * kIefKPer * kDecPad + o;
#pragma unroll 16
  for (int k = 0; k < kIefKPer; ++k) {
    const float gv = __ldg(gp + (size_t)k * kDecPad);
#pragma unroll
    for (int r = 0; r < kIefRows; ++r) acc[r] = fmaf(gv, xs[r][k], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kIefRows; ++r)
    if (m0 + r < M) partial[((size_t)ks * M + m0 + r) * kDecPad + o] = acc[r];
}

struct IefIterArgs {
  int B, iters;
  const float *bb0, *bb1, *pos0, *pos1;
  const float *th0, *th1; int th_stride;
  const float *sh0, *sh1; int sh_stride;
  const float *init_pose, *init_shape;
  const float *partial, *GuT, *g;
  float *out_pose0, *out_betas0, *out_pose1, *out_betas1;
};

// One CTA per frame pair, thread (v, o): view v in {0,1}, output o in [0,160).
// GuT (284 x 160 fp32 = 178 KB) is staged in shared memory once per CTA: the three iterations then read it
// conflict-free (consecutive o) instead of chasing 3 x 284 dependent L2 loads per thread.
constexpr int kIefIterSmem = kState * kDecPad * (int)sizeof(float);
__global__ void __launch_bounds__(2 * kDecPad) ief_iter_kernel(IefIterArgs a) {
  ptx::grid_dep_wait();      // launched through launch_chain (common.cuh)
  ptx::grid_dep_launch();
  extern __shared__ __align__(16) float gsm[];   // [kState][kDecPad]
  __shared__ float st[2][kDecPad];        // per view: pose[0..135) then shape[135..145)
  __shared__ __align__(16) float u[2][kState];
  const int b = blockIdx.x, v = threadIdx.x / kDecPad, o = threadIdx.x % kDecPad, M = 2 * a.B;
  const int m = v * a.B + b;
  {
    const float4* src = reinterpret_cast<const float4*>(a.GuT);
    float4* dst = reinterpret_cast<float4*>(gsm);
#pragma unroll 12
    for (int i = threadIdx.x; i < kState * kDecPad / 4; i += 2 * kDecPad) dst[i] = __ldg(src + i);
  }
  // base = g + fixed-order sum of the split-K partials
  float base = 0.f;
  if (o < kDec) {
    base = __ldg(a.g + o);
    for (int ks = 0; ks < kIefKSlices; ++ks) base += __ldg(a.partial + ((size_t)ks * M + m) * kDecPad + o);
  }
  // initial state (model_copenet.py:121-135): [position | init theta (132) | init shape]
  if (o < 3) st[v][o] = (v ? a.pos1 : a.pos0)[b * 3 + o];
  else if (o < 135) {
    const float* th = v ? a.th1 : a.th0;
    st[v][o] = th ? th[(size_t)b * a.th_stride + (o - 3)] : a.init_pose[o - 3];
  } else if (o < kDec) {
    const float* sh = v ? a.sh1 : a.sh0;
    st[v][o] = sh ? sh[(size_t)b * a.sh_stride + (o - 135)] : a.init_shape[o - 135];
  }
  __syncthreads();
  const float* gu = gsm + o;
  for (int it = 0; it < a.iters; ++it) {
    // u = [bb(3), pose_self(135), shape_self(10), art_other(126), shape_other(10)]   (:185,:192)
    for (int k = o; k < kState; k += kDecPad) {
      float x;
      if (k < 3) x = (v ? a.bb1 : a.bb0)[b * 3 + k];
      else if (k < 148) x = st[v][k - 3];
      else if (k < 274) x = st[1 - v][9 + (k - 148)];
      else x = st[1 - v][135 + (k - 274)];
      u[v][k] = x;
    }
    __syncthreads();
    float d = base;
    if (o < kDec) {
      // four partial sums (fixed order: the result does not depend on the staging)
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 4
      for (int k = 0; k < kState; k += 4) {
        const float4 uv = *reinterpret_cast<const float4*>(&u[v][k]);
        d0 = fmaf(gu[(k + 0) * kDecPad], uv.x, d0);
        d1 = fmaf(gu[(k + 1) * kDecPad], uv.y, d1);
        d2 = fmaf(gu[(k + 2) * kDecPad], uv.z, d2);
        d3 = fmaf(gu[(k + 3) * kDecPad], uv.w, d3);
      }
      d += (d0 + d1) + (d2 + d3);
    }
    __syncthreads();
    if (o < kDec) st[v][o] += d;          // pred = previous + decoder output  (:195-202)
    __syncthreads();
  }
  if (o < 135) (v ? a.out_pose1 : a.out_pose0)[(size_t)b * 135 + o] = st[v][o];
  else if (o < kDec) (v ? a.out_betas1 : a.out_betas0)[(size_t)b * 10 + (o - 135)] = st[v][o];
}

// hmr regressor (model_hmr.py:112-172): single view, state = [pose 132 | shape 10 | cam 3] is also the
// iterated part of the fc1 input (:161), so one iteration is  state += base + Gu state.
// One CTA per image, thread o = output; base = g + fixed-order sum of the split-K partials.
struct HmrIterArgs {
  int B, iters;
  const float *th; int th_stride;
  const float *sh; int sh_stride;
  const float *cam; int cam_stride;
  const float *init_pose, *init_shape, *init_cam;
  const float *partial, *GuT, *g;
  float *out_pose, *out_betas, *out_cam;
};

__global__ void __launch_bounds__(kDecPad) hmr_iter_kernel(HmrIterArgs a) {
  __shared__ float st[kDecPad];
  const int b = blockIdx.x, o = threadIdx.x;
  float base = 0.f;
  if (o < kDec) {
    base = __ldg(a.g + o);
    for (int ks = 0; ks < kIefKSlices; ++ks) base += __ldg(a.partial + ((size_t)ks * a.B + b) * kDecPad + o);
  }
  if (o < 132) st[o] = a.th ? a.th[(size_t)b * a.th_stride + o] : a.init_pose[o];                   // model_hmr.py:116-119
  else if (o < 142) st[o] = a.sh ? a.sh[(size_t)b * a.sh_stride + (o - 132)] : a.init_shape[o - 132];
  else if (o < kDec) st[o] = a.cam ? a.cam[(size_t)b * a.cam_stride + (o - 142)] : a.init_cam[o - 142];
  else st[o] = 0.f;
  __syncthreads();
  const float* gu = a.GuT + o;
  for (int it = 0; it < a.iters; ++it) {
    float d = base;
    if (o < kDec) {
#pragma unroll 5
      for (int k = 0; k < kDec; ++k) d = fmaf(__ldg(gu + (size_t)k * kDecPad), st[k], d);
    }
    __syncthreads();
    if (o < kDec) st[o] += d;               // pred = decoder output + previous (:168-170)
    __syncthreads();
  }
  if (o < 132) a.out_pose[(size_t)b * 132 + o] = st[o];
  else if (o < 142) a.out_betas[(size_t)b * 10 + (o - 132)] = st[o];
  else if (o < kDec) a.out_cam[(size_t)b * 3 + (o - 142)] = st[o];
}

static int ief_state_create(IefState& s, int state_width) {
  AP_CHECK_CUDA(cudaMalloc((void**)&s.GxT, (size_t)kFeat * kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.GuT, (size_t)state_width * kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.g, kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.T, (size_t)kDec * kHid * sizeof(double)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.wdec, (size_t)kDec * kHid * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.bdec, kDec * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.init_pose, 144 * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.init_shape, 10 * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.init_cam, 3 * sizeof(float)));
  AP_CHECK_CUDA(cudaMemset(s.GxT, 0, (size_t)kFeat * kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMemset(s.GuT, 0, (size_t)state_width * kDecPad * sizeof(float)));
  return 0;
}

static void ief_state_destroy(IefState& s) {
  void* ptrs[] = {s.GxT, s.GuT, s.g, s.T, s.wdec, s.bdec, s.init_pose, s.init_shape, s.init_cam, s.partial};
  for (void* p : ptrs) cudaFree(p);
  s = IefState();
}

// Wdec / bdec from up to three decoders, then G = Wdec W2 W1 and g (fp64 accumulation)
struct DecPart { const float* w; const float* b; int rows; };
static int ief_fold(IefState& s, const DecPart* parts, int nparts, const float* fc1_w, const float* fc1_b, int state_width,
                    const float* fc2_w, const float* fc2_b, cudaStream_t st) {
  int row = 0;
  for (int i = 0; i < nparts; ++i) {
    AP_CHECK_CUDA(cudaMemcpyAsync(s.wdec + (size_t)row * kHid, parts[i].w, (size_t)parts[i].rows * kHid * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
    AP_CHECK_CUDA(cudaMemcpyAsync(s.bdec + row, parts[i].b, parts[i].rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
    row += parts[i].rows;
  }
  AP_REQUIRE(row == kDec, "ief_fold: decoders have %d rows, expected %d", row, kDec);
  const int fc1_in = kFeat + state_width;
  ief_fold_t_kernel<<<dim3(ceil_div(kHid, 128), kDec), 128, 0, st>>>(s.wdec, fc2_w, s.T);
  AP_LAUNCH_CHECK();
  ief_fold_g_kernel<<<dim3(ceil_div(fc1_in, 128), kDec), 128, 0, st>>>(s.T, fc1_w, fc1_in, s.GxT, s.GuT);
  AP_LAUNCH_CHECK();
  ief_fold_bias_kernel<<<1, kDecPad, 0, st>>>(s.T, fc1_b, s.wdec, fc2_b, s.bdec, s.g);
  AP_LAUNCH_CHECK();
  return 0;
}

static int ensure_partial(IefState& s, int rows, cudaStream_t st) {
  if (rows > s.partial_rows) {
    AP_CHECK_CUDA(cudaStreamSynchronize(st));
    cudaFree(s.partial);
    s.partial = nullptr; s.partial_rows = 0;
    AP_CHECK_CUDA(cudaMalloc((void**)&s.partial, (size_t)kIefKSlices * rows * kDecPad * sizeof(float)));
    s.partial_rows = rows;
  }
  return 0;
}

int ief_create(airpose_net* h) {
  if (ief_state_create(h->ief, kState)) return 1;
  return ief_state_create(h->ief_hmr, kDec);
}

void ief_destroy(airpose_net* h) {
  ief_state_destroy(h->ief);
  ief_state_destroy(h->ief_hmr);
}

int ief_load(airpose_net* h, const airpose_net_params* p, cudaStream_t st) {
  AP_REQUIRE(p->fc1_w && p->fc1_b && p->fc2_w && p->fc2_b && p->decpose_w && p->decpose_b && p->decshape_w &&
             p->decshape_b && p->init_pose && p->init_shape, "airpose_net_load: regressor parameter is null");
  IefState& s = h->ief;
  const DecPart parts[2] = {{p->decpose_w, p->decpose_b, 135}, {p->decshape_w, p->decshape_b, 10}};
  if (ief_fold(s, parts, 2, p->fc1_w, p->fc1_b, kState, p->fc2_w, p->fc2_b, st)) return 1;
  AP_CHECK_CUDA(cudaMemcpyAsync(s.init_pose, p->init_pose, 144 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(s.init_shape, p->init_shape, 10 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int ief_load_hmr(airpose_net* h, const airpose_hmr_params* p, cudaStream_t st) {
  AP_REQUIRE(p->fc1_w && p->fc1_b && p->fc2_w && p->fc2_b && p->decpose_w && p->decpose_b && p->decshape_w &&
             p->decshape_b && p->deccam_w && p->deccam_b && p->init_pose && p->init_shape && p->init_cam,
             "airpose_hmr_load: regressor parameter is null");
  IefState& s = h->ief_hmr;
  const DecPart parts[3] = {{p->decpose_w, p->decpose_b, 132}, {p->decshape_w, p->decshape_b, 10}, {p->deccam_w, p->deccam_b, 3}};
  if (ief_fold(s, parts, 3, p->fc1_w, p->fc1_b, kDec, p->fc2_w, p->fc2_b, st)) return 1;
  AP_CHECK_CUDA(cudaMemcpyAsync(s.init_pose, p->init_pose, 144 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(s.init_shape, p->init_shape, 10 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(s.init_cam, p->init_cam, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}


}  // namespace airpose

using namespace airpose;

extern "C" int airpose_ief_fwd(airpose_net_t* h, const airpose_ief_args* a, void* stream_) {
  AP_REQUIRE(h && a, "airpose_ief_fwd: null argument");
  AP_REQUIRE(h->loaded && !h->hmr_loaded, "airpose_ief_fwd: two-view weights not loaded (call airpose_net_load)");
  AP_REQUIRE(a->batch >= 0 && a->iters >= 1, "airpose_ief_fwd: bad batch/iters");
  AP_REQUIRE(a->xf0 && a->xf1 && a->bb0 && a->bb1 && a->pos0 && a->pos1 && a->out_pose0 && a->out_pose1 &&
             a->out_betas0 && a->out_betas1, "airpose_ief_fwd: null tensor");
  AP_REQUIRE((a->init_theta0 == nullptr) == (a->init_theta1 == nullptr) && (a->init_shape0 == nullptr) == (a->init_shape1 == nullptr),
             "airpose_ief_fwd: init_theta / init_shape must be given for both views or neither");
  const int B = a->batch, M = 2 * B;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  IefState& s = h->ief;
  if (ensure_partial(s, M, st)) return 1;
  AP_CHECK_CUDA(launch_chain(ief_base_kernel, dim3(ceil_div(M, kIefRows), kIefKSlices), dim3(kDecPad), st, B, a->xf0, a->xf1, (const float*)s.GxT, s.partial));
  IefIterArgs k{};
  k.B = B; k.iters = a->iters;
  k.bb0 = a->bb0; k.bb1 = a->bb1; k.pos0 = a->pos0; k.pos1 = a->pos1;
  k.th0 = a->init_theta0; k.th1 = a->init_theta1; k.th_stride = a->init_theta_stride;
  k.sh0 = a->init_shape0; k.sh1 = a->init_shape1; k.sh_stride = a->init_shape_stride;
  k.init_pose = s.init_pose; k.init_shape = s.init_shape;
  k.partial = s.partial; k.GuT = s.GuT; k.g = s.g;
  k.out_pose0 = a->out_pose0; k.out_betas0 = a->out_betas0; k.out_pose1 = a->out_pose1; k.out_betas1 = a->out_betas1;
  AP_CHECK_CUDA(cudaFuncSetAttribute(ief_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kIefIterSmem));
  AP_CHECK_CUDA(launch_chain_smem(ief_iter_kernel, dim3(B), dim3(2 * kDecPad), kIefIterSmem, st, k));
  return 0;
}

extern "C" int airpose_hmr_ief_fwd(airpose_net_t* h, const airpose_hmr_ief_args* a, void* stream_) {
  AP_REQUIRE(h && a, "airpose_hmr_ief_fwd: null argument");
  AP_REQUIRE(h->hmr_loaded, "airpose_hmr_ief_fwd: hmr weights not loaded (call airpose_hmr_load)");
  AP_REQUIRE(a->batch >= 0 && a->iters >= 1, "airpose_hmr_ief_fwd: bad batch/iters");
  AP_REQUIRE(a->xf && a->out_pose && a->out_betas && a->out_cam, "airpose_hmr_ief_fwd: null tensor");
  const int B = a->batch;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  IefState& s = h->ief_hmr;
  if (ensure_partial(s, B, st)) return 1;
  // base product: the two-view kernel with "view 1" empty (rows [0,B) come from xf)
  ief_base_rows_kernel<<<dim3(ceil_div(B, kIefRows), kIefKSlices), kDecPad, 0, st>>>(B, a->xf, s.GxT, s.partial);
  AP_LAUNCH_CHECK();
  HmrIterArgs k{};
  k.B = B; k.iters = a->iters;
  k.th = a->init_theta; k.th_stride = a->init_theta_stride;
  k.sh = a->init_shape; k.sh_stride = a->init_shape_stride;
  k.cam = a->init_cam; k.cam_stride = a->init_cam_stride;
  k.init_pose = s.init_pose; k.init_shape = s.init_shape; k.init_cam = s.init_cam;
  k.partial = s.partial; k.GuT = s.GuT; k.g = s.g;
  k.out_pose = a->out_pose; k.out_betas = a->out_betas; k.out_cam = a->out_cam;
  hmr_iter_kernel<<<B, kDecPad, 0, st>>>(k);
  AP_LAUNCH_CHECK();
  return 0;
}
