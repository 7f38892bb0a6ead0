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
__device__ __forceinline__ const float* dec_row(const float* decpose_w, const float* decshape_w, int o) {
  return o < 135 ? decpose_w + (size_t)o * kHid : decshape_w + (size_t)(o - 135) * kHid;
}

// T[o][j] = sum_i Wdec[o][i] * W2[i][j]      (145 x 1024, fp64)
__global__ void ief_fold_t_kernel(const float* __restrict__ decpose_w, const float* __restrict__ decshape_w,
                                  const float* __restrict__ fc2_w, double* __restrict__ T) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
  if (j >= kHid) return;
  const float* wd = dec_row(decpose_w, decshape_w, o);
  double acc = 0.0;
  for (int i = 0; i < kHid; ++i) acc += (double)__ldg(wd + i) * (double)__ldg(fc2_w + (size_t)i * kHid + j);
  T[(size_t)o * kHid + j] = acc;
}

// G[o][k] = sum_j T[o][j] * W1[j][k]  -> GxT[k][o] (k < 2048) / GuT[k-2048][o]
__global__ void ief_fold_g_kernel(const double* __restrict__ T, const float* __restrict__ fc1_w, float* __restrict__ GxT,
                                  float* __restrict__ GuT) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
  const int fc1_in = kFeat + kState;
  if (k >= fc1_in) return;
  double acc = 0.0;
  for (int j = 0; j < kHid; ++j) acc += T[(size_t)o * kHid + j] * (double)__ldg(fc1_w + (size_t)j * fc1_in + k);
  if (k < kFeat) GxT[(size_t)k * kDecPad + o] = (float)acc;
  else GuT[(size_t)(k - kFeat) * kDecPad + o] = (float)acc;
}

// g[o] = T[o] . b1 + Wdec[o] . b2 + bdec[o]
__global__ void ief_fold_bias_kernel(const double* __restrict__ T, const float* __restrict__ fc1_b,
                                     const float* __restrict__ decpose_w, const float* __restrict__ decshape_w,
                                     const float* __restrict__ fc2_b, const float* __restrict__ decpose_b,
                                     const float* __restrict__ decshape_b, float* __restrict__ g) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= kDecPad) return;
  if (o >= kDec) { g[o] = 0.f; return; }
  const float* wd = dec_row(decpose_w, decshape_w, o);
  double acc = o < 135 ? (double)decpose_b[o] : (double)decshape_b[o - 135];
  for (int j = 0; j < kHid; ++j) acc += T[(size_t)o * kHid + j] * (double)fc1_b[j] + (double)wd[j] * (double)fc2_b[j];
  g[o] = (float)acc;
}

// ------------------------------------------------------------------------------ forward kernels
// partial[ks][m][o] = sum_{k in slice ks} GxT[k][o] * xf[m][k];  rows m in [0,2B): view m / B, pair m % B.
__global__ void __launch_bounds__(kDecPad) ief_base_kernel(int B, const float* __restrict__ xf0, const float* __restrict__ xf1,
                                                           const float* __restrict__ GxT, float* __restrict__ partial) {
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
#pragma unroll 4
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
__global__ void __launch_bounds__(2 * kDecPad) ief_iter_kernel(IefIterArgs a) {
  __shared__ float st[2][kDecPad];        // per view: pose[0..135) then shape[135..145)
  __shared__ float u[2][kState];
  const int b = blockIdx.x, v = threadIdx.x / kDecPad, o = threadIdx.x % kDecPad, M = 2 * a.B;
  const int m = v * a.B + b;
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
  const float* gu = a.GuT + o;
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
#pragma unroll 4
      for (int k = 0; k < kState; ++k) d = fmaf(__ldg(gu + (size_t)k * kDecPad), u[v][k], d);
    }
    __syncthreads();
    if (o < kDec) st[v][o] += d;          // pred = previous + decoder output  (:195-202)
    __syncthreads();
  }
  if (o < 135) (v ? a.out_pose1 : a.out_pose0)[(size_t)b * 135 + o] = st[v][o];
  else if (o < kDec) (v ? a.out_betas1 : a.out_betas0)[(size_t)b * 10 + (o - 135)] = st[v][o];
}

int ief_create(airpose_net* h) {
  IefState& s = h->ief;
  AP_CHECK_CUDA(cudaMalloc((void**)&s.GxT, (size_t)kFeat * kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.GuT, (size_t)kState * kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.g, kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.T, (size_t)kDec * kHid * sizeof(double)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.init_pose, 144 * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&s.init_shape, 10 * sizeof(float)));
  AP_CHECK_CUDA(cudaMemset(s.GxT, 0, (size_t)kFeat * kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMemset(s.GuT, 0, (size_t)kState * kDecPad * sizeof(float)));
  return 0;
}

void ief_destroy(airpose_net* h) {
  IefState& s = h->ief;
  void* ptrs[] = {s.GxT, s.GuT, s.g, s.T, s.init_pose, s.init_shape, s.partial};
  for (void* p : ptrs) cudaFree(p);
  s = IefState();
}

int ief_load(airpose_net* h, const airpose_net_params* p, cudaStream_t st) {
  AP_REQUIRE(p->fc1_w && p->fc1_b && p->fc2_w && p->fc2_b && p->decpose_w && p->decpose_b && p->decshape_w &&
             p->decshape_b && p->init_pose && p->init_shape, "airpose_net_load: regressor parameter is null");
  IefState& s = h->ief;
  ief_fold_t_kernel<<<dim3(ceil_div(kHid, 128), kDec), 128, 0, st>>>(p->decpose_w, p->decshape_w, p->fc2_w, s.T);
  AP_LAUNCH_CHECK();
  ief_fold_g_kernel<<<dim3(ceil_div(kFeat + kState, 128), kDec), 128, 0, st>>>(s.T, p->fc1_w, s.GxT, s.GuT);
  AP_LAUNCH_CHECK();
  ief_fold_bias_kernel<<<1, kDecPad, 0, st>>>(s.T, p->fc1_b, p->decpose_w, p->decshape_w, p->fc2_b, p->decpose_b,
                                              p->decshape_b, s.g);
  AP_LAUNCH_CHECK();
  AP_CHECK_CUDA(cudaMemcpyAsync(s.init_pose, p->init_pose, 144 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(s.init_shape, p->init_shape, 10 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace airpose

using namespace airpose;

extern "C" int airpose_ief_fwd(airpose_net_t* h, const airpose_ief_args* a, void* stream_) {
  AP_REQUIRE(h && a, "airpose_ief_fwd: null argument");
  AP_REQUIRE(h->loaded, "airpose_ief_fwd: weights not loaded (call airpose_net_load)");
  AP_REQUIRE(a->batch >= 0 && a->iters >= 1, "airpose_ief_fwd: bad batch/iters");
  AP_REQUIRE(a->xf0 && a->xf1 && a->bb0 && a->bb1 && a->pos0 && a->pos1 && a->out_pose0 && a->out_pose1 &&
             a->out_betas0 && a->out_betas1, "airpose_ief_fwd: null tensor");
  AP_REQUIRE((a->init_theta0 == nullptr) == (a->init_theta1 == nullptr) && (a->init_shape0 == nullptr) == (a->init_shape1 == nullptr),
             "airpose_ief_fwd: init_theta / init_shape must be given for both views or neither");
  const int B = a->batch, M = 2 * B;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  IefState& s = h->ief;
  if (M > s.partial_rows) {
    AP_CHECK_CUDA(cudaStreamSynchronize(st));
    cudaFree(s.partial);
    s.partial = nullptr; s.partial_rows = 0;
    AP_CHECK_CUDA(cudaMalloc((void**)&s.partial, (size_t)kIefKSlices * M * kDecPad * sizeof(float)));
    s.partial_rows = M;
  }
  ief_base_kernel<<<dim3(ceil_div(M, kIefRows), kIefKSlices), kDecPad, 0, st>>>(B, a->xf0, a->xf1, s.GxT, s.partial);
  AP_LAUNCH_CHECK();
  IefIterArgs k{};
  k.B = B; k.iters = a->iters;
  k.bb0 = a->bb0; k.bb1 = a->bb1; k.pos0 = a->pos0; k.pos1 = a->pos1;
  k.th0 = a->init_theta0; k.th1 = a->init_theta1; k.th_stride = a->init_theta_stride;
  k.sh0 = a->init_shape0; k.sh1 = a->init_shape1; k.sh_stride = a->init_shape_stride;
  k.init_pose = s.init_pose; k.init_shape = s.init_shape;
  k.partial = s.partial; k.GuT = s.GuT; k.g = s.g;
  k.out_pose0 = a->out_pose0; k.out_betas0 = a->out_betas0; k.out_pose1 = a->out_pose1; k.out_betas1 = a->out_betas1;
  ief_iter_kernel<<<B, 2 * kDecPad, 0, st>>>(k);
  AP_LAUNCH_CHECK();
  return 0;
}
