// ResNet-50 trunk and IEF regressor of copenet, eval mode.
//
// Replaces (paths relative to /root/reference/copenet/src/copenet/models):
//   model_copenet.py:161-176  copenet.forward_feat_ext  (stem, 16 Bottlenecks :27-47, AvgPool2d(7))
//   model_copenet.py:118-159,178-204  the 3-iteration regressor loop / forward_reg
//
// Data layout: activations NHWC bf16 in a per-handle workspace, processed in
// chunks of images so consecutive layers meet in L2; conv weights bf16 [Cout][tap][Cin];
// BatchNorm folded to per-channel fp32 scale/shift applied in the GEMM epilogue.
#include <cuda_bf16.h>

#include <cstdlib>
#include <map>
#include <utility>
#include <vector>

#include "common.cuh"
#include "gemm.cuh"

namespace airpose {

struct ConvSpec { int cout, cin, k, stride, pad; };

static std::vector<ConvSpec> resnet50_specs() {      // forward order, see synthetic.conv_specs()
  std::vector<ConvSpec> v;
  v.push_back({64, 3, 7, 2, 3});
  const int layers[4] = {3, 4, 6, 3}, planes[4] = {64, 128, 256, 512};
  int inpl = 64;
  for (int li = 0; li < 4; ++li)
    for (int b = 0; b < layers[li]; ++b) {
      const int s = (li > 0 && b == 0) ? 2 : 1;
      v.push_back({planes[li], inpl, 1, 1, 0});
      v.push_back({planes[li], planes[li], 3, s, 1});
      v.push_back({planes[li] * 4, planes[li], 1, 1, 0});
      if (b == 0) v.push_back({planes[li] * 4, inpl, 1, s, 0});
      inpl = planes[li] * 4;
    }
  return v;
}

// Stem operand layout (DESIGN.md "stem"): per input row h and output column q the 7 taps x 3 channels
// along W (21 values, padded to 32) are packed once; rows are split by parity so that for a fixed
// vertical tap r the 128 output pixels of a tile read 128 CONSECUTIVE packed rows.
constexpr int kStemTapK = 32;          // 7*3 = 21 padded to 32 bf16 = one 64-byte swizzle row
constexpr int kStemK = 7 * kStemTapK;  // 224
constexpr int kStemPlaneRows = 115;    // 2 zero rows + 112 + 1 zero row  (vertical taps reach p-2 .. p+1)
constexpr int kFeat = 2048;
constexpr int kState = 284, kStatePad = 320;
constexpr int kHid = 1024;
constexpr int kDec = 145, kDecPad = 160;

// ------------------------------------------------------------------------------ pack kernels
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin,
                                        int k, int kpad, int stem) {
  const int64_t total = (int64_t)cout * kpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / kpad), kk = (int)(i % kpad);
    float val = 0.f;
    if (stem) {                       // [o][r][s*3+c], 21 -> 32 zero padded per vertical tap r
      const int r = kk / kStemTapK, e = kk % kStemTapK;
      if (e < 21) val = w[(((int64_t)o * cin + (e % 3)) * k + r) * k + e / 3];
    } else {                          // (r, s, c): tap-major, channel-minor = the im2col K order
      const int tap = kk / cin, c = kk % cin;
      val = w[(((int64_t)o * cin + c) * k * k) + tap];
    }
    out[i] = __float2bfloat16_rn(val);
  }
}

__global__ void fold_bn_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, int c, float* __restrict__ scale,
                               float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float s = g[i] / sqrtf(var[i] + eps);
  scale[i] = s;
  shift[i] = b[i] - mean[i] * s;
}

// W [rows, src_ld] columns [col0, col0+ncols) -> bf16 [rows_pad, 3*kp] = [hi | hi | lo]
__global__ void pack_split_weight_kernel(const float* __restrict__ w, int rows, int64_t src_ld, int col0, int ncols,
                                         __nv_bfloat16* __restrict__ out, int rows_pad, int kp) {
  const int64_t total = (int64_t)rows_pad * kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / kp), k = (int)(i % kp);
    const float v = (r < rows && k < ncols) ? w[(int64_t)r * src_ld + col0 + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + (int64_t)r * 3 * kp;
    o[k] = hi; o[kp + k] = hi; o[2 * kp + k] = lo;
  }
}

// x fp32 [rows, ld] -> bf16 [rows, 3*kp] = [hi | lo | hi]
__global__ void split_act_kernel(const float* __restrict__ x, int rows, int64_t ld, int ncols,
                                 __nv_bfloat16* __restrict__ out, int kp) {
  const int64_t total = (int64_t)rows * kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / kp), k = (int)(i % kp);
    const float v = (k < ncols) ? x[(int64_t)r * ld + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + (int64_t)r * 3 * kp;
    o[k] = hi; o[kp + k] = lo; o[2 * kp + k] = hi;
  }
}

__global__ void concat_bias_kernel(const float* a, int na, const float* b, int nb, float* out, int npad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  out[i] = i < na ? a[i] : (i < na + nb ? b[i - na] : 0.f);
}

// ------------------------------------------------------------------------------ trunk kernels
// Stem operand pack: x fp32 NCHW [n,3,224,224] -> bf16 [n][2 parities][115 rows][112 q][32]:
//   out[n][par][hp][q][s*3+c] = x[n][c][2*(hp-2)+par][2q-3+s]   (zero outside the image / for hp in {0,1,114})
// (7x7, stride 2, pad 3; model_copenet.py:57-58).  One thread packs one (row, q): 64 bytes.
__global__ void stem_pack_kernel(const float* __restrict__ x, int n, __nv_bfloat16* __restrict__ out) {
  const int64_t total = (int64_t)n * 2 * kStemPlaneRows * 112;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % 112);
    const int hp = (int)((i / 112) % kStemPlaneRows);
    const int par = (int)((i / (112 * kStemPlaneRows)) % 2);
    const int img = (int)(i / (112 * kStemPlaneRows * 2));
    const int h = 2 * (hp - 2) + par;
    __align__(16) __nv_bfloat16 vals[kStemTapK];
#pragma unroll
    for (int e = 0; e < kStemTapK; ++e) vals[e] = __float2bfloat16_rn(0.f);
    if (hp >= 2 && hp < 114) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* row = x + (((int64_t)img * 3 + c) * 224 + h) * 224;
#pragma unroll
        for (int sx = 0; sx < 7; ++sx) {
          const int w = 2 * q - 3 + sx;
          if (w >= 0 && w < 224) vals[sx * 3 + c] = __float2bfloat16_rn(__ldg(row + w));
        }
      }
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * kStemTapK);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = reinterpret_cast<const uint4*>(vals)[j];
  }
}

// MaxPool2d(3, stride 2, pad 1) on NHWC bf16 [n,112,112,64] -> [n,56,56,64]; 8 channels per thread.
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ x, int n, __nv_bfloat16* __restrict__ y) {
  const int64_t total = (int64_t)n * 56 * 56 * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % 8);
    const int64_t pix = i / 8;
    const int q = (int)(pix % 56), pr = (int)((pix / 56) % 56), img = (int)(pix / (56 * 56));
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      const int h = pr * 2 - 1 + r;
      if (h < 0 || h >= 112) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = q * 2 - 1 + s;
        if (w < 0 || w >= 112) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)img * 112 + h) * 112 + w) * 64 + cg * 8));
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          m[2 * j] = fmaxf(m[2 * j], __uint_as_float(u[j] << 16));
          m[2 * j + 1] = fmaxf(m[2 * j + 1], __uint_as_float(u[j] & 0xFFFF0000u));
        }
      }
    }
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(m[j]);
    *reinterpret_cast<uint4*>(y + pix * 64 + cg * 8) = *reinterpret_cast<const uint4*>(o);
  }
}

// AvgPool2d(7) on NHWC bf16 [n,7,7,2048] -> fp32 [n,2048]  (model_copenet.py:173-174)
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ x, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * kFeat) return;
  const int c = i % kFeat, img = i / kFeat;
  float s = 0.f;
  for (int p = 0; p < 49; ++p) s += __bfloat162float(x[((int64_t)img * 49 + p) * kFeat + c]);
  out[i] = s / 49.f;
}

// ------------------------------------------------------------------------------ IEF kernels
// rows m in [0,2B): view v = m / B, sample b = m % B.
__global__ void ief_init_kernel(int B, const float* pos0, const float* pos1, const float* init_pose,
                                const float* th0, const float* th1, int th_stride, const float* init_shape,
                                const float* sh0, const float* sh1, int sh_stride, float* pose, float* shape) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * 145) return;
  const int m = i / 145, e = i % 145, v = m / B, b = m % B;
  if (e < 3) pose[m * 135 + e] = (v ? pos1 : pos0)[b * 3 + e];
  else if (e < 135) {
    const float* th = v ? th1 : th0;
    pose[m * 135 + e] = th ? th[(int64_t)b * th_stride + (e - 3)] : init_pose[e - 3];   // model_copenet.py:121-132
  } else {
    const float* sh = v ? sh1 : sh0;
    shape[m * 10 + (e - 135)] = sh ? sh[(int64_t)b * sh_stride + (e - 135)] : init_shape[e - 135];
  }
}

// fc1 input minus the image feature: [bb, pos, orient, art_self, shape_self, art_other, shape_other]
// (model_copenet.py:185,192), written as the split-bf16 operand [hi | lo | hi] of width 3*320.
__global__ void ief_state_kernel(int B, const float* bb0, const float* bb1, const float* pose, const float* shape,
                                 __nv_bfloat16* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * kStatePad) return;
  const int m = i / kStatePad, k = i % kStatePad, v = m / B, b = m % B, mo = (1 - v) * B + b;
  float x = 0.f;
  if (k < 3) x = (v ? bb1 : bb0)[b * 3 + k];
  else if (k < 138) x = pose[m * 135 + (k - 3)];
  else if (k < 148) x = shape[m * 10 + (k - 138)];
  else if (k < 274) x = pose[mo * 135 + 9 + (k - 148)];
  else if (k < kState) x = shape[mo * 10 + (k - 274)];
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  __nv_bfloat16* o = out + (int64_t)m * 3 * kStatePad;
  o[k] = hi; o[kStatePad + k] = lo; o[2 * kStatePad + k] = hi;
}

// pred_pose = cat(pos, orient, art) + decpose(xc); pred_shape = shape + decshape(xc)  (:195-202)
__global__ void ief_update_kernel(int B, const float* d, float* pose, float* shape, float* out_pose0, float* out_betas0,
                                  float* out_pose1, float* out_betas1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B * 145) return;
  const int m = i / 145, e = i % 145, v = m / B, b = m % B;
  if (e < 135) {
    const float val = pose[m * 135 + e] + d[m * kDecPad + e];
    pose[m * 135 + e] = val;
    if (out_pose0) (v ? out_pose1 : out_pose0)[b * 135 + e] = val;
  } else {
    const float val = shape[m * 10 + (e - 135)] + d[m * kDecPad + e];
    shape[m * 10 + (e - 135)] = val;
    if (out_betas0) (v ? out_betas1 : out_betas0)[b * 10 + (e - 135)] = val;
  }
}

}  // namespace airpose

using namespace airpose;

struct TrunkPlan {
  std::vector<GemmLaunch> gemms;       // in launch order: [stem,] then per block conv1, conv2, [down], conv3
  const __nv_bfloat16* final_act = nullptr;
};

struct IefPlan {
  GemmLaunch g0, g1, g2, g3;
};

struct airpose_net {
  int device = 0;
  int max_images = 0;
  int chunk = 0;
  bool loaded = false;
  std::vector<ConvSpec> specs;
  std::vector<__nv_bfloat16*> wq;       // packed conv weights
  std::vector<float*> scale, shift;     // folded BN
  // IEF
  __nv_bfloat16 *w1a = nullptr, *w1b = nullptr, *w2 = nullptr, *wd = nullptr;
  float *b1 = nullptr, *b2 = nullptr, *bd = nullptr, *init_pose = nullptr, *init_shape = nullptr;
  // workspaces
  __nv_bfloat16* col = nullptr;
  __nv_bfloat16* stem_out = nullptr;
  __nv_bfloat16* act[4] = {nullptr, nullptr, nullptr, nullptr};    // stage A (stem, layer1, layer2): `chunk` images
  __nv_bfloat16* actB[4] = {nullptr, nullptr, nullptr, nullptr};   // stage B (layer3, layer4): `group` images
  int group = 0;
  std::map<std::pair<int, int>, TrunkPlan> plansA;                 // (images, slot inside the group)
  std::map<int, TrunkPlan> plansB;                                 // images
  // IEF workspace (grown on demand)
  int ief_cap = 0;
  __nv_bfloat16 *xf_split = nullptr, *state_split = nullptr, *y1_split = nullptr, *y2_split = nullptr;
  float *hbuf = nullptr, *dbuf = nullptr, *pose = nullptr, *shape = nullptr;
  std::map<int, IefPlan> ief_plans;
};

// Stage A (56x56 / 28x28 activations) runs in chunks small enough that a layer's output is still in
// L2 when the next layer reads it; stage B (14x14 / 7x7) runs on a larger group so that its GEMMs
// have enough 128-row tiles to fill 148 SMs.  (DESIGN.md "trunk schedule")
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  const int c = e ? atoi(e) : dflt;
  return c > 0 ? c : dflt;
}
static int default_chunk() { return env_int("AIRPOSE_TRUNK_CHUNK", 32); }
static int default_group() { return env_int("AIRPOSE_TRUNK_GROUP", 128); }
constexpr size_t kStageBElems = 28 * 28 * 512;      // per image: the largest stage-B tensor (layer3 input)

extern "C" int airpose_net_create(airpose_net_t** out, int max_images, int device) {
  AP_REQUIRE(out && max_images > 0, "airpose_net_create: bad argument");
  AP_CHECK_CUDA(cudaSetDevice(device));
  auto* h = new airpose_net();
  h->device = device;
  h->max_images = max_images;
  h->chunk = std::min(max_images, default_chunk());
  h->group = std::max(h->chunk, std::min(max_images, default_group()) / h->chunk * h->chunk);
  h->specs = resnet50_specs();
  const size_t nconv = h->specs.size();
  h->wq.resize(nconv); h->scale.resize(nconv); h->shift.resize(nconv);
  for (size_t i = 0; i < nconv; ++i) {
    const ConvSpec& s = h->specs[i];
    const size_t kk = (i == 0) ? kStemK : (size_t)s.k * s.k * s.cin;
    AP_CHECK_CUDA(cudaMalloc((void**)&h->wq[i], (size_t)s.cout * kk * 2));
    AP_CHECK_CUDA(cudaMalloc((void**)&h->scale[i], s.cout * sizeof(float)));
    AP_CHECK_CUDA(cudaMalloc((void**)&h->shift[i], s.cout * sizeof(float)));
  }
  AP_CHECK_CUDA(cudaMalloc((void**)&h->w1a, (size_t)kHid * 3 * kFeat * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->w1b, (size_t)kHid * 3 * kStatePad * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->w2, (size_t)kHid * 3 * kHid * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->wd, (size_t)kDecPad * 3 * kHid * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->b1, kHid * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->b2, kHid * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->bd, kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->init_pose, 144 * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->init_shape, 10 * sizeof(float)));
  const size_t act_elems = (size_t)h->chunk * 112 * 112 * 64;       // == 56*56*256, the largest activation
  AP_CHECK_CUDA(cudaMalloc((void**)&h->col, (size_t)h->chunk * 2 * kStemPlaneRows * 112 * kStemTapK * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->stem_out, act_elems * 2));
  for (int i = 0; i < 4; ++i) AP_CHECK_CUDA(cudaMalloc((void**)&h->act[i], act_elems * 2));
  for (int i = 0; i < 4; ++i) AP_CHECK_CUDA(cudaMalloc((void**)&h->actB[i], (size_t)h->group * kStageBElems * 2));
  *out = h;
  return 0;
}

extern "C" int airpose_net_destroy(airpose_net_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (auto p : h->wq) cudaFree(p);
  for (auto p : h->scale) cudaFree(p);
  for (auto p : h->shift) cudaFree(p);
  void* ptrs[] = {h->w1a, h->w1b, h->w2, h->wd, h->b1, h->b2, h->bd, h->init_pose, h->init_shape, h->col, h->stem_out,
                  h->act[0], h->act[1], h->act[2], h->act[3], h->actB[0], h->actB[1], h->actB[2], h->actB[3], h->xf_split, h->state_split, h->y1_split, h->y2_split,
                  h->hbuf, h->dbuf, h->pose, h->shape};
  for (void* p : ptrs) cudaFree(p);
  delete h;
  return 0;
}

extern "C" int airpose_net_load(airpose_net_t* h, const airpose_net_params* p, void* stream_) {
  AP_REQUIRE(h && p, "airpose_net_load: null argument");
  cudaStream_t st = (cudaStream_t)stream_;
  for (size_t i = 0; i < h->specs.size(); ++i) {
    const ConvSpec& s = h->specs[i];
    const airpose_conv_params& c = p->conv[i];
    AP_REQUIRE(c.weight && c.bn_weight && c.bn_bias && c.bn_mean && c.bn_var, "airpose_net_load: conv %zu has a null parameter", i);
    const int kpad = (i == 0) ? kStemK : s.k * s.k * s.cin;
    pack_conv_weight_kernel<<<256, 256, 0, st>>>(c.weight, h->wq[i], s.cout, s.cin, s.k, kpad, i == 0);
    AP_LAUNCH_CHECK();
    fold_bn_kernel<<<ceil_div(s.cout, 256), 256, 0, st>>>(c.bn_weight, c.bn_bias, c.bn_mean, c.bn_var, p->bn_eps, s.cout,
                                                          h->scale[i], h->shift[i]);
    AP_LAUNCH_CHECK();
  }
  AP_REQUIRE(p->fc1_w && p->fc1_b && p->fc2_w && p->fc2_b && p->decpose_w && p->decpose_b && p->decshape_w &&
             p->decshape_b && p->init_pose && p->init_shape, "airpose_net_load: regressor parameter is null");
  const int fc1_in = kFeat + kState;
  pack_split_weight_kernel<<<512, 256, 0, st>>>(p->fc1_w, kHid, fc1_in, 0, kFeat, h->w1a, kHid, kFeat);
  AP_LAUNCH_CHECK();
  pack_split_weight_kernel<<<256, 256, 0, st>>>(p->fc1_w, kHid, fc1_in, kFeat, kState, h->w1b, kHid, kStatePad);
  AP_LAUNCH_CHECK();
  pack_split_weight_kernel<<<512, 256, 0, st>>>(p->fc2_w, kHid, kHid, 0, kHid, h->w2, kHid, kHid);
  AP_LAUNCH_CHECK();
  // decoder rows: 0..134 decpose, 135..144 decshape, 145..159 zero
  AP_CHECK_CUDA(cudaMemsetAsync(h->wd, 0, (size_t)kDecPad * 3 * kHid * 2, st));
  pack_split_weight_kernel<<<128, 256, 0, st>>>(p->decpose_w, 135, kHid, 0, kHid, h->wd, 135, kHid);
  AP_LAUNCH_CHECK();
  pack_split_weight_kernel<<<32, 256, 0, st>>>(p->decshape_w, 10, kHid, 0, kHid, h->wd + (size_t)135 * 3 * kHid, 10, kHid);
  AP_LAUNCH_CHECK();
  AP_CHECK_CUDA(cudaMemcpyAsync(h->b1, p->fc1_b, kHid * sizeof(float), cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(h->b2, p->fc2_b, kHid * sizeof(float), cudaMemcpyDeviceToDevice, st));
  concat_bias_kernel<<<1, 256, 0, st>>>(p->decpose_b, 135, p->decshape_b, 10, h->bd, kDecPad);
  AP_LAUNCH_CHECK();
  AP_CHECK_CUDA(cudaMemcpyAsync(h->init_pose, p->init_pose, 144 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  AP_CHECK_CUDA(cudaMemcpyAsync(h->init_shape, p->init_shape, 10 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  h->loaded = true;
  return 0;
}

static int conv_launch(airpose_net* h, int idx, const __nv_bfloat16* x, int n, int H, int W, const __nv_bfloat16* residual,
                       int relu, __nv_bfloat16* out, GemmLaunch* L) {
  const ConvSpec& s = h->specs[idx];
  ConvGeom& g = L->geom;
  g.n = n; g.H = H; g.W = W; g.Cin = s.cin; g.ksize = s.k; g.stride = s.stride; g.pad = s.pad;
  g.Ho = (H + 2 * s.pad - s.k) / s.stride + 1;
  g.Wo = (W + 2 * s.pad - s.k) / s.stride + 1;
  L->M = n * g.Ho * g.Wo; L->N = s.cout; L->K = s.k * s.k * s.cin;
  L->block_n = pick_block_n(L->M, L->N);
  if (s.k == 1 && s.stride == 1) {
    L->im2col = 0;
    if (make_tmap_tiled_bf16(&L->tmA, x, L->M, L->K, L->K, 128, 64)) return 1;
  } else {
    L->im2col = 1;
    if (make_tmap_im2col_bf16(&L->tmA, x, g, 64, 128)) return 1;
  }
  if (make_tmap_tiled_bf16(&L->tmB, h->wq[idx], L->N, L->K, L->K, L->block_n, 64)) return 1;
  L->epi.scale = h->scale[idx]; L->epi.shift = h->shift[idx];
  L->epi.residual = residual; L->epi.ldr = s.cout;
  L->epi.relu = relu;
  L->epi.out_bf16 = out; L->epi.ldd = s.cout;
  if (use_tma_epilogue() && tma_epilogue_eligible(*L) && enable_tma_epilogue(L)) return 1;
  L->pdl = use_pdl();
  return 0;
}

static int build_stem_gemm(airpose_net* h, int n, GemmLaunch* Lp) {
  // stem: 7 k-blocks (one per vertical tap) over the packed operand, BN + ReLU in the epilogue
  GemmLaunch& L = *Lp;
  L.M = n * 112 * 112; L.N = 64; L.K = kStemK; L.block_n = 64;
  L.stem = 1;
  L.stem_img_rows = 112 * 112;
  L.stem_img_stride = 2 * kStemPlaneRows * 112;
  for (int r = 0; r < 7; ++r) {       // input row 2p-3+r: parity (r+1)&1, plane row p + c_r + 2
    const int par = (r + 1) & 1;
    const int cr = (r - 3 - par) / 2;            // exact: r-3-par is even
    L.stem_tap_off[r] = (par * kStemPlaneRows + 2 + cr) * 112;
  }
  if (make_tmap_tiled_bf16(&L.tmA, h->col, (int64_t)n * 2 * kStemPlaneRows * 112, kStemTapK, kStemTapK, 128, kStemTapK, 64)) return 1;
  if (make_tmap_tiled_bf16(&L.tmB, h->wq[0], 64, kStemK, kStemK, 64, kStemTapK, 64)) return 1;
  L.epi.scale = h->scale[0]; L.epi.shift = h->shift[0]; L.epi.relu = 1;
  L.epi.out_bf16 = h->stem_out; L.epi.ldd = 64;
  return enable_tma_epilogue(&L);
}

// Bottleneck blocks of layers [l0, l1) on `n` images, input in buf[0] at H x H; the last block's
// output goes to `final_out` when given (else stays in one of buf[]).
static int build_blocks(airpose_net* h, int n, int l0, int l1, int H, __nv_bfloat16* const buf[4], __nv_bfloat16* final_out,
                        TrunkPlan* plan) {
  const int layers[4] = {3, 4, 6, 3};
  int idx = 1;
  for (int li = 0; li < l0; ++li) idx += 3 * layers[li] + 1;
  int a = 0, b = 1, c = 2, d = 3;          // buffer roles: X, T1/OUT, T2, DS
  for (int li = l0; li < l1; ++li)
    for (int blk = 0; blk < layers[li]; ++blk) {
      const bool down = blk == 0;
      const bool last = (li == l1 - 1) && (blk == layers[li] - 1);
      const int stride = h->specs[idx + 1].stride;
      const int Ho = H / stride;
      GemmLaunch L1{}, L2{}, L3{}, LD{};
      if (conv_launch(h, idx, buf[a], n, H, H, nullptr, 1, buf[b], &L1)) return 1;
      if (conv_launch(h, idx + 1, buf[b], n, H, H, nullptr, 1, buf[c], &L2)) return 1;
      plan->gemms.push_back(L1);
      plan->gemms.push_back(L2);
      const __nv_bfloat16* res = buf[a];
      if (down) {
        if (conv_launch(h, idx + 3, buf[a], n, H, H, nullptr, 0, buf[d], &LD)) return 1;
        plan->gemms.push_back(LD);
        res = buf[d];
      }
      __nv_bfloat16* out = (last && final_out) ? final_out : buf[b];
      if (conv_launch(h, idx + 2, buf[c], n, Ho, Ho, res, 1, out, &L3)) return 1;
      plan->gemms.push_back(L3);
      plan->final_act = out;
      std::swap(a, b);
      idx += down ? 4 : 3;
      H = Ho;
    }
  return 0;
}

// stage A: stem GEMM + layer1 + layer2 on `n` <= chunk images; output [n,28,28,512] lands in slot
// `slot` of the stage-B input buffer.
static int build_plan_a(airpose_net* h, int n, int slot, TrunkPlan* plan) {
  plan->gemms.clear();
  GemmLaunch L{};
  if (build_stem_gemm(h, n, &L)) return 1;
  plan->gemms.push_back(L);
  __nv_bfloat16* out = slot >= 0 ? h->actB[0] + (size_t)slot * h->chunk * kStageBElems : nullptr;
  return build_blocks(h, n, 0, 2, 56, h->act, out, plan);
}

// stage B: layer3 + layer4 on `n` <= group images, input in actB[0].
static int build_plan_b(airpose_net* h, int n, TrunkPlan* plan) {
  plan->gemms.clear();
  return build_blocks(h, n, 2, 4, 28, h->actB, nullptr, plan);
}

static int launch_stem_front(airpose_net* h, const float* x, int n, const GemmLaunch& stem, __nv_bfloat16* pooled, cudaStream_t st) {
  const int64_t work = (int64_t)n * 2 * kStemPlaneRows * 112;
  stem_pack_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(work, 128), 148 * 32), 128, 0, st>>>(x, n, h->col);
  AP_LAUNCH_CHECK();
  if (launch_gemm(stem, st)) return 1;
  maxpool_kernel<<<(unsigned)std::min<int64_t>(ceil_div64((int64_t)n * 56 * 56 * 8, 256), 148 * 16), 256, 0, st>>>(h->stem_out, n, pooled);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_backbone_fwd(airpose_net_t* h, const float* x, int n_images, float* out_feat, void* stream_) {
  AP_REQUIRE(h && x && out_feat, "airpose_backbone_fwd: null argument");
  AP_REQUIRE(h->loaded, "airpose_backbone_fwd: weights not loaded (call airpose_net_load)");
  AP_REQUIRE(n_images >= 0, "airpose_backbone_fwd: negative image count");
  cudaStream_t st = (cudaStream_t)stream_;
  for (int g0 = 0; g0 < n_images; g0 += h->group) {
    const int ng = std::min(h->group, n_images - g0);
    for (int i0 = 0, slot = 0; i0 < ng; i0 += h->chunk, ++slot) {
      const int n = std::min(h->chunk, ng - i0);
      auto key = std::make_pair(n, slot);
      auto it = h->plansA.find(key);
      if (it == h->plansA.end()) {
        TrunkPlan plan;
        if (build_plan_a(h, n, slot, &plan)) return 1;
        it = h->plansA.emplace(key, std::move(plan)).first;
      }
      const TrunkPlan& plan = it->second;
      if (launch_stem_front(h, x + (size_t)(g0 + i0) * 3 * 224 * 224, n, plan.gemms[0], h->act[0], st)) return 1;
      for (size_t g = 1; g < plan.gemms.size(); ++g)
        if (launch_gemm(plan.gemms[g], st)) return 1;
    }
    auto it = h->plansB.find(ng);
    if (it == h->plansB.end()) {
      TrunkPlan plan;
      if (build_plan_b(h, ng, &plan)) return 1;
      it = h->plansB.emplace(ng, std::move(plan)).first;
    }
    const TrunkPlan& plan = it->second;
    for (size_t g = 0; g < plan.gemms.size(); ++g)
      if (launch_gemm(plan.gemms[g], st)) return 1;
    avgpool_kernel<<<ceil_div(ng * kFeat, 256), 256, 0, st>>>(plan.final_act, ng, out_feat + (size_t)g0 * kFeat);
    AP_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int airpose_backbone_stem(airpose_net_t* h, const float* x, int n, void* out, void* stream_) {
  AP_REQUIRE(h && x && out, "airpose_backbone_stem: null argument");
  AP_REQUIRE(h->loaded, "airpose_backbone_stem: weights not loaded (call airpose_net_load)");
  AP_REQUIRE(n > 0 && n <= h->chunk, "airpose_backbone_stem: n=%d exceeds the chunk size %d", n, h->chunk);
  GemmLaunch stem{};
  if (build_stem_gemm(h, n, &stem)) return 1;
  return launch_stem_front(h, x, n, stem, (__nv_bfloat16*)out, (cudaStream_t)stream_);
}

static int ensure_ief_ws(airpose_net* h, int B, cudaStream_t st) {
  if (B <= h->ief_cap) return 0;
  AP_CHECK_CUDA(cudaStreamSynchronize(st));
  void* old[] = {h->xf_split, h->state_split, h->y1_split, h->y2_split, h->hbuf, h->dbuf, h->pose, h->shape};
  for (void* p : old) cudaFree(p);
  h->ief_plans.clear();
  const size_t M = (size_t)2 * B;
  AP_CHECK_CUDA(cudaMalloc((void**)&h->xf_split, M * 3 * kFeat * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->state_split, M * 3 * kStatePad * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->y1_split, M * 3 * kHid * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->y2_split, M * 3 * kHid * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->hbuf, M * kHid * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->dbuf, M * kDecPad * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->pose, M * 135 * sizeof(float)));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->shape, M * 10 * sizeof(float)));
  h->ief_cap = B;
  return 0;
}

static int make_ief_gemm(GemmLaunch* L, const __nv_bfloat16* A, int M, int K3, const __nv_bfloat16* W, int N) {
  L->M = M; L->N = N; L->K = K3;
  L->block_n = 64;                     // M is small: narrow tiles spread the N dimension over more SMs
  if (make_tmap_tiled_bf16(&L->tmA, A, M, K3, K3, 128, 64)) return 1;
  if (make_tmap_tiled_bf16(&L->tmB, W, N, K3, K3, 64, 64)) return 1;
  return 0;
}

extern "C" int airpose_ief_fwd(airpose_net_t* h, const airpose_ief_args* a, void* stream_) {
  AP_REQUIRE(h && a, "airpose_ief_fwd: null argument");
  AP_REQUIRE(h->loaded, "airpose_ief_fwd: weights not loaded (call airpose_net_load)");
  AP_REQUIRE(a->batch >= 0 && a->iters >= 1, "airpose_ief_fwd: bad batch/iters");
  AP_REQUIRE(a->xf0 && a->xf1 && a->bb0 && a->bb1 && a->pos0 && a->pos1 && a->out_pose0 && a->out_pose1 &&
             a->out_betas0 && a->out_betas1, "airpose_ief_fwd: null tensor");
  const int B = a->batch, M = 2 * B;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  if (ensure_ief_ws(h, B, st)) return 1;
  auto it = h->ief_plans.find(B);
  if (it == h->ief_plans.end()) {
    IefPlan p;
    if (make_ief_gemm(&p.g0, h->xf_split, M, 3 * kFeat, h->w1a, kHid)) return 1;
    p.g0.epi.shift = h->b1; p.g0.epi.out_f32 = h->hbuf; p.g0.epi.ldf = kHid;
    if (make_ief_gemm(&p.g1, h->state_split, M, 3 * kStatePad, h->w1b, kHid)) return 1;
    p.g1.epi.residual = h->hbuf; p.g1.epi.ldr = kHid; p.g1.epi.residual_f32 = 1;
    p.g1.epi.out_split = h->y1_split; p.g1.epi.lds = 3 * kHid;
    if (make_ief_gemm(&p.g2, h->y1_split, M, 3 * kHid, h->w2, kHid)) return 1;
    p.g2.epi.shift = h->b2; p.g2.epi.out_split = h->y2_split; p.g2.epi.lds = 3 * kHid;
    if (make_ief_gemm(&p.g3, h->y2_split, M, 3 * kHid, h->wd, kDecPad)) return 1;
    p.g3.epi.shift = h->bd; p.g3.epi.out_f32 = h->dbuf; p.g3.epi.ldf = kDecPad;
    it = h->ief_plans.emplace(B, p).first;
  }
  const IefPlan& p = it->second;
  const int nthr = 256;
  split_act_kernel<<<ceil_div(B * kFeat, nthr), nthr, 0, st>>>(a->xf0, B, kFeat, kFeat, h->xf_split, kFeat);
  AP_LAUNCH_CHECK();
  split_act_kernel<<<ceil_div(B * kFeat, nthr), nthr, 0, st>>>(a->xf1, B, kFeat, kFeat, h->xf_split + (size_t)B * 3 * kFeat, kFeat);
  AP_LAUNCH_CHECK();
  if (launch_gemm(p.g0, st)) return 1;
  ief_init_kernel<<<ceil_div(M * 145, nthr), nthr, 0, st>>>(B, a->pos0, a->pos1, h->init_pose, a->init_theta0, a->init_theta1,
                                                           a->init_theta_stride, h->init_shape, a->init_shape0,
                                                           a->init_shape1, a->init_shape_stride, h->pose, h->shape);
  AP_LAUNCH_CHECK();
  for (int iter = 0; iter < a->iters; ++iter) {
    ief_state_kernel<<<ceil_div(M * kStatePad, nthr), nthr, 0, st>>>(B, a->bb0, a->bb1, h->pose, h->shape, h->state_split);
    AP_LAUNCH_CHECK();
    if (launch_gemm(p.g1, st)) return 1;
    if (launch_gemm(p.g2, st)) return 1;
    if (launch_gemm(p.g3, st)) return 1;
    const bool last = iter == a->iters - 1;
    ief_update_kernel<<<ceil_div(M * 145, nthr), nthr, 0, st>>>(B, h->dbuf, h->pose, h->shape, last ? a->out_pose0 : nullptr,
                                                             last ? a->out_betas0 : nullptr, last ? a->out_pose1 : nullptr,
                                                             last ? a->out_betas1 : nullptr);
    AP_LAUNCH_CHECK();
  }
  return 0;
}
