// ResNet-50 trunk of copenet, eval mode (the IEF regressor lives in ief.cu).
//
// Replaces (paths relative to /root/reference/copenet/src/copenet/models):
//   model_copenet.py:161-176  copenet.forward_feat_ext  (stem, 16 Bottlenecks :27-47, AvgPool2d(7))
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
#include "net.cuh"
#include "ptx.cuh"

namespace airpose {

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

// The same two kernels for ALL 53 convs in one launch each (blockIdx.y = conv): the per-layer launches were ~6 us of fixed
// cost each for microseconds of work, 106 of them per weight reload (and one reload per training step).  Tables by value.
constexpr int kMaxConvs = 53;
struct PackTab {
  const float* w[kMaxConvs]; __nv_bfloat16* out[kMaxConvs];
  int cout[kMaxConvs], cin[kMaxConvs], k[kMaxConvs], kpad[kMaxConvs];
};
__global__ void pack_conv_weight_all_kernel(const __grid_constant__ PackTab t) {
  const int l = blockIdx.y;
  const float* __restrict__ w = t.w[l];
  __nv_bfloat16* __restrict__ out = t.out[l];
  const int cin = t.cin[l], k = t.k[l], kpad = t.kpad[l];
  const int64_t total = (int64_t)t.cout[l] * kpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / kpad), kk = (int)(i % kpad);
    float val = 0.f;
    if (l == 0) {                     // stem: [o][r][s*3+c], 21 -> 32 zero padded per vertical tap r
      const int r = kk / kStemTapK, e = kk % kStemTapK;
      if (e < 21) val = w[(((int64_t)o * cin + (e % 3)) * k + r) * k + e / 3];
    } else {
      const int tap = kk / cin, c = kk % cin;
      val = w[(((int64_t)o * cin + c) * k * k) + tap];
    }
    out[i] = __float2bfloat16_rn(val);
  }
}
struct FoldTab {
  const float* g[kMaxConvs]; const float* b[kMaxConvs]; const float* mean[kMaxConvs]; const float* var[kMaxConvs];
  float* scale[kMaxConvs]; float* shift[kMaxConvs];
  int c[kMaxConvs];
};
__global__ void fold_bn_all_kernel(const __grid_constant__ FoldTab t, float eps) {
  const int l = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t.c[l]; i += gridDim.x * blockDim.x) {
    const float s = t.g[l][i] / sqrtf(t.var[l][i] + eps);
    t.scale[l][i] = s;
    t.shift[l][i] = t.b[l][i] - t.mean[l][i] * s;
  }
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
// idx (optional, training tape): per output element the window position 3 r + s of its FIRST maximum in scan order
// (PyTorch's tie rule), which is where the backward routes the gradient.
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ x, int n, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx = nullptr) {
  const int64_t total = (int64_t)n * 56 * 56 * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % 8);
    const int64_t pix = i / 8;
    const int q = (int)(pix % 56), pr = (int)((pix / 56) % 56), img = (int)(pix / (56 * 56));
    float m[8];
    uint32_t am[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -INFINITY; am[j] = 0; }
    for (int r = 0; r < 3; ++r) {
      const int h = pr * 2 - 1 + r;
      if (h < 0 || h >= 112) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = q * 2 - 1 + s;
        if (w < 0 || w >= 112) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)img * 112 + h) * 112 + w) * 64 + cg * 8));
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float o = (j & 1) ? __uint_as_float(u[j >> 1] & 0xFFFF0000u) : __uint_as_float(u[j >> 1] << 16);
          if (o > m[j]) { m[j] = o; am[j] = 3 * r + s; }
        }
      }
    }
    if (idx) {
      uint2 pk;
      pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
      pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
      *reinterpret_cast<uint2*>(idx + pix * 64 + cg * 8) = pk;
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

// ------------------------------------------------------------------------------ training-mode BatchNorm
// nn.BatchNorm2d in train mode (model_copenet.py:16-21,57-58; torch semantics): normalise with the BATCH mean and
// biased variance over (n, h, w), update running_mean / running_var (momentum, unbiased variance).  The conv writes its
// raw output z (bf16 [M, C]); bn_stats sums z and z^2 per channel (fp32 per slab, fp64 across slabs: deterministic);
// bn_finalize turns them into scale/shift and updates the running statistics; bn_apply writes
// relu(z * scale + shift (+ residual)) in bf16.  All three are HBM-bound passes over [M, C].
constexpr int kBnSlabs = 148 * 2;
constexpr int kBnRedThreads = 256;                // threads of the two reduction passes (512 measured slower: stats 722 vs 575 us, backward reduce 1204 vs 1003 us per step)

// The BatchNorm kernels form dependent chains of short launches (statistics -> finalize -> apply, 318 launches per training step):
// each starts with griddepcontrol.wait / launch_dependents and goes through launch_chain (common.cuh), so that its CTAs are resident
// and waiting when the kernel before it drains instead of being launched after it: 11.99 -> 11.59 ms per step (gpurun r02t14).
// Every BatchNorm kernel below takes the two views of a two-view tape in ONE launch: blockIdx.y = view, whose rows start
// view_elems elements further on (per-view statistics / partials / coefficients follow the same index).  gridDim.y = 1 and
// view_elems = 0 is the one-view case.
__global__ void __launch_bounds__(kBnRedThreads) bn_stats_kernel(const __nv_bfloat16* __restrict__ z, int64_t M, int C, float* __restrict__ part,
                                                       int64_t view_elems) {
  ptx::grid_dep_wait();      // launched with programmatic stream serialization: nothing of the previous kernel is read before this
  ptx::grid_dep_launch();
  __shared__ float red[kBnRedThreads][17];
  z += (size_t)blockIdx.y * view_elems;
  part += (size_t)blockIdx.y * gridDim.x * C * 2;
  const int groups = C / 8;                         // 8 channels (16 B) per thread
  const int lanes = kBnRedThreads / groups;         // row lanes per block (C <= 2048)
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  const int64_t per = (M + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(M, r0 + per);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  auto add_row = [&](const uint4& v) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = __uint_as_float(u[j] << 16), b = __uint_as_float(u[j] & 0xFFFF0000u);
      s[2 * j] += a; q[2 * j] = fmaf(a, a, q[2 * j]);
      s[2 * j + 1] += b; q[2 * j + 1] = fmaf(b, b, q[2 * j + 1]);
    }
  };
  if (rl < lanes) {
    // four rows in flight per thread (the pass is latency-bound with one), accumulated in row order: same sums as a plain loop
    int64_t r = r0 + rl;
    for (; r + 3 * lanes < r1; r += 4 * lanes) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldg(reinterpret_cast<const uint4*>(z + (r + k * lanes) * C + cg * 8));
#pragma unroll
      for (int k = 0; k < 4; ++k) add_row(v[k]);
    }
    for (; r < r1; r += lanes) add_row(__ldg(reinterpret_cast<const uint4*>(z + r * C + cg * 8)));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][8 + j] = q[j]; }
  __syncthreads();
  if (threadIdx.x < groups) {                       // fixed-order sum over the row lanes
    float ts[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) ts[j] = 0.f;
    for (int l = 0; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 16; ++j) ts[j] += red[l * groups + threadIdx.x][j];
    // partials are laid out [channel][slab][2] so that the finalize kernel's warp reads the slabs of a channel as one
    // contiguous run (it was 32 scattered sectors per load with [slab][channel][2]: ~12 us per finalize launch)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float2*>(part + ((size_t)(threadIdx.x * 8 + j) * gridDim.x + blockIdx.x) * 2) = make_float2(ts[j], ts[8 + j]);
  }
}

// fixed-order sum of the slab partials of one channel by one warp: lane l takes slabs l, l+32, ..., then a shuffle tree
__device__ __forceinline__ void slab_sum(const float* __restrict__ part, int slabs, int C, int c, double& s, double& q) {
  const int lane = threadIdx.x & 31;
  s = 0.0; q = 0.0;
  for (int i = lane; i < slabs; i += 32) {
    const float2 v = *reinterpret_cast<const float2*>(part + ((size_t)c * slabs + i) * 2);
    s += v.x; q += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); q += __shfl_down_sync(0xffffffffu, q, o); }
  s = __shfl_sync(0xffffffffu, s, 0); q = __shfl_sync(0xffffffffu, q, 0);
}

// one warp per channel (blockDim = 128 -> 4 channels per CTA); the views one after the other, so that the running statistics
// receive view 0's update before view 1's (the reference calls forward_feat_ext once per view)
__global__ void bn_finalize_kernel(const float* __restrict__ part, int slabs, int64_t M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save0, float* __restrict__ save1, int views) {
  ptx::grid_dep_wait();      // launched with programmatic stream serialization: nothing of the previous kernel is read before this
  ptx::grid_dep_launch();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  for (int v = 0; v < views; ++v) {
    double s, q;
    slab_sum(part + (size_t)v * slabs * C * 2, slabs, C, c, s, q);
    if ((threadIdx.x & 31) != 0) continue;
    const double mean = s / (double)M;
    const double var = fmax(q / (double)M - mean * mean, 0.0);            // biased, as F.batch_norm normalises
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    // explicit roundings: bn_bwd_* rebuild scale / shift from (gamma, beta, saved mean, saved invstd) with these same two
    // operations to re-derive the ReLU mask from z instead of reading y
    const float sc = __fmul_rn(gamma[c], invstd);
    scale[v * 2048 + c] = sc;
    shift[v * 2048 + c] = __fmaf_rn(-(float)mean, sc, beta[c]);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * (double)M / (double)max((int64_t)1, M - 1));
    }
    float* save = v ? save1 : save0;
    if (save) { save[c] = (float)mean; save[C + c] = invstd; }
  }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ z, int64_t M, int C, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual,
                                                       int relu, __nv_bfloat16* __restrict__ y, int64_t view_elems) {
  ptx::grid_dep_wait();      // launched with programmatic stream serialization: nothing of the previous kernel is read before this
  ptx::grid_dep_launch();
  const int groups = C / 8;
  const int64_t total = M * groups;
  z += (size_t)blockIdx.y * view_elems; y += (size_t)blockIdx.y * view_elems;
  if (residual) residual += (size_t)blockIdx.y * view_elems;
  scale += blockIdx.y * 2048; shift += blockIdx.y * 2048;
  // a thread keeps its channel group for the whole loop (the grid stride is a multiple of 256 and groups divides 256)
  const int cg = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) % groups);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + cg * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + cg * 8 + 4));
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + cg * 8)), h1 = __ldg(reinterpret_cast<const float4*>(shift + cg * 8 + 4));
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(z) + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t ru[4] = {0, 0, 0, 0};
    if (residual) { const uint4 r = __ldg(reinterpret_cast<const uint4*>(residual) + i); ru[0] = r.x; ru[1] = r.y; ru[2] = r.z; ru[3] = r.w; }
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = fmaf(__uint_as_float(u[j] << 16), sc[2 * j], sh[2 * j]) + __uint_as_float(ru[j] << 16);
      float b = fmaf(__uint_as_float(u[j] & 0xFFFF0000u), sc[2 * j + 1], sh[2 * j + 1]) + __uint_as_float(ru[j] & 0xFFFF0000u);
      if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
      o[2 * j] = __float2bfloat16_rn(a); o[2 * j + 1] = __float2bfloat16_rn(b);
    }
    reinterpret_cast<uint4*>(y)[i] = *reinterpret_cast<const uint4*>(o);
  }
}

}  // namespace airpose

using namespace airpose;

// Stage A (56x56 / 28x28 activations) runs in chunks of images, stage B (14x14 / 7x7) on a larger group
// so that its GEMMs have enough 128-row tiles to fill 148 SMs.  Measured on B200 (profiles/): every
// launch costs ~5 us of ramp-up/drain even with PDL, which outweighs L2 residency of smaller chunks
// (128 images: chunk 8 -> 4.63 ms, 16 -> 3.38, 32 -> 3.01, 64 -> 2.73).  (DESIGN.md "trunk schedule")
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  const int c = e ? atoi(e) : dflt;
  return c > 0 ? c : dflt;
}
static bool use_fused_stem() {          // AIRPOSE_NO_FUSED_STEM=1: pack + stem GEMM + max-pool as three launches (A/B runs)
  static const bool on = getenv("AIRPOSE_NO_FUSED_STEM") == nullptr;
  return on;
}
static bool use_fused_tail() {          // AIRPOSE_NO_FUSED_TAIL=1: every conv as its own implicit-GEMM launch (A/B runs)
  static const bool on = getenv("AIRPOSE_NO_FUSED_TAIL") == nullptr;
  return on;
}
// AIRPOSE_TRUNK_SPLIT_B=1: each 64-image chunk runs the WHOLE trunk on its own stream (two independent pipelines whose
// kernels fill each other's ramp-up / drain), instead of joining the chunks into one 128-image stage B.
static bool use_split_b() {
  static const bool on = getenv("AIRPOSE_TRUNK_SPLIT_B") != nullptr && atoi(getenv("AIRPOSE_TRUNK_SPLIT_B")) != 0;
  return on;
}
static int default_chunk() { return env_int("AIRPOSE_TRUNK_CHUNK", 64); }
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
  AP_CHECK_CUDA(cudaMalloc((void**)&h->wq_stem_pairs, stem_pairs_weight_elems() * 2));
  if (ief_create(h)) return 1;
  const size_t act_elems = (size_t)h->chunk * 112 * 112 * 64;       // == 56*56*256, the largest activation
  // two stage-A buffer sets (and a side stream) only when a call can have more than one chunk
  h->sets = (max_images > h->chunk && !getenv("AIRPOSE_TRUNK_ONE_STREAM")) ? airpose_net::kSets : 1;
  for (int s = 0; s < h->sets; ++s) {
    AP_CHECK_CUDA(cudaMalloc((void**)&h->colS[s], (size_t)h->chunk * 2 * kStemPlaneRows * 112 * kStemTapK * 2));
    AP_CHECK_CUDA(cudaMalloc((void**)&h->stem_outS[s], act_elems * 2));
    for (int i = 0; i < 4; ++i) AP_CHECK_CUDA(cudaMalloc((void**)&h->actS[s][i], act_elems * 2));
  }
  if (h->sets > 1) {
    AP_CHECK_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    AP_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    AP_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  for (int i = 0; i < 4; ++i) AP_CHECK_CUDA(cudaMalloc((void**)&h->actB[i], (size_t)h->group * kStageBElems * 2));
  h->split_b = h->sets > 1 && use_split_b();
  if (h->split_b)
    for (int i = 0; i < 4; ++i) AP_CHECK_CUDA(cudaMalloc((void**)&h->actB1[i], (size_t)h->chunk * kStageBElems * 2));
  *out = h;
  return 0;
}

extern "C" int airpose_net_destroy(airpose_net_t* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (auto p : h->wq) cudaFree(p);
  cudaFree(h->wq_stem_pairs);
  for (auto p : h->scale) cudaFree(p);
  for (auto p : h->shift) cudaFree(p);
  ief_destroy(h);
  void* ptrs[] = {h->actB[0], h->actB[1], h->actB[2], h->actB[3], h->actB1[0], h->actB1[1], h->actB1[2], h->actB1[3], h->ztrain, h->bn_part, h->bn_scale, h->bn_shift};
  for (void* p : ptrs) cudaFree(p);
  for (int s = 0; s < airpose_net::kSets; ++s) {
    cudaFree(h->colS[s]); cudaFree(h->stem_outS[s]);
    for (int i = 0; i < 4; ++i) cudaFree(h->actS[s][i]);
  }
  // training tapes and the scratch of the backward pass
  for (auto& tp : h->tape) {
    for (auto p : tp.z) cudaFree(p);
    for (auto p : tp.y) cudaFree(p);
    cudaFree(tp.pooled); cudaFree(tp.pool_idx); cudaFree(tp.stats); cudaFree(tp.stats1);
  }
  for (auto p : h->bw) cudaFree(p);
  cudaFree(h->bw_wd); cudaFree(h->bw_wg);
  cudaFree(h->bw_t0); cudaFree(h->bw_t1); cudaFree(h->bw_w); cudaFree(h->bw_coef);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
  return 0;
}

static int load_trunk(airpose_net_t* h, const airpose_conv_params* conv, float bn_eps, cudaStream_t st) {
  AP_REQUIRE(h->specs.size() <= (size_t)kMaxConvs, "load_trunk: %zu convs exceed the table size", h->specs.size());
  PackTab pt{};
  FoldTab ft{};
  for (size_t i = 0; i < h->specs.size(); ++i) {
    const ConvSpec& s = h->specs[i];
    const airpose_conv_params& c = conv[i];
    AP_REQUIRE(c.weight && c.bn_weight && c.bn_bias && c.bn_mean && c.bn_var, "airpose_net_load: conv %zu has a null parameter", i);
    pt.w[i] = c.weight; pt.out[i] = h->wq[i]; pt.cout[i] = s.cout; pt.cin[i] = s.cin; pt.k[i] = s.k;
    pt.kpad[i] = (i == 0) ? kStemK : s.k * s.k * s.cin;
    ft.g[i] = c.bn_weight; ft.b[i] = c.bn_bias; ft.mean[i] = c.bn_mean; ft.var[i] = c.bn_var;
    ft.scale[i] = h->scale[i]; ft.shift[i] = h->shift[i]; ft.c[i] = s.cout;
  }
  const unsigned nl = (unsigned)h->specs.size();
  pack_conv_weight_all_kernel<<<dim3(64, nl), 256, 0, st>>>(pt);
  AP_LAUNCH_CHECK();
  if (stem_pack_pairs_weight(conv[0].weight, h->wq_stem_pairs, st)) return 1;
  fold_bn_all_kernel<<<dim3(2, nl), 256, 0, st>>>(ft, bn_eps);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_net_load(airpose_net_t* h, const airpose_net_params* p, void* stream_) {
  AP_REQUIRE(h && p, "airpose_net_load: null argument");
  cudaStream_t st = (cudaStream_t)stream_;
  if (load_trunk(h, p->conv, p->bn_eps, st)) return 1;
  if (ief_load(h, p, st)) return 1;
  h->loaded = true;
  h->hmr_loaded = false;
  return 0;
}

// Same trunk, the single-view hmr regressor (model_hmr.py:48-92).
extern "C" int airpose_hmr_load(airpose_net_t* h, const airpose_hmr_params* p, void* stream_) {
  AP_REQUIRE(h && p, "airpose_hmr_load: null argument");
  cudaStream_t st = (cudaStream_t)stream_;
  if (load_trunk(h, p->conv, p->bn_eps, st)) return 1;
  if (ief_load_hmr(h, p, st)) return 1;
  h->loaded = true;          // the trunk entry points work with either regressor
  h->hmr_loaded = true;
  return 0;
}

static int conv_launch(airpose_net* h, int idx, const __nv_bfloat16* x, int n, int H, int W, const __nv_bfloat16* residual,
                       int relu, __nv_bfloat16* out, GemmLaunch* L, bool raw = false) {
  const ConvSpec& s = h->specs[idx];
  ConvGeom& g = L->geom;
  g.n = n; g.H = H; g.W = W; g.Cin = s.cin; g.ksize = s.k; g.stride = s.stride; g.pad = s.pad;
  g.Ho = (H + 2 * s.pad - s.k) / s.stride + 1;
  g.Wo = (W + 2 * s.pad - s.k) / s.stride + 1;
  L->M = n * g.Ho * g.Wo; L->N = s.cout; L->K = s.k * s.k * s.cin;
  L->block_n = pick_block_n(L->M, L->N, L->K);
  if (s.k == 1 && s.stride == 1) {
    L->im2col = 0;
    if (make_tmap_tiled_bf16(&L->tmA, x, L->M, L->K, L->K, 128, 64)) return 1;
  } else {
    L->im2col = 1;
    if (make_tmap_im2col_bf16(&L->tmA, x, g, 64, 128)) return 1;
  }
  L->pair_b_box = use_tma_epilogue() && prefers_pair(L->M, L->N, L->K);
  if (make_tmap_tiled_bf16(&L->tmB, h->wq[idx], L->N, L->K, L->K, L->pair_b_box ? 128 : L->block_n, 64)) return 1;
  L->epi.scale = raw ? nullptr : h->scale[idx]; L->epi.shift = raw ? nullptr : h->shift[idx];   // raw: the conv output itself
  L->epi.residual = residual; L->epi.ldr = s.cout;
  L->epi.relu = relu;
  L->epi.out_bf16 = out; L->epi.ldd = s.cout;
  if (use_tma_epilogue() && tma_epilogue_eligible(*L) && enable_tma_epilogue(L)) return 1;
  if (L->pair_b_box && !L->tma_epi) {
    L->pair_b_box = 0;
    if (make_tmap_tiled_bf16(&L->tmB, h->wq[idx], L->N, L->K, L->K, L->block_n, 64)) return 1;
  }
  L->pdl = use_pdl();
  return 0;
}

static int build_stem_gemm(airpose_net* h, int n, int set, GemmLaunch* Lp, bool raw = false) {
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
  if (make_tmap_tiled_bf16(&L.tmA, h->colS[set], (int64_t)n * 2 * kStemPlaneRows * 112, kStemTapK, kStemTapK, 128, kStemTapK, 64)) return 1;
  if (make_tmap_tiled_bf16(&L.tmB, h->wq[0], 64, kStemK, kStemK, 64, kStemTapK, 64)) return 1;
  L.epi.scale = raw ? nullptr : h->scale[0]; L.epi.shift = raw ? nullptr : h->shift[0]; L.epi.relu = raw ? 0 : 1;
  L.epi.out_bf16 = h->stem_outS[set]; L.epi.ldd = 64;
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
      auto push = [&](const GemmLaunch& L) { plan->ops.push_back({0, (int)plan->gemms.size()}); plan->gemms.push_back(L); };
      if (conv_launch(h, idx, buf[a], n, H, H, nullptr, 1, buf[b], &L1)) return 1;
      push(L1);
      const __nv_bfloat16* res = buf[a];
      if (down) {
        if (conv_launch(h, idx + 3, buf[a], n, H, H, nullptr, 0, buf[d], &LD)) return 1;
        push(LD);
        res = buf[d];
      }
      __nv_bfloat16* out = (last && final_out) ? final_out : buf[c];
      const ConvSpec& s2 = h->specs[idx + 1];
      const ConvSpec& s3 = h->specs[idx + 2];
      if (use_fused_tail() && stride == 1 && bneck_tail_supported(H, H, s2.cout, s3.cout)) {
        // conv2 + conv3 in one launch (bneck.cu): conv1's output never comes back from HBM nine times, conv2's never leaves the SM
        TailLaunch T{};
        if (build_bneck_tail(&T, buf[b], h->wq[idx + 1], h->scale[idx + 1], h->shift[idx + 1], h->wq[idx + 2], h->scale[idx + 2],
                             h->shift[idx + 2], res, out, n, H, H)) return 1;
        plan->ops.push_back({1, (int)plan->tails.size()});
        plan->tails.push_back(T);
        // roles: X <- out; the old X and T1 become scratch
        if (out == buf[c]) { std::swap(a, c); }
      } else {
        if (out == buf[c]) out = buf[b];               // unfused: conv3 may overwrite T1
        if (conv3x3_slab_supported(H, H, s2.cin, s2.cout, s2.k, s2.stride, s2.pad)) {
          SlabLaunch S{};                              // input band resident in shared memory instead of nine im2col passes
          if (build_conv3x3_slab(&S, buf[b], h->wq[idx + 1], h->scale[idx + 1], h->shift[idx + 1], 1, buf[c], n, H, H)) return 1;
          plan->ops.push_back({2, (int)plan->slabs.size()});
          plan->slabs.push_back(S);
        } else {
          if (conv_launch(h, idx + 1, buf[b], n, H, H, nullptr, 1, buf[c], &L2)) return 1;
          push(L2);
        }
        if (conv_launch(h, idx + 2, buf[c], n, Ho, Ho, res, 1, out, &L3)) return 1;
        push(L3);
        if (out == buf[b]) std::swap(a, b);
      }
      plan->final_act = out;
      idx += down ? 4 : 3;
      H = Ho;
    }
  return 0;
}

// stage A: stem GEMM + layer1 + layer2 on `n` <= chunk images; output [n,28,28,512] lands at image
// offset `first` of the stage-B input buffer.
static int build_plan_a(airpose_net* h, int n, int first, int set, TrunkPlan* plan) {
  plan->gemms.clear(); plan->tails.clear(); plan->slabs.clear(); plan->ops.clear();
  GemmLaunch L{};
  if (build_stem_gemm(h, n, set, &L)) return 1;
  plan->gemms.push_back(L);
  if (use_fused_stem() && build_stem_pool(&plan->stem, h->colS[set], h->wq_stem_pairs, h->scale[0], h->shift[0], h->actS[set][0], n)) return 1;
  __nv_bfloat16* out = h->split_b ? (set ? h->actB1[0] : h->actB[0]) : h->actB[0] + (size_t)first * kStageBElems;
  return build_blocks(h, n, 0, 2, 56, h->actS[set], out, plan);
}

// stage B: layer3 + layer4 on `n` <= group images, input in actB[0].
static int build_plan_b(airpose_net* h, int n, int set, TrunkPlan* plan) {
  plan->gemms.clear(); plan->tails.clear(); plan->slabs.clear(); plan->ops.clear();
  return build_blocks(h, n, 2, 4, 28, set ? h->actB1 : h->actB, nullptr, plan);
}

static int launch_plan_ops(const TrunkPlan& plan, cudaStream_t st) {
  for (const PlanOp& op : plan.ops) {
    const int rc = op.kind == 0 ? launch_gemm(plan.gemms[op.idx], st)
                 : op.kind == 1 ? launch_bneck_tail(plan.tails[op.idx], st) : launch_conv3x3_slab(plan.slabs[op.idx], st);
    if (rc) return 1;
  }
  return 0;
}

static int launch_stem_front(airpose_net* h, const float* x, int n, int set, const GemmLaunch& stem, __nv_bfloat16* pooled, cudaStream_t st) {
  const int64_t work = (int64_t)n * 2 * kStemPlaneRows * 112;
  stem_pack_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(work, 128), 148 * 32), 128, 0, st>>>(x, n, h->colS[set]);
  AP_LAUNCH_CHECK();
  if (launch_gemm(stem, st)) return 1;
  maxpool_kernel<<<(unsigned)std::min<int64_t>(ceil_div64((int64_t)n * 56 * 56 * 8, 256), 148 * 16), 256, 0, st>>>(h->stem_outS[set], n, pooled);
  AP_LAUNCH_CHECK();
  return 0;
}

// Images [0, n0) come from x0, images [n0, n_images) from x1 (the two views of a batch of pairs).
static int backbone_fwd_segments(airpose_net_t* h, const float* x0, int n0, const float* x1, int n_images, float* out_feat,
                                 cudaStream_t st) {
  const size_t img = (size_t)3 * 224 * 224;
  auto src = [&](int i) { return i < n0 ? x0 + (size_t)i * img : x1 + (size_t)(i - n0) * img; };
  for (int g0 = 0; g0 < n_images; g0 += h->group) {
    const int ng = std::min(h->group, n_images - g0);
    // Stage A: consecutive chunks alternate between two buffer sets on two streams (fork/join with events).
    // The kernels are persistent, one CTA per SM, so two of them never share an SM -- but the CTAs of the
    // second stream's kernel start on every SM the first one's tail has already left, which hides the
    // ~5 us of ramp-up/drain each launch costs.  Stream order keeps each set's reuse safe.
    const bool fork = h->sets > 1 && ng > h->chunk;
    static const int stage_a_cap = env_int("AIRPOSE_STAGEA_GRID", 0);
    if (fork) {
      set_grid_cap(stage_a_cap);
      AP_CHECK_CUDA(cudaEventRecord(h->ev_fork, st));
      AP_CHECK_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    }
    int chunk_no = 0;
    for (int i0 = 0, n = 0; i0 < ng; i0 += n, ++chunk_no) {
      n = std::min(h->chunk, ng - i0);
      const int first = g0 + i0;
      if (first < n0 && first + n > n0) n = n0 - first;      // a chunk never straddles the two input tensors
      const int set = fork ? (chunk_no & 1) : 0;
      cudaStream_t cst = set ? h->side_stream : st;
      auto key = std::make_pair(n, 2 * i0 + set);
      auto it = h->plansA.find(key);
      if (it == h->plansA.end()) {
        TrunkPlan plan;
        if (build_plan_a(h, n, i0, set, &plan)) return 1;
        it = h->plansA.emplace(key, std::move(plan)).first;
      }
      const TrunkPlan& plan = it->second;
      if (plan.stem.valid) {
        if (launch_stem_pool(plan.stem, src(first), h->colS[set], n, cst)) return 1;
      } else if (launch_stem_front(h, src(first), n, set, plan.gemms[0], h->actS[set][0], cst)) return 1;
      if (launch_plan_ops(plan, cst)) return 1;
      if (h->split_b) {                     // this chunk's stage B right behind its stage A, on the same stream
        auto itb = h->plansB.find(2 * n + set);
        if (itb == h->plansB.end()) {
          TrunkPlan pb;
          if (build_plan_b(h, n, set, &pb)) return 1;
          itb = h->plansB.emplace(2 * n + set, std::move(pb)).first;
        }
        if (launch_plan_ops(itb->second, cst)) return 1;
        avgpool_kernel<<<ceil_div(n * kFeat, 256), 256, 0, cst>>>(itb->second.final_act, n, out_feat + (size_t)first * kFeat);
        AP_LAUNCH_CHECK();
      }
    }
    set_grid_cap(0);
    if (fork) {
      AP_CHECK_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
      AP_CHECK_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    }
    if (h->split_b) continue;
    auto it = h->plansB.find(2 * ng);
    if (it == h->plansB.end()) {
      TrunkPlan plan;
      if (build_plan_b(h, ng, 0, &plan)) return 1;
      it = h->plansB.emplace(2 * ng, std::move(plan)).first;
    }
    const TrunkPlan& plan = it->second;
    if (launch_plan_ops(plan, st)) return 1;
    avgpool_kernel<<<ceil_div(ng * kFeat, 256), 256, 0, st>>>(plan.final_act, ng, out_feat + (size_t)g0 * kFeat);
    AP_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int airpose_backbone_fwd(airpose_net_t* h, const float* x, int n_images, float* out_feat, void* stream_) {
  AP_REQUIRE(h && x && out_feat, "airpose_backbone_fwd: null argument");
  AP_REQUIRE(h->loaded, "airpose_backbone_fwd: weights not loaded (call airpose_net_load)");
  AP_REQUIRE(n_images >= 0, "airpose_backbone_fwd: negative image count");
  return backbone_fwd_segments(h, x, n_images, x, n_images, out_feat, (cudaStream_t)stream_);
}

extern "C" int airpose_backbone_fwd_pair(airpose_net_t* h, const float* x0, const float* x1, int n_pairs, float* out_feat,
                                         void* stream_) {
  AP_REQUIRE(h && x0 && x1 && out_feat, "airpose_backbone_fwd_pair: null argument");
  AP_REQUIRE(h->loaded, "airpose_backbone_fwd_pair: weights not loaded (call airpose_net_load)");
  AP_REQUIRE(n_pairs >= 0, "airpose_backbone_fwd_pair: negative pair count");
  return backbone_fwd_segments(h, x0, n_pairs, x1, 2 * n_pairs, out_feat, (cudaStream_t)stream_);
}

extern "C" int airpose_backbone_stem(airpose_net_t* h, const float* x, int n, void* out, void* stream_) {
  AP_REQUIRE(h && x && out, "airpose_backbone_stem: null argument");
  AP_REQUIRE(h->loaded, "airpose_backbone_stem: weights not loaded (call airpose_net_load)");
  AP_REQUIRE(n > 0 && n <= h->chunk, "airpose_backbone_stem: n=%d exceeds the chunk size %d", n, h->chunk);
  if (use_fused_stem()) {
    StemLaunch sl{};
    if (build_stem_pool(&sl, h->colS[0], h->wq_stem_pairs, h->scale[0], h->shift[0], out, n)) return 1;
    return launch_stem_pool(sl, x, h->colS[0], n, (cudaStream_t)stream_);
  }
  GemmLaunch stem{};
  if (build_stem_gemm(h, n, 0, &stem)) return 1;
  return launch_stem_front(h, x, n, 0, stem, (__nv_bfloat16*)out, (cudaStream_t)stream_);
}

// ------------------------------------------------------------------------------ training-mode trunk forward
// copenet.forward_feat_ext with the module in train() mode (what Lightning's training_step runs, and what the frozen
// trunk of `train_reg_only` runs too): every BatchNorm normalises with the statistics of THIS batch of n images and
// updates its running statistics.  Per conv: raw GEMM -> bn_stats -> bn_finalize -> bn_apply (+residual, ReLU).
// One call = one view (the reference calls forward_feat_ext once per view, so the statistics are per view).
// M = rows per view; views = 2: rows [M, 2M) of z / residual / y are view 1 (two-view tape), statistics saved to save1
static int bn_train(airpose_net* h, int idx, const __nv_bfloat16* z, int64_t M, int C, const airpose_bn_train_params* bn,
                    const __nv_bfloat16* residual, int relu, __nv_bfloat16* y, cudaStream_t st, int views = 1, float* save1_base = nullptr) {
  const int64_t ve = M * C;
  AP_REQUIRE(C % 8 == 0 && 256 % (C / 8) == 0, "bn_train: C=%d must be 8 x a power of two <= 2048 (the kernels keep one channel group per thread)", C);
  AP_CHECK_CUDA(launch_chain(bn_stats_kernel, dim3(kBnSlabs, views), dim3(kBnRedThreads), st, z, M, C, h->bn_part, ve));
  float* save = bn->saved_stats ? bn->saved_stats + h->bn_save_off[idx] : nullptr;
  float* save1 = save1_base ? save1_base + h->bn_save_off[idx] : nullptr;
  AP_CHECK_CUDA(launch_chain(bn_finalize_kernel, dim3(ceil_div(C, 4)), dim3(128), st, h->bn_part, kBnSlabs, M, C, bn->bn_weight[idx],
                             bn->bn_bias[idx], bn->eps, bn->momentum, bn->running_mean[idx], bn->running_var[idx], h->bn_scale,
                             h->bn_shift, save, save1, views));
  const int64_t total = M * (C / 8);
  AP_CHECK_CUDA(launch_chain(bn_apply_kernel, dim3((unsigned)std::min<int64_t>(ceil_div64(total, 256), 148 * 16), views), dim3(256), st, z, M,
                             C, h->bn_scale, h->bn_shift, residual, relu, y, ve));
  return 0;
}

static int backbone_fwd_train_tape(airpose_net* h, const float* x, const float* x1, int n, int views, const airpose_bn_train_params* bn,
                                   float* out_feat, cudaStream_t st);

static int train_scratch_reserve(airpose_net* h);
static int backbone_fwd_train_notape(airpose_net* h, const float* x, int n, const airpose_bn_train_params* bn, float* out_feat,
                                     __nv_bfloat16* Z, __nv_bfloat16* const* buf, cudaStream_t st);

extern "C" int64_t airpose_bn_saved_stats_floats(void) {
  int64_t n = 0;
  for (const ConvSpec& s : resnet50_specs()) n += 2 * s.cout;
  return n;
}

extern "C" int airpose_backbone_fwd_train(airpose_net_t* h, const float* x, int n, const airpose_bn_train_params* bn, float* out_feat,
                                          void* stream_) {
  AP_REQUIRE(h && x && bn && out_feat, "airpose_backbone_fwd_train: null argument");
  AP_REQUIRE(h->loaded, "airpose_backbone_fwd_train: weights not loaded (call airpose_net_load)");
  // n = 1 is legal, as in the reference: BatchNorm2d takes its statistics over N*H*W (49 values per channel at layer4)
  AP_REQUIRE(n >= 1 && n <= h->chunk, "airpose_backbone_fwd_train: n=%d must be in [1, %d] (one call holds one chunk of images on its "
             "tape; the batch statistics of a larger batch would span chunks)", n, h->chunk);
  for (size_t i = 0; i < h->specs.size(); ++i)
    AP_REQUIRE(bn->bn_weight[i] && bn->bn_bias[i], "airpose_backbone_fwd_train: BatchNorm %zu has a null parameter", i);
  cudaStream_t st = (cudaStream_t)stream_;
  if (train_scratch_reserve(h)) return 1;
  if (bn->tape >= 0) return backbone_fwd_train_tape(h, x, nullptr, n, 1, bn, out_feat, st);
  __nv_bfloat16* Z = h->ztrain;
  __nv_bfloat16* const* buf = h->actS[0];
  return backbone_fwd_train_notape(h, x, n, bn, out_feat, Z, buf, st);
}

static int train_scratch_reserve(airpose_net* h) {
  const size_t act_elems = (size_t)h->chunk * 112 * 112 * 64;
  if (!h->ztrain) {
    AP_CHECK_CUDA(cudaMalloc((void**)&h->ztrain, act_elems * 2));
    AP_CHECK_CUDA(cudaMalloc((void**)&h->bn_part, (size_t)2 * kBnSlabs * 2048 * 2 * sizeof(float)));
    AP_CHECK_CUDA(cudaMalloc((void**)&h->bn_scale, 2 * 2048 * sizeof(float)));
    AP_CHECK_CUDA(cudaMalloc((void**)&h->bn_shift, 2 * 2048 * sizeof(float)));
    int64_t off = 0;
    h->bn_save_off.clear();
    for (const ConvSpec& s : h->specs) { h->bn_save_off.push_back(off); off += 2 * s.cout; }
  }
  return 0;
}

static int backbone_fwd_train_notape(airpose_net* h, const float* x, int n, const airpose_bn_train_params* bn, float* out_feat,
                                     __nv_bfloat16* Z, __nv_bfloat16* const* buf, cudaStream_t st) {
  // stem: pack, raw 7x7 conv, BN + ReLU in place, max-pool
  {
    const int64_t work = (int64_t)n * 2 * kStemPlaneRows * 112;
    stem_pack_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(work, 128), 148 * 32), 128, 0, st>>>(x, n, h->colS[0]);
    AP_LAUNCH_CHECK();
    GemmLaunch L{};
    if (build_stem_gemm(h, n, 0, &L, true)) return 1;
    if (launch_gemm(L, st)) return 1;
    const int64_t M = (int64_t)n * 112 * 112;
    if (bn_train(h, 0, h->stem_outS[0], M, 64, bn, nullptr, 1, h->stem_outS[0], st)) return 1;
    maxpool_kernel<<<(unsigned)std::min<int64_t>(ceil_div64((int64_t)n * 56 * 56 * 8, 256), 148 * 16), 256, 0, st>>>(h->stem_outS[0], n,
                                                                                                                     buf[0]);
    AP_LAUNCH_CHECK();
  }
  const int layers[4] = {3, 4, 6, 3};
  int idx = 1, H = 56;
  int a = 0, b = 1, c = 2, d = 3;          // buffer roles: X, T1/OUT, T2, DS
  for (int li = 0; li < 4; ++li)
    for (int blk = 0; blk < layers[li]; ++blk) {
      const bool down = blk == 0;
      const int stride = h->specs[idx + 1].stride;
      const int Ho = H / stride;
      const int planes = h->specs[idx].cout;
      GemmLaunch L1{}, L2{}, L3{}, LD{};
      if (conv_launch(h, idx, buf[a], n, H, H, nullptr, 0, Z, &L1, true) || launch_gemm(L1, st)) return 1;
      if (bn_train(h, idx, Z, (int64_t)n * H * H, planes, bn, nullptr, 1, buf[b], st)) return 1;
      if (conv_launch(h, idx + 1, buf[b], n, H, H, nullptr, 0, Z, &L2, true) || launch_gemm(L2, st)) return 1;
      if (bn_train(h, idx + 1, Z, (int64_t)n * Ho * Ho, planes, bn, nullptr, 1, buf[c], st)) return 1;
      const __nv_bfloat16* res = buf[a];
      if (down) {
        if (conv_launch(h, idx + 3, buf[a], n, H, H, nullptr, 0, Z, &LD, true) || launch_gemm(LD, st)) return 1;
        if (bn_train(h, idx + 3, Z, (int64_t)n * Ho * Ho, planes * 4, bn, nullptr, 0, buf[d], st)) return 1;
        res = buf[d];
      }
      if (conv_launch(h, idx + 2, buf[c], n, Ho, Ho, nullptr, 0, Z, &L3, true) || launch_gemm(L3, st)) return 1;
      if (bn_train(h, idx + 2, Z, (int64_t)n * Ho * Ho, planes * 4, bn, res, 1, buf[b], st)) return 1;
      std::swap(a, b);
      idx += down ? 4 : 3;
      H = Ho;
    }
  avgpool_kernel<<<ceil_div(n * kFeat, 256), 256, 0, st>>>(buf[a], n, out_feat);
  AP_LAUNCH_CHECK();
  return 0;
}

// Only the trunk part of airpose_net_load (packed bf16 conv weights, folded eval BatchNorm): what a full training step
// invalidates every iteration; the collapsed regressor matrix is re-formed lazily by the next eval-mode regressor call.
extern "C" int airpose_net_load_trunk(airpose_net_t* h, const airpose_net_params* p, void* stream_) {
  AP_REQUIRE(h && p, "airpose_net_load_trunk: null argument");
  if (load_trunk(h, p->conv, p->bn_eps, (cudaStream_t)stream_)) return 1;
  h->loaded = true;
  return 0;
}

// Only the regressor part of airpose_net_load (the collapsed matrix G): what changes between the steps of a
// regressor-only training run.
extern "C" int airpose_net_load_regressor(airpose_net_t* h, const airpose_net_params* p, void* stream_) {
  AP_REQUIRE(h && p, "airpose_net_load_regressor: null argument");
  AP_REQUIRE(h->loaded && !h->hmr_loaded, "airpose_net_load_regressor: the two-view network is not loaded");
  return ief_load(h, p, (cudaStream_t)stream_);
}

// ================================================================================================ trunk backward
// (see airpose_backbone_bwd_train in include/airpose_b200.h)
namespace airpose {

// ---- small elementwise / layout kernels of the backward pass
// g_y[n,7,7,C] = g_feat[n,C] / 49   (AvgPool2d(7) backward)
__global__ void avgpool_bwd_kernel(const float* __restrict__ g_feat, int n, __nv_bfloat16* __restrict__ gy) {
  const int64_t total = (int64_t)n * 49 * kFeat;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % kFeat);
    const int img = (int)(i / (49 * kFeat));
    gy[i] = __float2bfloat16_rn(g_feat[(size_t)img * kFeat + c] * (1.f / 49.f));
  }
}

// partial sums of dpre = dy * [y > 0] and dpre * xhat per channel (xhat = (z - mean) * invstd)
template <bool kZMask>
__global__ void __launch_bounds__(kBnRedThreads) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                            const __nv_bfloat16* __restrict__ z, const float* __restrict__ stats0,
                                                            const float* __restrict__ stats1, int64_t M, int C, float* __restrict__ part,
                                                            int64_t view_elems, const float* __restrict__ gamma, const float* __restrict__ beta) {
  ptx::grid_dep_wait();      // launched with programmatic stream serialization: nothing of the previous kernel is read before this
  ptx::grid_dep_launch();
  __shared__ float red[kBnRedThreads][17];
  const float* __restrict__ stats = blockIdx.y ? stats1 : stats0;
  dy += (size_t)blockIdx.y * view_elems; z += (size_t)blockIdx.y * view_elems;
  if (y) y += (size_t)blockIdx.y * view_elems;
  part += (size_t)blockIdx.y * gridDim.x * C * 2;
  const int groups = C / 8, lanes = kBnRedThreads / groups;
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  const int64_t per = (M + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(M, r0 + per);
  float s[8], q[8], mean[8], istd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = q[j] = 0.f; mean[j] = stats[cg * 8 + j]; istd[j] = stats[C + cg * 8 + j]; }
  // gamma != null (a ReLU without residual): the mask [y > 0] is re-derived from z -- y = relu(bf16(z scale + shift)) with scale
  // and shift rebuilt exactly as bn_finalize_kernel rounds them -- and y is not read at all
  constexpr bool zmask = kZMask;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = zmask ? __fmul_rn(gamma[cg * 8 + j], istd[j]) : 0.f;
    sh[j] = zmask ? __fmaf_rn(-mean[j], sc[j], beta[cg * 8 + j]) : 0.f;
  }
  auto add_row = [&](const uint4& vd, const uint4& vz, const uint4& vy) {
    const uint32_t ud[4] = {vd.x, vd.y, vd.z, vd.w}, uz[4] = {vz.x, vz.y, vz.z, vz.w}, uy[4] = {vy.x, vy.y, vy.z, vy.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z0 = __uint_as_float(uz[j] << 16), z1 = __uint_as_float(uz[j] & 0xFFFF0000u);
      const float y0 = zmask ? __bfloat162float(__float2bfloat16_rn(fmaf(z0, sc[2 * j], sh[2 * j]))) : __uint_as_float(uy[j] << 16);
      const float y1 = zmask ? __bfloat162float(__float2bfloat16_rn(fmaf(z1, sc[2 * j + 1], sh[2 * j + 1]))) : __uint_as_float(uy[j] & 0xFFFF0000u);
      const float d0 = y0 > 0.f ? __uint_as_float(ud[j] << 16) : 0.f;
      const float d1 = y1 > 0.f ? __uint_as_float(ud[j] & 0xFFFF0000u) : 0.f;
      const float x0 = (z0 - mean[2 * j]) * istd[2 * j];
      const float x1 = (z1 - mean[2 * j + 1]) * istd[2 * j + 1];
      s[2 * j] += d0; q[2 * j] = fmaf(d0, x0, q[2 * j]);
      s[2 * j + 1] += d1; q[2 * j + 1] = fmaf(d1, x1, q[2 * j + 1]);
    }
  };
  const uint4 pos = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);     // "positive" when there is no ReLU
  // (three rows in flight per thread was measured SLOWER here: 1452 vs 983 us over the 53 launches of a step, profiles/r02t4)
  if (rl < lanes)
    for (int64_t r = r0 + rl; r < r1; r += lanes) {
      const int64_t o = r * C + cg * 8;
      add_row(__ldg(reinterpret_cast<const uint4*>(dy + o)), __ldg(reinterpret_cast<const uint4*>(z + o)),
              (y && !zmask) ? __ldg(reinterpret_cast<const uint4*>(y + o)) : pos);
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][8 + j] = q[j]; }
  __syncthreads();
  if (threadIdx.x < groups) {
    float ts[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) ts[j] = 0.f;
    for (int l = 0; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 16; ++j) ts[j] += red[l * groups + threadIdx.x][j];
    // partials are laid out [channel][slab][2] so that the finalize kernel's warp reads the slabs of a channel as one
    // contiguous run (it was 32 scattered sectors per load with [slab][channel][2]: ~12 us per finalize launch)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float2*>(part + ((size_t)(threadIdx.x * 8 + j) * gridDim.x + blockIdx.x) * 2) = make_float2(ts[j], ts[8 + j]);
  }
}

// dgamma, dbeta (fp32, overwritten or accumulated; the views add in order) and per view the coefficients of dz = c1 (dpre - c2 - xhat c3)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int slabs, int64_t M, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ stats0, const float* __restrict__ stats1, float* __restrict__ g_gamma,
                                       float* __restrict__ g_beta, int accumulate, float* __restrict__ coef, int views) {
  ptx::grid_dep_wait();      // launched with programmatic stream serialization: nothing of the previous kernel is read before this
  ptx::grid_dep_launch();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per channel
  if (c >= C) return;
  for (int v = 0; v < views; ++v) {
    double s, q;
    slab_sum(part + (size_t)v * slabs * C * 2, slabs, C, c, s, q);
    if ((threadIdx.x & 31) != 0) continue;
    const bool acc = accumulate || v > 0;
    if (g_gamma) { g_gamma[c] = (float)q + (acc ? g_gamma[c] : 0.f); g_beta[c] = (float)s + (acc ? g_beta[c] : 0.f); }
    float* cf = coef + v * 3 * 2048;
    cf[c] = gamma[c] * (v ? stats1 : stats0)[C + c];
    cf[2048 + c] = (float)(s / (double)M);
    cf[4096 + c] = (float)(q / (double)M);
  }
}

template <bool kZMask>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                           const __nv_bfloat16* __restrict__ z, const float* __restrict__ stats0,
                                                           const float* __restrict__ stats1, const float* __restrict__ coef, int64_t M, int C,
                                                           __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ dpre_out,
                                                           int64_t view_elems, const float* __restrict__ gamma, const float* __restrict__ beta) {
  ptx::grid_dep_wait();      // launched with programmatic stream serialization: nothing of the previous kernel is read before this
  ptx::grid_dep_launch();
  constexpr bool zmask = kZMask;                     // as in bn_bwd_reduce_kernel: the ReLU mask from z, y is not read
  const int groups = C / 8;
  const int64_t total = M * groups;
  const float* __restrict__ stats = blockIdx.y ? stats1 : stats0;
  coef += blockIdx.y * 3 * 2048;
  dy += (size_t)blockIdx.y * view_elems; z += (size_t)blockIdx.y * view_elems; dz += (size_t)blockIdx.y * view_elems;
  if (y) y += (size_t)blockIdx.y * view_elems;
  if (dpre_out) dpre_out += (size_t)blockIdx.y * view_elems;
  // a thread keeps its channel group for the whole loop (the grid stride is a multiple of 256 and groups divides 256): the five
  // per-channel tables are loaded once (they were ten 16-byte loads and a 64-bit modulo per 8 elements)
  const int cg = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) % groups);
  float mean[8], istd[8], c1[8], c2[8], c3[8], gm[8], bt[8];
  {
    const float* tabs[5] = {stats + cg * 8, stats + C + cg * 8, coef + cg * 8, coef + 2048 + cg * 8, coef + 4096 + cg * 8};
    float* dsts[5] = {mean, istd, c1, c2, c3};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(tabs[k])), b = __ldg(reinterpret_cast<const float4*>(tabs[k]) + 1);
      dsts[k][0] = a.x; dsts[k][1] = a.y; dsts[k][2] = a.z; dsts[k][3] = a.w;
      dsts[k][4] = b.x; dsts[k][5] = b.y; dsts[k][6] = b.z; dsts[k][7] = b.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { gm[j] = zmask ? gamma[cg * 8 + j] : 0.f; bt[j] = zmask ? beta[cg * 8 + j] : 0.f; }
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 vd = __ldg(reinterpret_cast<const uint4*>(dy) + i);
    const uint4 vz = __ldg(reinterpret_cast<const uint4*>(z) + i);
    uint4 vy = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    if (y && !zmask) vy = __ldg(reinterpret_cast<const uint4*>(y) + i);
    const uint32_t ud[4] = {vd.x, vd.y, vd.z, vd.w}, uz[4] = {vz.x, vz.y, vz.z, vz.w}, uy[4] = {vy.x, vy.y, vy.z, vy.w};
    __align__(16) __nv_bfloat16 o[8], p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t wd = ud[j >> 1], wz = uz[j >> 1], wy = uy[j >> 1];
      const float dv = (j & 1) ? __uint_as_float(wd & 0xFFFF0000u) : __uint_as_float(wd << 16);
      const float zv = (j & 1) ? __uint_as_float(wz & 0xFFFF0000u) : __uint_as_float(wz << 16);
      float yv = (j & 1) ? __uint_as_float(wy & 0xFFFF0000u) : __uint_as_float(wy << 16);
      if (zmask) {
        const float sc = __fmul_rn(gm[j], istd[j]);
        yv = __bfloat162float(__float2bfloat16_rn(fmaf(zv, sc, __fmaf_rn(-mean[j], sc, bt[j]))));
      }
      const float dp = yv > 0.f ? dv : 0.f;
      const float xh = (zv - mean[j]) * istd[j];
      o[j] = __float2bfloat16_rn(c1[j] * (dp - c2[j] - xh * c3[j]));
      p[j] = __float2bfloat16_rn(dp);
    }
    reinterpret_cast<uint4*>(dz)[i] = *reinterpret_cast<const uint4*>(o);
    if (dpre_out) reinterpret_cast<uint4*>(dpre_out)[i] = *reinterpret_cast<const uint4*>(p);
  }
}

// K-major operands of the weight-gradient GEMMs.  One kernel, two sources:
//   kIm2col = false   out[c][m]        = in[m][c]                                      (a plain transpose)
//   kIm2col = true    out[(tap, c)][m] = x[n, ho*s + r - pad, wo*s + sx - pad, c]      (zero outside; tap = blockIdx.z)
// bf16, 64 pixels x 64 channels per CTA through shared memory: 16-byte loads along the channels, 16-byte stores along the
// pixels (128 contiguous bytes per 8 lanes on both sides).  ld = row pitch of out (a multiple of 8, >= M).
template <bool kIm2col>
__global__ void __launch_bounds__(256) transpose64_kernel(const __nv_bfloat16* __restrict__ in, int64_t M, int C, __nv_bfloat16* __restrict__ out,
                                                          int64_t ld, int H, int W, int k, int stride, int pad, int Ho, int Wo) {
  __shared__ __align__(16) uint32_t tileT[64][33];          // [channel][pixel pair] (+1 word: conflict-free 4-byte reads below)
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int tap = kIm2col ? blockIdx.z : 0;
  const int t = threadIdx.x;
  {
    const int cg = t & 7;
    __nv_bfloat16* tp = reinterpret_cast<__nv_bfloat16*>(&tileT[0][0]);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int ml = (t >> 3) + rr * 32;
      const int64_t m = m0 + ml;
      const int c = c0 + cg * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (m < M && c < C) {
        if (!kIm2col) {
          v = __ldg(reinterpret_cast<const uint4*>(in + m * C + c));
        } else {
          const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho), img = (int)(m / ((int64_t)Wo * Ho));
          const int h = ho * stride + tap / k - pad, w = wo * stride + tap % k - pad;
          if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(reinterpret_cast<const uint4*>(in + (((int64_t)img * H + h) * W + w) * C + c));
        }
      }
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 8; ++j)
        tp[(cg * 8 + j) * 66 + ml] = __ushort_as_bfloat16((unsigned short)((j & 1) ? (u[j >> 1] >> 16) : (u[j >> 1] & 0xFFFFu)));
    }
  }
  __syncthreads();
  {
    const int mg = t & 7;
    const int64_t m = m0 + mg * 8;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int cl = (t >> 3) + rr * 32;
      const int c = c0 + cl;
      if (c >= C || m >= M) continue;
      __nv_bfloat16* dst = out + ((int64_t)tap * C + c) * ld + m;
      const uint4 v = make_uint4(tileT[cl][mg * 4], tileT[cl][mg * 4 + 1], tileT[cl][mg * 4 + 2], tileT[cl][mg * 4 + 3]);
      if (m + 8 <= M) {
        *reinterpret_cast<uint4*>(dst) = v;
      } else {
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        for (int j = 0; j < (int)(M - m); ++j)
          dst[j] = __ushort_as_bfloat16((unsigned short)((j & 1) ? (u[j >> 1] >> 16) : (u[j >> 1] & 0xFFFFu)));
      }
    }
  }
}

static void launch_transpose(const __nv_bfloat16* in, int64_t M, int C, __nv_bfloat16* out, int64_t ld, cudaStream_t st) {
  transpose64_kernel<false><<<dim3((unsigned)ceil_div64(M, 64), ceil_div(C, 64)), 256, 0, st>>>(in, M, C, out, ld, 0, 0, 1, 1, 0, 1, 1);
}
static void launch_im2colT(const __nv_bfloat16* x, int n, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                           __nv_bfloat16* out, int64_t ld, cudaStream_t st) {
  const int64_t M = (int64_t)n * Ho * Wo;
  transpose64_kernel<true><<<dim3((unsigned)ceil_div64(M, 64), ceil_div(C, 64), k * k), 256, 0, st>>>(x, M, C, out, ld, H, W, k, stride, pad, Ho, Wo);
}

// stem: out[(r*7+s)*3 + c][m] = x_nchw[n, c, 2p - 3 + r, 2q - 3 + s]  (147 rows; rows 147..191 are zero).  One thread = 8
// consecutive pixels of one output row (112 = 14 x 8) = one 16-byte store; ld = row pitch of out.
__global__ void stem_im2colT_kernel(const float* __restrict__ x, int n, __nv_bfloat16* __restrict__ out, int64_t ld) {
  const int64_t M8 = (int64_t)n * 112 * 14;
  const int kk = blockIdx.y;                      // 0..191
  const int c = kk % 3, tap = kk / 3, r = tap / 7, s = tap % 7;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < M8; g += (int64_t)gridDim.x * blockDim.x) {
    const int q0 = (int)(g % 14) * 8, p = (int)((g / 14) % 112), img = (int)(g / (14 * 112));
    __align__(16) __nv_bfloat16 v[8];
    const int h = 2 * p - 3 + r;
    const bool row_ok = kk < 147 && h >= 0 && h < 224;
    const float* xr = x + (((int64_t)img * 3 + c) * 224 + (row_ok ? h : 0)) * 224;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int w = 2 * (q0 + j) - 3 + s;
      v[j] = __float2bfloat16_rn((row_ok && w >= 0 && w < 224) ? __ldg(xr + w) : 0.f);
    }
    *reinterpret_cast<uint4*>(out + (int64_t)kk * ld + g * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// [n,Ho,Wo,C] -> [n,2Ho,2Wo,C] with the values at the even positions and zeros elsewhere
__global__ void dilate2_kernel(const __nv_bfloat16* __restrict__ in, int n, int Ho, int Wo, int C, __nv_bfloat16* __restrict__ out) {
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  const int groups = C / 8;
  const int64_t total = (int64_t)n * 2 * Ho * 2 * Wo * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    const int64_t pix = i / groups;
    const int w = (int)(pix % (2 * Wo)), hh = (int)((pix / (2 * Wo)) % (2 * Ho)), img = (int)(pix / ((int64_t)4 * Wo * Ho));
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(w & 1) && !(hh & 1)) v = __ldg(reinterpret_cast<const uint4*>(in + ((((int64_t)img * Ho + (hh >> 1)) * Wo + (w >> 1)) * C)) + cg);
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// dgrad operand: out[cin][(r', s')][cout] = w[cout][cin][k-1-r'][k-1-s']   (bf16; for k = 1 a plain transpose)
__global__ void pack_dgrad_weight_kernel(const float* __restrict__ w, int cout, int cin, int k, __nv_bfloat16* __restrict__ out) {
  const int64_t total = (int64_t)cin * k * k * cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % cout), tap = (int)((i / cout) % (k * k)), c = (int)(i / ((int64_t)cout * k * k));
    const int r = tap / k, s = tap % k;
    out[i] = __float2bfloat16_rn(w[(((int64_t)o * cin + c) * k + (k - 1 - r)) * k + (k - 1 - s)]);
  }
}

// g_weight[cout][c][tap] (+)= D[cout][tap * cin + c]   (D bf16 with row pitch ldd)
__global__ void wgrad_unpack_kernel(const __nv_bfloat16* __restrict__ D, int cout, int cin, int kk, int ldd, float* __restrict__ g, int accumulate) {
  const int64_t total = (int64_t)cout * cin * kk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kk), c = (int)((i / kk) % cin), o = (int)(i / ((int64_t)kk * cin));
    const float v = __bfloat162float(D[(int64_t)o * ldd + tap * cin + c]);
    g[i] = v + (accumulate ? g[i] : 0.f);
  }
}

// Both for ALL convs of a backward pass in one launch each (blockIdx.y = conv): dgrad operands packed before the pass, wgrad GEMM
// outputs unpacked after it (each conv keeps its own slot of one scratch buffer) -- 105 launches of ~6 us fixed cost less.
struct DgradTab { const float* w[kMaxConvs]; __nv_bfloat16* out[kMaxConvs]; int cout[kMaxConvs], cin[kMaxConvs], k[kMaxConvs]; };
// out[(c, k*k-1-tap)][o] = w[o][(c, tap)]: a transpose of the [cout] x [cin k k] matrix with the taps of every channel reversed.
// 32 x 32 tiles through shared memory: 128-byte reads along (c, tap), 64-byte writes along o (one thread per element going
// down the cout stride touched a 32-byte sector per 4 bytes read: 218 us per backward pass).
__global__ void __launch_bounds__(256) pack_dgrad_weight_all_kernel(const __grid_constant__ DgradTab t) {
  __shared__ float tile[32][33];
  const int l = blockIdx.y + 1;                       // the stem has no data gradient
  const float* __restrict__ w = t.w[l];
  __nv_bfloat16* __restrict__ out = t.out[l];
  const int cout = t.cout[l], kk = t.k[l] * t.k[l], J = t.cin[l] * kk;
  const int tj = (J + 31) / 32, to = (cout + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int tt = blockIdx.x; tt < tj * to; tt += gridDim.x) {
    const int j0 = (tt % tj) * 32, o0 = (tt / tj) * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int o = o0 + r, j = j0 + tx;
      tile[r][tx] = (o < cout && j < J) ? w[(int64_t)o * J + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int j = j0 + r, o = o0 + tx;
      if (j < J && o < cout) {
        const int c = j / kk, tap = j - c * kk;
        out[((int64_t)c * kk + (kk - 1 - tap)) * cout + o] = __float2bfloat16_rn(tile[tx][r]);
      }
    }
    __syncthreads();
  }
}
struct WgradTab { const __nv_bfloat16* D[kMaxConvs]; float* g[kMaxConvs]; int cout[kMaxConvs], cin[kMaxConvs], kk[kMaxConvs], ldd[kMaxConvs]; };
__global__ void wgrad_unpack_all_kernel(const __grid_constant__ WgradTab t, int accumulate) {
  const int l = blockIdx.y;
  const __nv_bfloat16* __restrict__ D = t.D[l];
  float* __restrict__ g = t.g[l];
  if (!g) return;
  const int cin = t.cin[l], kk = t.kk[l], ldd = t.ldd[l];
  const int64_t total = (int64_t)t.cout[l] * cin * kk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kk), c = (int)((i / kk) % cin), o = (int)(i / ((int64_t)kk * cin));
    const float v = __bfloat162float(D[(int64_t)o * ldd + tap * cin + c]);
    g[i] = v + (accumulate ? g[i] : 0.f);
  }
}

// MaxPool2d(3, 2, 1) backward on NHWC bf16: every input pixel collects the gradient of the (at most four) windows whose recorded
// argmax (maxpool_kernel's idx: the first maximum in scan order, PyTorch's tie rule) it is.  Gather form, 8 channels per thread.
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ gy, int n,
                                                          __nv_bfloat16* __restrict__ gx) {
  const int64_t total = (int64_t)n * 112 * 112 * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % 8);
    const int64_t pix = i / 8;
    const int w = (int)(pix % 112), h = (int)((pix / 112) % 112), img = (int)(pix / (112 * 112));
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    // windows (pr, q) that contain (h, w): 2 pr - 1 <= h <= 2 pr + 1
    for (int pr = h / 2; pr <= min(55, (h + 1) / 2); ++pr)
      for (int q = w / 2; q <= min(55, (w + 1) / 2); ++q) {
        const uint32_t me = 3 * (h - (2 * pr - 1)) + (w - (2 * q - 1));          // this pixel's position inside that window
        const int64_t o = (((int64_t)img * 56 + pr) * 56 + q) * 64 + cg * 8;
        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(idx + o));
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(gy + o));
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t a = ((j < 4 ? pk.x : pk.y) >> (8 * (j & 3))) & 0xFFu;
          if (a == me) g[j] += (j & 1) ? __uint_as_float(u[j >> 1] & 0xFFFF0000u) : __uint_as_float(u[j >> 1] << 16);
        }
      }
    __align__(16) __nv_bfloat16 ov[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ov[j] = __float2bfloat16_rn(g[j]);
    *reinterpret_cast<uint4*>(gx + pix * 64 + cg * 8) = *reinterpret_cast<const uint4*>(ov);
  }
}

}  // namespace airpose

using namespace airpose;

// ---- per-conv geometry of the trunk in forward order (index = conv index of resnet50_specs)
struct ConvIO {
  int in_src;      // index of the conv whose y is this conv's input; -1 = max-pooled stem output; -2 = the image (stem)
  int res_src;     // residual added before the ReLU: conv index, -1 = pooled, -3 = none
  int relu;
  int Hin, Hout;   // input / output spatial size
};
static std::vector<ConvIO> resnet50_io() {
  std::vector<ConvIO> io;
  io.push_back({-2, -3, 1, 224, 112});
  const int layers[4] = {3, 4, 6, 3};
  int idx = 1, H = 56, x_src = -1;
  for (int li = 0; li < 4; ++li)
    for (int blk = 0; blk < layers[li]; ++blk) {
      const bool down = blk == 0;
      const int stride = (li > 0 && blk == 0) ? 2 : 1;
      const int Ho = H / stride;
      io.push_back({x_src, -3, 1, H, H});                       // conv1
      io.push_back({idx, -3, 1, H, Ho});                        // conv2 (stride on the 3x3)
      io.push_back({idx + 1, down ? idx + 3 : x_src, 1, Ho, Ho});   // conv3 + residual
      if (down) io.push_back({x_src, -3, 0, H, Ho});            // downsample conv + BN, no ReLU
      x_src = idx + 2;
      idx += down ? 4 : 3;
      H = Ho;
    }
  return io;
}

static int tape_reserve(airpose_net* h, int t, int n) {
  airpose_net::Tape& tp = h->tape[t];
  const std::vector<ConvIO> io = resnet50_io();
  if (tp.cap < n) {
    for (auto p : tp.z) cudaFree(p);
    for (auto p : tp.y) cudaFree(p);
    cudaFree(tp.pooled); cudaFree(tp.pool_idx); cudaFree(tp.stats); cudaFree(tp.stats1);
    tp.z.assign(io.size(), nullptr); tp.y.assign(io.size(), nullptr);
    for (size_t i = 0; i < io.size(); ++i) {
      const size_t elems = (size_t)n * io[i].Hout * io[i].Hout * h->specs[i].cout;
      AP_CHECK_CUDA(cudaMalloc((void**)&tp.z[i], elems * 2));
      AP_CHECK_CUDA(cudaMalloc((void**)&tp.y[i], elems * 2));
    }
    AP_CHECK_CUDA(cudaMalloc((void**)&tp.pooled, (size_t)n * 56 * 56 * 64 * 2));
    AP_CHECK_CUDA(cudaMalloc((void**)&tp.pool_idx, (size_t)n * 56 * 56 * 64));
    size_t ns = 0;
    for (const ConvSpec& s : h->specs) ns += 2 * s.cout;
    AP_CHECK_CUDA(cudaMalloc((void**)&tp.stats, ns * sizeof(float)));
    AP_CHECK_CUDA(cudaMalloc((void**)&tp.stats1, ns * sizeof(float)));
    tp.cap = n;
  }
  tp.n = n;
  return 0;
}

// training-mode forward that keeps every layer's z and y (called from airpose_backbone_fwd_train when bn->tape >= 0, and from
// airpose_backbone_fwd_train_pair).  views == 2: the tape holds BOTH views of a batch of pairs, images [0, n) from x and [n, 2n)
// from x1.  Every conv GEMM, pooling and (in the backward) every dgrad / wgrad GEMM then runs ONCE over the 2n images, while
// BatchNorm -- whose batch statistics are per forward_feat_ext call in the reference (model_copenet.py:140-141) -- runs per
// view on its half of the rows, view 0 first, so the running statistics see the same two updates in the same order.
static int backbone_fwd_train_tape(airpose_net* h, const float* x, const float* x1, int n, int views, const airpose_bn_train_params* bn,
                                   float* out_feat, cudaStream_t st) {
  const int t = bn->tape;
  AP_REQUIRE(t == 0 || t == 1, "airpose_backbone_fwd_train: tape must be -1, 0 or 1");
  const int nt = views * n;
  if (tape_reserve(h, t, nt)) return 1;
  airpose_net::Tape& tp = h->tape[t];
  tp.views = views;
  const std::vector<ConvIO> io = resnet50_io();
  airpose_bn_train_params b0 = *bn;
  b0.saved_stats = tp.stats;
  // BatchNorm of conv i: each view's half of the rows against its own batch statistics, both views in one set of launches
  auto bn_views = [&](int i, int64_t M_total, int C, const __nv_bfloat16* res, int relu) -> int {
    return bn_train(h, i, tp.z[i], M_total / views, C, &b0, res, relu, tp.y[i], st, views, tp.stats1);
  };
  {
    const int64_t work = (int64_t)n * 2 * kStemPlaneRows * 112;
    const size_t per_view = (size_t)n * 2 * kStemPlaneRows * 112 * kStemTapK;
    for (int v = 0; v < views; ++v) {
      stem_pack_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(work, 128), 148 * 32), 128, 0, st>>>(v ? x1 : x, n, h->colS[0] + v * per_view);
      AP_LAUNCH_CHECK();
    }
    GemmLaunch L{};
    if (build_stem_gemm(h, nt, 0, &L, true)) return 1;
    L.epi.out_bf16 = tp.z[0];
    if (enable_tma_epilogue(&L)) return 1;
    if (launch_gemm(L, st)) return 1;
    if (bn_views(0, (int64_t)nt * 112 * 112, 64, nullptr, 1)) return 1;
    maxpool_kernel<<<(unsigned)std::min<int64_t>(ceil_div64((int64_t)nt * 56 * 56 * 8, 256), 148 * 16), 256, 0, st>>>(tp.y[0], nt, tp.pooled, tp.pool_idx);
    AP_LAUNCH_CHECK();
  }
  auto src = [&](int s) -> const __nv_bfloat16* { return s == -1 ? tp.pooled : tp.y[s]; };
  auto run = [&](int i) -> int {
    GemmLaunch L{};
    if (conv_launch(h, i, src(io[i].in_src), nt, io[i].Hin, io[i].Hin, nullptr, 0, tp.z[i], &L, true) || launch_gemm(L, st)) return 1;
    const __nv_bfloat16* res = io[i].res_src == -3 ? nullptr : src(io[i].res_src);
    return bn_views(i, (int64_t)nt * io[i].Hout * io[i].Hout, h->specs[i].cout, res, io[i].relu);
  };
  for (int i = 1; i < (int)io.size(); ++i) {
    if (io[i].res_src > i) {                       // conv3 of a block with a downsample branch: the branch (next index) runs first
      if (run(io[i].res_src) || run(i)) return 1;
      ++i;                                         // the downsample entry has been done
    } else {
      if (run(i)) return 1;
    }
  }
  // the last block's output: conv3 of layer4.2 = index 51 (52 convs + stem = 53; the last entry in forward order is conv3 of the
  // last block because downsample entries follow conv3 only in the first block of a layer)
  const int last = (int)io.size() - 1;
  avgpool_kernel<<<ceil_div(nt * kFeat, 256), 256, 0, st>>>(tp.y[last], nt, out_feat);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_backbone_fwd_train_pair(airpose_net_t* h, const float* x0, const float* x1, int n, const airpose_bn_train_params* bn,
                                               float* out_feat, void* stream_) {
  AP_REQUIRE(h && x0 && x1 && bn && out_feat, "airpose_backbone_fwd_train_pair: null argument");
  AP_REQUIRE(h->loaded, "airpose_backbone_fwd_train_pair: weights not loaded (call airpose_net_load)");
  AP_REQUIRE(n >= 2 && 2 * n <= h->chunk, "airpose_backbone_fwd_train_pair: n=%d pairs must be in [2, %d] (one call handles one chunk of "
             "images; larger batches go through airpose_backbone_fwd_train per view)", n, h->chunk / 2);
  AP_REQUIRE(bn->tape == 0 || bn->tape == 1, "airpose_backbone_fwd_train_pair: a tape (0 / 1) is required");
  for (size_t i = 0; i < h->specs.size(); ++i)
    AP_REQUIRE(bn->bn_weight[i] && bn->bn_bias[i], "airpose_backbone_fwd_train_pair: BatchNorm %zu has a null parameter", i);
  if (train_scratch_reserve(h)) return 1;
  return backbone_fwd_train_tape(h, x0, x1, n, 2, bn, out_feat, (cudaStream_t)stream_);
}

static int bw_reserve(airpose_net* h, int n) {
  if (h->bw_cap >= n) return 0;
  for (auto& p : h->bw) { cudaFree(p); p = nullptr; }
  cudaFree(h->bw_t0); cudaFree(h->bw_t1); cudaFree(h->bw_w); cudaFree(h->bw_coef);
  const size_t act = (size_t)n * 112 * 112 * 64;                 // the largest activation (== 56*56*256)
  for (auto& p : h->bw) AP_CHECK_CUDA(cudaMalloc((void**)&p, act * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->bw_t0, (act + 8 * 2048) * 2));                       // + the pitch padding of conv_wgrad
  AP_CHECK_CUDA(cudaMalloc((void**)&h->bw_t1, (std::max((size_t)576 * 3136, (size_t)192 * 12544) * n + 8 * 4608) * 2));
  AP_CHECK_CUDA(cudaMalloc((void**)&h->bw_w, (size_t)512 * 4608 * 2 * 2));
  if (!h->bw_wd) {                                   // one slot per conv for the dgrad operand and for the wgrad GEMM output
    size_t off = 0;
    for (size_t i = 0; i < h->specs.size(); ++i) {
      const ConvSpec& s = h->specs[i];
      h->bw_w_off[i] = off;
      off += ((size_t)s.cout * (i == 0 ? 192 : s.k * s.k * s.cin) + 63) & ~(size_t)63;       // 128-byte aligned slots
    }
    AP_CHECK_CUDA(cudaMalloc((void**)&h->bw_wd, off * 2));
    AP_CHECK_CUDA(cudaMalloc((void**)&h->bw_wg, off * 2));
  }
  AP_CHECK_CUDA(cudaMalloc((void**)&h->bw_coef, 2 * 3 * 2048 * sizeof(float)));
  if (!h->bn_part) AP_CHECK_CUDA(cudaMalloc((void**)&h->bn_part, (size_t)2 * kBnSlabs * 2048 * 2 * sizeof(float)));
  h->bw_cap = n;
  return 0;
}

static unsigned ew_grid(int64_t n) { return (unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 16); }

// BatchNorm (+ReLU) backward of conv i: dy -> dz (and dpre when asked), dgamma / dbeta into the output struct.  M = rows of the whole
// tape; with a two-view tape each view's half is reduced against its own statistics and the second view's dgamma / dbeta add.
static int bn_bwd(airpose_net* h, const airpose_net::Tape& tp, int i, int64_t M, const airpose_bn_train_params* bn,
                  const airpose_trunk_grads* g, const __nv_bfloat16* dy, bool relu, __nv_bfloat16* dz, __nv_bfloat16* dpre,
                  cudaStream_t st) {
  const int C = h->specs[i].cout;
  const int views = tp.views;
  const int64_t Mv = M / views, ve = Mv * C;
  const float* stats0 = tp.stats + h->bn_save_off[i];
  const float* stats1 = tp.stats1 + h->bn_save_off[i];
  const __nv_bfloat16* y = relu ? tp.y[i] : nullptr;
  // AIRPOSE_BN_BWD_MASK_FROM_Z=1: for a ReLU without a residual (bn1 / bn2 of every block and the stem) the mask is re-derived
  // from z and y is not read (a third less traffic in both passes, bit-identical gradients).  Measured SLOWER on B200 --
  // 12.75 vs 12.63 ms per step, same box (gpurun r02t11): the extra fma + bf16 rounding per element costs more than the 16-byte
  // load it saves -- so it is off by default.
  static const std::vector<ConvIO> io = resnet50_io();
  static const bool from_z = getenv("AIRPOSE_BN_BWD_MASK_FROM_Z") != nullptr;
  const bool zmask = relu && from_z && io[i].res_src == -3;
  const float* zg = zmask ? bn->bn_weight[i] : nullptr;
  const float* zb = zmask ? bn->bn_bias[i] : nullptr;
  const __nv_bfloat16* zt = tp.z[i];
  float* part = h->bn_part;
  AP_CHECK_CUDA(launch_chain(zmask ? bn_bwd_reduce_kernel<true> : bn_bwd_reduce_kernel<false>, dim3(kBnSlabs, views), dim3(kBnRedThreads), st, dy,
                             y, zt, stats0, stats1, Mv, C, part, ve, zg, zb));
  AP_CHECK_CUDA(launch_chain(bn_bwd_finalize_kernel, dim3(ceil_div(C, 4)), dim3(128), st, (const float*)part, kBnSlabs, Mv, C,
                             (const float*)bn->bn_weight[i], stats0, stats1, g->g_bn_weight[i], g->g_bn_bias[i], g->accumulate ? 1 : 0,
                             h->bw_coef, views));
  AP_CHECK_CUDA(launch_chain(zmask ? bn_bwd_apply_kernel<true> : bn_bwd_apply_kernel<false>, dim3(ew_grid(Mv * (C / 8)), views), dim3(256), st, dy,
                             y, zt, stats0, stats1, (const float*)h->bw_coef, Mv, C, dz, dpre, ve, zg, zb));
  return 0;
}

// Fixed-order sum of the split-K partial tiles of a weight-gradient GEMM (gemm.cuh SplitKInfo), written straight into the fp32
// gradient in the parameter's own layout: g[cout][cin][tap] (+)= sum_r P_r[cout][tap * cin + c].  An item = 4 columns of one row;
// a CTA takes 32 items x 8 range groups (group k sums ranges k, k+8, ... in order, the groups are added in order through shared
// memory): the early layers have few outputs and up to 148 partials each, one thread per item left most SMs idle.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int R, int tiles_n, int BN, int slot_floats, int cout,
                                                           int cin, int kk, float* __restrict__ g, int accumulate) {
  __shared__ float4 red[8][32];
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  const int N = cin * kk, n4 = (N + 3) / 4;
  const int64_t total = (int64_t)n4 * cout;
  const int il = threadIdx.x & 31, rg = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + il;
    const bool valid = i < total;
    const int o = valid ? (int)(i % cout) : 0, col0 = valid ? (int)(i / cout) * 4 : 0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const int n_blk = col0 / BN, cl = col0 - n_blk * BN, m_blk = o >> 7, row = o & 127;
      const int tile = m_blk * tiles_n + n_blk;
      const float4* src = reinterpret_cast<const float4*>(part + (size_t)tile * R * slot_floats) + (((cl >> 6) * 16 + ((cl & 63) >> 2)) * 128 + row);
      for (int r = rg; r < R; r += 8) {
        const float4 v = __ldcg(src + (size_t)r * (slot_floats / 4));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    red[rg][il] = a;
    __syncthreads();
    if (rg == 0 && valid) {
#pragma unroll
      for (int k = 1; k < 8; ++k) { const float4 v = red[k][il]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = col0 + e;
        if (n >= N) break;
        const int tap = n / cin, c = n - tap * cin;
        const int64_t idx = ((int64_t)o * cin + c) * kk + tap;
        g[idx] = av[e] + (accumulate ? g[idx] : 0.f);
      }
    }
    __syncthreads();
  }
}

// Runs one weight-gradient GEMM whose operands are described by L (dz as the transposed A operand).  Few tiles and a very long K
// (the early layers): plain split-K over all SMs + wgrad_reduce_kernel straight into the fp32 gradient; otherwise stream-K into
// the conv's bf16 slot, unpacked later (batched pass) or right away.
static int run_wgrad_gemm(airpose_net* h, GemmLaunch& L, int i, int cin, int kk, const airpose_trunk_grads* g, cudaStream_t st) {
  const int cout = L.M, ldd = L.N;
  L.epi.out_bf16 = h->bw_batched ? h->bw_wg + h->bw_w_off[i] : h->bw_w; L.epi.ldd = ldd;
  L.pdl = use_pdl();
  if (enable_tma_epilogue(&L)) return 1;
  const int ranges = splitk_ranges(L.M, L.N, L.K, L.block_n);
  // AIRPOSE_WGRAD_MT2=1: 256 x 256 tiles (a third less operand bytes per flop) for the large layers -- measured slower, as on the
  // forward convs: 12.51 vs 12.05 ms per step (gpurun r02t9), so off by default
  static const bool use_mt2 = getenv("AIRPOSE_WGRAD_MT2") != nullptr;
  L.mt2 = (use_mt2 && !ranges && L.M >= 256 && L.block_n == 256) ? 1 : 0;
  if (ranges) {
    SplitKInfo sk{};
    sk.ranges = ranges;
    if (launch_gemm_sk(L, st, &sk)) return 1;
    const int64_t total = (int64_t)((cin * kk + 3) / 4) * cout;
    AP_CHECK_CUDA(launch_chain(wgrad_reduce_kernel, dim3((unsigned)std::min<int64_t>(ceil_div64(total, 32), 148 * 8)), dim3(256), st, sk.part, ranges,
                               sk.tiles_n, sk.block_n, sk.slot_floats, cout, cin, kk, g->g_weight[i], (int)g->accumulate));
    if (h->bw_batched) h->bw_reduced[i] = true;       // the batched unpack at the end of the pass skips this conv
    return 0;
  }
  if (launch_gemm_sk(L, st)) return 1;
  if (h->bw_batched) return 0;                      // unpacked with every other conv by ONE launch at the end of the pass
  wgrad_unpack_kernel<<<ew_grid((int64_t)cout * cin * kk), 256, 0, st>>>(h->bw_w, cout, cin, kk, ldd, g->g_weight[i], g->accumulate);
  AP_LAUNCH_CHECK();
  return 0;
}

// weight gradient of conv i: g_weight (+)= dz^T . im2col(x_in)  on the tensor cores.  Both operands contract over the pixels, the
// OUTER dimension of the NHWC tensors: dz [pixels][cout] and, for 1x1 stride-1 convs, x_in [pixels][cin] are read as they lie
// through MN-major descriptors; 3x3 / strided convs read im2col(x_in)^T through an im2col tensor map with 64-pixel boxes.
// AIRPOSE_WGRAD_TRANSPOSE=1 restores the explicit transposes to K-major operands (A/B runs).
static int conv_wgrad(airpose_net* h, int i, const __nv_bfloat16* dz, const __nv_bfloat16* x_in, int n, int Hin, int Hout,
                      const airpose_trunk_grads* g, cudaStream_t st) {
  const ConvSpec& s = h->specs[i];
  const int64_t M = (int64_t)n * Hout * Hout;
  const int Kdim = s.k * s.k * s.cin;
  static const bool direct = getenv("AIRPOSE_WGRAD_TRANSPOSE") == nullptr;
  if (direct && s.cin % 64 == 0) {
    GemmLaunch L{};
    L.M = s.cout; L.N = Kdim; L.K = (int)M;
    L.block_n = Kdim % 256 == 0 ? 256 : (Kdim > 64 ? 128 : 64);
    L.mn_a = 1;
    if (make_tmap_tiled_bf16(&L.tmA, dz, M, s.cout, s.cout, 64, 64)) return 1;
    if (s.k == 1 && s.stride == 1) {
      L.mn_b = 1;
      if (make_tmap_tiled_bf16(&L.tmB, x_in, M, s.cin, s.cin, 64, 64)) return 1;
    } else {
      L.mn_b = 2;
      ConvGeom& cg = L.geom;
      cg.n = n; cg.H = Hin; cg.W = Hin; cg.Cin = s.cin; cg.ksize = s.k; cg.stride = s.stride; cg.pad = s.pad; cg.Ho = Hout; cg.Wo = Hout;
      if (make_tmap_im2col_bf16(&L.tmB, x_in, cg, 64, 64)) return 1;
    }
    return run_wgrad_gemm(h, L, i, s.cin, s.k * s.k, g, st);
  }
  // K-major operands [channels][pixels]: the row pitch is rounded up to 8 elements (16-byte TMA pitch) so that any image
  // count works (196 and 49 pixels per image are not multiples of 8); the GEMM's K extent stays M, the pad is never read
  const int64_t ld = (M + 7) & ~(int64_t)7;
  launch_transpose(dz, M, s.cout, h->bw_t0, ld, st);
  AP_LAUNCH_CHECK();
  if (s.k == 1 && s.stride == 1) launch_transpose(x_in, M, s.cin, h->bw_t1, ld, st);
  else launch_im2colT(x_in, n, Hin, Hin, s.cin, s.k, s.stride, s.pad, Hout, Hout, h->bw_t1, ld, st);
  AP_LAUNCH_CHECK();
  airpose_gemm_args ga{};
  ga.A = h->bw_t0; ga.lda = ld; ga.B = h->bw_t1; ga.ldb = ld;
  ga.M = s.cout; ga.N = Kdim; ga.K = (int)M;
  ga.out_bf16 = h->bw_batched ? h->bw_wg + h->bw_w_off[i] : h->bw_w; ga.ldd = Kdim;
  if (airpose_gemm_bf16(&ga, st)) return 1;
  if (h->bw_batched) return 0;                      // unpacked with every other conv by ONE launch at the end of the pass
  const int64_t total = (int64_t)s.cout * Kdim;
  wgrad_unpack_kernel<<<ew_grid(total), 256, 0, st>>>(h->bw_w, s.cout, s.cin, s.k * s.k, Kdim, g->g_weight[i], g->accumulate);
  AP_LAUNCH_CHECK();
  return 0;
}

// data gradient of conv i: dx = conv_transpose(dz, W) (+ add), as an implicit-GEMM conv of (dilated) dz with the flipped weights
static int conv_dgrad(airpose_net* h, int i, const float* w_f32, const __nv_bfloat16* dz, int n, int Hin, int Hout, const __nv_bfloat16* add,
                      __nv_bfloat16* dx, __nv_bfloat16* scratch, cudaStream_t st) {
  const ConvSpec& s = h->specs[i];
  __nv_bfloat16* wd = h->bw_batched ? h->bw_wd + h->bw_w_off[i] : h->bw_w + (size_t)512 * 4608;   // pre-packed / second half of the scratch
  if (!h->bw_batched) {
    pack_dgrad_weight_kernel<<<ew_grid((int64_t)s.cin * s.k * s.k * s.cout), 256, 0, st>>>(w_f32, s.cout, s.cin, s.k, wd);
    AP_LAUNCH_CHECK();
  }
  if (s.k == 1 && s.stride == 1) {
    airpose_gemm_args ga{};
    ga.A = dz; ga.lda = s.cout; ga.B = wd; ga.ldb = s.cout;
    ga.M = n * Hout * Hout; ga.N = s.cin; ga.K = s.cout;
    ga.residual = add; ga.ldr = s.cin;
    ga.out_bf16 = dx; ga.ldd = s.cin;
    return airpose_gemm_bf16(&ga, st);
  }
  if (s.k == 1) {          // 1x1 stride 2 (downsample): GEMM on the strided pixels, then scatter to the even positions
    AP_REQUIRE(add == nullptr, "conv_dgrad: strided 1x1 with an addend is not used");
    airpose_gemm_args ga{};
    ga.A = dz; ga.lda = s.cout; ga.B = wd; ga.ldb = s.cout;
    ga.M = n * Hout * Hout; ga.N = s.cin; ga.K = s.cout;
    ga.out_bf16 = scratch; ga.ldd = s.cin;
    if (airpose_gemm_bf16(&ga, st)) return 1;
    AP_CHECK_CUDA(launch_chain(dilate2_kernel, dim3(ew_grid((int64_t)n * Hin * Hin * (s.cin / 8))), dim3(256), st, (const __nv_bfloat16*)scratch, n, Hout,
                               Hout, s.cin, dx));
    return 0;
  }
  const __nv_bfloat16* src = dz;
  if (s.stride == 2) {
    AP_CHECK_CUDA(launch_chain(dilate2_kernel, dim3(ew_grid((int64_t)n * Hin * Hin * (s.cout / 8))), dim3(256), st, dz, n, Hout, Hout, s.cout, scratch));
    src = scratch;
  }
  airpose_conv_args ca{};
  ca.x = src; ca.n = n; ca.H = Hin; ca.W = Hin; ca.Cin = s.cout;
  ca.w = wd; ca.Cout = s.cin; ca.ksize = 3; ca.stride = 1; ca.pad = 1;
  ca.residual = add; ca.relu = 0; ca.out = dx;
  return airpose_conv_bf16(&ca, st);
}

// x / x1: the images of view 0 / view 1 (x1 only for a two-view tape); n_view images per view; the tape holds views * n_view.
static int backbone_bwd_train_impl(airpose_net_t* h, const float* x, const float* x1, int n_view, int views, int t,
                                   const airpose_bn_train_params* bn, const float* g_feat, const airpose_trunk_grads* g,
                                   const float* const* conv_weights, cudaStream_t st) {
  const airpose_net::Tape& tp = h->tape[t];
  const int n = views * n_view;          // everything below except BatchNorm and the stem's image operand sees one batch of n images
  AP_REQUIRE(tp.n == n && n > 0 && tp.views == views, "airpose_backbone_bwd_train: tape %d holds a forward of %d images in %d view(s), not %d in %d",
             t, tp.n, tp.views, n, views);
  for (size_t i = 0; i < h->specs.size(); ++i)
    AP_REQUIRE(g->g_weight[i] && g->g_bn_weight[i] && g->g_bn_bias[i] && conv_weights[i], "airpose_backbone_bwd_train: null buffer (conv %zu)", i);
  if (bw_reserve(h, n)) return 1;
  {                                                  // every conv's dgrad operand (flipped, transposed bf16 weights), one launch
    DgradTab dt{};
    for (size_t i = 1; i < h->specs.size(); ++i) {
      const ConvSpec& s = h->specs[i];
      dt.w[i] = conv_weights[i]; dt.out[i] = h->bw_wd + h->bw_w_off[i]; dt.cout[i] = s.cout; dt.cin[i] = s.cin; dt.k[i] = s.k;
    }
    pack_dgrad_weight_all_kernel<<<dim3(96, (unsigned)h->specs.size() - 1), 256, 0, st>>>(dt);
    AP_LAUNCH_CHECK();
  }
  h->bw_batched = true;                              // conv_wgrad / conv_dgrad use the per-conv slots; reset at the end of the pass
  const std::vector<ConvIO> io = resnet50_io();
  auto src = [&](int s) -> const __nv_bfloat16* { return s == -1 ? tp.pooled : tp.y[s]; };
  __nv_bfloat16 *G = h->bw[0], *G2 = h->bw[1], *DZ = h->bw[2], *DPRE = h->bw[3], *SCR = h->bw[4], *T = h->bw[5];
  // gradient of the last block's output from the average pool
  avgpool_bwd_kernel<<<ew_grid((int64_t)n * 49 * kFeat), 256, 0, st>>>(g_feat, n, G);
  AP_LAUNCH_CHECK();
  // walk the blocks in reverse
  const int layers[4] = {3, 4, 6, 3};
  std::vector<int> first_idx;          // conv1 index of every block, in forward order
  { int idx = 1; for (int li = 0; li < 4; ++li) for (int b = 0; b < layers[li]; ++b) { first_idx.push_back(idx); idx += (b == 0) ? 4 : 3; } }
  // weight gradients of convs [lo, hi) out of their bf16 slots (those not already written by wgrad_reduce_kernel), one launch
  auto unpack_range = [&](int lo, int hi) -> int {
    WgradTab wt{};
    for (int i = lo; i < hi; ++i) {
      const ConvSpec& s = h->specs[i];
      wt.D[i] = h->bw_wg + h->bw_w_off[i]; wt.g[i] = h->bw_reduced[i] ? nullptr : g->g_weight[i]; wt.cout[i] = s.cout; wt.cin[i] = s.cin;
      wt.kk[i] = s.k * s.k; wt.ldd[i] = i == 0 ? 192 : s.k * s.k * s.cin;
      h->bw_reduced[i] = false;
    }
    wgrad_unpack_all_kernel<<<dim3(48, (unsigned)h->specs.size()), 256, 0, st>>>(wt, g->accumulate);
    AP_LAUNCH_CHECK();
    return 0;
  };
  const int upper_block = layers[0] + layers[1];      // first block of layer3
  int unpacked_from = (int)h->specs.size();           // convs [unpacked_from, end) have been unpacked
  for (int bi = (int)first_idx.size() - 1; bi >= 0; --bi) {
    const int i1 = first_idx[bi], i2 = i1 + 1, i3 = i1 + 2;
    // a block has a downsample branch iff its conv3 takes its residual from conv index i1 + 3
    const bool has_ds = io[i3].res_src == i1 + 3;
    const int Hin = io[i1].Hin, Ho = io[i3].Hout;
    const int64_t Mo = (int64_t)n * Ho * Ho;
    // conv3 + bn3 (+ residual, ReLU):  G = dL/d(block output)
    if (bn_bwd(h, tp, i3, Mo, bn, g, G, true, DZ, DPRE, st)) return 1;                       // DPRE = gradient of the residual branch
    if (conv_wgrad(h, i3, DZ, tp.y[i2], n, Ho, Ho, g, st)) return 1;
    if (conv_dgrad(h, i3, conv_weights[i3], DZ, n, Ho, Ho, nullptr, G2, SCR, st)) return 1;  // G2 = dL/d y2
    // conv2 + bn2 + ReLU
    if (bn_bwd(h, tp, i2, Mo, bn, g, G2, true, DZ, nullptr, st)) return 1;
    if (conv_wgrad(h, i2, DZ, tp.y[i1], n, Hin, Ho, g, st)) return 1;
    if (conv_dgrad(h, i2, conv_weights[i2], DZ, n, Hin, Ho, nullptr, G2, SCR, st)) return 1; // G2 = dL/d y1   [n,Hin,Hin,planes]
    // conv1 + bn1 + ReLU
    const int64_t Mi = (int64_t)n * Hin * Hin;
    const __nv_bfloat16* xin = src(io[i1].in_src);
    if (bn_bwd(h, tp, i1, Mi, bn, g, G2, true, DZ, nullptr, st)) return 1;
    if (conv_wgrad(h, i1, DZ, xin, n, Hin, Hin, g, st)) return 1;
    const __nv_bfloat16* addend = DPRE;                                                      // identity residual: dL/dx += dpre3
    if (has_ds) {
      const int id = i1 + 3;
      if (bn_bwd(h, tp, id, Mo, bn, g, DPRE, false, T, nullptr, st)) return 1;               // T = dz of the downsample conv
      if (conv_wgrad(h, id, T, xin, n, Hin, Ho, g, st)) return 1;
      if (conv_dgrad(h, id, conv_weights[id], T, n, Hin, Ho, nullptr, G, SCR, st)) return 1; // G = downsample path of dL/dx
      addend = G;
      // conv1's data gradient adds the downsample path
      if (conv_dgrad(h, i1, conv_weights[i1], DZ, n, Hin, Hin, addend, G2, SCR, st)) return 1;
      std::swap(G, G2);
    } else {
      if (conv_dgrad(h, i1, conv_weights[i1], DZ, n, Hin, Hin, addend, G, SCR, st)) return 1;
    }
    // G now holds dL/d(block input)
    if (bi == upper_block && g->upper_done) {
      // every gradient of layer3 and layer4 (95 % of the trunk's parameters) is final once the work enqueued so far has run: the
      // caller can start its all-reduce of that part while layer2, layer1 and the stem are still being differentiated
      if (unpack_range(i1, unpacked_from)) return 1;
      unpacked_from = i1;
      g->upper_done(g->user);
    }
  }
  // stem: max-pool backward, bn1 + ReLU backward, weight gradient of the 7x7 conv (no data gradient: the input is the image)
  maxpool_bwd_kernel<<<ew_grid((int64_t)n * 112 * 112 * 8), 256, 0, st>>>(tp.pool_idx, G, n, G2);
  AP_LAUNCH_CHECK();
  const int64_t M0 = (int64_t)n * 112 * 112;
  if (bn_bwd(h, tp, 0, M0, bn, g, G2, true, DZ, nullptr, st)) return 1;
  {
    static const bool direct = getenv("AIRPOSE_WGRAD_TRANSPOSE") == nullptr;      // as in conv_wgrad
    if (!direct) {
      launch_transpose(DZ, M0, 64, h->bw_t0, M0, st);
      AP_LAUNCH_CHECK();
    }
    for (int v = 0; v < views; ++v) {                       // columns [v * M0 / views, ...) of the K-major operand come from view v's images
      stem_im2colT_kernel<<<dim3(148 * 4, 192), 256, 0, st>>>(v ? x1 : x, n_view, h->bw_t1 + (size_t)v * (M0 / views), M0);
      AP_LAUNCH_CHECK();
    }
    if (direct) {
      GemmLaunch L{};
      L.M = 64; L.N = 192; L.K = (int)M0;
      L.block_n = 128;
      L.mn_a = 1;
      if (make_tmap_tiled_bf16(&L.tmA, DZ, M0, 64, 64, 64, 64)) return 1;
      if (make_tmap_tiled_bf16(&L.tmB, h->bw_t1, 192, M0, M0, L.block_n, 64)) return 1;
      if (run_wgrad_gemm(h, L, 0, 3, 49, g, st)) return 1;
    } else {
      airpose_gemm_args ga{};
      ga.A = h->bw_t0; ga.lda = M0; ga.B = h->bw_t1; ga.ldb = M0;
      ga.M = 64; ga.N = 192; ga.K = (int)M0;
      ga.out_bf16 = h->bw_wg + h->bw_w_off[0]; ga.ldd = 192;
      if (airpose_gemm_bf16(&ga, st)) return 1;
    }
  }
  if (unpack_range(0, unpacked_from)) return 1;        // every remaining conv's weight gradient out of its slot, one launch
  h->bw_batched = false;
  return 0;
}

extern "C" int airpose_backbone_bwd_train(airpose_net_t* h, const float* x, int n, int t, const airpose_bn_train_params* bn,
                                          const float* g_feat, const airpose_trunk_grads* g, const float* const* conv_weights,
                                          void* stream_) {
  AP_REQUIRE(h && x && bn && g_feat && g && conv_weights, "airpose_backbone_bwd_train: null argument");
  AP_REQUIRE(t == 0 || t == 1, "airpose_backbone_bwd_train: tape must be 0 or 1");
  return backbone_bwd_train_impl(h, x, nullptr, n, 1, t, bn, g_feat, g, conv_weights, (cudaStream_t)stream_);
}

extern "C" int airpose_backbone_bwd_train_pair(airpose_net_t* h, const float* x0, const float* x1, int n, int t,
                                               const airpose_bn_train_params* bn, const float* g_feat, const airpose_trunk_grads* g,
                                               const float* const* conv_weights, void* stream_) {
  AP_REQUIRE(h && x0 && x1 && bn && g_feat && g && conv_weights, "airpose_backbone_bwd_train_pair: null argument");
  AP_REQUIRE(t == 0 || t == 1, "airpose_backbone_bwd_train_pair: tape must be 0 or 1");
  return backbone_bwd_train_impl(h, x0, x1, n, 2, t, bn, g_feat, g, conv_weights, (cudaStream_t)stream_);
}

// ---- building blocks of the trunk backward exported for the per-layer parity tests
extern "C" int airpose_debug_conv_bwd(airpose_net_t* h, int conv_idx, int n, const void* dz, const void* x_in, const float* w_f32,
                                      const void* add, void* out_dx, float* out_gw, int accumulate, void* stream_) {
  AP_REQUIRE(h && dz && x_in && w_f32 && out_gw, "airpose_debug_conv_bwd: null argument");
  AP_REQUIRE(conv_idx >= 1 && conv_idx < (int)h->specs.size() && n > 0, "airpose_debug_conv_bwd: bad index / n");
  cudaStream_t st = (cudaStream_t)stream_;
  if (bw_reserve(h, n)) return 1;
  h->bw_batched = false;                             // per-layer entry point: pack / unpack immediately (a failed whole pass may have left it set)
  const std::vector<ConvIO> io = resnet50_io();
  airpose_trunk_grads g{};
  g.g_weight[conv_idx] = out_gw;
  g.accumulate = accumulate;
  if (conv_wgrad(h, conv_idx, (const __nv_bfloat16*)dz, (const __nv_bfloat16*)x_in, n, io[conv_idx].Hin, io[conv_idx].Hout, &g, st)) return 1;
  if (out_dx && conv_dgrad(h, conv_idx, w_f32, (const __nv_bfloat16*)dz, n, io[conv_idx].Hin, io[conv_idx].Hout, (const __nv_bfloat16*)add,
                           (__nv_bfloat16*)out_dx, h->bw[4], st)) return 1;
  return 0;
}

extern "C" int airpose_debug_bn_bwd(airpose_net_t* h, int64_t M, int C, const void* dy, const void* y, const void* z, const float* stats,
                                    const float* gamma, void* out_dz, void* out_dpre, float* g_gamma, float* g_beta, int accumulate,
                                    void* stream_) {
  AP_REQUIRE(h && dy && z && stats && gamma && out_dz && g_gamma && g_beta, "airpose_debug_bn_bwd: null argument");
  AP_REQUIRE(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "airpose_debug_bn_bwd: unsupported channel count %d", C);
  cudaStream_t st = (cudaStream_t)stream_;
  if (bw_reserve(h, 8)) return 1;
  bn_bwd_reduce_kernel<false><<<kBnSlabs, kBnRedThreads, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, stats, stats, M, C,
                                                 h->bn_part, 0, nullptr, nullptr);
  AP_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<ceil_div(C, 4), 128, 0, st>>>(h->bn_part, kBnSlabs, M, C, gamma, stats, stats, g_gamma, g_beta, accumulate, h->bw_coef, 1);
  AP_LAUNCH_CHECK();
  bn_bwd_apply_kernel<false><<<ew_grid(M * (C / 8)), 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, stats, stats,
                                                            h->bw_coef, M, C, (__nv_bfloat16*)out_dz, (__nv_bfloat16*)out_dpre, 0, nullptr, nullptr);
  AP_LAUNCH_CHECK();
  return 0;
}

// copies one tensor of a training tape into a caller buffer (tests): which = 0 raw conv output z, 1 normalised output y, 2 pooled stem
extern "C" int airpose_debug_tape_get(airpose_net_t* h, int t, int conv_idx, int which, void* dst, int64_t dst_elems, void* stream_) {
  AP_REQUIRE(h && dst && (t == 0 || t == 1), "airpose_debug_tape_get: bad argument");
  const airpose_net::Tape& tp = h->tape[t];
  AP_REQUIRE(tp.n > 0, "airpose_debug_tape_get: tape %d is empty", t);
  const std::vector<ConvIO> io = resnet50_io();
  const __nv_bfloat16* src;
  int64_t elems;
  if (which == 2) { src = tp.pooled; elems = (int64_t)tp.n * 56 * 56 * 64; }
  else {
    AP_REQUIRE(conv_idx >= 0 && conv_idx < (int)io.size(), "airpose_debug_tape_get: bad conv index");
    src = which == 0 ? tp.z[conv_idx] : tp.y[conv_idx];
    elems = (int64_t)tp.n * io[conv_idx].Hout * io[conv_idx].Hout * h->specs[conv_idx].cout;
  }
  AP_REQUIRE(dst_elems == elems, "airpose_debug_tape_get: destination holds %lld elements, the tensor has %lld", (long long)dst_elems, (long long)elems);
  AP_CHECK_CUDA(cudaMemcpyAsync(dst, src, elems * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
  return 0;
}
