// Adam (amsgrad) step of the training configuration, one launch over a flat parameter buffer.
//
// Replaces torch.optim.Adam(model.parameters(), lr, weight_decay=0, amsgrad=True)
// (/root/reference/copenet/src/copenet/copenet_twoview.py:416-425) with the arithmetic of torch's
// single-tensor implementation:  m = lerp(m, g, 1-b1);  v = b2 v + (1-b2) g g;  vmax = max(vmax, v);
// p -= (lr / (1-b1^t)) * m / (sqrt(vmax) / sqrt(1-b2^t) + eps).
// 27.1 M parameters x 36 B of traffic per element (5 reads, 4 writes): HBM-bound, float4 accesses,
// grid sized in multiples of the SM count.  The flat layout also makes the gradient all-reduce of the
// data-parallel step ONE collective (airpose_b200/parallel.py).
#include "common.cuh"

namespace airpose {
namespace {

struct AdamK {
  float* p; const float* g; float* m; float* v; float* vmax;
  int64_t n;
  float b1, b2, eps, step_size, inv_bc2_sqrt, grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* vmax, const AdamK& k) {
  g *= k.grad_scale;
  m = m + (1.f - k.b1) * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = v * k.b2 + (1.f - k.b2) * g * g;                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
  float d = v;
  if (vmax) { *vmax = fmaxf(*vmax, v); d = *vmax; }     // torch.maximum(max_exp_avg_sq, exp_avg_sq)
  const float denom = sqrtf(d) * k.inv_bc2_sqrt + k.eps;
  p = p - k.step_size * (m / denom);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(AdamK k) {
  const int64_t n4 = k.n / 4;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(k.p);
  const float4* g4 = reinterpret_cast<const float4*>(k.g);
  float4* m4 = reinterpret_cast<float4*>(k.m);
  float4* v4 = reinterpret_cast<float4*>(k.v);
  float4* x4 = reinterpret_cast<float4*>(k.vmax);
  for (int64_t i = tid; i < n4; i += nth) {
    float4 p = p4[i], m = m4[i], v = v4[i];
    const float4 g = __ldg(g4 + i);
    float4 x = k.vmax ? x4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    adam_one(p.x, g.x, m.x, v.x, k.vmax ? &x.x : nullptr, k);
    adam_one(p.y, g.y, m.y, v.y, k.vmax ? &x.y : nullptr, k);
    adam_one(p.z, g.z, m.z, v.z, k.vmax ? &x.z : nullptr, k);
    adam_one(p.w, g.w, m.w, v.w, k.vmax ? &x.w : nullptr, k);
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (k.vmax) x4[i] = x;
  }
  for (int64_t i = n4 * 4 + tid; i < k.n; i += nth)
    adam_one(k.p[i], k.g[i], k.m[i], k.v[i], k.vmax ? k.vmax + i : nullptr, k);
}

}  // namespace
}  // namespace airpose

using namespace airpose;

extern "C" int airpose_adam_step(const airpose_adam_args* a, void* stream) {
  AP_REQUIRE(a && a->param && a->grad && a->exp_avg && a->exp_avg_sq, "airpose_adam_step: null argument");
  AP_REQUIRE(a->n >= 0 && a->step >= 1, "airpose_adam_step: bad n/step");
  AP_REQUIRE(a->beta1 >= 0.f && a->beta1 < 1.f && a->beta2 >= 0.f && a->beta2 < 1.f, "airpose_adam_step: betas out of range");
  const uintptr_t al = (uintptr_t)a->param | (uintptr_t)a->grad | (uintptr_t)a->exp_avg | (uintptr_t)a->exp_avg_sq |
                       (uintptr_t)a->max_exp_avg_sq;
  AP_REQUIRE((al & 15) == 0, "airpose_adam_step: buffers must be 16-byte aligned");
  if (a->n == 0) return 0;
  AdamK k;
  k.p = a->param; k.g = a->grad; k.m = a->exp_avg; k.v = a->exp_avg_sq; k.vmax = a->max_exp_avg_sq;
  k.n = a->n; k.b1 = a->beta1; k.b2 = a->beta2; k.eps = a->eps;
  const double bc1 = 1.0 - pow((double)a->beta1, (double)a->step), bc2 = 1.0 - pow((double)a->beta2, (double)a->step);
  k.step_size = (float)((double)a->lr / bc1);
  k.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  k.grad_scale = a->grad_scale == 0.f ? 1.f : a->grad_scale;
  const int64_t want = ceil_div64(a->n / 4 + 1, 256);
  const int grid = (int)std::min<int64_t>(want, 148 * 8);
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(k);
  AP_LAUNCH_CHECK();
  return 0;
}
