// Test-mode outputs and metrics of copenet_twoview:
//   /root/reference/copenet/src/copenet/copenet_twoview.py:323-326   pred/gt rotation matrices -> angle-axis
//   /root/reference/copenet/src/copenet/copenet_twoview.py:556-559   angle-axis -> rotation matrices (test_epoch_end)
//   /root/reference/copenet/src/copenet/copenet_twoview.py:541-551,583-586   MPE / MPJPE reductions
// The two conversions are torchgeometry 0.1.2's `rotation_matrix_to_angle_axis` and `angle_axis_to_rotation_matrix`
// (requirements.txt pins torchgeometry==0.1.2; the package is NOT in /root/reference and not installable offline).
// PARITY UNPINNED: the arithmetic below restates the published source of torchgeometry/core/conversions.py at that version
// (rotation matrix -> quaternion through the TRANSPOSED matrix with the four-branch trace test and eps = 1e-6, quaternion ->
// angle-axis with the atan2 form; Rodrigues with a first-order Taylor branch below theta^2 = 1e-6), reading its `1 - mask`
// on boolean masks as logical NOT (what every maintained fork does; the literal expression raises under torch >= 1.2).
// It is pinned only by known-answer identities (tests): identity, quarter turns, all four branches, aa -> R -> aa round trips.
#include <algorithm>
#include "common.cuh"

namespace airpose {
namespace {

// one thread per matrix; R row-major 3x3 at R + i * stride (stride 9, or 12 for the reference's [N,3,4] layout: then the rows are
// 4 floats apart -- `row` below)
__global__ void rotmat_to_angle_axis_kernel(const float* __restrict__ R, int64_t n, int mat_stride, int row, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* m = R + i * mat_stride;
  // rmat_t = transpose(rotation_matrix): t[a][b] = m[b][a]
  const float t00 = m[0], t01 = m[row], t02 = m[2 * row];
  const float t10 = m[1], t11 = m[row + 1], t12 = m[2 * row + 1];
  const float t20 = m[2], t21 = m[row + 2], t22 = m[2 * row + 2];
  const bool d2 = t22 < 1e-6f, d0_d1 = t00 > t11, d0_nd1 = t00 < -t11;
  float q0, q1, q2, q3, tt;
  if (d2 && d0_d1) {
    tt = 1.f + t00 - t11 - t22;
    q0 = t12 - t21; q1 = tt; q2 = t01 + t10; q3 = t20 + t02;
  } else if (d2) {
    tt = 1.f - t00 + t11 - t22;
    q0 = t20 - t02; q1 = t01 + t10; q2 = tt; q3 = t12 + t21;
  } else if (d0_nd1) {
    tt = 1.f - t00 - t11 + t22;
    q0 = t01 - t10; q1 = t20 + t02; q2 = t12 + t21; q3 = tt;
  } else {
    tt = 1.f + t00 + t11 + t22;
    q0 = tt; q1 = t12 - t21; q2 = t20 - t02; q3 = t01 - t10;
  }
  const float s = 0.5f / sqrtf(tt);                 // q /= sqrt(t); q *= 0.5
  q0 *= s; q1 *= s; q2 *= s; q3 *= s;
  // quaternion_to_angle_axis
  const float sin2 = q1 * q1 + q2 * q2 + q3 * q3;
  const float sn = sqrtf(sin2);
  const float two_theta = 2.f * (q0 < 0.f ? atan2f(-sn, -q0) : atan2f(sn, q0));
  const float k = sin2 > 0.f ? two_theta / sn : 2.f;
  out[i * 3] = q1 * k; out[i * 3 + 1] = q2 * k; out[i * 3 + 2] = q3 * k;
}

// angle_axis_to_rotation_matrix, top-left 3x3 of the 4x4 it returns
__global__ void angle_axis_to_rotmat_kernel(const float* __restrict__ aa, int64_t n, float* __restrict__ R) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float rx = aa[i * 3], ry = aa[i * 3 + 1], rz = aa[i * 3 + 2];
  const float theta2 = rx * rx + ry * ry + rz * rz;
  float* o = R + i * 9;
  if (theta2 > 1e-6f) {
    const float theta = sqrtf(theta2);
    const float inv = 1.f / (theta + 1e-6f);
    const float wx = rx * inv, wy = ry * inv, wz = rz * inv;
    const float c = cosf(theta), s = sinf(theta), k = 1.f - c;
    o[0] = c + wx * wx * k;       o[1] = wx * wy * k - wz * s;  o[2] = wy * s + wx * wz * k;
    o[3] = wz * s + wx * wy * k;  o[4] = c + wy * wy * k;       o[5] = -wx * s + wy * wz * k;
    o[6] = -wy * s + wx * wz * k; o[7] = wx * s + wy * wz * k;  o[8] = c + wz * wz * k;
  } else {
    o[0] = 1.f; o[1] = -rz; o[2] = ry;
    o[3] = rz;  o[4] = 1.f; o[5] = -rx;
    o[6] = -ry; o[7] = rx;  o[8] = 1.f;
  }
}

// mean over rows of || a[row] - b[row] ||_2, rows = (item, first `used` of `per_item` points); one CTA, fixed-order reduction
__global__ void __launch_bounds__(256) mean_distance_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t items, int per_item,
                                                            int used, float* __restrict__ out) {
  __shared__ double red[256];
  const int64_t rows = items * used;
  double acc = 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += 256) {
    const int64_t o = ((r / used) * per_item + (r % used)) * 3;
    const float dx = a[o] - b[o], dy = a[o + 1] - b[o + 1], dz = a[o + 2] - b[o + 2];
    acc += (double)sqrtf(dx * dx + dy * dy + dz * dz);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = rows > 0 ? (float)(red[0] / (double)rows) : 0.f;
}

}  // namespace
}  // namespace airpose

using namespace airpose;

extern "C" int airpose_rotmat_to_angle_axis(const float* R, int64_t n, int32_t mat_stride, int32_t row_stride, float* out, void* stream) {
  AP_REQUIRE(R && out, "airpose_rotmat_to_angle_axis: null argument");
  AP_REQUIRE(n >= 0 && row_stride >= 3 && mat_stride >= 2 * row_stride + 3, "airpose_rotmat_to_angle_axis: bad strides (%d, %d)", mat_stride, row_stride);
  if (n == 0) return 0;
  rotmat_to_angle_axis_kernel<<<(unsigned)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(R, n, mat_stride, row_stride, out);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_angle_axis_to_rotmat(const float* aa, int64_t n, float* R, void* stream) {
  AP_REQUIRE(aa && R, "airpose_angle_axis_to_rotmat: null argument");
  AP_REQUIRE(n >= 0, "airpose_angle_axis_to_rotmat: negative count");
  if (n == 0) return 0;
  angle_axis_to_rotmat_kernel<<<(unsigned)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(aa, n, R);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_mean_distance(const float* a, const float* b, int64_t items, int32_t points_per_item, int32_t points_used, float* out,
                                     void* stream) {
  AP_REQUIRE(a && b && out, "airpose_mean_distance: null argument");
  AP_REQUIRE(items >= 0 && points_per_item > 0 && points_used > 0 && points_used <= points_per_item, "airpose_mean_distance: bad sizes");
  mean_distance_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a, b, items, points_per_item, points_used, out);
  AP_LAUNCH_CHECK();
  return 0;
}
