// Input preprocessing on the device: the step immediately before the hot path.
//
// (1) airpose_preprocess_bgr8: what the drone server does with a stage-0 message
//     (/root/reference/catkin_ws/src/aircap/packages/flight/airpose_server/server.py:93-98):
//       u8 BGR [224,224,3] -> RGB -> CHW -> float * (1/255) -> (x - mean) / std        (three separately rounded fp32 ops)
// (2) airpose_preprocess_crop_resize: what the dataset does per camera
//     (/root/reference/copenet/src/copenet/dsets/aerialpeople.py:125-141,174 + utils/utils.py:214-235):
//       u8 BGR frame -> RGB / 255. (float64) -> crop -> cv2.resize(INTER_LINEAR) so that the longer side is 224 ->
//       zero letterbox to 224 x 224 -> CHW float32 -> torchvision Normalize(mean, std)
//     cv2.resize's bilinear arithmetic for CV_64F (OpenCV imgproc resize.cpp, HResizeLinear / VResizeLinear; all double --
//     measured against cv2 4.13.0, oracle/airpose_oracle.py:_cv_linear_coef): scale = 1 / (dst / src), fx = (dx + 0.5) * scale - 0.5,
//     sx = floor(fx), fx -= sx, clamped at the borders (sx < 0 -> sx = 0, fx = 0; sx >= w - 1 -> sx = w - 1, fx = 0), horizontal
//     pass first, then vertical, both in double.  Restated here in double so that the result agrees with cv2 to the final
//     float32 rounding.  (For an exact 2x reduction cv2 switches INTER_LINEAR to its INTER_AREA fast path, which computes
//     the same 2x2 mean.)
//
// Both kernels are pure HBM streaming: one thread per output pixel (all three channels), coalesced writes per plane.
#include <algorithm>
#include "common.cuh"

namespace airpose {
namespace {

struct Norm3 { float mean[3], std[3]; };

// thread = one output pixel; reads 3 bytes (BGR), writes 3 planes
__global__ void __launch_bounds__(256) preprocess_bgr8_kernel(const uint8_t* __restrict__ in, int64_t n_pixels_total, int hw, Norm3 nm,
                                                              float* __restrict__ out) {
  const float inv255 = (float)(1.0 / 255);             // the Python scalar 1.0/255 cast to the tensor's dtype
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels_total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = i / hw;
    const int pix = (int)(i - img * hw);
    const uint8_t* p = in + i * 3;
    const float bgr[3] = {(float)p[0], (float)p[1], (float)p[2]};
#pragma unroll
    for (int c = 0; c < 3; ++c) {                        // output channel c = R, G, B = input byte 2 - c
      const float v = __fmul_rn(bgr[2 - c], inv255);
      out[(img * 3 + c) * hw + pix] = __fdiv_rn(__fsub_rn(v, nm.mean[c]), nm.std[c]);
    }
  }
}

// cv2's linear coefficient for destination index d (see the header comment)
__device__ __forceinline__ void cv_linear_coef(int d, double scale, int src_size, int& s, double& f) {
  double fx = ((double)d + 0.5) * scale - 0.5;
  s = (int)floor(fx);
  fx -= (double)s;
  if (s < 0) { fx = 0.0; s = 0; }
  if (s >= src_size - 1) { fx = 0.0; s = src_size - 1; }
  f = fx;
}

// rect = (y0, y1, x0, x1) in frame pixels (the crop is frame[y0:y1, x0:x1]); geom = (dst_h, dst_w, pad_top, pad_left)
__global__ void __launch_bounds__(256) crop_resize_kernel(const uint8_t* __restrict__ frames, int64_t frame_stride, int W,
                                                          const int32_t* __restrict__ rects, int n, int size, Norm3 nm,
                                                          float* __restrict__ out) {
  const int hw = size * size;
  const int64_t total = (int64_t)n * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int img = (int)(i / hw);
    const int pix = (int)(i - (int64_t)img * hw);
    const int oy = pix / size, ox = pix - oy * size;
    const int y0 = rects[img * 4], y1 = rects[img * 4 + 1], x0 = rects[img * 4 + 2], x1 = rects[img * 4 + 3];
    const int sh = y1 - y0, sw = x1 - x0;
    // resize_with_pad (utils.py:214-235): scale = size / max(h, w) as a Python float, dst = (int(scale*w), int(scale*h))
    const double scale = (double)size / (double)(sh > sw ? sh : sw);
    const int dw = (int)(scale * (double)sw), dh = (int)(scale * (double)sh);
    const int pad_top = (size - dh) / 2, pad_left = (size - dw) / 2;
    const int dy = oy - pad_top, dx = ox - pad_left;
    double val[3] = {0.0, 0.0, 0.0};                     // BORDER_CONSTANT, value 0
    if (dy >= 0 && dy < dh && dx >= 0 && dx < dw) {
      // cv::resize: inv_scale = dsize / ssize, scale = 1 / inv_scale (both double)
      const double scale_x = 1.0 / ((double)dw / (double)sw), scale_y = 1.0 / ((double)dh / (double)sh);
      int sx, sy; double fx, fy;
      cv_linear_coef(dx, scale_x, sw, sx, fx);
      cv_linear_coef(dy, scale_y, sh, sy, fy);
      const int sx1 = min(sx + 1, sw - 1), sy1 = min(sy + 1, sh - 1);     // weight 0 whenever clamped
      const double a0 = 1.0 - fx, a1 = fx, b0 = 1.0 - fy, b1 = fy;
      const uint8_t* f = frames + (int64_t)img * frame_stride;
      const uint8_t* p00 = f + ((int64_t)(y0 + sy) * W + (x0 + sx)) * 3;
      const uint8_t* p01 = f + ((int64_t)(y0 + sy) * W + (x0 + sx1)) * 3;
      const uint8_t* p10 = f + ((int64_t)(y0 + sy1) * W + (x0 + sx)) * 3;
      const uint8_t* p11 = f + ((int64_t)(y0 + sy1) * W + (x0 + sx1)) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int b = 2 - c;                             // RGB channel c is byte 2 - c of the BGR frame
        const double r0 = __dadd_rn(__dmul_rn((double)p00[b] / 255.0, a0), __dmul_rn((double)p01[b] / 255.0, a1));
        const double r1 = __dadd_rn(__dmul_rn((double)p10[b] / 255.0, a0), __dmul_rn((double)p11[b] / 255.0, a1));
        val[c] = __dadd_rn(__dmul_rn(r0, b0), __dmul_rn(r1, b1));
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[((int64_t)img * 3 + c) * hw + pix] = __fdiv_rn(__fsub_rn((float)val[c], nm.mean[c]), nm.std[c]);
  }
}

}  // namespace
}  // namespace airpose

using namespace airpose;

static unsigned stream_grid(int64_t n) { return (unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 16); }

extern "C" int airpose_preprocess_bgr8(const uint8_t* bgr_hwc, int32_t n_images, int32_t size, const float* mean3, const float* std3,
                                       float* out_nchw, void* stream) {
  AP_REQUIRE(bgr_hwc && out_nchw && mean3 && std3, "airpose_preprocess_bgr8: null argument");
  AP_REQUIRE(n_images > 0 && size > 0, "airpose_preprocess_bgr8: bad sizes (n=%d size=%d)", n_images, size);
  Norm3 nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3[c]; nm.std[c] = std3[c]; }
  const int64_t total = (int64_t)n_images * size * size;
  preprocess_bgr8_kernel<<<stream_grid(total), 256, 0, (cudaStream_t)stream>>>(bgr_hwc, total, size * size, nm, out_nchw);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int airpose_preprocess_crop_resize(const uint8_t* frames_bgr, int64_t frame_stride_bytes, int32_t frame_h, int32_t frame_w,
                                              const int32_t* rects_dev, int32_t n_images, int32_t size, const float* mean3,
                                              const float* std3, float* out_nchw, void* stream) {
  AP_REQUIRE(frames_bgr && rects_dev && out_nchw && mean3 && std3, "airpose_preprocess_crop_resize: null argument");
  AP_REQUIRE(n_images > 0 && size > 0 && frame_h > 0 && frame_w > 0, "airpose_preprocess_crop_resize: bad sizes");
  AP_REQUIRE(frame_stride_bytes == 0 || frame_stride_bytes >= (int64_t)frame_h * frame_w * 3,
             "airpose_preprocess_crop_resize: frame stride smaller than a frame (0 = every crop comes from the same frame)");
  Norm3 nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3[c]; nm.std[c] = std3[c]; }
  const int64_t total = (int64_t)n_images * size * size;
  crop_resize_kernel<<<stream_grid(total), 256, 0, (cudaStream_t)stream>>>(frames_bgr, frame_stride_bytes, frame_w, rects_dev, n_images,
                                                                            size, nm, out_nchw);
  AP_LAUNCH_CHECK();
  return 0;
}
