// Host-side interface of the tcgen05 GEMM / implicit-conv kernel (gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace airpose {

// Epilogue of one GEMM: v = acc*scale[n] + shift[n] (+ residual[m,n]); relu; store.
struct Epilogue {
  const float* scale = nullptr;           // [N] or null (=1)
  const float* shift = nullptr;           // [N] or null (=0)
  const void* residual = nullptr;         // [M,N] bf16 (or fp32 when residual_f32) or null
  int64_t ldr = 0;
  int residual_f32 = 0;
  int relu = 0;
  void* out_bf16 = nullptr; int64_t ldd = 0;     // bf16 [M,N]
  float* out_f32 = nullptr; int64_t ldf = 0;     // fp32 [M,N]
  void* out_split = nullptr; int64_t lds = 0;    // bf16 [M,3N]: hi | lo | hi  (split-bf16 operand of the next GEMM)
};

// Geometry of the A operand when it is an NHWC activation read through TMA im2col.
struct ConvGeom {
  int n = 0, H = 0, W = 0, Cin = 0;
  int Ho = 0, Wo = 0;
  int ksize = 1, stride = 1, pad = 0;
};

struct GemmLaunch {
  CUtensorMap tmA;     // tiled [M,K] map, or im2col map over (C,W,H,N)
  CUtensorMap tmB;     // tiled [N,K] map
  int M = 0, N = 0, K = 0;
  int block_n = 128;   // 64, 128 or 256
  int im2col = 0;
  ConvGeom geom;
  Epilogue epi;
};

// Tensor maps (cuTensorMapEncode* resolved through cudaGetDriverEntryPoint; no libcuda link).
int make_tmap_tiled_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld_elems,
                         int box_rows, int box_cols);
int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, const ConvGeom& g, int channels_per_pixel,
                          int pixels_per_column);

int pick_block_n(int M, int N);
int launch_gemm(const GemmLaunch& L, cudaStream_t stream);
int num_sms();

}  // namespace airpose
