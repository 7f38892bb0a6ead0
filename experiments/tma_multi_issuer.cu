// EXPERIMENT (not part of libairpose_b200.so): does the per-SM TMA delivery rate (~15-17 B/clk/SM in tma_box_rate.cu, ~30 B/clk/SM
// for the A + B streams of the conv GEMMs) rise when SEVERAL warps of a CTA issue boxes concurrently, each with its own ring?
// If it does, the conv kernels should split their operand loads over more producer warps; if not, the limit is the SM's
// ingest path and only fewer bytes per MMA help.  L2-resident operands: a [M, 64] bf16 matrix of 51 MB read repeatedly
// (tiled, 128-row x 128-byte boxes) and the 3x3 im2col view of a 25.7 MB NHWC tensor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I airpose_b200/csrc experiments/tma_multi_issuer.cu -lcuda -o /tmp/tma_multi && /tmp/tma_multi
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"

using namespace airpose;

constexpr int kBoxBytes = 128 * 128;
constexpr int kMaxProd = 4, kMaxStages = 6;

struct Args { int mode, stages, nprod, tiles, W, HW, passes, box_rows; };

__global__ void __launch_bounds__(32 * kMaxProd) stream_kernel(const __grid_constant__ CUtensorMap tm, Args a) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kMaxProd][kMaxStages];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < a.nprod) {
    for (int s = 0; s < a.stages; ++s) ptx::mbar_init(&full[w][s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if ((threadIdx.x & 31) != 0 || w >= a.nprod) return;
  const int box_bytes = a.box_rows * 128;
  uint8_t* ring = smem + (size_t)w * a.stages * kBoxBytes;
  const int per_tile = (a.mode == 2 ? 9 : 1) * (128 / a.box_rows);
  long long issued = 0, waited = 0;
  auto wait_one = [&]() {
    const int slot = (int)(waited % a.stages);
    ptx::mbar_wait(&full[w][slot], (uint32_t)((waited / a.stages) & 1), 7);
    ++waited;
  };
  for (int pass = 0; pass < a.passes; ++pass)
    for (int tile = blockIdx.x * a.nprod + w; tile < a.tiles; tile += gridDim.x * a.nprod)
      for (int j = 0; j < per_tile; ++j) {
        if (issued - waited == a.stages) wait_one();
        const int slot = (int)(issued % a.stages);
        ptx::mbar_arrive_expect_tx(&full[w][slot], box_bytes);
        void* dst = ring + (size_t)slot * kBoxBytes;
        if (a.mode == 2) {
          const int m0 = tile * 128;
          const int img = m0 / a.HW, rem = m0 % a.HW;
          ptx::tma_load_im2col_4d(&tm, &full[w][slot], dst, 0, rem % a.W - 1, rem / a.W - 1, img, (uint16_t)(j % 3), (uint16_t)(j / 3));
        } else {
          ptx::tma_load_2d(&tm, &full[w][slot], dst, 0, tile * 128 + j * a.box_rows);
        }
        ++issued;
      }
  while (waited < issued) wait_one();
}

static void check(CUresult r, const char* what) { if (r != CUDA_SUCCESS) { printf("%s failed: %d\n", what, (int)r); exit(1); } }

int main() {
  cudaFree(0);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n = 128, H = 56, W = 56;
  const size_t M = (size_t)n * H * W;                  // 401408 rows x 128 B = 51 MB
  __nv_bfloat16* x;
  cudaMalloc(&x, M * 64 * 2);
  cudaMemset(x, 0, M * 64 * 2);
  CUtensorMap tiled128, tiled64, tiled32, im2col;
  cuuint64_t dims[2] = {64, M}; cuuint64_t strides[1] = {128}; cuuint32_t es[2] = {1, 1};
  for (int br : {128, 64, 32}) {
    cuuint32_t box[2] = {64, (cuuint32_t)br};
    check(cuTensorMapEncodeTiled(br == 128 ? &tiled128 : (br == 64 ? &tiled64 : &tiled32), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE), "tiled");
  }
  {
    cuuint64_t d4[4] = {64, (cuuint64_t)W, (cuuint64_t)H, 64};
    cuuint64_t s3[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1};
    cuuint32_t e4[4] = {1, 1, 1, 1};
    check(cuTensorMapEncodeIm2col(&im2col, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, d4, s3, lower, upper, 64, 128, e4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE), "im2col");
  }
  const int smem = 1024 + 208 * 1024;                             // issuers x stages x 16 KB is capped at 200 KB below
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Case { const char* name; int mode; const CUtensorMap* tm; int box_rows; int tiles; int W, HW; };
  const Case cases[] = {{"tiled 128-row boxes, 51 MB", 0, &tiled128, 128, (int)(M / 128), 0, 0},
                        {"tiled  64-row boxes, 51 MB", 0, &tiled64, 64, (int)(M / 128), 0, 0},
                        {"tiled  32-row boxes, 51 MB", 0, &tiled32, 32, (int)(M / 128), 0, 0},
                        {"im2col 3x3, 25.7 MB tensor", 2, &im2col, 128, (int)(64 * H * W / 128), W, H * W}};
  for (const Case& c : cases)
    for (int nprod : {1, 2, 4})
      for (int stages : {2, 3, 6}) {
        if (nprod * stages * kBoxBytes > 200 * 1024) continue;
        const int bytes_smem = smem;
        cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes_smem);
        Args a{c.mode, stages, nprod, c.tiles, c.W, c.HW, 3, c.box_rows};
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          cudaEventRecord(e0);
          stream_kernel<<<sms, 32 * kMaxProd, bytes_smem>>>(*c.tm, a);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("kernel failed (%s nprod %d stages %d): %s\n", c.name, nprod, stages, cudaGetErrorString(e)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (rep > 0 && ms < best) best = ms;
        }
        const double bytes = (double)a.passes * c.tiles * (c.mode == 2 ? 9 : 1) * kBoxBytes;
        printf("%s  issuers %d x %d stages (%3d KB in flight): %8.1f us  %6.2f TB/s chip  %5.1f B/clk/SM\n", c.name, nprod, stages,
               nprod * stages * c.box_rows * 128 / 1024, best * 1e3, bytes / (best * 1e-3) / 1e12, bytes / sms / (best * 1e-3 * 1.965e9));
      }
  return 0;
}
