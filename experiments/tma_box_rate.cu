// EXPERIMENT (not part of libairpose_b200.so, not run by tests or bench): how fast does TMA deliver 128-row x 128-byte boxes to
// one SM from an L2-resident activation, by addressing mode and by the number of boxes in flight?
//
// Why: in the ncu capture of the trunk (profiles/r01s_ncu_full_trunk_128img.csv) every 24 KB k-block of the layer1/2 convs takes
// ~1000 clk to arrive although no unit is saturated.  This isolates the load side: persistent CTAs (one per SM)
// stream boxes through an S-stage mbarrier ring and drop them -- no MMA, no epilogue -- so the number is the ceiling any
// conv kernel with that operand layout can reach.
//   modes: 0 tiled, contiguous rows (a [M, 64] matrix: 1x1 conv with Cin = 64)
//          1 tiled, strided rows    (64 of the 256 columns of a [M, 256] matrix: 1x1 conv with Cin = 256, one k-block)
//          2 im2col 3x3, pad 1      (TMA im2col map over NHWC [n, 56, 56, 64], all nine taps of each 128-pixel tile)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I airpose_b200/csrc experiments/tma_box_rate.cu -lcuda -o /tmp/tma_rate && /tmp/tma_rate
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"

using namespace airpose;

constexpr int kBoxBytes = 128 * 128;
constexpr int kMaxStages = 12;

struct Args { int mode, stages, tiles, W, HW; };

__global__ void __launch_bounds__(32) stream_boxes_kernel(const __grid_constant__ CUtensorMap tm, Args a, unsigned long long* clocks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kMaxStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  const int per_tile = a.mode == 2 ? 9 : (a.mode == 1 ? 4 : 1);          // boxes per 128-row tile
  long long issued = 0, waited = 0;
  const long long t0 = clock64();
  auto issue = [&](int tile, int j) {
    const int slot = (int)(issued % a.stages);
    ptx::mbar_arrive_expect_tx(&full[slot], kBoxBytes);
    void* dst = smem + (size_t)slot * kBoxBytes;
    if (a.mode == 2) {
      const int m0 = tile * 128;
      const int img = m0 / a.HW, rem = m0 % a.HW;
      ptx::tma_load_im2col_4d(&tm, &full[slot], dst, 0, rem % a.W - 1, rem / a.W - 1, img, (uint16_t)(j % 3), (uint16_t)(j / 3));
    } else {
      ptx::tma_load_2d(&tm, &full[slot], dst, j * 64, tile * 128);
    }
    ++issued;
  };
  auto wait_one = [&]() {
    const int slot = (int)(waited % a.stages);
    ptx::mbar_wait(&full[slot], (uint32_t)((waited / a.stages) & 1), 7);
    ++waited;
  };
  for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x)
    for (int j = 0; j < per_tile; ++j) {
      if (issued - waited == a.stages) wait_one();      // the slot about to be reused has landed (nothing reads it: dropped)
      issue(tile, j);
    }
  while (waited < issued) wait_one();
  clocks[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

static void check(CUresult r, const char* what) { if (r != CUDA_SUCCESS) { printf("%s failed: %d\n", what, (int)r); exit(1); } }

int main() {
  cudaFree(0);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n = 64, H = 56, W = 56;
  const size_t M = (size_t)n * H * W;
  __nv_bfloat16* x;
  cudaMalloc(&x, M * 256 * 2);                 // large enough for the [M, 256] view
  cudaMemset(x, 0, M * 256 * 2);
  unsigned long long* dclk;
  cudaMalloc(&dclk, sms * sizeof(unsigned long long));
  CUtensorMap maps[3];
  {
    cuuint64_t dims[2] = {64, M}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    check(cuTensorMapEncodeTiled(&maps[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE), "tiled 64");
    cuuint64_t dims2[2] = {256, M}; cuuint64_t strides2[1] = {512};
    check(cuTensorMapEncodeTiled(&maps[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims2, strides2, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE), "tiled 256");
    cuuint64_t d4[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t s3[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1};
    cuuint32_t e4[4] = {1, 1, 1, 1};
    check(cuTensorMapEncodeIm2col(&maps[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, d4, s3, lower, upper, 64, 128, e4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE), "im2col");
  }
  const char* names[3] = {"tiled, contiguous rows", "tiled, 512-byte pitch ", "im2col 3x3            "};
  const int smem = 1024 + kMaxStages * kBoxBytes;
  cudaFuncSetAttribute(stream_boxes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode)
    for (int stages : {1, 2, 4, 6, 8, 12}) {
      Args a{mode, stages, (int)(M / 128), W, H * W};
      const double boxes = (double)a.tiles * (mode == 2 ? 9 : (mode == 1 ? 4 : 1));
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {       // the first repetition also brings the tensor into L2
        cudaEventRecord(e0);
        stream_boxes_kernel<<<sms, 32, smem>>>(maps[mode], a, dclk);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed (mode %d stages %d): %s\n", mode, stages, cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      const double bytes = boxes * kBoxBytes;
      printf("%s stages %2d: %8.1f us  %6.2f TB/s chip  %5.1f B/clk/SM @1.965 GHz  %6.0f clk per box per SM\n", names[mode], stages, best * 1e3,
             bytes / (best * 1e-3) / 1e12, bytes / sms / (best * 1e-3 * 1.965e9), best * 1e-3 * 1.965e9 / (boxes / sms));
    }
  return 0;
}
