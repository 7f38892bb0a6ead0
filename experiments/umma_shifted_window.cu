// EXPERIMENT (not part of libairpose_b200.so, not run by tests or bench): can a tcgen05 A operand be a SHIFTED window of a
// TMA-written, 128B-swizzled shared-memory slab?
//
// Why: the 3x3 convolutions of layer1/2 re-read every activation nine times through the L2->SM crossbar (DESIGN.md 3.1:
// 7.4 TB/s of crossbar traffic at 7 % DRAM utilisation).  If the MMA can read rows [s, s+128) of ONE slab for any row shift s,
// a zero-padded halo tile loaded once serves all nine taps (tap (r, c) = shift r * padded_width + c).
//
// What it does: loads S [160 rows][64 bf16] and W [64][64 bf16] with TMA (SWIZZLE_128B), then for each shift s in a list issues
// D_s = S[s : s+128] . W^T (128 x 64 x 64, four K=16 MMAs) with the A descriptor's start address advanced by s * 128 bytes,
// once with the descriptor's base-offset field left 0 and once set to (address >> 7) & 7, and compares with the host result.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I airpose_b200/csrc experiments/umma_shifted_window.cu -lcuda -o /tmp/umma_shift && /tmp/umma_shift
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace airpose;

constexpr int kRows = 160, kK = 64, kN = 64, kM = 128;
constexpr int kNumShifts = 10;
__constant__ int c_shifts[kNumShifts] = {0, 1, 2, 3, 7, 8, 9, 17, 30, 32};

// descriptor with an explicit base offset (bits 49..51: "matrix base offset" for start addresses that are not aligned to the
// 1024-byte swizzle repeat)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_offset) {
  return ptx::make_kmajor_sw128_desc(addr) | ((uint64_t)(base_offset & 7u) << 49);
}

__global__ void __launch_bounds__(128) shifted_window_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmW,
                                                            float* __restrict__ out /* [2][kNumShifts][128][64] */) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sS = smem;                       // 160 x 128 B = 20480 B (a multiple of 1024)
  uint8_t* sW = smem + kRows * 128;         // 64 x 128 B
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_load, 1);
    ptx::mbar_init(&bar_mma, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar_load, (kRows + kN) * 128);
    ptx::tma_load_2d(&tmS, &bar_load, sS, 0, 0);
    ptx::tma_load_2d(&tmW, &bar_load, sW, 0, 0);
  }
  ptx::mbar_wait(&bar_load, 0, 1);
  const uint32_t idesc = ptx::make_idesc_bf16(kM, kN);
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant)
    for (int si = 0; si < kNumShifts; ++si) {
      if (threadIdx.x == 0) {
        const uint32_t a0 = ptx::smem_u32(sS) + (uint32_t)c_shifts[si] * 128u;
        const uint32_t bo = variant ? ((a0 >> 7) & 7u) : 0u;
        ptx::tc_fence_after();
        for (int k = 0; k < kK / 16; ++k)
          ptx::umma_bf16(tmem, desc_sw128(a0 + k * 32, bo), ptx::make_kmajor_sw128_desc(ptx::smem_u32(sW) + k * 32), idesc, k > 0);
        ptx::umma_commit(&bar_mma);
      }
      ptx::mbar_wait(&bar_mma, phase, 2);
      phase ^= 1;
      ptx::tc_fence_after();
      // 4 warps x 32 lanes = 128 rows; 64 columns
      for (int c0 = 0; c0 < kN; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        ptx::tmem_ld_wait();
        float* o = out + (((size_t)variant * kNumShifts + si) * kM + warp * 32 + lane) * kN + c0;
        for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
      }
      ptx::tc_fence_before();
      __syncthreads();
    }
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

static CUtensorMap make_map(void* base, int rows, int cols, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  cudaFree(0);
  std::vector<__nv_bfloat16> hS(kRows * kK), hW(kN * kK);
  std::vector<float> fS(kRows * kK), fW(kN * kK);
  srand(1);
  for (size_t i = 0; i < hS.size(); ++i) { hS[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); fS[i] = __bfloat162float(hS[i]); }
  for (size_t i = 0; i < hW.size(); ++i) { hW[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); fW[i] = __bfloat162float(hW[i]); }
  __nv_bfloat16 *dS, *dW; float* dOut;
  cudaMalloc(&dS, hS.size() * 2); cudaMalloc(&dW, hW.size() * 2); cudaMalloc(&dOut, (size_t)2 * kNumShifts * kM * kN * 4);
  cudaMemcpy(dS, hS.data(), hS.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dOut, 0xff, (size_t)2 * kNumShifts * kM * kN * 4);
  const CUtensorMap tmS = make_map(dS, kRows, kK, kRows), tmW = make_map(dW, kN, kK, kN);
  const int smem = 1024 + (kRows + kN) * 128;
  cudaFuncSetAttribute(shifted_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  shifted_window_kernel<<<1, 128, smem>>>(tmS, tmW, dOut);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hOut((size_t)2 * kNumShifts * kM * kN);
  cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost);
  const int shifts[kNumShifts] = {0, 1, 2, 3, 7, 8, 9, 17, 30, 32};
  for (int variant = 0; variant < 2; ++variant) {
    printf("base offset %s:", variant ? "(addr >> 7) & 7" : "0              ");
    for (int si = 0; si < kNumShifts; ++si) {
      double worst = 0;
      for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < kK; ++k) ref += (double)fS[(m + shifts[si]) * kK + k] * fW[n * kK + k];
          worst = fmax(worst, fabs(ref - hOut[(((size_t)variant * kNumShifts + si) * kM + m) * kN + n]));
        }
      printf("  s=%d %s", shifts[si], worst < 1e-3 ? "ok" : "WRONG");
    }
    printf("\n");
  }
  return 0;
}
