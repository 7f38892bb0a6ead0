// Fused tail of a ResNet bottleneck for the 56x56 stage (layer1, 64 mid channels):
//     conv2 (3x3, stride 1, pad 1) + BN + ReLU  ->  conv3 (1x1) + BN + residual + ReLU
// (Bottleneck.forward, copenet/src/copenet/models/model_copenet.py:33-46) in ONE persistent tcgen05 kernel.
//
// Why: as separate implicit-GEMM launches the 3x3 conv pulls every activation nine times through the
// L2->SM crossbar (347 MB for a 25.7 MB tensor, 47 us per 64 images at 16 % tensor-pipe) and its output makes a round trip
// through HBM before conv3 reads it.  Here a CTA owns a band of R = MT * RM image rows of one image:
//   * ONE 4-d TMA box (64 ch, W+2, R+2, 1) starting at (w, h) = (-1, h0-1) lands the zero-padded halo slab of conv1's output
//     in shared memory in PADDED pixel coordinates p = row * (W+2) + col (out-of-bounds elements are zero-filled by TMA:
//     that IS conv2's zero padding);
//   * tap (dr, dc) of the 3x3 is the same slab read through an A descriptor whose start address is shifted by
//     (dr * (W+2) + dc) rows -- tcgen05 applies the 128-byte swizzle to absolute shared-memory address bits, so any row
//     shift of a TMA-written SWIZZLE_128B slab is a valid operand (experiments/umma_shifted_window.cu, run on B200:
//     profiles/r02e_experiments.txt);
//   * conv2's accumulator goes TMEM -> registers (BN + ReLU) -> bf16 swizzled staging tile, which is conv3's A operand;
//   * conv3's accumulator (N = 128 per step, double buffered) gets BN + the residual, which a second TMA producer streams
//     into a 3-slot ring as (64 ch, W+2, RM, 1) boxes; the result overwrites the residual in place and leaves through a
//     4-d TMA store (the two junk columns per padded row fall outside the tensor and are clipped by the store).
// Both weight matrices stay resident in shared memory (72 + 32 KB) for the life of the CTA.
// Measured and rejected (round 2, gpurun r02g2): the folded BatchNorm vectors as by-value kernel parameters read through the
// constant bank (LDC.64) instead of warp-uniform 16-byte shared loads -- 52.1 -> 58.0 us per launch.
//
// CTA = 12 warps: 0 slab/weight TMA producer, 1 MMA issuer, 2 residual TMA producer, 3 TMEM allocator,
//                 4-11 two epilogue warpgroups (each owns 32 resp. 64 of the columns of every accumulator).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace airpose {

namespace {

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kCM = 64;                    // mid channels
constexpr int kCO = 256;                   // output channels
constexpr int kRow = 128;                  // bytes per pixel row of a 64-channel bf16 tile
constexpr int kSlabRows = 256;             // (RM + 2) * Wp rows loaded; windows reach row 2 * Wp + 2 + 127
constexpr int kSlabBytes = kSlabRows * kRow;          // 32768
constexpr int kW2Bytes = 9 * kCM * kRow;              // 73728: nine [64 cout][64 cin] tap blocks
constexpr int kW3Bytes = kCO * kRow;                  // 32768: [256 cout][64 cin]
constexpr int kTileBytes = 128 * kRow;                // 16384: staging tile / residual slot / output slot
constexpr int kSlabOff = 0;
constexpr int kW2Off = kSlabOff + kSlabBytes;
constexpr int kW3Off = kW2Off + kW2Bytes;
constexpr int kStgOff = kW3Off + kW3Bytes;
constexpr int kResOff = kStgOff + kTileBytes;         // [2]: one residual slot per epilogue warpgroup
constexpr int kOutOff = kResOff + 2 * kTileBytes;     // [2]: one output slot per epilogue warpgroup
constexpr int kBarOff = kOutOff + 2 * kTileBytes;
// barriers: w_full, slab_full, slab_empty, acc2_full, stg_full, stg_empty, acc3_full[2], acc3_empty[2], res_full[2], res_empty[2]
constexpr int kNumBars = 14;
constexpr int kScaleOff = (kBarOff + kNumBars * 8 + 16 + 15) & ~15;        // read with 16-byte shared loads
constexpr int kSmemBytes = 1024 + kScaleOff + (2 * kCM + 2 * kCO) * 4;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
constexpr int kTmemCols = 512;             // acc2: 64 columns from 0; acc3: 2 x 128 columns from 128
constexpr int kAcc3Col = 128;

struct TailParams {
  int H, W, Wp;            // Wp = W + 2
  int RM;                  // image rows per tile = per 128-row M-tile (RM * Wp <= 128)
  int tiles_per_img, num_tiles;
  const float* scale2; const float* shift2;
  const float* scale3; const float* shift3;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
using ptx::lds128;
using ptx::lds_f4;
using ptx::sts128;

// mbarrier arrive that carries a (otherwise unused) register operand: the instructions producing `dep` stay in the program
// and are ordered before the arrive.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, uint32_t dep) {
  asm volatile("{\n\t.reg .b32 t;\n\tmov.b32 t, %1;\n\tmbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(ptx::smem_u32(bar)), "r"(dep) : "memory");
}

// 32 accumulator columns (r) * scale + shift + residual (four 16-byte groups q of bf16) -> ReLU -> bf16 (four 16-byte groups).
// s3 / h3: shared-space addresses of the 32 scales / shifts of these columns.
__device__ __forceinline__ void bn_res_relu(const uint32_t (&r)[32], const uint4* q, uint32_t s3, uint32_t h3, uint4 (&out)[4]) {
  // per 16 columns: the eight scale / shift loads first (volatile asm: they would otherwise queue behind the previous stores)
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    float4 S[4], Hs[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) { S[g] = lds_f4(s3 + jj * 64 + g * 16); Hs[g] = lds_f4(h3 + jj * 64 + g * 16); }
#pragma unroll
    for (int jl = 0; jl < 2; ++jl) {
      const int j = jj * 2 + jl;
      const float2 sc[4] = {make_float2(S[2 * jl].x, S[2 * jl].y), make_float2(S[2 * jl].z, S[2 * jl].w),
                            make_float2(S[2 * jl + 1].x, S[2 * jl + 1].y), make_float2(S[2 * jl + 1].z, S[2 * jl + 1].w)};
      const float2 sh[4] = {make_float2(Hs[2 * jl].x, Hs[2 * jl].y), make_float2(Hs[2 * jl].z, Hs[2 * jl].w),
                            make_float2(Hs[2 * jl + 1].x, Hs[2 * jl + 1].y), make_float2(Hs[2 * jl + 1].z, Hs[2 * jl + 1].w)};
      const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
      uint32_t o[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {     // (acc * scale + shift) + residual, each rounded to fp32 like the reference's two ops
        const float2 a = make_float2(__uint_as_float(r[j * 8 + 2 * h]), __uint_as_float(r[j * 8 + 2 * h + 1]));
        const float2 v = ptx::fadd2(ptx::ffma2(a, sc[h], sh[h]), ptx::bf16x2_to_f2(w[h]));
        o[h] = ptx::cvt_bf16x2_relu(v.x, v.y);
      }
      out[j] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
bneck_tail_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmR,
                  const __grid_constant__ CUtensorMap tmO, const TailParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* w_full = bars + 0;
  uint64_t* slab_full = bars + 1;
  uint64_t* slab_empty = bars + 2;
  uint64_t* acc2_full = bars + 3;
  uint64_t* stg_full = bars + 4;
  uint64_t* stg_empty = bars + 5;
  uint64_t* acc3_full = bars + 6;       // [2]
  uint64_t* acc3_empty = bars + 8;      // [2]
  uint64_t* res_full = bars + 10;       // [2] per warpgroup
  uint64_t* res_empty = bars + 12;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* sc2 = reinterpret_cast<float*>(smem + kScaleOff);
  float* sh2 = sc2 + kCM;
  float* sc3 = sh2 + kCM;
  float* sh3 = sc3 + kCO;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt_rows = p.RM * p.Wp;                 // valid rows of the M-tile (<= 128)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmT); ptx::prefetch_tmap(&tmW2); ptx::prefetch_tmap(&tmW3); ptx::prefetch_tmap(&tmR); ptx::prefetch_tmap(&tmO);
    ptx::mbar_init(w_full, 1);
    ptx::mbar_init(slab_full, 1); ptx::mbar_init(slab_empty, 1);
    ptx::mbar_init(acc2_full, 1);
    ptx::mbar_init(stg_full, kEpiWarps); ptx::mbar_init(stg_empty, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc3_full[i], 1); ptx::mbar_init(&acc3_empty[i], kEpiWarps);
      ptx::mbar_init(&res_full[i], 1); ptx::mbar_init(&res_empty[i], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 3) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kCM; i += kThreads) { sc2[i] = __ldg(p.scale2 + i); sh2[i] = __ldg(p.shift2 + i); }
  for (int i = threadIdx.x; i < kCO; i += kThreads) { sc3[i] = __ldg(p.scale3 + i); sh3[i] = __ldg(p.shift3 + i); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {          // the weights do not depend on the previous kernel
    ptx::mbar_arrive_expect_tx(w_full, kW2Bytes + kW3Bytes);
    for (int tap = 0; tap < 9; ++tap) ptx::tma_load_2d(&tmW2, w_full, smem + kW2Off + tap * (kCM * kRow), tap * kCM, 0);
    ptx::tma_load_2d(&tmW3, w_full, smem + kW3Off, 0, 0);
  }
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ slab producer
    if (lane == 0) {
      const uint32_t slab_tx = (uint32_t)((p.RM + 2) * p.Wp * kRow);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int n = tile / p.tiles_per_img, h0 = (tile - n * p.tiles_per_img) * p.RM;
        ptx::mbar_wait(slab_empty, (uint32_t)((it & 1) ^ 1), 100);
        ptx::mbar_arrive_expect_tx(slab_full, slab_tx);
        ptx::tma_load_4d(&tmT, slab_full, smem + kSlabOff, 0, -1, h0 - 1, n);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc2 = ptx::make_idesc_bf16(128, kCM);
      constexpr uint32_t idesc3 = ptx::make_idesc_bf16(128, 128);
      const uint32_t slab_a = ptx::smem_u32(smem + kSlabOff);
      const uint32_t w2_a = ptx::smem_u32(smem + kW2Off);
      const uint32_t w3_a = ptx::smem_u32(smem + kW3Off);
      const uint32_t stg_a = ptx::smem_u32(smem + kStgOff);
      ptx::mbar_wait(w_full, 0, 200);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        ptx::mbar_wait(slab_full, (uint32_t)(it & 1), 201);
        ptx::tc_fence_after();
        // conv2: nine row-shifted windows of the slab.  acc2 is free: stg_full of the previous tile was waited on below,
        // and the epilogue arrives on it only after it has read acc2.
        for (int tap = 0; tap < 9; ++tap) {
          const int dr = tap / 3, dc = tap - dr * 3;
          const uint32_t a0 = slab_a + (uint32_t)((dr * p.Wp + dc) * kRow);
          const uint32_t b0 = w2_a + (uint32_t)(tap * (kCM * kRow));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(tmem_base, ptx::make_kmajor_sw128_desc(a0 + k * 32), ptx::make_kmajor_sw128_desc(b0 + k * 32), idesc2,
                           (tap | k) != 0);
        }
        ptx::umma_commit(slab_empty);
        ptx::umma_commit(acc2_full);
        ptx::mbar_wait(stg_full, (uint32_t)(it & 1), 202);
        ptx::tc_fence_after();
        for (int half = 0; half < 2; ++half) {
          ptx::mbar_wait(&acc3_empty[half], (uint32_t)((it & 1) ^ 1), 203 + half);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(tmem_base + kAcc3Col + half * 128, ptx::make_kmajor_sw128_desc(stg_a + k * 32),
                           ptx::make_kmajor_sw128_desc(w3_a + half * (128 * kRow) + k * 32), idesc3, k != 0);
          ptx::umma_commit(&acc3_full[half]);
        }
        ptx::umma_commit(stg_empty);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual producer: chunk (half, wg) -> slot wg
    if (lane == 0) {
      const uint32_t res_tx = (uint32_t)(mt_rows * kRow);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int n = tile / p.tiles_per_img, h0 = (tile - n * p.tiles_per_img) * p.RM;
        for (int half = 0; half < 2; ++half)
          for (int wg = 0; wg < 2; ++wg) {
            const int j = it * 2 + half;               // this warpgroup's running chunk number
            ptx::mbar_wait(&res_empty[wg], (uint32_t)((j & 1) ^ 1), 300 + wg);
            ptx::mbar_arrive_expect_tx(&res_full[wg], res_tx);
            ptx::tma_load_4d(&tmR, &res_full[wg], smem + kResOff + wg * kTileBytes, (2 * half + wg) * 64, 0, h0, n);
          }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue warpgroups
    const int e = warp - kEpiWarp0;
    const int wg = e >> 2;                           // 0 / 1: which columns of every accumulator
    const int quad = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = quad * 32 + lane;                // row of the M-tile == TMEM lane
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int wgt = threadIdx.x - (kEpiWarp0 + 4 * wg) * 32;     // 0..127 inside the warpgroup
    const uint32_t srow = ptx::smem_u32(smem + kStgOff + row * kRow);
    const uint32_t rrow = ptx::smem_u32(smem + kResOff + wg * kTileBytes + row * kRow);
    const uint32_t orow = ptx::smem_u32(smem + kOutOff + wg * kTileBytes + row * kRow);
    const uint32_t sc2_a = ptx::smem_u32(sc2), sh2_a = ptx::smem_u32(sh2), sc3_a = ptx::smem_u32(sc3), sh3_a = ptx::smem_u32(sh3);
    // The residual of the NEXT chunk sits in registers while the current one is processed: the slot is handed back to the
    // producer as soon as it has been read, so a load is always in flight (the kernel is HBM-bound, not MMA-bound).
    uint4 resA[8], resB[8];
    int jr = 0;                                      // chunks of this warpgroup read from the residual slot so far
    auto read_res = [&](uint4 (&dst)[8]) {
      ptx::mbar_wait(&res_full[wg], (uint32_t)(jr & 1), 410 + wg);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = lds128(rrow + (((uint32_t)j ^ swz) << 4));
      // The slot may be refilled by TMA as soon as res_empty completes, so every lane's loads must have RETURNED (not just
      // issued) before lane 0 arrives: the vote below reads a value derived from all eight loads of every lane.  (Under a
      // busy shared-memory pipe -- UMMA operand reads, TMA writes -- an issued LDS can trail an mbarrier arrive by hundreds of
      // cycles; seen on B200 as rare single-tile corruption, tools/diag_determinism.py.)
      uint32_t acc = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc |= dst[j].x ^ dst[j].w;
      const uint32_t dep = __ballot_sync(0xffffffffu, acc == 0x7fc1a5e3u);
      if (lane == 0) mbar_arrive_after(&res_empty[wg], dep);
      ++jr;
    };
    if ((int)blockIdx.x < p.num_tiles) read_res(resA);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int n = tile / p.tiles_per_img, h0 = (tile - n * p.tiles_per_img) * p.RM;
      const bool more = tile + (int)gridDim.x < p.num_tiles;
      // ---- conv2 epilogue: BN + ReLU -> bf16 staging tile (conv3's A operand)
      ptx::mbar_wait(acc2_full, (uint32_t)(it & 1), 400);
      ptx::tc_fence_after();
      {
        uint32_t r[32];
        ptx::tmem_ld_32x32(lane_addr + wg * 32, r);
        ptx::tmem_ld_wait();
        ptx::mbar_wait(stg_empty, (uint32_t)((it & 1) ^ 1), 401);     // conv3 of the previous tile has read the staging tile
        float4 S[8], Hs[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) { S[g] = lds_f4(sc2_a + wg * 128 + g * 16); Hs[g] = lds_f4(sh2_a + wg * 128 + g * 16); }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 sc[4] = {make_float2(S[2 * j].x, S[2 * j].y), make_float2(S[2 * j].z, S[2 * j].w),
                                make_float2(S[2 * j + 1].x, S[2 * j + 1].y), make_float2(S[2 * j + 1].z, S[2 * j + 1].w)};
          const float2 sh[4] = {make_float2(Hs[2 * j].x, Hs[2 * j].y), make_float2(Hs[2 * j].z, Hs[2 * j].w),
                                make_float2(Hs[2 * j + 1].x, Hs[2 * j + 1].y), make_float2(Hs[2 * j + 1].z, Hs[2 * j + 1].w)};
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 v = ptx::ffma2(make_float2(__uint_as_float(r[j * 8 + 2 * i]), __uint_as_float(r[j * 8 + 2 * i + 1])), sc[i], sh[i]);
            o[i] = ptx::cvt_bf16x2_relu(v.x, v.y);
          }
          sts128(srow + (((uint32_t)(wg * 4 + j) ^ swz) << 4), make_uint4(o[0], o[1], o[2], o[3]));
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(stg_full);
      }
      // ---- conv3 epilogue: BN + residual + ReLU -> output slot -> 4-d TMA store (junk columns / rows are clipped)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = 2 * half + wg;                 // 64-channel chunk of the 256 outputs
        uint4 (&cur)[8] = half ? resB : resA;
        if (half == 0) read_res(resB);               // next chunk of this warpgroup (same tile, always exists)
        else if (more) read_res(resA);               // first chunk of the next tile
        ptx::mbar_wait(&acc3_full[half], (uint32_t)(it & 1), 402 + half);
        ptx::tc_fence_after();
        const uint32_t taddr = lane_addr + kAcc3Col + half * 128 + wg * 64;
        const uint32_t s3 = sc3_a + c * 256, h3 = sh3_a + c * 256;
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr, r);
        ptx::tmem_ld_wait();
        uint4 o[4];
        bn_res_relu(r, &cur[0], s3, h3, o);
        // the previous store of this warpgroup must have read the output slot before it is overwritten
        if (wgt == 0) ptx::tma_store_wait_read<0>();
        ptx::named_bar_sync(1 + wg, 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(orow + (((uint32_t)j ^ swz) << 4), o[j]);
        ptx::tmem_ld_32x32(taddr + 32, r);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc3_empty[half]);
        bn_res_relu(r, &cur[4], s3 + 128, h3 + 128, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(orow + (((uint32_t)(4 + j) ^ swz) << 4), o[j]);
        ptx::fence_proxy_async();
        ptx::named_bar_sync(3 + wg, 128);
        if (wgt == 0) {
          ptx::tma_store_4d(&tmO, smem + kOutOff + wg * kTileBytes, c * 64, 0, h0, n);
          ptx::tma_store_commit();
        }
      }
    }
    if (wgt == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

struct TailLaunchImpl {
  CUtensorMap tmT, tmW2, tmW3, tmR, tmO;
  TailParams p;
};
static_assert(sizeof(TailLaunchImpl) <= sizeof(TailLaunch::storage), "TailLaunch::storage too small");

bool bneck_tail_supported(int H, int W, int Cm, int Cout) {
  return Cm == kCM && Cout == kCO && 2 * (W + 2) <= 128 && H >= 1;
}

int build_bneck_tail(TailLaunch* L, const void* t1, const void* w2, const float* scale2, const float* shift2, const void* w3,
                     const float* scale3, const float* shift3, const void* residual, void* out, int n, int H, int W) {
  AP_REQUIRE(bneck_tail_supported(H, W, kCM, kCO), "build_bneck_tail: unsupported geometry H=%d W=%d", H, W);
  TailLaunchImpl& I = *reinterpret_cast<TailLaunchImpl*>(L->storage);
  TailParams& p = I.p;
  p.H = H; p.W = W; p.Wp = W + 2;
  p.RM = std::min(128 / p.Wp, kSlabRows / p.Wp - 2);
  AP_REQUIRE(p.RM >= 1 && 2 * p.Wp + 2 + 128 <= kSlabRows && (p.RM + 2) * p.Wp <= kSlabRows,
             "build_bneck_tail: the halo slab of a %dx%d band does not fit", p.RM, W);
  p.tiles_per_img = ceil_div(H, p.RM);
  p.num_tiles = n * p.tiles_per_img;
  p.scale2 = scale2; p.shift2 = shift2; p.scale3 = scale3; p.shift3 = shift3;
  if (make_tmap_nhwc4d_bf16(&I.tmT, t1, n, H, W, kCM, p.Wp, p.RM + 2)) return 1;
  if (make_tmap_tiled_bf16(&I.tmW2, w2, kCM, 9 * kCM, 9 * kCM, kCM, 64)) return 1;
  if (make_tmap_tiled_bf16(&I.tmW3, w3, kCO, kCM, kCM, kCO, 64)) return 1;
  if (make_tmap_nhwc4d_bf16(&I.tmR, residual, n, H, W, kCO, p.Wp, p.RM)) return 1;
  if (make_tmap_nhwc4d_bf16(&I.tmO, out, n, H, W, kCO, p.Wp, p.RM)) return 1;
  L->valid = 1;
  L->pdl = use_pdl();
  return 0;
}

int launch_bneck_tail(const TailLaunch& L, cudaStream_t stream) {
  AP_REQUIRE(L.valid, "launch_bneck_tail: launch was not built");
  const TailLaunchImpl& I = *reinterpret_cast<const TailLaunchImpl*>(L.storage);
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(bneck_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  if (I.p.num_tiles == 0) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(I.p.num_tiles, grid_limit()));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = L.pdl ? 1 : 0;
  AP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, bneck_tail_kernel, I.tmT, I.tmW2, I.tmW3, I.tmR, I.tmO, I.p));
  count_launch();
  return 0;
}

}  // namespace airpose

using namespace airpose;

extern "C" int airpose_bneck_tail_bf16(const airpose_bneck_tail_args* a, void* stream) {
  AP_REQUIRE(a && a->t1 && a->w2 && a->w3 && a->residual && a->out && a->scale2 && a->shift2 && a->scale3 && a->shift3,
             "airpose_bneck_tail_bf16: null argument");
  AP_REQUIRE(bneck_tail_supported(a->H, a->W, a->Cm, 4 * a->Cm), "airpose_bneck_tail_bf16: unsupported geometry (H=%d W=%d Cm=%d)", a->H,
             a->W, a->Cm);
  AP_REQUIRE(a->n >= 0, "airpose_bneck_tail_bf16: negative image count");
  if (a->n == 0) return 0;
  TailLaunch L{};
  if (build_bneck_tail(&L, a->t1, a->w2, a->scale2, a->shift2, a->w3, a->scale3, a->shift3, a->residual, a->out, a->n, a->H, a->W)) return 1;
  return launch_bneck_tail(L, (cudaStream_t)stream);
}
