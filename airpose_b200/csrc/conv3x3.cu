// 3x3 / stride 1 / pad 1 convolution + BN + ReLU for the 128-channel stage (layer2's conv2: Bottleneck.forward,
// copenet/src/copenet/models/model_copenet.py:37-39) as ONE persistent tcgen05 kernel that keeps the input band in shared memory.
//
// Why: as an implicit GEMM over a TMA im2col map this layer pulls every activation nine times and every weight
// k-block once per 128-pixel tile through the SM's L2 port -- 576 KB per tile for 4608 clk of MMA, 29 B/clk against the ~30 B/clk
// a conv kernel gets out of TMA: 28 us per 64 images at 32 % tensor pipe with NO unit above a third of its peak.  Here a work unit
// is a band of R image rows of one image = two 128-row M-tiles in PADDED pixel coordinates p = row * (W + 2) + col:
//   * the band's halo slab arrives ONCE: per 64-channel half one 4-d TMA box (64 ch, W + 2, R + 2, 1) starting at (w, h) =
//     (-1, h0 - 1); out-of-bounds elements are zero-filled by TMA: that IS the conv's padding;
//   * tap (dr, dc) of M-tile t is the same slab through an A descriptor shifted by (128 t + dr (W + 2) + dc) rows (tcgen05 applies
//     the 128-byte swizzle to absolute shared-memory address bits: any whole-row shift is a valid operand, see bneck.cu);
//   * the weights stream through a ring of [128 cout x 64 cin] blocks, one block per (channel half, tap), each used by BOTH
//     M-tiles: 288 KB of weights + 68 KB of slab per 420 output pixels instead of 1.15 MB;
//   * the slab halves are released one after the other (channel-half-major tap order), so the next band's first half loads
//     while this band's second half is still being multiplied;
//   * epilogue (two warpgroups, one per M-tile): TMEM -> BN + ReLU -> bf16 -> swizzled staging chunk -> 4-d TMA store of
//     (64 ch, W + 2, R, 1): the two junk columns per padded row fall outside the tensor and are clipped by the store.
//
// CTA = 12 warps: 0 slab TMA producer, 1 MMA issuer, 2-3 weight TMA producers (3 also allocates TMEM), 4-11 epilogue.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace airpose {

namespace {

constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kC = 128;                           // input = output channels
constexpr int kRowB = 128;                        // bytes per 64-channel pixel row
constexpr int kSlabRows = 320;                    // windows of M-tile 1 reach row 128 + 2 Wp + 2 + 127
constexpr int kSlabHalf = kSlabRows * kRowB;      // 40960: one 64-channel half of the slab
constexpr int kWBlk = kC * kRowB;                 // 16384: [128 cout][64 cin]
constexpr int kWStages = 6;
constexpr int kStgBytes = 256 * kRowB;            // 32768: one 64-channel chunk of both M-tiles
constexpr int kSlabOff = 0;
constexpr int kWOff = kSlabOff + 2 * kSlabHalf;   // 81920
constexpr int kStgOff = kWOff + kWStages * kWBlk; // 180224
constexpr int kBarOff = kStgOff + kStgBytes;      // 212992
// barriers: slab_full[2], slab_empty[2], w_full[S], w_empty[S], acc_full[2], acc_empty[2]
constexpr int kNumBars = 4 + 2 * kWStages + 4;
constexpr int kScaleOff = (kBarOff + kNumBars * 8 + 16 + 15) & ~15;
constexpr int kSmemBytes = 1024 + kScaleOff + 2 * kC * 4;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
static_assert(kSlabHalf % 1024 == 0 && kWBlk % 1024 == 0 && kStgOff % 1024 == 0, "swizzle atom alignment");
constexpr int kTmemCols = 512;                    // 2 accumulator sets x 2 M-tiles x 128 columns

struct SlabParams {
  int H, W, Wp;            // Wp = W + 2
  int R;                   // image rows per unit: R * Wp <= 256
  int units_per_img, num_units;
  int relu;
  const float* scale; const float* shift;
};

using ptx::lds_f4;
using ptx::sts128;

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_slab_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO, const SlabParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* slab_full = bars;                     // [2] per channel half
  uint64_t* slab_empty = bars + 2;                // [2]
  uint64_t* w_full = bars + 4;                    // [kWStages]
  uint64_t* w_empty = w_full + kWStages;          // [kWStages]
  uint64_t* acc_full = w_empty + kWStages;        // [2] per accumulator set
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* sc = reinterpret_cast<float*>(smem + kScaleOff);
  float* sh = sc + kC;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmO);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&slab_full[i], 1); ptx::mbar_init(&slab_empty[i], 1);
      ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], kEpiWarps);
    }
    for (int s = 0; s < kWStages; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 3) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kC; i += kThreads) {
    sc[i] = p.scale ? __ldg(p.scale + i) : 1.f;
    sh[i] = p.shift ? __ldg(p.shift + i) : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ slab producer: two 64-channel halves per unit
    if (lane == 0) {
      const uint32_t slab_tx = (uint32_t)((p.R + 2) * p.Wp * kRowB);
      int it = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
        const int n = u / p.units_per_img, h0 = (u - n * p.units_per_img) * p.R;
        for (int kh = 0; kh < 2; ++kh) {
          ptx::mbar_wait(&slab_empty[kh], (uint32_t)((it & 1) ^ 1), 100 + kh);
          ptx::mbar_arrive_expect_tx(&slab_full[kh], slab_tx);
          ptx::tma_load_4d(&tmX, &slab_full[kh], smem + kSlabOff + kh * kSlabHalf, kh * 64, -1, h0 - 1, n);
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ------------------------------------------------------------------ weight producers: 18 blocks per unit, (half, tap) order.
    // TWO issuing warps, alternate blocks: one thread gets ~one 16 KB box per 1000 clk out of TMA whatever the ring depth
    // (experiments/tma_multi_issuer.cu: 15.6 B/clk/SM per issuer, issuers in different warps add up) -- with a single issuer the
    // 18 weight blocks of a unit took 19 k clk against 9 k clk of MMA and the kernel was no faster than the im2col GEMM.
    if (lane == 0) {
      const int me = warp - 2;
      long long b = 0;                               // running block number of this CTA: stage = b % S, phase = (b / S) & 1
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x)
        for (int kh = 0; kh < 2; ++kh)
          for (int tap = 0; tap < 9; ++tap, ++b) {
            if ((int)(b & 1) != me) continue;
            const int s = (int)(b % kWStages);
            const uint32_t ph = (uint32_t)((b / kWStages) & 1);
            ptx::mbar_wait(&w_empty[s], ph ^ 1, 300 + s);
            ptx::mbar_arrive_expect_tx(&w_full[s], kWBlk);
            ptx::tma_load_2d(&tmW, &w_full[s], smem + kWOff + s * kWBlk, tap * kC + kh * 64, 0);
          }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, kC);
      const uint32_t slab_a = ptx::smem_u32(smem + kSlabOff);
      const uint32_t w_a = ptx::smem_u32(smem + kWOff);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
        const int ab = it & 1;
        ptx::mbar_wait(&acc_empty[ab], (uint32_t)(((it >> 1) & 1) ^ 1), 200 + ab);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + ab * 256;
        for (int kh = 0; kh < 2; ++kh) {
          ptx::mbar_wait(&slab_full[kh], (uint32_t)(it & 1), 210 + kh);
          ptx::tc_fence_after();
          for (int tap = 0; tap < 9; ++tap) {
            const int dr = tap / 3, dc = tap - dr * 3;
            ptx::mbar_wait(&w_full[s], ph, 220 + s);
            ptx::tc_fence_after();
            const uint32_t b0 = w_a + (uint32_t)(s * kWBlk);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint32_t a0 = slab_a + (uint32_t)(kh * kSlabHalf) + (uint32_t)((mt * 128 + dr * p.Wp + dc) * kRowB);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_bf16(d0 + mt * 128, ptx::make_kmajor_sw128_desc(a0 + k * 32), ptx::make_kmajor_sw128_desc(b0 + k * 32), idesc,
                               (kh | tap | k) != 0);
            }
            ptx::umma_commit(&w_empty[s]);
            if (++s == kWStages) { s = 0; ph ^= 1; }
          }
          ptx::umma_commit(&slab_empty[kh]);      // this half of the slab may be refilled with the next band
        }
        ptx::umma_commit(&acc_full[ab]);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue: warpgroup = M-tile
    const int wg = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = quad * 32 + lane;                // row of the M-tile == TMEM lane
    const uint32_t swz = (uint32_t)(row & 7);
    const int et = threadIdx.x - kEpiWarp0 * 32;     // 0..255
    const uint32_t srow = ptx::smem_u32(smem + kStgOff) + (uint32_t)((wg * 128 + row) * kRowB);
    const uint32_t sc_a = ptx::smem_u32(sc), sh_a = ptx::smem_u32(sh);
    int it = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
      const int ab = it & 1;
      const int n = u / p.units_per_img, h0 = (u - n * p.units_per_img) * p.R;
      ptx::mbar_wait(&acc_full[ab], (uint32_t)((it >> 1) & 1), 400 + ab);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {                  // 64-channel chunks
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ab * 256 + wg * 128 + c * 64;
        ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        ptx::tmem_ld_wait();
        if (c == 1) {                                // this warp has drained its part of the accumulator set
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_empty[ab]);
        }
        uint4 O[8];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float4 S[8], Hs[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            S[g] = lds_f4(sc_a + (uint32_t)(c * 64 + hh * 32 + g * 4) * 4);
            Hs[g] = lds_f4(sh_a + (uint32_t)(c * 64 + hh * 32 + g * 4) * 4);
          }
#pragma unroll
          for (int g = 0; g < 8; g += 2) {           // two float4 groups = 8 channels = one 16-byte store
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int e = hh * 32 + (g + q) * 4;
              const float2 v0 = ptx::ffma2(make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), make_float2(S[g + q].x, S[g + q].y),
                                           make_float2(Hs[g + q].x, Hs[g + q].y));
              const float2 v1 = ptx::ffma2(make_float2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), make_float2(S[g + q].z, S[g + q].w),
                                           make_float2(Hs[g + q].z, Hs[g + q].w));
              o[2 * q] = p.relu ? ptx::cvt_bf16x2_relu(v0.x, v0.y) : ptx::cvt_bf16x2(v0.x, v0.y);
              o[2 * q + 1] = p.relu ? ptx::cvt_bf16x2_relu(v1.x, v1.y) : ptx::cvt_bf16x2(v1.x, v1.y);
            }
            O[hh * 4 + g / 2] = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        // the previous TMA store out of the staging chunk must have read it before it is overwritten
        if (et == 0) ptx::tma_store_wait_read<0>();
        ptx::named_bar_sync(1, kEpiWarps * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) sts128(srow + (((uint32_t)j ^ swz) << 4), O[j]);
        ptx::fence_proxy_async();
        ptx::named_bar_sync(2, kEpiWarps * 32);
        if (et == 0) {
          ptx::tma_store_4d(&tmO, smem + kStgOff, c * 64, 0, h0, n);
          ptx::tma_store_commit();
        }
      }
    }
    if (et == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

struct SlabLaunchImpl {
  CUtensorMap tmX, tmW, tmO;
  SlabParams p;
};
static_assert(sizeof(SlabLaunchImpl) <= sizeof(SlabLaunch::storage), "SlabLaunch::storage too small");

bool conv3x3_slab_supported(int H, int W, int Cin, int Cout, int ksize, int stride, int pad) {
  static const bool on = getenv("AIRPOSE_NO_SLAB_CONV") == nullptr;
  // windows of M-tile 1 reach row 128 + 2 (W + 2) + 2 + 127 of the slab
  return on && Cin == kC && Cout == kC && ksize == 3 && stride == 1 && pad == 1 && H >= 1 && W >= 1 && 2 * (W + 2) + 2 + 256 <= kSlabRows;
}

int build_conv3x3_slab(SlabLaunch* L, const void* x, const void* w, const float* scale, const float* shift, int relu, void* out, int n, int H,
                       int W) {
  AP_REQUIRE(conv3x3_slab_supported(H, W, kC, kC, 3, 1, 1), "build_conv3x3_slab: unsupported geometry H=%d W=%d", H, W);
  SlabLaunchImpl& I = *reinterpret_cast<SlabLaunchImpl*>(L->storage);
  SlabParams& p = I.p;
  p.H = H; p.W = W; p.Wp = W + 2;
  const int rmax = std::min(256 / p.Wp, kSlabRows / p.Wp - 2);      // rows of two M-tiles; the halo slab must fit too
  AP_REQUIRE(rmax >= 1, "build_conv3x3_slab: a %d-wide row does not fit", W);
  p.units_per_img = ceil_div(H, rmax);
  p.R = ceil_div(H, p.units_per_img);                               // equal bands (28 rows -> 4 x 7, not 3 x 8 + 4)
  p.num_units = n * p.units_per_img;
  p.relu = relu;
  p.scale = scale; p.shift = shift;
  if (make_tmap_nhwc4d_bf16(&I.tmX, x, n, H, W, kC, p.Wp, p.R + 2)) return 1;
  if (make_tmap_tiled_bf16(&I.tmW, w, kC, 9 * kC, 9 * kC, kC, 64)) return 1;
  if (make_tmap_nhwc4d_bf16(&I.tmO, out, n, H, W, kC, p.Wp, p.R)) return 1;
  L->valid = 1;
  L->pdl = use_pdl();
  return 0;
}

int launch_conv3x3_slab(const SlabLaunch& L, cudaStream_t stream) {
  AP_REQUIRE(L.valid, "launch_conv3x3_slab: launch was not built");
  const SlabLaunchImpl& I = *reinterpret_cast<const SlabLaunchImpl*>(L.storage);
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  if (I.p.num_units == 0) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(I.p.num_units, grid_limit()));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = L.pdl ? 1 : 0;
  AP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_slab_kernel, I.tmX, I.tmW, I.tmO, I.p));
  count_launch();
  return 0;
}

}  // namespace airpose
