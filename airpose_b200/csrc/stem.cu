// Fused ResNet stem: conv 7x7 stride 2 pad 3 (3 -> 64) + BN + ReLU + MaxPool 3x3 stride 2 pad 1
// (copenet/src/copenet/models/model_copenet.py:57-60,163-166) in ONE persistent tcgen05 kernel; the 112x112x64 conv output
// (103 MB per 64 images, written and read back by the unfused path) never exists in memory.
//
// Operand layout ("column pairs"): the packed input X2p[n][parity][115][56][32] holds, per input row h = 2 (hp - 2) + parity
// and per PAIR u of output columns (2u, 2u+1), the 9 input columns 4u-3 .. 4u+5 x 3 channels (27 values, padded to 32
// = one 64-byte swizzle row).  One GEMM row (p, u) computes BOTH output pixels of the pair: N = 128 = 64 channels of pixel 2u
// | 64 channels of pixel 2u+1, with the 7x7 weights laid out twice in a [128][7][32] matrix (shifted by two columns for the
// odd pixel).  Half the operand bytes of the one-pixel-per-row layout, N = 128 MMAs instead of N = 64, and the two pixels a
// thread needs for the horizontal max are in its own registers.
//
// A CTA owns kRP = 2 pooled rows of one image = 5 conv rows = 280 GEMM rows (three 128-row M-tiles).  Input rows are split
// by parity so that vertical tap r reads plane rows p + c_r + 2: with the 8 (parity 1) / 7 (parity 0) plane rows of the band
// in shared memory ONCE (15 TMA boxes of 56 x 64 B, double buffered), tap r of M-tile t is the band shifted by
// (c_r - c_min) * 56 + 128 t rows -- whole 512-byte swizzle atoms, so every shift is a valid SWIZZLE_64B operand.
// Epilogue (two warpgroups, 32 channels of both pixels each): TMEM -> BN + ReLU -> bf16x2 -> horizontal 3-max (the left
// neighbour's odd pixel comes from the previous lane by shuffle; lane 0 takes it from a small shared-memory edge buffer)
// -> h-pooled band in shared memory -> vertical 3-max -> 16-byte coalesced global stores of the pooled NHWC tensor.
// Pool padding: all values are >= 0 after the ReLU, so an out-of-image neighbour can be ignored (equivalently taken as 0).
//
// CTA = 12 warps: 0 band TMA producer, 1 MMA issuer, 3 TMEM allocator, 4-11 epilogue.
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
constexpr int kRP = 2;                         // pooled rows per tile
constexpr int kNR = 2 * kRP + 1;               // conv rows per tile
constexpr int kUW = 56;                        // column pairs per conv row
constexpr int kMRows = kNR * kUW;              // 280 GEMM rows
constexpr int kMT = (kMRows + 127) / 128;      // 3 M-tiles
constexpr int kPlaneRows = 115;                // plane rows per (image, parity): 2 zero rows + 112 + 1 zero row
constexpr int kRowB = 64;                      // bytes per packed row (32 bf16)
constexpr int kPRB = kUW * kRowB;              // 3584 bytes per plane row
constexpr int kRows1 = kNR + 3;                // plane rows of the parity-1 band (taps r = 0, 2, 4, 6: c_r = -2 .. 1)
constexpr int kRows0 = kNR + 2;                // parity-0 band (taps r = 1, 3, 5: c_r = -1 .. 1)
constexpr int kBand1 = kRows1 * kPRB;          // 28672
constexpr int kBand0 = kRows0 * kPRB;          // 25088
constexpr int kStage = kBand1 + kBand0;        // 53760 (a multiple of 512: every band starts on a swizzle atom)
constexpr int kBandPad = 8192;                 // the last M-tile's junk rows read up to 104 rows past the last band
constexpr int kWBytes = 7 * 128 * kRowB;       // 57344
constexpr int kHbBytes = kMRows * 128;         // 35840: h-pooled band, 64 channels bf16 per (conv row, pooled column)
constexpr int kEdgeBytes = 2 * 3 * 4 * 64;     // [warpgroup][M-tile][warp] x 32 channels
constexpr int kBandOff = 0;
constexpr int kWOff = kBandOff + 2 * kStage + kBandPad;
constexpr int kHbOff = kWOff + kWBytes;
constexpr int kEdgeOff = kHbOff + kHbBytes;
constexpr int kBarOff = kEdgeOff + kEdgeBytes;
// barriers: w_full, band_full[2], band_empty[2], acc_full[3], acc_empty[3]
constexpr int kNumBars = 11;
constexpr int kScaleOff = (kBarOff + kNumBars * 8 + 16 + 15) & ~15;        // read with 16-byte shared loads
constexpr int kSmemBytes = 1024 + kScaleOff + 2 * 64 * 4;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
static_assert(kStage % 1024 == 0 || kStage % 512 == 0, "bands must start on a swizzle atom");
constexpr int kTmemCols = 512;                 // three 128-column accumulators

struct StemParams {
  int n_img, num_tiles;
  const float* scale; const float* shift;
  __nv_bfloat16* out;                          // [n, 56, 56, 64]
};

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a), y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 m = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&m);
}
__device__ __forceinline__ uint4 hmax2_u4(const uint4& a, const uint4& b) {
  return make_uint4(hmax2_u32(a.x, b.x), hmax2_u32(a.y, b.y), hmax2_u32(a.z, b.z), hmax2_u32(a.w, b.w));
}

__global__ void __launch_bounds__(kThreads, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* w_full = bars + 0;
  uint64_t* band_full = bars + 1;       // [2]
  uint64_t* band_empty = bars + 3;      // [2]
  uint64_t* acc_full = bars + 5;        // [3]
  uint64_t* acc_empty = bars + 8;       // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  float* sc_s = reinterpret_cast<float*>(smem + kScaleOff);
  float* sh_s = sc_s + 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW);
    ptx::mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&band_full[i], 1); ptx::mbar_init(&band_empty[i], 1); }
    for (int i = 0; i < kMT; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], kEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 3) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x < 64) { sc_s[threadIdx.x] = __ldg(p.scale + threadIdx.x); sh_s[threadIdx.x] = __ldg(p.shift + threadIdx.x); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {          // the weights do not depend on the previous kernel
    ptx::mbar_arrive_expect_tx(w_full, kWBytes);
    for (int r = 0; r < 7; ++r) ptx::tma_load_2d(&tmW, w_full, smem + kWOff + r * (128 * kRowB), r * 32, 0);
  }
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ band producer
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int n = tile / 28, p0 = 4 * (tile - n * 28) - 1;       // first conv row of the tile (-1 for the top tile)
        const int s = it & 1;
        ptx::mbar_wait(&band_empty[s], (uint32_t)(((it >> 1) & 1) ^ 1), 100 + s);
        ptx::mbar_arrive_expect_tx(&band_full[s], kStage);
        uint8_t* b1 = smem + kBandOff + s * kStage;
        uint8_t* b0 = b1 + kBand1;
        // parity 1: plane rows p0 .. p0 + 7 (p0 = -1 reads the zero row that closes the previous plane, or out of bounds: zeros)
        const int base1 = ((n * 2 + 1) * kPlaneRows + p0) * kUW, base0 = ((n * 2 + 0) * kPlaneRows + p0 + 1) * kUW;
        for (int j = 0; j < kRows1; ++j) ptx::tma_load_2d(&tmX, &band_full[s], b1 + j * kPRB, 0, base1 + j * kUW);
        for (int j = 0; j < kRows0; ++j) ptx::tma_load_2d(&tmX, &band_full[s], b0 + j * kPRB, 0, base0 + j * kUW);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 128);
      const uint32_t band_a = ptx::smem_u32(smem + kBandOff), w_a = ptx::smem_u32(smem + kWOff);
      ptx::mbar_wait(w_full, 0, 200);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&band_full[s], (uint32_t)((it >> 1) & 1), 201 + s);
        ptx::tc_fence_after();
        for (int t = 0; t < kMT; ++t) {
          ptx::mbar_wait(&acc_empty[t], (uint32_t)((it & 1) ^ 1), 210 + t);
          ptx::tc_fence_after();
#pragma unroll
          for (int r = 0; r < 7; ++r) {
            const int par = (r + 1) & 1;                       // input row 2p - 3 + r has parity (r + 1) & 1
            const int cs = (r - 3 - par) / 2 + (par ? 2 : 1);  // c_r - c_min of that parity's band
            const uint32_t a0 = band_a + (uint32_t)(s * kStage + (par ? 0 : kBand1) + (cs * kUW + 128 * t) * kRowB);
            const uint32_t b0 = w_a + (uint32_t)(r * (128 * kRowB));
#pragma unroll
            for (int k = 0; k < 2; ++k)
              ptx::umma_bf16(tmem_base + t * 128, ptx::make_kmajor_desc(a0 + k * 32, 64), ptx::make_kmajor_desc(b0 + k * 32, 64), idesc,
                             (r | k) != 0);
          }
          ptx::umma_commit(&acc_full[t]);
        }
        ptx::umma_commit(&band_empty[s]);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int e = warp - kEpiWarp0;
    const int wg = e >> 2;                           // channels [32 wg, 32 wg + 32) of both pixels of the pair
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int et = threadIdx.x - kEpiWarp0 * 32;     // 0..255
    const uint32_t sc_a = ptx::smem_u32(sc_s) + wg * 128, sh_a = ptx::smem_u32(sh_s) + wg * 128;
    const uint32_t hb_a = ptx::smem_u32(smem + kHbOff);
    const uint32_t edge_a = ptx::smem_u32(smem + kEdgeOff) + wg * (3 * 4 * 64);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int n = tile / 28, i0 = 2 * (tile - n * 28);
#pragma unroll 1
      for (int t = 0; t < kMT; ++t) {
        const int m = t * 128 + row;                 // GEMM row = conv row (m / 56) of the tile, column pair m % 56
        const int u = m % kUW;
        ptx::mbar_wait(&acc_full[t], (uint32_t)(it & 1), 400 + t);
        ptx::tc_fence_after();
        uint32_t re[32], ro[32];
        ptx::tmem_ld_32x32(lane_addr + t * 128 + wg * 32, re);
        ptx::tmem_ld_32x32(lane_addr + t * 128 + 64 + wg * 32, ro);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[t]);
        // BN + ReLU -> bf16x2: E = pixel 2u, O = pixel 2u + 1 (16 words = 32 channels each)
        uint32_t E[16], O[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 s4 = ptx::lds_f4(sc_a + g * 16), h4 = ptx::lds_f4(sh_a + g * 16);
          const float2 s01 = make_float2(s4.x, s4.y), s23 = make_float2(s4.z, s4.w), h01 = make_float2(h4.x, h4.y), h23 = make_float2(h4.z, h4.w);
          float2 v;
          v = ptx::ffma2(make_float2(__uint_as_float(re[4 * g]), __uint_as_float(re[4 * g + 1])), s01, h01); E[2 * g] = ptx::cvt_bf16x2_relu(v.x, v.y);
          v = ptx::ffma2(make_float2(__uint_as_float(re[4 * g + 2]), __uint_as_float(re[4 * g + 3])), s23, h23); E[2 * g + 1] = ptx::cvt_bf16x2_relu(v.x, v.y);
          v = ptx::ffma2(make_float2(__uint_as_float(ro[4 * g]), __uint_as_float(ro[4 * g + 1])), s01, h01); O[2 * g] = ptx::cvt_bf16x2_relu(v.x, v.y);
          v = ptx::ffma2(make_float2(__uint_as_float(ro[4 * g + 2]), __uint_as_float(ro[4 * g + 3])), s23, h23); O[2 * g + 1] = ptx::cvt_bf16x2_relu(v.x, v.y);
        }
        // left neighbour's odd pixel (conv column 2u - 1): previous lane; lane 0 reads the previous warp's lane 31 from the edge
        // buffer (the previous M-tile's last warp for warp 0 -- written before that M-tile's barrier)
        const uint32_t my_edge = edge_a + (uint32_t)((t * 4 + quad) * 64);
        uint4 pe[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (lane == 0 && quad == 0 && t > 0) {       // written before the previous M-tile's barrier
#pragma unroll
          for (int j = 0; j < 4; ++j) pe[j] = ptx::lds128(edge_a + (uint32_t)(((t - 1) * 4 + 3) * 64) + j * 16);
        }
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < 4; ++j) ptx::sts128(my_edge + j * 16, make_uint4(O[4 * j], O[4 * j + 1], O[4 * j + 2], O[4 * j + 3]));
        }
        ptx::named_bar_sync(1 + wg, 128);
        if (lane == 0 && quad > 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) pe[j] = ptx::lds128(my_edge - 64 + j * 16);
        }
        uint32_t Hm[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint32_t left = __shfl_up_sync(0xffffffffu, O[i], 1);
          if (lane == 0) left = (i & 3) == 0 ? pe[i >> 2].x : ((i & 3) == 1 ? pe[i >> 2].y : ((i & 3) == 2 ? pe[i >> 2].z : pe[i >> 2].w));
          if (u == 0) left = 0u;                     // conv column -1: pool padding
          Hm[i] = hmax2_u32(hmax2_u32(E[i], O[i]), left);
        }
        if (m < kMRows) {
          const uint32_t hrow = hb_a + (uint32_t)m * 128u;
          const uint32_t swz = (uint32_t)(m & 7);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ptx::sts128(hrow + (((uint32_t)(wg * 4 + j) ^ swz) << 4), make_uint4(Hm[4 * j], Hm[4 * j + 1], Hm[4 * j + 2], Hm[4 * j + 3]));
        }
      }
      // ---- vertical 3-max over the h-pooled band and store: task = (pooled row, pooled column, 8 channels)
      ptx::named_bar_sync(3, kEpiWarps * 32);
      for (int idx = et; idx < kRP * kUW * 8; idx += kEpiWarps * 32) {
        const int ch = idx & 7, j = (idx >> 3) % kUW, il = idx / (8 * kUW);
        const int r0 = (2 * il) * kUW + j;           // conv row 2 i - 1 of pooled row i = i0 + il  (band row 2 il)
        const uint4 b = ptx::lds128(hb_a + (uint32_t)(r0 + kUW) * 128u + (((uint32_t)ch ^ (uint32_t)((r0 + kUW) & 7)) << 4));
        const uint4 c = ptx::lds128(hb_a + (uint32_t)(r0 + 2 * kUW) * 128u + (((uint32_t)ch ^ (uint32_t)((r0 + 2 * kUW) & 7)) << 4));
        uint4 o = hmax2_u4(b, c);
        if (i0 + il > 0) {                           // conv row -1 does not exist (pool padding)
          const uint4 a = ptx::lds128(hb_a + (uint32_t)r0 * 128u + (((uint32_t)ch ^ (uint32_t)(r0 & 7)) << 4));
          o = hmax2_u4(o, a);
        }
        *reinterpret_cast<uint4*>(p.out + ((((size_t)n * 56 + i0 + il) * 56 + j) * 64 + ch * 8)) = o;
      }
      ptx::named_bar_sync(4, kEpiWarps * 32);        // the band is rewritten by the next tile
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// x fp32 NCHW [n,3,224,224] -> X2p bf16 [n][2 parities][115][56][32]:
//   X2p[n][par][hp][u][s*3 + c] = x[n][c][2 (hp - 2) + par][4u - 3 + s],  s = 0..8  (zero outside the image, for hp in {0, 1, 114}
//   and in the five padding slots).  One thread packs one (plane row, column pair) = 64 bytes.
__global__ void stem_pack_pairs_kernel(const float* __restrict__ x, int n, __nv_bfloat16* __restrict__ out) {
  const int64_t total = (int64_t)n * 2 * kPlaneRows * kUW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % kUW);
    const int hp = (int)((i / kUW) % kPlaneRows);
    const int par = (int)((i / (kUW * kPlaneRows)) % 2);
    const int img = (int)(i / (kUW * kPlaneRows * 2));
    const int h = 2 * (hp - 2) + par;
    __align__(16) __nv_bfloat16 vals[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) vals[e] = __float2bfloat16_rn(0.f);
    if (hp >= 2 && hp < 114) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* rowp = x + (((int64_t)img * 3 + c) * 224 + h) * 224;
#pragma unroll
        for (int s = 0; s < 9; ++s) {
          const int w = 4 * u - 3 + s;
          if (w >= 0 && w < 224) vals[s * 3 + c] = __float2bfloat16_rn(__ldg(rowp + w));
        }
      }
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = reinterpret_cast<const uint4*>(vals)[j];
  }
}

// conv1.weight fp32 [64][3][7][7] -> Wp bf16 [128][7][32]: row o < 64 (pixel 2u): Wp[o][r][s*3+c] = w[o][c][r][s], s = 0..6;
// row 64 + o (pixel 2u + 1, two input columns further right): Wp[64+o][r][s*3+c] = w[o][c][r][s-2], s = 2..8; zero elsewhere.
__global__ void stem_pack_pairs_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 7 * 32) return;
  const int e = i % 32, r = (i / 32) % 7, row = i / (32 * 7);
  const int o = row & 63, odd = row >> 6;
  float v = 0.f;
  if (e < 27) {
    const int s = e / 3 - 2 * odd, c = e % 3;
    if (s >= 0 && s < 7) v = w[((o * 3 + c) * 7 + r) * 7 + s];
  }
  out[i] = __float2bfloat16_rn(v);
}

}  // namespace

struct StemLaunchImpl {
  CUtensorMap tmX, tmW;
  StemParams p;
};
static_assert(sizeof(StemLaunchImpl) <= sizeof(StemLaunch::storage), "StemLaunch::storage too small");

size_t stem_pairs_operand_elems(int n) { return (size_t)n * 2 * kPlaneRows * kUW * 32; }
size_t stem_pairs_weight_elems() { return (size_t)128 * 7 * 32; }

int stem_pack_pairs_weight(const float* w_f32, void* out_bf16, cudaStream_t st) {
  stem_pack_pairs_weight_kernel<<<ceil_div(128 * 7 * 32, 256), 256, 0, st>>>(w_f32, (__nv_bfloat16*)out_bf16);
  AP_LAUNCH_CHECK();
  return 0;
}

int build_stem_pool(StemLaunch* L, const void* x2p, const void* wp, const float* scale, const float* shift, void* out, int n) {
  StemLaunchImpl& I = *reinterpret_cast<StemLaunchImpl*>(L->storage);
  I.p.n_img = n; I.p.num_tiles = n * 28;
  I.p.scale = scale; I.p.shift = shift; I.p.out = (__nv_bfloat16*)out;
  if (make_tmap_tiled_bf16(&I.tmX, x2p, (int64_t)n * 2 * kPlaneRows * kUW, 32, 32, kUW, 32, 64)) return 1;
  if (make_tmap_tiled_bf16(&I.tmW, wp, 128, 7 * 32, 7 * 32, 128, 32, 64)) return 1;
  L->valid = 1;
  L->pdl = use_pdl();
  return 0;
}

int launch_stem_pool(const StemLaunch& L, const float* x_nchw, void* x2p, int n, cudaStream_t stream) {
  AP_REQUIRE(L.valid, "launch_stem_pool: launch was not built");
  const StemLaunchImpl& I = *reinterpret_cast<const StemLaunchImpl*>(L.storage);
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  if (n == 0) return 0;
  const int64_t work = (int64_t)n * 2 * kPlaneRows * kUW;
  stem_pack_pairs_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(work, 128), 148 * 32), 128, 0, stream>>>(x_nchw, n, (__nv_bfloat16*)x2p);
  AP_LAUNCH_CHECK();
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
  AP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, stem_pool_kernel, I.tmX, I.tmW, I.p));
  count_launch();
  return 0;
}

}  // namespace airpose
