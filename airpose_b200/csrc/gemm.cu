// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T.
//
// This one kernel is every dense contraction of the AirPose trunk and regressor
// (copenet/src/copenet/models/model_copenet.py:161-204): 1x1 convs read NHWC activations
// as a plain [M,K] matrix through a tiled TMA map; 3x3 / strided convs read the same NHWC
// tensor through a TMA *im2col* map (implicit GEMM, nothing is materialised); the IEF
// linears run as split-bf16 GEMMs.  BatchNorm (folded scale/shift), residual add, ReLU and
// the bf16 rounding happen in the epilogue, straight out of tensor memory.
//
// CTA = 6 warps, one CTA per SM, static round-robin over 128 x BN output tiles:
//   warp 0    TMA producer   (one lane): A/B k-blocks of 64 bf16 -> 128B-swizzled smem ring
//   warp 1    MMA issuer     (one lane): tcgen05.mma 128 x BN x 16, accumulators in TMEM,
//                                         double-buffered so tile i+1 overlaps epilogue i
//   warps 2-5 epilogue       (128 thr):  tcgen05.ld -> scale/shift/residual/relu -> global
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace airpose {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;              // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kABytes = kBlockM * kBlockK * 2;

struct KParams {
  int M, N, K;
  int num_kb, tiles_m, tiles_n;
  int im2col, cblks, ksize, stride, pad, Wo, HoWo;
  Epilogue epi;
};

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;                       // two accumulator stages
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kBarBytes + 2 * 2 * BN * 4;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sc_s = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + C::kBarBytes);   // [2][BN]
  float* sh_s = sc_s + 2 * BN;                                                                 // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], kEpiThreads / 32); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, C::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
        const int m0 = m_blk * kBlockM, n0 = n_blk * BN;
        int cw = 0, ch = 0, cn = 0;
        if (p.im2col) {                       // window origin of the tile's first output pixel
          cn = m0 / p.HoWo;
          const int rem = m0 - cn * p.HoWo;
          const int po = rem / p.Wo, qo = rem - po * p.Wo;
          cw = qo * p.stride - p.pad;
          ch = po * p.stride - p.pad;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + kABytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          if (p.im2col) {
            const int tap = kb / p.cblks, cb = kb - tap * p.cblks;
            const int r = tap / p.ksize, s = tap - r * p.ksize;
            ptx::tma_load_im2col_4d(&tmA, &full_bar[stage], sa, cb * kBlockK, cw, ch, cn, (uint16_t)s, (uint16_t)r);
          } else {
            ptx::tma_load_2d(&tmA, &full_bar[stage], sa, kb * kBlockK, m0);
          }
          ptx::tma_load_2d(&tmB, &full_bar[stage], sb, kb * kBlockK, n0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1, 200 + as);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase, 300 + stage);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
          const uint64_t adesc = ptx::make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = ptx::make_kmajor_sw128_desc(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in 16-byte units
            ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);      // frees the smem slot when these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[as]);           // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int et = threadIdx.x - 64;                // 0..127
    const int quad = warp & 3;                      // TMEM lane quarter this warp may read
    const Epilogue& e = p.epi;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      const int n0 = n_blk * BN;
      for (int i = et; i < BN; i += kEpiThreads) {
        const int n = n0 + i;
        sc_s[as * BN + i] = (n < p.N) ? (e.scale ? __ldg(e.scale + n) : 1.f) : 0.f;
        sh_s[as * BN + i] = (n < p.N && e.shift) ? __ldg(e.shift + n) : 0.f;
      }
      ptx::named_bar_sync(1, kEpiThreads);
      ptx::mbar_wait(&tfull_bar[as], aphase, 400 + as);
      ptx::tc_fence_after();
      const int64_t row = (int64_t)m_blk * kBlockM + quad * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + c * 32, r);
        ptx::tmem_ld_wait();
        const int col0 = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
          v[i] = fmaf(__uint_as_float(r[i]), sc_s[as * BN + c * 32 + i], sh_s[as * BN + c * 32 + i]);
        if (row_ok) {
          if (e.residual) {
            if (e.residual_f32) {
              const float* rp = reinterpret_cast<const float*>(e.residual) + row * e.ldr + col0;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                if (col0 + g * 4 < p.N) {
                  const float4 q = __ldg(reinterpret_cast<const float4*>(rp) + g);
                  v[g * 4] += q.x; v[g * 4 + 1] += q.y; v[g * 4 + 2] += q.z; v[g * 4 + 3] += q.w;
                }
            } else {
              const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(e.residual) + row * e.ldr + col0;
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (col0 + g * 8 < p.N) {
                  const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp) + g);
                  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                  for (int h = 0; h < 4; ++h) {
                    v[g * 8 + 2 * h] += __uint_as_float(w[h] << 16);
                    v[g * 8 + 2 * h + 1] += __uint_as_float(w[h] & 0xFFFF0000u);
                  }
                }
            }
          }
          if (e.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (e.out_bf16) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(e.out_bf16) + row * e.ldd + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (col0 + g * 8 < p.N) {
                uint4 q;
                q.x = pack_bf16(v[g * 8], v[g * 8 + 1]); q.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                q.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]); q.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                reinterpret_cast<uint4*>(op)[g] = q;
              }
          }
          if (e.out_f32) {
            float* op = e.out_f32 + row * e.ldf + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (col0 + g * 4 < p.N)
                reinterpret_cast<float4*>(op)[g] = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          }
          if (e.out_split) {     // x = hi + lo (+ O(2^-17 x)); next GEMM consumes [hi | lo | hi]
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(e.out_split) + row * e.lds + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (col0 + g * 8 < p.N) {
                float lo[8];
                uint32_t hw[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  const __nv_bfloat16 a = __float2bfloat16_rn(v[g * 8 + 2 * h]), b = __float2bfloat16_rn(v[g * 8 + 2 * h + 1]);
                  lo[2 * h] = v[g * 8 + 2 * h] - __bfloat162float(a);
                  lo[2 * h + 1] = v[g * 8 + 2 * h + 1] - __bfloat162float(b);
                  hw[h] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
                }
                const uint4 qh = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                const uint4 ql = make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]),
                                            pack_bf16(lo[6], lo[7]));
                reinterpret_cast<uint4*>(op)[g] = qh;
                reinterpret_cast<uint4*>(op + p.N)[g] = ql;
                reinterpret_cast<uint4*>(op + 2 * (int64_t)p.N)[g] = qh;
              }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ host

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void* driver_fn(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return fn;
}

int make_tmap_tiled_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld_elems,
                         int box_rows, int box_cols, int swizzle_bytes) {
  static EncodeTiledFn fn = (EncodeTiledFn)driver_fn("cuTensorMapEncodeTiled");
  AP_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  AP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems % 8) == 0,
             "TMA operand must be 16-byte aligned with a leading dimension that is a multiple of 8 (ld=%lld)",
             (long long)ld_elems);
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  AP_REQUIRE((swizzle_bytes == 128 || swizzle_bytes == 64) && box_cols * 2 == swizzle_bytes,
             "make_tmap_tiled_bf16: box of %d columns does not match a %d-byte swizzle", box_cols, swizzle_bytes);
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld box=%dx%d)", (int)r,
             (long long)rows, (long long)cols, (long long)ld_elems, box_rows, box_cols);
  return 0;
}

int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, const ConvGeom& g, int channels_per_pixel,
                          int pixels_per_column) {
  static EncodeIm2colFn fn = (EncodeIm2colFn)driver_fn("cuTensorMapEncodeIm2col");
  AP_REQUIRE(fn, "cuTensorMapEncodeIm2col is not available from the driver");
  AP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (g.Cin % 8) == 0, "im2col operand misaligned");
  const cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.n};
  const cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.W * g.Cin * 2, (cuuint64_t)g.H * g.W * g.Cin * 2};
  // fprop corners: the window origin ranges over [-pad, dim - 1 + pad - (k-1)]
  const int lower[2] = {-g.pad, -g.pad};
  const int upper[2] = {g.pad - (g.ksize - 1), g.pad - (g.ksize - 1)};
  const cuuint32_t estr[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper,
                        (cuuint32_t)channels_per_pixel, (cuuint32_t)pixels_per_column, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed with %d (n=%d H=%d W=%d C=%d k=%d s=%d p=%d)", (int)r, g.n,
             g.H, g.W, g.Cin, g.ksize, g.stride, g.pad);
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB; CUTLASS
  // (cute/atom/copy_traits_sm90_im2col.hpp) clears bit 21 of the second descriptor word for them.
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 &&
      (int64_t)g.n * g.H * g.W * g.Cin * 2 < 131072 && !getenv("AIRPOSE_NO_IM2COL_WORKAROUND"))
    reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  return 0;
}

int make_tmap_nhwc4d_bf16(CUtensorMap* out, const void* base, int n, int H, int W, int C, int box_w, int box_h) {
  static EncodeTiledFn fn = (EncodeTiledFn)driver_fn("cuTensorMapEncodeTiled");
  AP_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  AP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && C % 64 == 0 && box_w >= 1 && box_w <= 256 && box_h >= 1 && box_h <= 256,
             "make_tmap_nhwc4d_bf16: bad operand (C=%d box %dx%d)", C, box_w, box_h);
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (4-d) failed with %d (n=%d H=%d W=%d C=%d box %dx%d)", (int)r, n, H, W, C,
             box_w, box_h);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static thread_local int tl_grid_cap = 0;
void set_grid_cap(int cap) { tl_grid_cap = cap; }
int grid_limit() { return tl_grid_cap > 0 ? std::min(tl_grid_cap, num_sms()) : num_sms(); }

bool use_tma_epilogue() {
  static const bool on = getenv("AIRPOSE_NO_TMA_EPI") == nullptr;
  return on;
}

bool use_pdl() {
  static const bool on = getenv("AIRPOSE_NO_PDL") == nullptr;
  return on;
}

bool use_stream_k() {
  static const bool on = getenv("AIRPOSE_GEMM_V1") == nullptr;
  return on;
}

// Which kernel runs a [M,N,K] problem with a bf16 TMA epilogue (measured per layer on B200,
// profiles/r01d_layers_*.txt): the stream-K kernel (gemm_sk.cu, 128x256 tiles cut at k-block granularity)
// wins where tiles are long in K and few enough to quantise badly on 148 SMs -- the 3x3 convs of
// layer3/layer4 -- and loses 5-15 % elsewhere to its per-cut-tile fix-up, so everything else keeps the
// round-robin kernel (gemm_tma.cu).  AIRPOSE_GEMM_V1=1 / AIRPOSE_GEMM_SK=1 force one or the other.
bool prefers_stream_k(int M, int N, int K) {
  static const bool force_sk = getenv("AIRPOSE_GEMM_SK") != nullptr;
  if (!use_stream_k()) return false;
  if (force_sk) return true;
  static const int min_kb = getenv("AIRPOSE_SK_MINKB") ? atoi(getenv("AIRPOSE_SK_MINKB")) : 36;
  const int bn = N % 256 == 0 ? 256 : (N > 64 ? 128 : 64);
  const int tiles = ceil_div(M, kBlockM) * ceil_div(N, bn);
  return ceil_div(K, 64) >= min_kb && tiles < 8 * num_sms() && tiles % num_sms() != 0;
}

// The cta_group::2 pair kernel (gemm_sk2.cu) needs N % 256 == 0 and pays off where the mainloop dominates: long K and
// enough rows.  Measured (profiles/r01g_layers_pair.txt): 8192^3 1166 -> 1413 TFLOP/s; layer3 3x3 (M=25088, K=2304)
// 40.8 -> 39.5 us; layer4 3x3 (M=6272) 42.0 -> 43.0 us and the K<=2048 1x1 layers 24.6 -> 28.7 us, which therefore stay
// on the cta_group::1 kernels.  AIRPOSE_GEMM_2CTA=0 disables it; AIRPOSE_2CTA_MINKB / AIRPOSE_2CTA_MINM move the thresholds.
bool prefers_pair(int M, int N, int K) {
  static const bool on = getenv("AIRPOSE_GEMM_2CTA") == nullptr || atoi(getenv("AIRPOSE_GEMM_2CTA")) != 0;
  static const int min_kb = getenv("AIRPOSE_2CTA_MINKB") ? atoi(getenv("AIRPOSE_2CTA_MINKB")) : 36;
  static const int min_m = getenv("AIRPOSE_2CTA_MINM") ? atoi(getenv("AIRPOSE_2CTA_MINM")) : 16384;
  return on && use_stream_k() && N % 256 == 0 && K % 64 == 0 && K / 64 >= min_kb && M >= min_m;
}

int b_box_rows(int M, int N, int K, int block_n) { return prefers_pair(M, N, K) ? 128 : block_n; }

int pick_block_n(int M, int N, int K) {
  if (prefers_pair(M, N, K)) return 256;
  if (prefers_stream_k(M, N, K)) {   // stream-K removes the wave-quantisation penalty of wide tiles
    if (N % 256 == 0) return 256;
    if (N > 64) return 128;
    return 64;
  }
  const int tiles_m = ceil_div(M, kBlockM);
  if (N % 256 == 0 && (int64_t)tiles_m * (N / 256) >= 2 * num_sms()) return 256;
  if (N > 64) return 128;
  return 64;
}

template <int BN>
static int launch_bn(const GemmLaunch& L, const KParams& kp, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes));
    configured = true;
  }
  const int tiles = kp.tiles_m * kp.tiles_n;
  const int grid = std::min(tiles, num_sms());
  gemm_bf16_kernel<BN><<<grid, kThreads, Cfg<BN>::kSmemBytes, stream>>>(L.tmA, L.tmB, kp);
  AP_LAUNCH_CHECK();
  return 0;
}

int launch_gemm(const GemmLaunch& L, cudaStream_t stream) {
  if (L.tma_epi && L.pair_b_box) return launch_gemm_sk2(L, stream);
  if (L.tma_epi) return (!L.stem && prefers_stream_k(L.M, L.N, L.K)) ? launch_gemm_sk(L, stream) : launch_gemm_tma(L, stream);
  AP_REQUIRE(L.M > 0 && L.N > 0 && L.K > 0, "launch_gemm: empty problem %dx%dx%d", L.M, L.N, L.K);
  AP_REQUIRE(L.N % 8 == 0, "launch_gemm: N=%d must be a multiple of 8", L.N);
  const Epilogue& e = L.epi;
  AP_REQUIRE(!e.out_bf16 || (e.ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(e.out_bf16) & 15) == 0), "launch_gemm: out_bf16 misaligned");
  AP_REQUIRE(!e.out_f32 || (e.ldf % 4 == 0 && (reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0), "launch_gemm: out_f32 misaligned");
  AP_REQUIRE(!e.out_split || (e.lds % 8 == 0 && (reinterpret_cast<uintptr_t>(e.out_split) & 15) == 0), "launch_gemm: out_split misaligned");
  AP_REQUIRE(!e.residual || ((e.ldr % (e.residual_f32 ? 4 : 8)) == 0 && (reinterpret_cast<uintptr_t>(e.residual) & 15) == 0),
             "launch_gemm: residual misaligned");
  KParams kp{};
  kp.M = L.M; kp.N = L.N; kp.K = L.K;
  kp.num_kb = ceil_div(L.K, kBlockK);
  kp.tiles_m = ceil_div(L.M, kBlockM);
  kp.tiles_n = ceil_div(L.N, L.block_n);
  kp.im2col = L.im2col;
  if (L.im2col) {
    const ConvGeom& g = L.geom;
    AP_REQUIRE(g.Cin % kBlockK == 0, "launch_gemm: im2col needs Cin %% 64 == 0 (Cin=%d)", g.Cin);
    AP_REQUIRE(L.K == g.ksize * g.ksize * g.Cin, "launch_gemm: K=%d does not match the conv geometry", L.K);
    kp.cblks = g.Cin / kBlockK; kp.ksize = g.ksize; kp.stride = g.stride; kp.pad = g.pad;
    kp.Wo = g.Wo; kp.HoWo = g.Ho * g.Wo;
  }
  kp.epi = e;
  switch (L.block_n) {
    case 64: return launch_bn<64>(L, kp, stream);
    case 128: return launch_bn<128>(L, kp, stream);
    case 256: return launch_bn<256>(L, kp, stream);
    default: AP_REQUIRE(false, "launch_gemm: unsupported block_n %d", L.block_n);
  }
  return 0;
}

}  // namespace airpose

using namespace airpose;

extern "C" int airpose_gemm_bf16(const airpose_gemm_args* g, void* stream) {
  AP_REQUIRE(g && g->A && g->B, "airpose_gemm_bf16: null argument");
  GemmLaunch L{};
  L.M = g->M; L.N = g->N; L.K = g->K;
  if (g->a_t || g->b_t) {                 // transposed operand(s): MN-major descriptors in the stream-K kernel
    AP_REQUIRE(g->out_bf16 && !g->out_f32 && !g->residual && g->N % 64 == 0, "airpose_gemm_bf16: a_t / b_t need a bf16 output with N %% 64 == 0 and no residual");
    AP_REQUIRE(!g->a_t || g->M % 8 == 0, "airpose_gemm_bf16: a_t needs M %% 8 == 0 (M=%d)", g->M);
    L.block_n = g->N % 256 == 0 ? 256 : (g->N > 64 ? 128 : 64);
    L.mn_a = g->a_t ? 1 : 0; L.mn_b = g->b_t ? 1 : 0;
    if (L.mn_a ? make_tmap_tiled_bf16(&L.tmA, g->A, g->K, g->M, g->lda, 64, 64) : make_tmap_tiled_bf16(&L.tmA, g->A, g->M, g->K, g->lda, kBlockM, kBlockK)) return 1;
    if (L.mn_b ? make_tmap_tiled_bf16(&L.tmB, g->B, g->K, g->N, g->ldb, 64, 64) : make_tmap_tiled_bf16(&L.tmB, g->B, g->N, g->K, g->ldb, L.block_n, kBlockK)) return 1;
    L.epi.scale = g->scale; L.epi.shift = g->shift; L.epi.relu = g->relu;
    L.epi.out_bf16 = g->out_bf16; L.epi.ldd = g->ldd;
    AP_REQUIRE(tma_epilogue_eligible(L), "airpose_gemm_bf16: a_t / b_t need the TMA epilogue (aligned bf16 output)");
    if (enable_tma_epilogue(&L)) return 1;
    L.pdl = use_pdl();
    return launch_gemm_sk(L, (cudaStream_t)stream);
  }
  L.block_n = pick_block_n(g->M, g->N, g->K);
  if (make_tmap_tiled_bf16(&L.tmA, g->A, g->M, g->K, g->lda, kBlockM, kBlockK)) return 1;
  const bool tma_ok = use_tma_epilogue() && g->out_bf16 && !g->out_f32 && g->N % 64 == 0;
  L.pair_b_box = tma_ok && prefers_pair(g->M, g->N, g->K) && g->ldb == g->K;
  if (make_tmap_tiled_bf16(&L.tmB, g->B, g->N, g->K, g->ldb, L.pair_b_box ? 128 : L.block_n, kBlockK)) return 1;
  L.epi.scale = g->scale; L.epi.shift = g->shift;
  L.epi.residual = g->residual; L.epi.ldr = g->ldr;
  L.epi.relu = g->relu;
  L.epi.out_bf16 = g->out_bf16; L.epi.ldd = g->ldd;
  L.epi.out_f32 = g->out_f32; L.epi.ldf = g->ldf;
  if (use_tma_epilogue() && tma_epilogue_eligible(L) && enable_tma_epilogue(&L)) return 1;
  if (L.pair_b_box && !L.tma_epi) {      // the pair kernel needs the TMA epilogue: fall back to block_n-row B boxes
    L.pair_b_box = 0;
    if (make_tmap_tiled_bf16(&L.tmB, g->B, g->N, g->K, g->ldb, L.block_n, kBlockK)) return 1;
  }
  L.pdl = use_pdl();           // the kernels wait (griddepcontrol.wait) before touching anything the previous kernel wrote
  return launch_gemm(L, (cudaStream_t)stream);
}

extern "C" int airpose_conv_bf16(const airpose_conv_args* c, void* stream) {
  AP_REQUIRE(c && c->x && c->w && c->out, "airpose_conv_bf16: null argument");
  if (!c->residual && conv3x3_slab_supported(c->H, c->W, c->Cin, c->Cout, c->ksize, c->stride, c->pad)) {      // as trunk.cu dispatches it
    SlabLaunch S{};
    if (build_conv3x3_slab(&S, c->x, c->w, c->scale, c->shift, c->relu, c->out, c->n, c->H, c->W)) return 1;
    return launch_conv3x3_slab(S, (cudaStream_t)stream);
  }
  GemmLaunch L{};
  ConvGeom& g = L.geom;
  g.n = c->n; g.H = c->H; g.W = c->W; g.Cin = c->Cin;
  g.ksize = c->ksize; g.stride = c->stride; g.pad = c->pad;
  g.Ho = (c->H + 2 * c->pad - c->ksize) / c->stride + 1;
  g.Wo = (c->W + 2 * c->pad - c->ksize) / c->stride + 1;
  L.M = c->n * g.Ho * g.Wo; L.N = c->Cout; L.K = c->ksize * c->ksize * c->Cin;
  L.block_n = pick_block_n(L.M, L.N, L.K);
  L.im2col = 1;
  if (make_tmap_im2col_bf16(&L.tmA, c->x, g, kBlockK, kBlockM)) return 1;
  L.pair_b_box = use_tma_epilogue() && L.N % 64 == 0 && prefers_pair(L.M, L.N, L.K);
  if (make_tmap_tiled_bf16(&L.tmB, c->w, L.N, L.K, L.K, L.pair_b_box ? 128 : L.block_n, kBlockK)) return 1;
  L.epi.scale = c->scale; L.epi.shift = c->shift;
  L.epi.residual = c->residual; L.epi.ldr = c->Cout;
  L.epi.relu = c->relu;
  L.epi.out_bf16 = c->out; L.epi.ldd = c->Cout;
  if (use_tma_epilogue() && tma_epilogue_eligible(L) && enable_tma_epilogue(&L)) return 1;
  if (L.pair_b_box && !L.tma_epi) {
    L.pair_b_box = 0;
    if (make_tmap_tiled_bf16(&L.tmB, c->w, L.N, L.K, L.K, L.block_n, kBlockK)) return 1;
  }
  L.pdl = use_pdl();           // the kernels wait (griddepcontrol.wait) before touching anything the previous kernel wrote
  return launch_gemm(L, (cudaStream_t)stream);
}
