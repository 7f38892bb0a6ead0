// tcgen05 GEMM / implicit conv with a TMA epilogue -- the kernel every conv of the ResNet-50
// trunk runs on (copenet/src/copenet/models/model_copenet.py:27-47,161-176).
//
// Same mainloop as gemm.cu (TMA -> 128B-swizzled smem ring -> tcgen05.mma into a
// double-buffered TMEM accumulator); what changes is everything that touches HBM in the
// epilogue, because most of the trunk's layers are bound by bytes, not flops:
//   - the bf16 residual tile arrives through TMA into swizzled smem (its own 2-deep ring),
//   - the output tile is packed to bf16 in 128x64 chunks in swizzled smem and leaves through
//     TMA stores (full 128-byte lines, M tail clipped by the tensor map),
//   - the kernel is launched with programmatic stream serialization: barrier init, TMEM
//     allocation and descriptor prefetch overlap the previous layer's tail.
//
// CTA = 7 warps, one CTA per SM, static round-robin over 128 x BN tiles:
//   warp 0    A/B TMA producer     warp 1   MMA issuer     warp 2   residual TMA producer
//   warps 3-6 epilogue: tcgen05.ld -> scale/shift (+residual) -> relu -> bf16 -> smem -> TMA store
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace airpose {

namespace {

constexpr int kBlockM = 128;
constexpr int kUmmaK = 16;
constexpr int kThreads = 224;
constexpr int kEpiWarp0 = 3;
constexpr int kEpiThreads = 128;
constexpr int kChunkN = 64;                       // epilogue chunk: 128 rows x 64 bf16 = one 128B-swizzle box
constexpr int kChunkBytes = kBlockM * kChunkN * 2;
constexpr int kResBufs = 2;

struct KP {
  int M, N, K;
  int num_kb, tiles_m, tiles_n;
  int im2col, cblks, ksize, stride, pad, Wo, HoWo;
  int stem, stem_img_rows, stem_img_stride;
  int stem_tap_off[8];
  int has_res, relu;
  const float* scale;
  const float* shift;
};

template <int BN, int BK>
struct Cfg {
  static constexpr int kStages = (BK == 32) ? 8 : ((BN == 256) ? 3 : (BN == 128 ? 4 : 6));
  static constexpr int kABytes = kBlockM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  // 3 output buffers need one barrier per chunk; BN=256 has room for 2 only and pays a second barrier
  static constexpr int kOutBufs = (BN == 256) ? 2 : 3;
  static constexpr int kRing = kStages * kStageBytes;
  static constexpr int kOutOff = kRing;
  static constexpr int kResOff = kOutOff + kOutBufs * kChunkBytes;
  static constexpr int kBarOff = kResOff + kResBufs * kChunkBytes;
  static constexpr int kNumBars = 2 * kStages + 4 + 2 * kResBufs;
  static constexpr int kScaleOff = (kBarOff + kNumBars * 8 + 16 + 15) & ~15;   // read with 16-byte shared loads
  static constexpr int kSmemBytes = 1024 + kScaleOff + 2 * 2 * BN * 4;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

template <int BN, int BK>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR, const KP p) {
  using C = Cfg<BN, BK>;
  constexpr int kBlockK = BK;
  constexpr int kABytes = C::kABytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;
  uint64_t* rempty_bar = rfull_bar + kResBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty_bar + kResBufs);
  float* sc_s = reinterpret_cast<float*>(smem + C::kScaleOff);   // [2][BN]
  float* sh_s = sc_s + 2 * BN;                                   // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmD);
    if (p.has_res) ptx::prefetch_tmap(&tmR);
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], kEpiThreads / 32); }
    for (int s = 0; s < kResBufs; ++s) { ptx::mbar_init(&rfull_bar[s], 1); ptx::mbar_init(&rempty_bar[s], kEpiThreads / 32); }
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

  // Everything above overlapped the previous kernel; its output (our A operand / residual) is
  // visible only after this point.
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ A/B TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
        const int m0 = m_blk * kBlockM, n0 = n_blk * BN;
        int cw = 0, ch = 0, cn = 0;
        if (p.im2col) {
          cn = m0 / p.HoWo;
          const int rem = m0 - cn * p.HoWo;
          const int po = rem / p.Wo, qo = rem - po * p.Wo;
          cw = qo * p.stride - p.pad;
          ch = po * p.stride - p.pad;
        }
        int stem_row = 0;
        if (BK == 32) {                                // stem: tiles never straddle images (12544 = 98 * 128)
          const int img = m0 / p.stem_img_rows;
          stem_row = img * p.stem_img_stride + (m0 - img * p.stem_img_rows);
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + kABytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          if (BK == 32) {
            ptx::tma_load_2d(&tmA, &full_bar[stage], sa, 0, stem_row + p.stem_tap_off[kb]);
          } else if (p.im2col) {
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
          const uint64_t adesc = ptx::make_kmajor_desc(sa, BK * 2);
          const uint64_t bdesc = ptx::make_kmajor_desc(sa + kABytes, BK * 2);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[as]);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual TMA producer
    if (lane == 0 && p.has_res) {
      int rs = 0; uint32_t rphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
        const int m0 = m_blk * kBlockM, n0 = n_blk * BN;
        for (int c = 0; c < BN / kChunkN; ++c) {
          if (n0 + c * kChunkN >= p.N) break;
          ptx::mbar_wait(&rempty_bar[rs], rphase ^ 1, 500 + rs);
          ptx::mbar_arrive_expect_tx(&rfull_bar[rs], kChunkBytes);
          ptx::tma_load_2d(&tmR, &rfull_bar[rs], smem + C::kResOff + rs * kChunkBytes, n0 + c * kChunkN, m0);
          if (++rs == kResBufs) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int et = threadIdx.x - kEpiWarp0 * 32;     // 0..127
    const int quad = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = quad * 32 + lane;                // row of the tile == TMEM lane
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t smem_a = ptx::smem_u32(smem), sc_a = ptx::smem_u32(sc_s), sh_a = ptx::smem_u32(sh_s);
    int it = 0, oc = 0, rs = 0;
    uint32_t rphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      const int m0 = m_blk * kBlockM, n0 = n_blk * BN;
      for (int i = et; i < BN; i += kEpiThreads) {
        const int n = n0 + i;
        ptx::sts_f32(sc_a + (as * BN + i) * 4, (n < p.N) ? (p.scale ? __ldg(p.scale + n) : 1.f) : 0.f);
        ptx::sts_f32(sh_a + (as * BN + i) * 4, (n < p.N && p.shift) ? __ldg(p.shift + n) : 0.f);
      }
      ptx::named_bar_sync(1, kEpiThreads);
      ptx::mbar_wait(&tfull_bar[as], aphase, 400 + as);
      ptx::tc_fence_after();
      int nchunks = (p.N - n0 + kChunkN - 1) / kChunkN;
      if (nchunks > BN / kChunkN) nchunks = BN / kChunkN;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + c * kChunkN;
        ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        ptx::tmem_ld_wait();
        if (c == nchunks - 1) {                      // accumulator drained: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
        }
        const uint32_t sc4 = sc_a + (as * BN + c * kChunkN) * 4, sh4 = sh_a + (as * BN + c * kChunkN) * 4;
        // v = acc * scale + shift with packed fp32 math (FFMA2); the shared-memory loads of each 32 columns go first (they are
        // volatile asm and would otherwise queue behind other shared-memory traffic)
        float2 V[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float4 S[8], Hs[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) { S[g] = ptx::lds_f4(sc4 + hh * 128 + g * 16); Hs[g] = ptx::lds_f4(sh4 + hh * 128 + g * 16); }
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int e = hh * 32 + g * 4;
            V[e / 2] = ptx::ffma2(make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), make_float2(S[g].x, S[g].y),
                                  make_float2(Hs[g].x, Hs[g].y));
            V[e / 2 + 1] = ptx::ffma2(make_float2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), make_float2(S[g].z, S[g].w),
                                      make_float2(Hs[g].z, Hs[g].w));
          }
        }
        if (p.has_res) {                       // the residual tile is waited for only now: the arithmetic above overlaps its arrival
          ptx::mbar_wait(&rfull_bar[rs], rphase, 600 + rs);
          const uint32_t rb = smem_a + C::kResOff + rs * kChunkBytes + row * 128;
          uint4 Q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) Q[j] = ptx::lds128(rb + (((uint32_t)j ^ swz) << 4));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            V[4 * j] = ptx::fadd2(V[4 * j], ptx::bf16x2_to_f2(Q[j].x));
            V[4 * j + 1] = ptx::fadd2(V[4 * j + 1], ptx::bf16x2_to_f2(Q[j].y));
            V[4 * j + 2] = ptx::fadd2(V[4 * j + 2], ptx::bf16x2_to_f2(Q[j].z));
            V[4 * j + 3] = ptx::fadd2(V[4 * j + 3], ptx::bf16x2_to_f2(Q[j].w));
          }
        }
        uint4 O[8];
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            O[j] = make_uint4(ptx::cvt_bf16x2_relu(V[4 * j].x, V[4 * j].y), ptx::cvt_bf16x2_relu(V[4 * j + 1].x, V[4 * j + 1].y),
                              ptx::cvt_bf16x2_relu(V[4 * j + 2].x, V[4 * j + 2].y), ptx::cvt_bf16x2_relu(V[4 * j + 3].x, V[4 * j + 3].y));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            O[j] = make_uint4(ptx::cvt_bf16x2(V[4 * j].x, V[4 * j].y), ptx::cvt_bf16x2(V[4 * j + 1].x, V[4 * j + 1].y),
                              ptx::cvt_bf16x2(V[4 * j + 2].x, V[4 * j + 2].y), ptx::cvt_bf16x2(V[4 * j + 3].x, V[4 * j + 3].y));
        }
        uint8_t* ob = smem + C::kOutOff + (oc % C::kOutBufs) * kChunkBytes;
        const uint32_t orow = smem_a + C::kOutOff + (oc % C::kOutBufs) * kChunkBytes + row * 128;
        if (C::kOutBufs == 2) ptx::named_bar_sync(2, kEpiThreads);   // thread 0 arrives after its wait_read<1>
#pragma unroll
        for (int j = 0; j < 8; ++j) ptx::sts128(orow + (((uint32_t)j ^ swz) << 4), O[j]);
        if (p.has_res) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&rempty_bar[rs]);
          if (++rs == kResBufs) { rs = 0; rphase ^= 1; }
        }
        ptx::fence_proxy_async();
        ptx::named_bar_sync(1, kEpiThreads);
        if (et == 0) {
          ptx::tma_store_2d(&tmD, ob, n0 + c * kChunkN, m0);
          ptx::tma_store_commit();
          // 3 buffers: chunk k+1 reuses the buffer of store k-2 -- all but the newest store have left smem
          // before thread 0 reaches the next barrier.  2 buffers: the extra barrier above orders it.
          ptx::tma_store_wait_read<1>();
        }
        ++oc;
      }
    }
    if (et == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int BN, int BK>
int launch_bn(const GemmLaunch& L, const KP& kp, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, BK>::kSmemBytes));
    configured = true;
  }
  const int tiles = kp.tiles_m * kp.tiles_n;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(tiles, grid_limit()));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg<BN, BK>::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = L.pdl ? 1 : 0;
  AP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tma_kernel<BN, BK>, L.tmA, L.tmB, L.tmD, L.epi.residual ? L.tmR : L.tmD, kp));
  count_launch();
  return 0;
}

}  // namespace

bool tma_epilogue_eligible(const GemmLaunch& L) {
  const Epilogue& e = L.epi;
  return e.out_bf16 && !e.out_f32 && !e.out_split && L.N % kChunkN == 0 && e.ldd % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(e.out_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(e.scale) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(e.shift) & 15) == 0 &&
         (!e.residual || (!e.residual_f32 && e.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(e.residual) & 15) == 0));
}

int enable_tma_epilogue(GemmLaunch* L) {
  AP_REQUIRE(tma_epilogue_eligible(*L), "enable_tma_epilogue: epilogue is not eligible for the TMA path");
  if (make_tmap_tiled_bf16(&L->tmD, L->epi.out_bf16, L->M, L->N, L->epi.ldd, kBlockM, kChunkN)) return 1;
  if (L->epi.residual && make_tmap_tiled_bf16(&L->tmR, L->epi.residual, L->M, L->N, L->epi.ldr, kBlockM, kChunkN)) return 1;
  L->tma_epi = 1;
  return 0;
}

int launch_gemm_tma(const GemmLaunch& L, cudaStream_t stream) {
  AP_REQUIRE(L.M > 0 && L.N > 0 && L.K > 0, "launch_gemm_tma: empty problem %dx%dx%d", L.M, L.N, L.K);
  AP_REQUIRE(L.tma_epi, "launch_gemm_tma: tensor maps of the epilogue were not built");
  KP kp{};
  kp.M = L.M; kp.N = L.N; kp.K = L.K;
  const int kBlockK = L.stem ? 32 : 64;
  kp.num_kb = ceil_div(L.K, kBlockK);
  kp.tiles_m = ceil_div(L.M, kBlockM);
  kp.tiles_n = ceil_div(L.N, L.block_n);
  kp.im2col = L.im2col;
  if (L.im2col) {
    const ConvGeom& g = L.geom;
    AP_REQUIRE(g.Cin % kBlockK == 0, "launch_gemm_tma: im2col needs Cin %% 64 == 0 (Cin=%d)", g.Cin);
    AP_REQUIRE(L.K == g.ksize * g.ksize * g.Cin, "launch_gemm_tma: K=%d does not match the conv geometry", L.K);
    kp.cblks = g.Cin / kBlockK; kp.ksize = g.ksize; kp.stride = g.stride; kp.pad = g.pad;
    kp.Wo = g.Wo; kp.HoWo = g.Ho * g.Wo;
  }
  if (L.stem) {
    AP_REQUIRE(L.block_n == 64 && kp.num_kb <= 8 && L.stem_img_rows % kBlockM == 0, "launch_gemm_tma: bad stem geometry");
    kp.stem = 1; kp.stem_img_rows = L.stem_img_rows; kp.stem_img_stride = L.stem_img_stride;
    for (int i = 0; i < 8; ++i) kp.stem_tap_off[i] = L.stem_tap_off[i];
    kp.has_res = L.epi.residual != nullptr;
    kp.relu = L.epi.relu;
    kp.scale = L.epi.scale; kp.shift = L.epi.shift;
    return launch_bn<64, 32>(L, kp, stream);
  }
  kp.has_res = L.epi.residual != nullptr;
  kp.relu = L.epi.relu;
  kp.scale = L.epi.scale; kp.shift = L.epi.shift;
  switch (L.block_n) {
    case 64: return launch_bn<64, 64>(L, kp, stream);
    case 128: return launch_bn<128, 64>(L, kp, stream);
    case 256: return launch_bn<256, 64>(L, kp, stream);
    default: AP_REQUIRE(false, "launch_gemm_tma: unsupported block_n %d", L.block_n);
  }
  return 0;
}

}  // namespace airpose
