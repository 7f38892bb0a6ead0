// Stream-K tcgen05 GEMM / implicit conv with a TMA epilogue -- second generation of the kernel every
// conv of the ResNet-50 trunk runs on (copenet/src/copenet/models/model_copenet.py:27-47,161-176).
//
// What the ncu captures of the first TMA-epilogue kernel (gemm_tma.cu; profiles/r01c_*) showed and
// what changes here:
//   * K-heavy layers (3x3 convs, layer3/4) moved 7-10 TB/s through the L2->SM crossbar with
//     128x128 tiles and lost up to a third of a wave to tile quantisation (196 or 392 tiles on 148
//     SMs).  Work is now split STREAM-K: the tiles x k-blocks iteration space is cut into one
//     contiguous range per CTA, so every SM gets the same number of k-blocks whatever the tile count
//     and 128x256 tiles (half the operand traffic per flop) can be used whenever Cout allows.  A tile
//     cut by a range boundary is finished by the CTA that owns its first k-block; the CTAs holding the
//     rest add their fp32 partial accumulators through an L2-resident workspace (at most one partial
//     per CTA, written at the START of that CTA's range, consumed at the END of the owner's range, so
//     nobody waits in practice).
//   * memory-bound layers (1x1 convs of layer1/2 with the residual) were latency-bound in a 4-warp
//     epilogue with a 2-chunk residual ring: there are now two epilogue warpgroups working on
//     alternate 64-column chunks, the shared-memory partition is chosen per launch (deep residual
//     ring when K is short, deep operand ring when K is long), and weights that fit stay RESIDENT in
//     shared memory for the whole launch instead of being re-fetched per tile.
//   * shared memory is addressed through the shared window (LDS/STS), not generic LD/ST.
//   * what finally bounds the K-heavy layers is the L2->SM fabric (~45-50 B/clk per SM with all 148 SMs
//     pulling): a 128xBN tile needs 64*(128+BN)/BN B/clk at full tensor rate, i.e. at most 22 / 33 / 44 % of
//     the tensor peak for BN = 64 / 128 / 256 -- exactly the tensor-pipe numbers ncu reports.  MT = 2 gives a
//     CTA two 128-row sub-tiles that share every B k-block (a 256xBN tile: half the B traffic per flop,
//     accumulators MT*BN TMEM columns, double-buffered only while 2*MT*BN <= 512).
//
//   * round 2, for the weight-gradient GEMMs of the training step (dW = dZ^T . im2col(X), both operands contracting over the
//     pixels = the OUTER dimension of NHWC tensors): operands given transposed are read through MN-major descriptors from
//     64-column x 64-row TMA boxes (KP::mn_a / mn_b; mn_b == 2: the boxes come from an im2col tensor map), warp 2 -- idle without a
//     residual -- issues the B boxes as a second TMA issuer, and GEMMs with a handful of tiles and thousands of k-blocks run as
//     plain split-K over all SMs (KP::splitk_r): every range leaves its fp32 partial in the workspace and the caller's reduce
//     kernel sums them in a fixed order (stream-K's one-owner gather caps a tile at 16 ranges).
//
// CTA = 11 warps, one CTA per SM:
//   warp 0    A/B TMA producer     warp 1   MMA issuer (TMEM double-buffered)    warp 2   residual TMA producer
//   warps 3-6 / 7-10   epilogue groups 0 / 1: tcgen05.ld -> scale/shift (+residual) -> relu -> bf16 ->
//                      swizzled smem -> TMA store; chunk q of the launch belongs to group q & 1.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <utility>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace airpose {

namespace {

constexpr int kBlockM = 128;
constexpr int kUmmaK = 16;
constexpr int kEpiWarp0 = 3;
constexpr int kEpiGroups = 2;
constexpr int kGroupThreads = 128;
constexpr int kThreads = 32 * (kEpiWarp0 + 4 * kEpiGroups);     // 352
constexpr int kChunkN = 64;                       // epilogue chunk: 128 rows x 64 bf16 = one 128B-swizzle box
constexpr int kChunkBytes = kBlockM * kChunkN * 2;
constexpr int kMaxStages = 8;
constexpr int kMaxRes = 6;
constexpr int kBarBytes = 512;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxGrid = 512;                     // flag slots

struct KP {
  int M, N, K;
  int num_kb, tiles_m, tiles_n;
  int im2col, cblks, ksize, stride, pad, Wo, HoWo;
  int stem, stem_img_rows, stem_img_stride;
  int stem_tap_off[8];
  int has_res, relu;
  int stages, res_bufs, b_res;      // shared-memory partition of this launch
  int split;                        // 1: stream-K (ranges of k-blocks)  0: whole tiles, round-robin over the CTAs
  int mn_a, mn_b;                   // operand given as [K][M] / [K][N]: 64 x 64 boxes, MN-major descriptors
  int b_warp;                       // transposed B: its boxes are issued by warp 2 (a second TMA issuer)
  int splitk_r;                     // > 0: plain split-K, every tile cut into splitk_r ranges, EVERY range leaves its fp32 partial in ws (slot = CTA)
  const float* scale;
  const float* shift;
  float* ws;                        // [grid][128 x BN] fp32 partial accumulators
  uint32_t* flags;                  // [grid][2]  == epoch once that CTA's partial (per epilogue group) is in ws
  uint32_t epoch;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// explicit shared-window accesses (the dynamic smem base is realigned by hand, which hides the
// address space from the compiler)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One contiguous piece of a CTA's range: k-blocks [kb0, kb1) of `tile`.
struct Seg { int tile, kb0, kb1; };
struct SegIter {
  int u, u1, num_kb, step;          // step == 0: stream-K range [u, u1) of k-block units; else tiles u, u+step, ... < u1
  __device__ SegIter(int cta, int grid, int units, int nkb, int split, int splitk_r = 0)
      : u(split ? (int)((int64_t)cta * units / grid) : cta), u1(split ? (int)((int64_t)(cta + 1) * units / grid) : units / nkb),
        num_kb(nkb), step(split ? 0 : grid) {
    if (splitk_r) {                   // CTA = (tile, r): k-blocks [r nkb / R, (r+1) nkb / R) of its tile
      const int tile = cta / splitk_r, r = cta - tile * splitk_r;
      u = tile * nkb + (int)((int64_t)r * nkb / splitk_r);
      u1 = tile * nkb + (int)((int64_t)(r + 1) * nkb / splitk_r);
      step = 0;
    }
  }
  __device__ bool next(Seg& s) {
    if (u >= u1) return false;
    if (step) {
      s.tile = u; s.kb0 = 0; s.kb1 = num_kb;
      u += step;
      return true;
    }
    s.tile = u / num_kb;
    s.kb0 = u - s.tile * num_kb;
    const int len = min(num_kb - s.kb0, u1 - u);
    s.kb1 = s.kb0 + len;
    u += len;
    return true;
  }
};

template <int BN, int BK, int MT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR, const KP p) {
  constexpr int kASub = kBlockM * BK * 2;           // one 128-row sub-tile of A
  constexpr int kABytes = MT * kASub;
  constexpr int kBBytes = BN * BK * 2;
  constexpr int kAccCols = MT * BN;                 // TMEM columns of one accumulator set
  constexpr int kAccBufs = (2 * kAccCols <= 512) ? 2 : 1;
  constexpr int kTmemCols = kAccBufs * kAccCols;
  constexpr int kChunks = BN / kChunkN;
  constexpr int kTileM = MT * kBlockM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);
  const uint32_t smem_base = ptx::smem_u32(smem);

  const int stage_bytes = p.b_res ? kABytes : kABytes + kBBytes;
  const int ring_bytes = p.stages * stage_bytes;
  const int bres_off = ring_bytes;
  const int out_off = bres_off + (p.b_res ? p.num_kb * kBBytes : 0);
  const int res_off = out_off + kEpiGroups * kChunkBytes;
  const int bar_off = res_off + p.res_bufs * kChunkBytes;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + bar_off);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;
  uint64_t* rempty_bar = rfull_bar + kMaxRes;
  uint64_t* bfull_bar = rempty_bar + kMaxRes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = p.tiles_m * p.tiles_n * p.num_kb;
  const int cta = blockIdx.x, grid = gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmD);
    if (p.has_res) ptx::prefetch_tmap(&tmR);
    for (int s = 0; s < kMaxStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], 4 * kEpiGroups); }
    for (int s = 0; s < kMaxRes; ++s) { ptx::mbar_init(&rfull_bar[s], 1); ptx::mbar_init(&rempty_bar[s], 4); }
    ptx::mbar_init(bfull_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Everything above overlapped the previous kernel; its output (our A operand / residual) and the
  // stream-K workspace are ours only after this point.
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ A/B TMA producer
    if (lane == 0) {
      if (p.b_res) {                                   // the whole weight matrix, once
        ptx::mbar_arrive_expect_tx(bfull_bar, p.num_kb * kBBytes);
        for (int kb = 0; kb < p.num_kb; ++kb) ptx::tma_load_2d(&tmB, bfull_bar, smem + bres_off + kb * kBBytes, kb * BK, 0);
      }
      int stage = 0; uint32_t phase = 0;
      SegIter it(cta, grid, units, p.num_kb, p.split, p.splitk_r);
      Seg s;
      while (it.next(s)) {
        const int m_blk = s.tile / p.tiles_n, n_blk = s.tile - m_blk * p.tiles_n;
        const int m0 = m_blk * kTileM, n0 = n_blk * BN;
        const int vmt = min(MT, (p.M - m0 + kBlockM - 1) / kBlockM);     // sub-tiles that hold rows of the problem
        int cw[MT], ch[MT], cn[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          cw[mt] = ch[mt] = cn[mt] = 0;
          if (p.im2col) {
            const int mm = m0 + mt * kBlockM;
            cn[mt] = mm / p.HoWo;
            const int rem = mm - cn[mt] * p.HoWo;
            const int po = rem / p.Wo, qo = rem - po * p.Wo;
            cw[mt] = qo * p.stride - p.pad;
            ch[mt] = po * p.stride - p.pad;
          }
        }
        int stem_row = 0;
        if (BK == 32) {                                // stem: tiles never straddle images (12544 = 98 * 128)
          const int img = m0 / p.stem_img_rows;
          stem_row = img * p.stem_img_stride + (m0 - img * p.stem_img_rows);
        }
        // MN-major operands arrive as boxes of 64 MN columns x BK rows; boxes past the extent are not loaded (their rows / columns
        // of the accumulator are never stored)
        constexpr int kBoxBytes = 64 * BK * 2;
        const int a_boxes = p.mn_a ? min(MT * 2, (p.M - m0 + 63) / 64) : 0;
        const int b_boxes = p.mn_b ? min(BN / 64, (p.N - n0 + 63) / 64) : 0;
        const int tx_bytes = (p.mn_a ? a_boxes * kBoxBytes : vmt * kASub) + (p.b_res ? 0 : (p.mn_b ? b_boxes * kBoxBytes : kBBytes));
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          uint8_t* sa = smem + stage * stage_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          if (p.mn_a) {
            for (int j = 0; j < a_boxes; ++j) ptx::tma_load_2d(&tmA, &full_bar[stage], sa + j * kBoxBytes, m0 + j * 64, kb * BK);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt >= vmt || p.mn_a) break;
            uint8_t* sam = sa + mt * kASub;
            if (BK == 32) {
              ptx::tma_load_2d(&tmA, &full_bar[stage], sam, 0, stem_row + p.stem_tap_off[kb]);
            } else if (p.im2col) {
              const int tap = kb / p.cblks, cb = kb - tap * p.cblks;
              const int r = tap / p.ksize, sx = tap - r * p.ksize;
              ptx::tma_load_im2col_4d(&tmA, &full_bar[stage], sam, cb * BK, cw[mt], ch[mt], cn[mt], (uint16_t)sx, (uint16_t)r);
            } else {
              ptx::tma_load_2d(&tmA, &full_bar[stage], sam, kb * BK, m0 + mt * kBlockM);
            }
          }
          if (p.mn_b && p.b_warp) {
            // the B boxes of this k-block are issued by warp 2 (below)
          } else if (p.mn_b == 2) {
            // B = im2col(x)^T: k-block kb is 64 output pixels, column block j of the tile is (tap, 64 input channels) -- the box the
            // forward conv loads as its A operand, 64 pixels tall
            const int mm = kb * BK;
            const int img = mm / p.HoWo, rem = mm - img * p.HoWo;
            const int po = rem / p.Wo, qo = rem - po * p.Wo;
            for (int j = 0; j < b_boxes; ++j) {
              const int nb = (n0 >> 6) + j, tap = nb / p.cblks, cb = nb - tap * p.cblks;
              const int r = tap / p.ksize, sx = tap - r * p.ksize;
              ptx::tma_load_im2col_4d(&tmB, &full_bar[stage], sa + kABytes + j * kBoxBytes, cb * 64, qo * p.stride - p.pad,
                                      po * p.stride - p.pad, img, (uint16_t)sx, (uint16_t)r);
            }
          } else if (p.mn_b) {
            for (int j = 0; j < b_boxes; ++j) ptx::tma_load_2d(&tmB, &full_bar[stage], sa + kABytes + j * kBoxBytes, n0 + j * 64, kb * BK);
          } else if (!p.b_res) {
            ptx::tma_load_2d(&tmB, &full_bar[stage], sa + kABytes, kb * BK, n0);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(kBlockM, BN) | (p.mn_a ? ptx::kIdescAMn : 0u) | (p.mn_b ? ptx::kIdescBMn : 0u);
      constexpr int kBoxBytes = 64 * BK * 2;
      const uint32_t a_kstep = p.mn_a ? (uint32_t)(kUmmaK * 128) >> 4 : 2u, b_kstep = p.mn_b ? (uint32_t)(kUmmaK * 128) >> 4 : 2u;
      if (p.b_res) ptx::mbar_wait(bfull_bar, 0, 250);
      int stage = 0; uint32_t phase = 0;
      int n = 0;
      SegIter it(cta, grid, units, p.num_kb, p.split, p.splitk_r);
      Seg s;
      while (it.next(s)) {
        const int as = n % kAccBufs; const uint32_t aphase = (n / kAccBufs) & 1;
        ++n;
        const int m_blk = s.tile / p.tiles_n;
        const int vmt = min(MT, (p.M - m_blk * kTileM + kBlockM - 1) / kBlockM);
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1, 200 + as);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccCols;
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase, 300 + stage);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = p.b_res ? smem_base + bres_off + kb * kBBytes : sa + kABytes;
          const uint64_t bdesc = p.mn_b ? ptx::make_mnmajor_desc(sb, kBoxBytes) : ptx::make_kmajor_desc(sb, BK * 2);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt >= vmt) break;
            const uint64_t adesc = p.mn_a ? ptx::make_mnmajor_desc(sa + mt * kASub, kBoxBytes) : ptx::make_kmajor_desc(sa + mt * kASub, BK * 2);
#pragma unroll
            for (int k = 0; k < BK / kUmmaK; ++k)
              ptx::umma_bf16(d_tmem + mt * BN, adesc + a_kstep * k, bdesc + b_kstep * k, idesc, (kb != s.kb0 || k != 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[as]);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual TMA producer
    if (lane == 0 && p.mn_b && p.b_warp) {
      // second TMA issuer (transposed operands: a k-block is up to 6 boxes of 8 KB, and one thread issues a box per ~1000 clk):
      // this warp loads the B boxes of every k-block, warp 0 the A boxes; both wait for the slot, warp 0 arms the barrier with
      // the bytes of both (a complete_tx may precede the expect_tx of its phase: the pending arrival keeps the phase open)
      constexpr int kBoxBytes = 64 * BK * 2;
      int stage = 0; uint32_t phase = 0;
      SegIter it(cta, grid, units, p.num_kb, p.split, p.splitk_r);
      Seg s;
      while (it.next(s)) {
        const int n_blk = s.tile % p.tiles_n, n0 = n_blk * BN;
        const int b_boxes = min(BN / 64, (p.N - n0 + 63) / 64);
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 700 + stage);
          uint8_t* sb = smem + stage * stage_bytes + kABytes;
          if (p.mn_b == 2) {
            const int mm = kb * BK;
            const int img = mm / p.HoWo, rem = mm - img * p.HoWo;
            const int po = rem / p.Wo, qo = rem - po * p.Wo;
            for (int j = 0; j < b_boxes; ++j) {
              const int nb = (n0 >> 6) + j, tap = nb / p.cblks, cb = nb - tap * p.cblks;
              const int r = tap / p.ksize, sx = tap - r * p.ksize;
              ptx::tma_load_im2col_4d(&tmB, &full_bar[stage], sb + j * kBoxBytes, cb * 64, qo * p.stride - p.pad,
                                      po * p.stride - p.pad, img, (uint16_t)sx, (uint16_t)r);
            }
          } else {
            for (int j = 0; j < b_boxes; ++j) ptx::tma_load_2d(&tmB, &full_bar[stage], sb + j * kBoxBytes, n0 + j * 64, kb * BK);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 0 && p.has_res) {
      int rq = 0;
      SegIter it(cta, grid, units, p.num_kb, p.split, p.splitk_r);
      Seg s;
      while (it.next(s)) {
        if (s.kb0 != 0 || p.splitk_r) continue;        // a partial handed to the tile's owner: no epilogue here
        const int m_blk = s.tile / p.tiles_n, n_blk = s.tile - m_blk * p.tiles_n;
        const int m0 = m_blk * kTileM, n0 = n_blk * BN;
        const int vmt = min(MT, (p.M - m0 + kBlockM - 1) / kBlockM);
        const int ncn = min(kChunks, (p.N - n0) / kChunkN);
        for (int c = 0; c < vmt * ncn; ++c, ++rq) {
          const int mt = c / ncn, cn = c - mt * ncn;
          const int rs = rq % p.res_bufs; const uint32_t rphase = (rq / p.res_bufs) & 1;
          ptx::mbar_wait(&rempty_bar[rs], rphase ^ 1, 500 + rs);
          ptx::mbar_arrive_expect_tx(&rfull_bar[rs], kChunkBytes);
          ptx::tma_load_2d(&tmR, &rfull_bar[rs], smem + res_off + rs * kChunkBytes, n0 + cn * kChunkN, m0 + mt * kBlockM);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue groups
    const int g = (warp - kEpiWarp0) >> 2;           // epilogue group
    const int quad = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = quad * 32 + lane;                // row of the sub-tile == TMEM lane
    const bool elected = ((warp - kEpiWarp0) & 3) == 0 && lane == 0;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t ob = smem_base + out_off + g * kChunkBytes;
    const uint32_t orow = ob + row * 128;
    const int bar_id = 1 + g;
    constexpr int kSlotF4 = kTileM * BN / 4;         // float4s per CTA slot of the workspace
    float4* ws_mine = reinterpret_cast<float4*>(p.ws) + (size_t)cta * kSlotF4;
    int n = 0, q = 0, rq = 0;
    SegIter it(cta, grid, units, p.num_kb, p.split, p.splitk_r);
    Seg s;
    while (it.next(s)) {
      const int as = n % kAccBufs; const uint32_t aphase = (n / kAccBufs) & 1;
      ++n;
      const int m_blk = s.tile / p.tiles_n, n_blk = s.tile - m_blk * p.tiles_n;
      const int m0 = m_blk * kTileM, n0 = n_blk * BN;
      const int vmt = min(MT, (p.M - m0 + kBlockM - 1) / kBlockM);
      const int ncn = min(kChunks, (p.N - n0) / kChunkN);
      const int nchunks = vmt * ncn;
      const bool contributor = p.splitk_r ? true : s.kb0 != 0;
      const bool gather = !contributor && s.kb1 < p.num_kb;     // owner of a tile other CTAs finish
      ptx::mbar_wait(&tfull_bar[as], aphase, 400 + as);
      ptx::tc_fence_after();
      int mine_left = 0;
      for (int c = 0; c < nchunks; ++c) mine_left += (((q + c) & 1) == g);
      if (mine_left == 0) {                          // nothing of this accumulator is ours: release it
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      }
      if (gather) {                                  // wait for every CTA that holds the rest of this tile
        int need = p.num_kb - s.kb1, cc = cta + 1;
        while (need > 0) {
          const int a0 = (int)((int64_t)cc * units / grid), a1 = (int)((int64_t)(cc + 1) * units / grid);
          if (a1 > a0) {
            uint32_t spins = 0;
            while (ld_acquire(p.flags + 2 * cc) != p.epoch || ld_acquire(p.flags + 2 * cc + 1) != p.epoch) {
              if (++spins > (1u << 22)) {
                printf("airpose: stream-K flag timeout cta=%d waits for %d\n", cta, cc);
                __trap();
              }
            }
            need -= min(a1 - a0, need);
          }
          ++cc;
        }
      }
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        if (((q + c) & 1) != g) continue;
        const int mt = c / ncn, cn = c - mt * ncn;
        const int slot_c = (mt * kChunks + cn) * 16 * kBlockM + row;     // float4 index of this thread's part of chunk c
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kAccCols + mt * BN + cn * kChunkN;
        ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        ptx::tmem_ld_wait();
        if (--mine_left == 0) {                      // our part of the accumulator is in registers
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
        }
        if (contributor) {                           // fp32 partial -> workspace (coalesced: 16 B x 32 lanes)
          float4* dst = ws_mine + slot_c;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            __stcg(dst + j * kBlockM, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                  __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
          continue;
        }
        float v[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
        if (gather) {
          int need = p.num_kb - s.kb1, cc = cta + 1;
          while (need > 0) {
            const int a0 = (int)((int64_t)cc * units / grid), a1 = (int)((int64_t)(cc + 1) * units / grid);
            if (a1 > a0) {
              const float4* src = reinterpret_cast<const float4*>(p.ws) + (size_t)cc * kSlotF4 + slot_c;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 t = __ldcg(src + j * kBlockM);
                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
              }
              need -= min(a1 - a0, need);
            }
            ++cc;
          }
        }
        {
          const int nb = n0 + cn * kChunkN;
          const float4* sc4 = reinterpret_cast<const float4*>(p.scale + nb);
          const float4* sh4 = reinterpret_cast<const float4*>(p.shift + nb);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 s4 = p.scale ? __ldg(sc4 + j) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 h4 = p.shift ? __ldg(sh4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * j + 0] = fmaf(v[4 * j + 0], s4.x, h4.x);
            v[4 * j + 1] = fmaf(v[4 * j + 1], s4.y, h4.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], s4.z, h4.z);
            v[4 * j + 3] = fmaf(v[4 * j + 3], s4.w, h4.w);
          }
        }
        if (p.has_res) {
          // chunks of owner segments are numbered rq, rq+1, ... in the producer's order
          const int rc = rq + c;
          const int rs = rc % p.res_bufs; const uint32_t rphase = (rc / p.res_bufs) & 1;
          ptx::mbar_wait(&rfull_bar[rs], rphase, 600 + rs);
          const uint32_t rb = smem_base + res_off + rs * kChunkBytes + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 w4 = lds128(rb + (((uint32_t)j ^ swz) << 4));
            const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              v[j * 8 + 2 * h] += __uint_as_float(w[h] << 16);
              v[j * 8 + 2 * h + 1] += __uint_as_float(w[h] & 0xFFFF0000u);
            }
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&rempty_bar[rs]);
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        // the group's staging buffer is free once its previous store has been read out of smem
        if (elected) ptx::tma_store_wait_read<0>();
        ptx::named_bar_sync(bar_id, kGroupThreads);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o;
          o.x = pack_bf16(v[j * 8 + 0], v[j * 8 + 1]); o.y = pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
          o.z = pack_bf16(v[j * 8 + 4], v[j * 8 + 5]); o.w = pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
          sts128(orow + (((uint32_t)j ^ swz) << 4), o);
        }
        ptx::fence_proxy_async();
        ptx::named_bar_sync(bar_id, kGroupThreads);
        if (elected) {
          ptx::tma_store_2d(&tmD, smem + out_off + g * kChunkBytes, n0 + cn * kChunkN, m0 + mt * kBlockM);
          ptx::tma_store_commit();
        }
      }
      if (contributor) {                             // publish this group's part of the partial
        __threadfence();
        ptx::named_bar_sync(bar_id, kGroupThreads);
        if (elected) st_release(p.flags + 2 * cta + g, p.epoch);
      }
      q += nchunks;
      if (!contributor && p.has_res) rq += nchunks;
    }
    if (elected) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------- host side
struct SkWorkspace {
  float* ws = nullptr;
  uint32_t* flags = nullptr;
  uint32_t epoch = 0;
  int device = -1;
};
// one workspace per (device, stream): stream-K launches on different streams may run concurrently
std::map<std::pair<int, cudaStream_t>, SkWorkspace> g_ws;
constexpr int kMaxMT = 2;

int get_workspace(cudaStream_t stream, SkWorkspace** out) {
  int dev = 0;
  AP_CHECK_CUDA(cudaGetDevice(&dev));
  SkWorkspace& w = g_ws[std::make_pair(dev, stream)];
  if (!w.ws) {
    AP_CHECK_CUDA(cudaMalloc((void**)&w.ws, (size_t)kMaxGrid * kMaxMT * kBlockM * 256 * sizeof(float)));
    AP_CHECK_CUDA(cudaMalloc((void**)&w.flags, (size_t)kMaxGrid * 2 * sizeof(uint32_t)));
    AP_CHECK_CUDA(cudaMemset(w.flags, 0, (size_t)kMaxGrid * 2 * sizeof(uint32_t)));
    AP_CHECK_CUDA(cudaDeviceSynchronize());
    w.device = dev;
  }
  *out = &w;
  return 0;
}

template <int BN, int BK, int MT>
int launch_bn(const GemmLaunch& L, KP& kp, cudaStream_t stream) {
  constexpr int kABytes = MT * kBlockM * BK * 2, kBBytes = BN * BK * 2;
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(gemm_sk_kernel<BN, BK, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    configured = true;
  }
  kp.tiles_m = ceil_div(kp.M, MT * kBlockM);
  // shared-memory partition: [operand ring | resident B | 2 output chunks | residual ring | barriers]
  const int fixed = 1024 + kBarBytes + kEpiGroups * kChunkBytes;
  kp.res_bufs = kp.has_res ? (kp.num_kb <= 2 ? 4 : 2) : 0;
  int avail = kSmemLimit - fixed - kp.res_bufs * kChunkBytes;
  const int b_total = kp.num_kb * kBBytes;
  kp.b_res = (!kp.mn_b && kp.tiles_n == 1 && kp.tiles_m > num_sms() && b_total <= 80 * 1024 && avail - b_total >= 3 * kABytes &&
              !getenv("AIRPOSE_NO_BRES")) ? 1 : 0;
  if (kp.b_res) avail -= b_total;
  const int stage_bytes = kp.b_res ? kABytes : kABytes + kBBytes;
  kp.stages = std::min(kMaxStages, avail / stage_bytes);
  AP_REQUIRE(kp.stages >= 2, "gemm_sk: shared memory partition failed (BN=%d MT=%d)", BN, MT);
  const int smem_bytes = 1024 + kp.stages * stage_bytes + (kp.b_res ? b_total : 0) + kEpiGroups * kChunkBytes +
                         kp.res_bufs * kChunkBytes + kBarBytes;
  SkWorkspace* w = nullptr;
  if (get_workspace(stream, &w)) return 1;
  kp.ws = w->ws; kp.flags = w->flags; kp.epoch = ++w->epoch;
  {
    // Under stream capture the epoch is frozen into the graph: a replay would find the flags of the previous replay (same
    // value) already set.  Clear them inside the graph before every stream-K launch, so each replay starts from zero flags.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    AP_CHECK_CUDA(cudaStreamIsCapturing(stream, &cs));
    if (cs != cudaStreamCaptureStatusNone) AP_CHECK_CUDA(cudaMemsetAsync(w->flags, 0, (size_t)kMaxGrid * 2 * sizeof(uint32_t), stream));
  }
  const int units = kp.tiles_m * kp.tiles_n * kp.num_kb;
  cudaLaunchConfig_t cfg{};
  // Stream-K pays a fixed price per cut tile (an fp32 partial tile through L2 and a gather at the end of
  // the owner's range): worth it when tiles are long in K and the tile count quantises badly on the SMs
  // (measured: layer3/4 3x3 convs -20..30 %, short-K layers +10..30 %).  Everything else runs whole
  // tiles round-robin.  Tiny problems are never split finer than 8 k-blocks per CTA.
  const int tiles = kp.tiles_m * kp.tiles_n;
  const int sms = std::min(num_sms(), kMaxGrid);
  static const int min_kb = getenv("AIRPOSE_SK_SPLIT_MINKB") ? atoi(getenv("AIRPOSE_SK_SPLIT_MINKB")) : 8;
  kp.split = (kp.num_kb >= min_kb && tiles < 8 * sms && tiles % sms != 0) ? 1 : 0;
  if (kp.splitk_r) kp.split = 1;
  // a tile is never cut into more than 16 ranges: its owner gathers the partials one after the other, which at 70+
  // partials per tile (weight-gradient GEMMs: 2 tiles, 1500 k-blocks) cost more than the mainloop they parallelised
  cfg.gridDim = dim3((unsigned)(kp.splitk_r ? tiles * kp.splitk_r
                                : kp.split ? std::min(sms, std::max(std::min(tiles, sms), std::min(units / 8, tiles * 16)))
                                           : std::min(tiles, sms)));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = L.pdl ? 1 : 0;
  AP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_sk_kernel<BN, BK, MT>, L.tmA, L.tmB, L.tmD, L.epi.residual ? L.tmR : L.tmD, kp));
  count_launch();
  return 0;
}

}  // namespace

int splitk_ranges(int M, int N, int K, int block_n) {
  const int tiles = ceil_div(M, kBlockM) * ceil_div(N, block_n), num_kb = ceil_div(K, 64);
  const int sms = std::min(num_sms(), kMaxGrid);
  if (tiles * 16 >= sms || getenv("AIRPOSE_NO_SPLITK")) return 0;        // stream-K's 16 ranges per tile already fill the GPU
  const int r = std::min(sms / tiles, num_kb / 4);
  return r >= 2 ? r : 0;
}

int launch_gemm_sk(const GemmLaunch& L, cudaStream_t stream, SplitKInfo* sk) {
  AP_REQUIRE(L.M > 0 && L.N > 0 && L.K > 0, "launch_gemm_sk: empty problem %dx%dx%d", L.M, L.N, L.K);
  AP_REQUIRE(L.tma_epi, "launch_gemm_sk: tensor maps of the epilogue were not built");
  KP kp{};
  kp.M = L.M; kp.N = L.N; kp.K = L.K;
  const int kBlockK = L.stem ? 32 : 64;
  kp.num_kb = ceil_div(L.K, kBlockK);
  kp.tiles_n = ceil_div(L.N, L.block_n);
  if (sk) {
    AP_REQUIRE(!L.stem && !L.epi.residual && !getenv("AIRPOSE_SK_MT2"), "launch_gemm_sk: split-K partials are for plain GEMMs");
    kp.splitk_r = sk->ranges;
    AP_REQUIRE(kp.splitk_r >= 2 && kp.splitk_r <= kp.num_kb && ceil_div(L.M, kBlockM) * kp.tiles_n * kp.splitk_r <= kMaxGrid,
               "launch_gemm_sk: bad split-K range count %d", kp.splitk_r);
    SkWorkspace* w = nullptr;
    if (get_workspace(stream, &w)) return 1;
    sk->part = w->ws; sk->tiles_n = kp.tiles_n; sk->block_n = L.block_n; sk->slot_floats = kBlockM * L.block_n;
  }
  kp.im2col = L.im2col;
  kp.mn_a = L.mn_a; kp.mn_b = L.mn_b;
  kp.b_warp = (L.mn_b && !L.epi.residual && !getenv("AIRPOSE_SK_ONE_ISSUER")) ? 1 : 0;
  AP_REQUIRE(!(L.mn_a || L.mn_b) || (!L.stem && !L.im2col), "launch_gemm_sk: MN-major operands are for plain GEMMs");
  if (L.mn_b == 2) {                      // B = im2col(x)^T through an im2col tensor map with 64-pixel boxes (weight gradients)
    const ConvGeom& g = L.geom;
    AP_REQUIRE(g.Cin % 64 == 0 && L.N == g.ksize * g.ksize * g.Cin && L.K == g.n * g.Ho * g.Wo, "launch_gemm_sk: im2col B does not match the geometry");
    kp.cblks = g.Cin / 64; kp.ksize = g.ksize; kp.stride = g.stride; kp.pad = g.pad;
    kp.Wo = g.Wo; kp.HoWo = g.Ho * g.Wo;
  }
  kp.has_res = L.epi.residual != nullptr;
  kp.relu = L.epi.relu;
  kp.scale = L.epi.scale; kp.shift = L.epi.shift;
  if (L.im2col) {
    const ConvGeom& g = L.geom;
    AP_REQUIRE(g.Cin % kBlockK == 0, "launch_gemm_sk: im2col needs Cin %% 64 == 0 (Cin=%d)", g.Cin);
    AP_REQUIRE(L.K == g.ksize * g.ksize * g.Cin, "launch_gemm_sk: K=%d does not match the conv geometry", L.K);
    kp.cblks = g.Cin / kBlockK; kp.ksize = g.ksize; kp.stride = g.stride; kp.pad = g.pad;
    kp.Wo = g.Wo; kp.HoWo = g.Ho * g.Wo;
  }
  if (L.stem) {
    AP_REQUIRE(L.block_n == 64 && kp.num_kb <= 8 && L.stem_img_rows % kBlockM == 0, "launch_gemm_sk: bad stem geometry");
    kp.stem = 1; kp.stem_img_rows = L.stem_img_rows; kp.stem_img_stride = L.stem_img_stride;
    for (int i = 0; i < 8; ++i) kp.stem_tap_off[i] = L.stem_tap_off[i];
    return launch_bn<64, 32, 1>(L, kp, stream);
  }
  // MT = 2 (two 128-row sub-tiles sharing each B k-block, 256 x BN tiles) is kept for experiments only:
  // measured SLOWER than MT = 1 on every trunk layer (layer3 3x3: 47.7 vs 39.6 us, layer4 3x3: 54 vs 41 us;
  // profiles/r01e_layers_mt2.txt).  It trades the L2->SM traffic for a single-buffered accumulator, and the
  // binding limit of a cta_group::1 tile is shared-memory bandwidth (MMA operand reads + TMA writes share
  // 128 B/clk: a 128x256 tile needs 192 B/clk at full tensor rate), which MT = 2 barely changes (160 B/clk).
  static const bool want_mt2 = getenv("AIRPOSE_SK_MT2") != nullptr;
  const bool mt2 = (want_mt2 || L.mt2) && L.M > kBlockM && !sk;
  switch (L.block_n) {
    case 64: return mt2 ? launch_bn<64, 64, 2>(L, kp, stream) : launch_bn<64, 64, 1>(L, kp, stream);
    case 128: return mt2 ? launch_bn<128, 64, 2>(L, kp, stream) : launch_bn<128, 64, 1>(L, kp, stream);
    case 256: return mt2 ? launch_bn<256, 64, 2>(L, kp, stream) : launch_bn<256, 64, 1>(L, kp, stream);
    default: AP_REQUIRE(false, "launch_gemm_sk: unsupported block_n %d", L.block_n);
  }
  return 0;
}

}  // namespace airpose
