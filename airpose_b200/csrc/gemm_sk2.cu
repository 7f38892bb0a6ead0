// cta_group::2 variant of the stream-K conv GEMM (gemm_sk.cu): a CTA PAIR (cluster of 2, one TPC) owns a 256 x 256
// tile.  Each CTA loads its own 128 rows of A and HALF of the B k-block (128 of the 256 weight rows); one thread of the
// leader CTA issues `tcgen05.mma.cta_group::2` (M = 256, N = 256, K = 16), which feeds both tensor cores from the two
// shared memories and accumulates 128 x 256 fp32 in EACH CTA's TMEM.  Per SM and k-block that is 16 KB of A + 16 KB of
// B written by TMA and read once by the MMA -- 128 B/clk of shared-memory traffic at the full tensor rate instead of
// the 192 B/clk of a cta_group::1 128 x 256 tile, and half the L2->SM operand traffic per flop.
//
// Protocol (barriers at the same shared-memory offsets in both CTAs):
//   full[s]    lives in the LEADER: the leader's producer arrives with expect_tx for both CTAs' bytes; both CTAs' TMA
//              loads (cp.async.bulk.tensor...cta_group::2) complete_tx on it (peer bit of the barrier address cleared).
//   empty[s]   one per CTA: the leader's `tcgen05.commit.cta_group::2 ... multicast::cluster` arrives on both.
//   tfull[a]   one per CTA, same multicast commit: each CTA's epilogue drains its own 128 accumulator rows.
//   tempty[a]  lives in the LEADER, count = 2 CTAs x 8 epilogue warps; the peer's warps arrive remotely (mapa).
// Everything else (stream-K ranges over PAIRS, fp32 partials through L2 -- one slot per CTA --, two epilogue groups,
// TMA stores, residual ring, PDL) is gemm_sk.cu's.
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

// ---- cluster / cta_group::2 primitives (forms as in cute/arch/copy_sm100_tma.hpp, cutlass/arch/barrier.h)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address -> the leader's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w, int h, int n,
                                                    uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w),
      "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[256 x 16] * B[256 x 16]^T; A rows / B rows are split across the two CTAs' shared memories
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(ptx::smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
// arrive on the LEADER's copy of a barrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(ptx::smem_u32(bar))
      : "memory");
}

// One contiguous piece of a CTA's range: k-blocks [kb0, kb1) of `tile`.
struct Seg { int tile, kb0, kb1; };
struct SegIter {
  int u, u1, num_kb, step;          // step == 0: stream-K range [u, u1) of k-block units; else tiles u, u+step, ... < u1
  __device__ SegIter(int cta, int grid, int units, int nkb, int split)
      : u(split ? (int)((int64_t)cta * units / grid) : cta), u1(split ? (int)((int64_t)(cta + 1) * units / grid) : units / nkb),
        num_kb(nkb), step(split ? 0 : grid) {}
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

template <int BN, int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_sk2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR, const KP p) {
  static_assert(BN == 256 && BK == 64, "the pair kernel is built for 256 x 256 pair tiles");
  constexpr int kABytes = kBlockM * BK * 2;         // this CTA's 128 rows of A
  constexpr int kBHalf = (BN / 2) * BK * 2;         // this CTA's half of the B k-block
  constexpr int kStageBytes = kABytes + kBHalf;     // 32 KB
  constexpr int kAccCols = BN;
  constexpr int kTmemCols = 2 * kAccCols;           // double-buffered
  constexpr int kChunks = BN / kChunkN;
  constexpr int kTileM = 2 * kBlockM;               // rows of a PAIR tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);
  const uint32_t smem_base = ptx::smem_u32(smem);

  const int out_off = p.stages * kStageBytes;
  const int res_off = out_off + kEpiGroups * kChunkBytes;
  const int bar_off = res_off + p.res_bufs * kChunkBytes;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + bar_off);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;
  uint64_t* rempty_bar = rfull_bar + kMaxRes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty_bar + kMaxRes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
  const int units = p.tiles_m * p.tiles_n * p.num_kb;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmD);
    if (p.has_res) ptx::prefetch_tmap(&tmR);
    for (int s = 0; s < kMaxStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull_bar[s], 1); ptx::mbar_init(&tempty_bar[s], 2 * 4 * kEpiGroups); }
    for (int s = 0; s < kMaxRes; ++s) { ptx::mbar_init(&rfull_bar[s], 1); ptx::mbar_init(&rempty_bar[s], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {                                   // the same warp in both CTAs
    tmem2_alloc(tmem_slot, kTmemCols);
    tmem2_relinquish();
  }
  ptx::tc_fence_before();
  cluster_sync_all();                                // barriers of both CTAs are initialised before anyone signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ A / B-half TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      SegIter it(pair, npairs, units, p.num_kb, p.split);
      Seg s;
      while (it.next(s)) {
        const int m_blk = s.tile / p.tiles_n, n_blk = s.tile - m_blk * p.tiles_n;
        const int m0 = m_blk * kTileM + (int)rank * kBlockM, n0 = n_blk * BN + (int)rank * (BN / 2);
        const bool a_valid = m0 < p.M;               // the peer's half of the last pair tile may hold no rows at all
        const bool peer_valid = m_blk * kTileM + kBlockM < p.M;
        int cw = 0, ch = 0, cn = 0;
        if (p.im2col && a_valid) {
          cn = m0 / p.HoWo;
          const int rem = m0 - cn * p.HoWo;
          const int po = rem / p.Wo, qo = rem - po * p.Wo;
          cw = qo * p.stride - p.pad;
          ch = po * p.stride - p.pad;
        }
        // the leader announces the bytes of BOTH CTAs on its own barrier
        const uint32_t tx_bytes = (uint32_t)(kABytes + kBHalf) + (uint32_t)kBHalf + (peer_valid ? (uint32_t)kABytes : 0u);
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          uint8_t* sa = smem + stage * kStageBytes;
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          if (a_valid) {
            if (p.im2col) {
              const int tap = kb / p.cblks, cb = kb - tap * p.cblks;
              const int r = tap / p.ksize, sx = tap - r * p.ksize;
              tma2_load_im2col_4d(&tmA, &full_bar[stage], sa, cb * BK, cw, ch, cn, (uint16_t)sx, (uint16_t)r);
            } else {
              tma2_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m0);
            }
          }
          tma2_load_2d(&tmB, &full_bar[stage], sa + kABytes, kb * BK, n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * kBlockM, BN);
      int stage = 0; uint32_t phase = 0;
      int n = 0;
      SegIter it(pair, npairs, units, p.num_kb, p.split);
      Seg s;
      while (it.next(s)) {
        const int as = n & 1; const uint32_t aphase = (n >> 1) & 1;
        ++n;
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1, 200 + as);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccCols;
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase, 300 + stage);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint64_t adesc = ptx::make_kmajor_desc(sa, BK * 2);
          const uint64_t bdesc = ptx::make_kmajor_desc(sa + kABytes, BK * 2);
#pragma unroll
          for (int k = 0; k < BK / kUmmaK; ++k)
            umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb != s.kb0 || k != 0) ? 1u : 0u);
          umma2_commit_mc(&empty_bar[stage]);          // frees this stage in both CTAs
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma2_commit_mc(&tfull_bar[as]);               // both CTAs' epilogues may drain their accumulator rows
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual TMA producer (per CTA, its own rows)
    if (lane == 0 && p.has_res) {
      int rq = 0;
      SegIter it(pair, npairs, units, p.num_kb, p.split);
      Seg s;
      while (it.next(s)) {
        if (s.kb0 != 0) continue;
        const int m_blk = s.tile / p.tiles_n, n_blk = s.tile - m_blk * p.tiles_n;
        const int m0 = m_blk * kTileM + (int)rank * kBlockM, n0 = n_blk * BN;
        if (m0 >= p.M) continue;
        const int ncn = min(kChunks, (p.N - n0) / kChunkN);
        for (int c = 0; c < ncn; ++c, ++rq) {
          const int rs = rq % p.res_bufs; const uint32_t rphase = (rq / p.res_bufs) & 1;
          ptx::mbar_wait(&rempty_bar[rs], rphase ^ 1, 500 + rs);
          ptx::mbar_arrive_expect_tx(&rfull_bar[rs], kChunkBytes);
          ptx::tma_load_2d(&tmR, &rfull_bar[rs], smem + res_off + rs * kChunkBytes, n0 + c * kChunkN, m0);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue groups (per CTA, its own 128 rows)
    const int g = (warp - kEpiWarp0) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool elected = ((warp - kEpiWarp0) & 3) == 0 && lane == 0;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t ob = smem_base + out_off + g * kChunkBytes;
    const uint32_t orow = ob + row * 128;
    const int bar_id = 1 + g;
    constexpr int kSlotF4 = kBlockM * BN / 4;        // float4s per CTA slot of the workspace
    float4* ws_mine = reinterpret_cast<float4*>(p.ws) + (size_t)blockIdx.x * kSlotF4;
    int n = 0, q = 0, rq = 0;
    SegIter it(pair, npairs, units, p.num_kb, p.split);
    Seg s;
    while (it.next(s)) {
      const int as = n & 1; const uint32_t aphase = (n >> 1) & 1;
      ++n;
      const int m_blk = s.tile / p.tiles_n, n_blk = s.tile - m_blk * p.tiles_n;
      const int m0 = m_blk * kTileM + (int)rank * kBlockM, n0 = n_blk * BN;
      const bool rows_valid = m0 < p.M;
      const int ncn = min(kChunks, (p.N - n0) / kChunkN);
      const int nchunks = rows_valid ? ncn : 0;
      const bool contributor = s.kb0 != 0;
      const bool gather = !contributor && s.kb1 < p.num_kb;
      ptx::mbar_wait(&tfull_bar[as], aphase, 400 + as);
      ptx::tc_fence_after();
      int mine_left = 0;
      for (int c = 0; c < nchunks; ++c) mine_left += (((q + c) & 1) == g);
      if (mine_left == 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
      }
      if (gather && nchunks > 0) {                   // wait for the same-rank CTA of every pair that holds the rest of this tile
        int need = p.num_kb - s.kb1, cc = pair + 1;
        while (need > 0) {
          const int a0 = (int)((int64_t)cc * units / npairs), a1 = (int)((int64_t)(cc + 1) * units / npairs);
          if (a1 > a0) {
            const uint32_t* fl = p.flags + 2 * (2 * cc + (int)rank);
            uint32_t spins = 0;
            while (ld_acquire(fl) != p.epoch || ld_acquire(fl + 1) != p.epoch) {
              if (++spins > (1u << 22)) {
                printf("airpose: stream-K (pair) flag timeout pair=%d rank=%u waits for %d\n", pair, rank, cc);
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
        const int slot_c = c * 16 * kBlockM + row;
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kAccCols + c * kChunkN;
        ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        ptx::tmem_ld_wait();
        if (--mine_left == 0) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
        }
        if (contributor) {
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
          int need = p.num_kb - s.kb1, cc = pair + 1;
          while (need > 0) {
            const int a0 = (int)((int64_t)cc * units / npairs), a1 = (int)((int64_t)(cc + 1) * units / npairs);
            if (a1 > a0) {
              const float4* src = reinterpret_cast<const float4*>(p.ws) + (size_t)(2 * cc + (int)rank) * kSlotF4 + slot_c;
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
          const int nb = n0 + c * kChunkN;
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
          ptx::tma_store_2d(&tmD, smem + out_off + g * kChunkBytes, n0 + c * kChunkN, m0);
          ptx::tma_store_commit();
        }
      }
      if (contributor) {                             // publish this group's part of the partial (also when it had no rows)
        __threadfence();
        ptx::named_bar_sync(bar_id, kGroupThreads);
        if (elected) st_release(p.flags + 2 * blockIdx.x + g, p.epoch);
      }
      q += nchunks;
      if (!contributor && p.has_res) rq += nchunks;
    }
    if (elected) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  cluster_sync_all();                                // the leader's MMAs read the peer's shared memory until the very end
  if (warp == 1) {
    __syncwarp();
    tmem2_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------- host side
struct SkWorkspace {
  float* ws = nullptr;
  uint32_t* flags = nullptr;
  uint32_t epoch = 0;
};
std::map<std::pair<int, cudaStream_t>, SkWorkspace> g_ws2;

int get_workspace2(cudaStream_t stream, SkWorkspace** out) {
  int dev = 0;
  AP_CHECK_CUDA(cudaGetDevice(&dev));
  SkWorkspace& w = g_ws2[std::make_pair(dev, stream)];
  if (!w.ws) {
    AP_CHECK_CUDA(cudaMalloc((void**)&w.ws, (size_t)kMaxGrid * kBlockM * 256 * sizeof(float)));
    AP_CHECK_CUDA(cudaMalloc((void**)&w.flags, (size_t)kMaxGrid * 2 * sizeof(uint32_t)));
    AP_CHECK_CUDA(cudaMemset(w.flags, 0, (size_t)kMaxGrid * 2 * sizeof(uint32_t)));
    AP_CHECK_CUDA(cudaDeviceSynchronize());
  }
  *out = &w;
  return 0;
}

}  // namespace

bool pair_kernel_eligible(const GemmLaunch& L) {
  return L.tma_epi && !L.stem && L.block_n == 256 && L.N % 256 == 0 && L.K % 64 == 0;
}

// B tensor map of a launch planned for the pair kernel: box = 128 weight rows (each CTA loads half of the 256)
int launch_gemm_sk2(const GemmLaunch& L, cudaStream_t stream) {
  AP_REQUIRE(pair_kernel_eligible(L), "launch_gemm_sk2: problem %dx%dx%d is not eligible for the pair kernel", L.M, L.N, L.K);
  AP_REQUIRE(L.pair_b_box, "launch_gemm_sk2: the B tensor map must be built with 128-row boxes");
  constexpr int BN = 256, BK = 64;
  KP kp{};
  kp.M = L.M; kp.N = L.N; kp.K = L.K;
  kp.num_kb = L.K / BK;
  kp.tiles_m = ceil_div(L.M, 2 * kBlockM);
  kp.tiles_n = L.N / BN;
  kp.im2col = L.im2col;
  kp.has_res = L.epi.residual != nullptr;
  kp.relu = L.epi.relu;
  kp.scale = L.epi.scale; kp.shift = L.epi.shift;
  if (L.im2col) {
    const ConvGeom& g = L.geom;
    AP_REQUIRE(g.Cin % BK == 0, "launch_gemm_sk2: im2col needs Cin %% 64 == 0 (Cin=%d)", g.Cin);
    AP_REQUIRE(L.K == g.ksize * g.ksize * g.Cin, "launch_gemm_sk2: K=%d does not match the conv geometry", L.K);
    kp.cblks = g.Cin / BK; kp.ksize = g.ksize; kp.stride = g.stride; kp.pad = g.pad;
    kp.Wo = g.Wo; kp.HoWo = g.Ho * g.Wo;
  }
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(gemm_sk2_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    configured = true;
  }
  constexpr int kStageBytes = kBlockM * BK * 2 + (BN / 2) * BK * 2;
  const int fixed = 1024 + kBarBytes + kEpiGroups * kChunkBytes;
  kp.res_bufs = kp.has_res ? (kp.num_kb <= 2 ? 4 : 2) : 0;
  const int avail = kSmemLimit - fixed - kp.res_bufs * kChunkBytes;
  kp.stages = std::min(kMaxStages, avail / kStageBytes);
  AP_REQUIRE(kp.stages >= 2, "launch_gemm_sk2: shared memory partition failed");
  const int smem_bytes = 1024 + kp.stages * kStageBytes + kEpiGroups * kChunkBytes + kp.res_bufs * kChunkBytes + kBarBytes;
  SkWorkspace* w = nullptr;
  if (get_workspace2(stream, &w)) return 1;
  kp.ws = w->ws; kp.flags = w->flags; kp.epoch = ++w->epoch;
  {
    // Under stream capture the epoch is frozen into the graph: a replay would find the flags of the previous replay (same
    // value) already set.  Clear them inside the graph before every stream-K launch, so each replay starts from zero flags.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    AP_CHECK_CUDA(cudaStreamIsCapturing(stream, &cs));
    if (cs != cudaStreamCaptureStatusNone) AP_CHECK_CUDA(cudaMemsetAsync(w->flags, 0, (size_t)kMaxGrid * 2 * sizeof(uint32_t), stream));
  }
  const int units = kp.tiles_m * kp.tiles_n * kp.num_kb;
  const int tiles = kp.tiles_m * kp.tiles_n;
  const int pairs_max = std::min(num_sms(), kMaxGrid) / 2;
  static const int min_kb = getenv("AIRPOSE_SK_SPLIT_MINKB") ? atoi(getenv("AIRPOSE_SK_SPLIT_MINKB")) : 8;
  kp.split = (kp.num_kb >= min_kb && tiles < 8 * pairs_max && tiles % pairs_max != 0) ? 1 : 0;
  const int pairs = kp.split ? std::min(pairs_max, std::max(std::min(tiles, pairs_max), units / 8)) : std::min(tiles, pairs_max);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = L.pdl ? 1 : 0;
  AP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_sk2_kernel<BN, BK>, L.tmA, L.tmB, L.tmD, L.epi.residual ? L.tmR : L.tmD, kp));
  count_launch();
  return 0;
}

}  // namespace airpose
