// SMPL-X vertex kernel of the hot path on tcgen05: pose-corrective blend shapes as a tensor-core
// contraction, shape blend + linear blend skinning + camera transform in its epilogue.
//
// Replaces lbs.py:179 (blend_shapes), :197-203 (pose offsets), :209-220 (skinning),
// body_models.py:980-982 (+transl) and utils/utils.py:237-239 (transform_smpl on the vertices) of
// /root/reference/copenet/src/copenet[/smplx/smplx] for the call pattern of copenet_twoview.py:281-292:
// 21 body rotations given, joints 22..54 identity, <= 10 betas, zero expression.
//
// The pose-corrective term  offsets[b, 3v+c] = sum_p feat[b,p] * posedirs[p, 3v+c]  (189 x 3 FMAs per
// vertex and mesh) is what bound the first kernel (2 % of the HBM roofline).  Here
//   D_c[v, b] = P_c[v, :] . F[b, :]^T      c = x,y,z;  M = 128 vertices, N = 32 meshes, K = 192 + 32
// runs on tcgen05 with P = fp16(posedirs * 2^10) RESIDENT in shared memory (168 KB per 128-vertex
// tile) and the feature F = R - I split into two fp16 terms (hi + lo, 22 significant bits) streamed
// through a TMA ring: 2 MMAs per k-step, fp32 accumulation in TMEM.  The only rounding beyond fp32 is
// the 11-bit mantissa of the stored posedirs: <= 1e-5 of the vertex scale (tests; north_star 1e-3).
// Round 2: the template and the SHAPE BLEND ride the same contraction as one more 32-wide k-block of
// split-fp16 products (smplx.cuh: T_hi T_lo | S_hi | S_hi | S_lo  against  1 1 | b_hi | b_lo | b_hi), so
// D * 2^-10 IS v_posed: 30 FMAs, 5 shared loads and 30 registers per vertex and mesh leave the epilogue,
// which is what lets it run on 16 warps instead of 8 (it was issue/latency-bound at 17 % warp occupancy).
//
// Everything else stays fp32 on the CUDA cores, in the epilogue, straight out of TMEM (TMEM lane =
// vertex, so per-vertex constants live in registers and a warp's stores of one mesh are contiguous):
//   v_posed = D * 2^-10
//   T = sum_k w_k A[b, joint_k]  over the vertex's non-zero skinning weights (A per mesh in smem)
//   v = T [v_posed; 1] + transl;  v_cam = R v + t
// On this call pattern A_j == A_ancestor(j) for every joint j >= 22 (identity local rotation:
// G_j = G_p [I | J_j - J_p]  =>  A_j = [R_p | t_p + R_p (J_j - J_p) - R_p J_j] = A_p), so only 22
// matrices per mesh are staged and the weights of folded joints are merged at create time.
//
// CTA = (128-vertex tile, range of mesh tiles), 19 warps:
//   warp 0   TMA: P once, then F k-blocks (2-stage ring)      warp 1   MMA issuer (TMEM double-buffered)
//   warp 2   bulk copies of per-mesh records (A, camera, transl; 8 meshes per copy, 4-stage ring)
//   warps 3-18  epilogue: lane quarter = warp % 4, mesh quarter (8 of the tile's 32 meshes) = (warp - 3) / 4
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"
#include "smplx.cuh"

namespace airpose {

namespace {

constexpr int kThreads = 608;
constexpr int kEpiWarp0 = 3;
constexpr int kEpiWarps = 16;
constexpr int kEpiGroups = kEpiWarps / 4;              // mesh quarters of a tile
constexpr int kPTileBytes = 128 * 128;                 // 128 vertices x 64 fp16
constexpr int kPxTileBytes = 128 * 64;                 // 128 vertices x 32 fp16 (template / shape k-block, 64-byte swizzle)
constexpr int kPxOff = 9 * kPTileBytes;
constexpr int kPBytes = 9 * kPTileBytes + 3 * kPxTileBytes;   // 3 coordinates x (3 pose k-blocks + the shape k-block)
constexpr int kFStages = 2;
constexpr int kFHalfBytes = kTcMeshTile * 128;         // 32 meshes x 64 fp16
constexpr int kFStageBytes = 2 * kFHalfBytes;          // hi + lo
constexpr int kFxBytes = kTcMeshTile * 64;             // shape k-block: 32 meshes x 32 fp16, hi only
constexpr int kRStages = 4;
static_assert(kTcMeshTile == kEpiGroups * kTcSub, "one record sub-batch per epilogue group and tile");
constexpr int kRecBytes = kTcRecFloats * 4;
constexpr int kRStageBytes = kTcSub * kRecBytes;
constexpr int kFOff = kPBytes;
constexpr int kROff = kFOff + kFStages * kFStageBytes;
constexpr int kBarOff = kROff + kRStages * kRStageBytes;
constexpr int kNumBars = 1 + 2 * kFStages + 4 + 2 * kRStages;
constexpr int kSmemBytes = 1024 + kBarOff + kNumBars * 8 + 16;
constexpr int kTmemCols = 512;                         // 2 buffers x 3 coordinates x (32 hi + 32 lo) columns = 384
constexpr int kAccCols = 6 * kTcMeshTile;              // columns of one accumulator buffer
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
static_assert(kRStageBytes % 16 == 0 && kROff % 16 == 0, "bulk copies need 16-byte alignment");

struct KArgs {
  int V, B, nb, NS, KW, has_transl;
  int tiles_total, tiles_per_cta;
  const int* sk_off;
  const float* sk_w;
  const int* sk_cnt;
  const float* rec;
  float* out;
  float* out_cam;
  int vrows;                 // rows per coordinate plane of P
};

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32xN(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32xN(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

// packed fp32 FMA (sm_100: SASS FFMA2): two independent fp32 FMAs per issue slot -- here the two meshes of a pair
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }
// explicit shared-window loads (the hand-aligned dynamic smem base hides the address space: generic LD.E otherwise)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// kind::f16 instruction descriptor with fp16 A/B (format 0), fp32 accumulator.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <bool kHasCam, int UN>
__global__ void __launch_bounds__(kThreads, 1)
smplx_vertex_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmPx,
                       const __grid_constant__ CUtensorMap tmFh, const __grid_constant__ CUtensorMap tmFl,
                       const __grid_constant__ CUtensorMap tmFx, const KArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* p_full = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* f_full = p_full + 1;
  uint64_t* f_empty = f_full + kFStages;
  uint64_t* tfull = f_empty + kFStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* r_full = tempty + 2;
  uint64_t* r_empty = r_full + kRStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_empty + kRStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vtile = blockIdx.x;
  const int t0 = blockIdx.y * a.tiles_per_cta;
  const int t1 = min(a.tiles_total, t0 + a.tiles_per_cta);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmP); ptx::prefetch_tmap(&tmPx);
    ptx::prefetch_tmap(&tmFh); ptx::prefetch_tmap(&tmFl); ptx::prefetch_tmap(&tmFx);
    ptx::mbar_init(p_full, 1);
    for (int s = 0; s < kFStages; ++s) { ptx::mbar_init(&f_full[s], 1); ptx::mbar_init(&f_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], kEpiWarps); }
    for (int s = 0; s < kRStages; ++s) { ptx::mbar_init(&r_full[s], 1); ptx::mbar_init(&r_empty[s], 4); }
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
  // everything above overlapped the previous kernel (smplx_pose_kernel); its output (the F operand, the per-mesh records) is read below
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ P (once) + F ring
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(p_full, kPBytes);
      for (int c = 0; c < 3; ++c) {
        for (int kb = 0; kb < 3; ++kb)
          ptx::tma_load_2d(&tmP, p_full, smem + (c * 3 + kb) * kPTileBytes, kb * 64, c * a.vrows + vtile * 128);
        ptx::tma_load_2d(&tmPx, p_full, smem + kPxOff + c * kPxTileBytes, kTcKPose, c * a.vrows + vtile * 128);
      }
      int stage = 0; uint32_t phase = 0;
      for (int t = t0; t < t1; ++t)
        for (int kb = 0; kb < 4; ++kb) {
          ptx::mbar_wait(&f_empty[stage], phase ^ 1, 100 + stage);
          uint8_t* fs = smem + kFOff + stage * kFStageBytes;
          if (kb < 3) {
            ptx::mbar_arrive_expect_tx(&f_full[stage], kFStageBytes);
            ptx::tma_load_2d(&tmFh, &f_full[stage], fs, kb * 64, t * kTcMeshTile);
            ptx::tma_load_2d(&tmFl, &f_full[stage], fs + kFHalfBytes, kb * 64, t * kTcMeshTile);
          } else {                                      // template / shape block: hi only, 64-byte rows
            ptx::mbar_arrive_expect_tx(&f_full[stage], kFxBytes);
            ptx::tma_load_2d(&tmFx, &f_full[stage], fs, kTcKPose, t * kTcMeshTile);
          }
          if (++stage == kFStages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      // F_hi | F_lo of a k-block are 64 consecutive K-major rows of the stage: ONE N = 64 MMA per k-step and coordinate writes the
      // hi products into columns 0..31 and the lo products into 32..63 (the epilogue adds them) -- the 128-row P operand is read
      // from shared memory once per k-step instead of twice (UMMA operand reads were 39 % of the shared-memory data pipe)
      constexpr uint32_t idesc = make_idesc_f16(128, 2 * kTcMeshTile);
      constexpr uint32_t idesc_x = make_idesc_f16(128, kTcMeshTile);
      ptx::mbar_wait(p_full, 0, 200);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty[as], aphase ^ 1, 210 + as);
        ptx::tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) {
          ptx::mbar_wait(&f_full[stage], phase, 220 + stage);
          ptx::tc_fence_after();
          const uint32_t fs = ptx::smem_u32(smem + kFOff + stage * kFStageBytes);
          if (kb < 3) {
            const uint64_t bhl = ptx::make_kmajor_sw128_desc(fs);          // rows 0..31 = hi, 32..63 = lo
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const uint64_t ad = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem + (c * 3 + kb) * kPTileBytes));
              const uint32_t d_tmem = tmem_base + as * kAccCols + c * 2 * kTcMeshTile;
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_bf16(d_tmem, ad + 2 * k, bhl + 2 * k, idesc, (kb | k) != 0);
            }
          } else {                                      // K = 32: template + shape blend (64-byte swizzled rows)
            const uint64_t bx = ptx::make_kmajor_desc(fs, 64);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const uint64_t ad = ptx::make_kmajor_desc(ptx::smem_u32(smem + kPxOff + c * kPxTileBytes), 64);
              const uint32_t d_tmem = tmem_base + as * kAccCols + c * 2 * kTcMeshTile;      // into the hi columns
#pragma unroll
              for (int k = 0; k < 2; ++k) ptx::umma_bf16(d_tmem, ad + 2 * k, bx + 2 * k, idesc_x, 1);
            }
          }
          ptx::umma_commit(&f_empty[stage]);
          if (++stage == kFStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull[as]);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ per-mesh records: 16 meshes per bulk copy
    if (lane == 0) {
      int s = 0;
      for (int t = t0; t < t1; ++t)
        for (int hf = 0; hf < kEpiGroups; ++hf, ++s) {
          const int slot = s % kRStages; const uint32_t ph = (s / kRStages) & 1;
          ptx::mbar_wait(&r_empty[slot], ph ^ 1, 300 + slot);
          ptx::mbar_arrive_expect_tx(&r_full[slot], kRStageBytes);
          bulk_load(smem + kROff + slot * kRStageBytes, a.rec + (size_t)(t * kTcMeshTile + hf * kTcSub) * kTcRecFloats,
                    kRStageBytes, &r_full[slot]);
        }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;                        // TMEM lane quarter this warp may read
    const int hf = (warp - kEpiWarp0) >> 2;           // which 8 meshes of the tile
    const int v = vtile * 128 + quad * 32 + lane;
    const bool valid = v < a.V;
    const int vc = valid ? v : a.V - 1;
    // per-vertex constants
    const int cnt = __ldg(a.sk_cnt + vc);
    int joff[kTcMaxKW]; float jw[kTcMaxKW];
#pragma unroll
    for (int k = 0; k < kTcMaxKW; ++k) {
      const bool on = k < cnt;
      joff[k] = on ? __ldg(a.sk_off + (size_t)k * a.V + vc) : 0;
      jw[k] = on ? __ldg(a.sk_w + (size_t)k * a.V + vc) : 0.f;
    }
    const int kw_warp = __reduce_max_sync(0xffffffffu, cnt);     // this warp's 32 vertices need at most this many joints
    const float inv_scale = 1.f / kTcPScale;
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
      const int s = kEpiGroups * it + hf;
      const int slot = s % kRStages; const uint32_t rph = (s / kRStages) & 1;
      ptx::mbar_wait(&tfull[as], aphase, 400 + as);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kAccCols + hf * kTcSub;
      ptx::mbar_wait(&r_full[slot], rph, 410 + slot);
      const uint32_t rbase = ptx::smem_u32(smem + kROff + slot * kRStageBytes);
      const int mesh0 = t * kTcMeshTile + hf * kTcSub;
      // UN (2 or 4) meshes per TMEM load and per unrolled body: the fully unrolled 16-mesh body was ~5700 SASS instructions
      // (91 KB) and spent 12 % of its samples in instruction-fetch stalls (profiles/r01c: stall_no_inst)
#pragma unroll 1
      for (int mq = 0; mq < kTcSub / UN; ++mq) {
        uint32_t dx[UN], dy[UN], dz[UN];
        tmem_ld_32xN(taddr + mq * UN, dx);
        tmem_ld_32xN(taddr + 2 * kTcMeshTile + mq * UN, dy);
        tmem_ld_32xN(taddr + 4 * kTcMeshTile + mq * UN, dz);
        uint32_t lx[UN], ly[UN], lz[UN];                  // the lo products
        tmem_ld_32xN(taddr + kTcMeshTile + mq * UN, lx);
        tmem_ld_32xN(taddr + 3 * kTcMeshTile + mq * UN, ly);
        tmem_ld_32xN(taddr + 5 * kTcMeshTile + mq * UN, lz);
        ptx::tmem_ld_wait();
        if (mq == kTcSub / UN - 1) {                      // accumulator is in registers: MMA may refill it
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty[as]);
        }
        if (mesh0 + mq * UN >= a.B) continue;            // uniform across the CTA (the release above still happens)
#pragma unroll
        for (int pi = 0; pi < UN / 2; ++pi) {            // two meshes per pass: every fp32 FMA below is an FFMA2
          const int m = mq * UN + 2 * pi;
          const int b = mesh0 + m;
          if (b >= a.B) break;                           // uniform across the CTA
          const bool second = b + 1 < a.B;
          // pair record: field f of meshes (b, b+1) = one float2 at index f
          const uint32_t rec = rbase + (uint32_t)(m >> 1) * 2 * kRecBytes;      // byte address; field f of the pair at rec + 8 f
          // v_posed = v_template + shapedirs . beta + pose offsets (lbs.py:179, :203): all of it is the accumulator
          const float2 x = make_float2((__uint_as_float(dx[2 * pi]) + __uint_as_float(lx[2 * pi])) * inv_scale,
                                       (__uint_as_float(dx[2 * pi + 1]) + __uint_as_float(lx[2 * pi + 1])) * inv_scale);
          const float2 y = make_float2((__uint_as_float(dy[2 * pi]) + __uint_as_float(ly[2 * pi])) * inv_scale,
                                       (__uint_as_float(dy[2 * pi + 1]) + __uint_as_float(ly[2 * pi + 1])) * inv_scale);
          const float2 z = make_float2((__uint_as_float(dz[2 * pi]) + __uint_as_float(lz[2 * pi])) * inv_scale,
                                       (__uint_as_float(dz[2 * pi + 1]) + __uint_as_float(lz[2 * pi + 1])) * inv_scale);
          // T = sum_k w_k A_k (lbs.py:209-213)
          float2 T[12];
#pragma unroll
          for (int e = 0; e < 12; ++e) T[e] = make_float2(0.f, 0.f);
#pragma unroll
          for (int k = 0; k < kTcMaxKW; ++k) {
            if (k >= kw_warp) break;                     // warp-uniform
            const float2 w = dup2(jw[k]);
            const uint32_t Aj = rec + 2 * joff[k];
#pragma unroll
            for (int q = 0; q < 6; ++q) {                // 12 matrix entries x 2 meshes = 6 x 16 bytes
              const float4 r = lds_f4(Aj + 16 * q);
              T[2 * q] = ffma2(w, make_float2(r.x, r.y), T[2 * q]);
              T[2 * q + 1] = ffma2(w, make_float2(r.z, r.w), T[2 * q + 1]);
            }
          }
          // v = T [v_posed; 1] (lbs.py:215-220) + transl (body_models.py:980-982)
          float2 ox = ffma2(T[0], x, ffma2(T[1], y, ffma2(T[2], z, T[3])));
          float2 oy = ffma2(T[4], x, ffma2(T[5], y, ffma2(T[6], z, T[7])));
          float2 oz = ffma2(T[8], x, ffma2(T[9], y, ffma2(T[10], z, T[11])));
          if (a.has_transl) {
            const float2 tx = lds_f2(rec + 8 * kTcRecTransl), ty = lds_f2(rec + 8 * (kTcRecTransl + 1)), tz = lds_f2(rec + 8 * (kTcRecTransl + 2));
            ox.x += tx.x; ox.y += tx.y; oy.x += ty.x; oy.y += ty.y; oz.x += tz.x; oz.y += tz.y;
          }
          if (valid) {
            float* o = a.out + ((size_t)b * a.V + v) * 3;
            o[0] = ox.x; o[1] = oy.x; o[2] = oz.x;
            if (second) {
              float* o1 = o + (size_t)a.V * 3;
              o1[0] = ox.y; o1[1] = oy.y; o1[2] = oz.y;
            }
            if (kHasCam) {         // transform_smpl (utils.py:237-239): R v + t about the origin; record: R row-major (9) then t (3)
              float2 cm[12];
#pragma unroll
              for (int q = 0; q < 6; ++q) {
                const float4 r = lds_f4(rec + 8 * kTcRecCam + 16 * q);
                cm[2 * q] = make_float2(r.x, r.y); cm[2 * q + 1] = make_float2(r.z, r.w);
              }
              const float2 cx = ffma2(cm[0], ox, ffma2(cm[1], oy, ffma2(cm[2], oz, cm[9])));
              const float2 cy = ffma2(cm[3], ox, ffma2(cm[4], oy, ffma2(cm[5], oz, cm[10])));
              const float2 cz = ffma2(cm[6], ox, ffma2(cm[7], oy, ffma2(cm[8], oz, cm[11])));
              float* oc = a.out_cam + ((size_t)b * a.V + v) * 3;
              oc[0] = cx.x; oc[1] = cy.x; oc[2] = cz.x;
              if (second) {
                float* oc1 = oc + (size_t)a.V * 3;
                oc1[0] = cx.y; oc1[1] = cy.y; oc1[2] = cz.y;
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&r_empty[slot]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int smplx_tc_create(const airpose_smplx_model_host* mh, const SmplxDev& d, SmplxTc* tc, std::vector<void*>* owned) {
  tc->ok = false;
  const int V = d.V, J = d.J;
  if (J <= kTcBodyJoints || d.P != (J - 1) * 9 || getenv("AIRPOSE_SMPLX_GENERIC")) return 0;
  // fold joints >= 22 onto their nearest body ancestor and merge the weights
  std::vector<int> anc(J);
  for (int j = 0; j < J; ++j) anc[j] = j < kTcBodyJoints ? j : anc[(int)mh->parents[j]];
  std::vector<int> cnt(V, 0);
  std::vector<std::vector<std::pair<int, float>>> rows(V);
  int KW = 1;
  for (int v = 0; v < V; ++v) {
    for (int j = 0; j < J; ++j) {
      const float w = mh->lbs_weights[(size_t)v * J + j];
      if (w == 0.f) continue;
      bool merged = false;
      for (auto& e : rows[v])
        if (e.first == anc[j]) { e.second += w; merged = true; break; }
      if (!merged) rows[v].push_back({anc[j], w});
    }
    cnt[v] = (int)rows[v].size();
    KW = std::max(KW, cnt[v]);
  }
  if (KW > kTcMaxKW) return 0;          // unusually dense skinning: the generic kernel handles it
  // Slot assignment per WARP of the vertex kernel (32 consecutive vertices): when the union of the joints its vertices use fits
  // the slot count, slot k means the SAME joint for all 32 lanes (weight 0 where a vertex does not use it), so every shared-
  // memory load of A_j in the epilogue is a warp-uniform broadcast (1 wavefront instead of 4+: the kernel was bound by the
  // shared-memory data pipe, profiles/r02y).  Consecutive vertices share their joints, so the union is barely larger than the
  // densest row (synthetic model: 2.7 vs 2.5 slots on average).  Groups whose union does not fit keep per-lane slots.
  for (int g0 = 0; g0 < V; g0 += 32) {
    const int g1 = std::min(V, g0 + 32);
    std::vector<int> uni;
    for (int v = g0; v < g1; ++v)
      for (auto& e : rows[v])
        if (std::find(uni.begin(), uni.end(), e.first) == uni.end()) uni.push_back(e.first);
    if ((int)uni.size() > kTcMaxKW) continue;
    std::sort(uni.begin(), uni.end());
    for (int v = g0; v < g1; ++v) {
      std::vector<std::pair<int, float>> r;
      for (int j : uni) {
        float w = 0.f;
        for (auto& e : rows[v])
          if (e.first == j) w = e.second;
        r.push_back({j, w});
      }
      rows[v] = r;
      cnt[v] = (int)r.size();
    }
    KW = std::max(KW, (int)uni.size());
  }
  std::vector<int> off((size_t)KW * V, 0);
  std::vector<float> w((size_t)KW * V, 0.f);
  for (int v = 0; v < V; ++v)
    for (int k = 0; k < cnt[v]; ++k) { off[(size_t)k * V + v] = rows[v][k].first * 48; w[(size_t)k * V + v] = rows[v][k].second; }
  // P[c][v][p] = fp16(posedirs[p][3v+c] * 2^10), p < 189; zero padded to 192 columns / vtiles*128 rows; then the
  // template / shape k-block (smplx.cuh): T_hi T_lo | S_hi[10] | S_hi[10] | S_lo[10], all x 2^10
  const int vtiles = ceil_div(V, 128), vrows = vtiles * 128;
  std::vector<__half> P((size_t)3 * vrows * kTcK, __float2half_rn(0.f));
  for (int p = 0; p < 189; ++p) {
    const float* src = mh->posedirs + (size_t)p * V * 3;
    for (int v = 0; v < V; ++v)
      for (int c = 0; c < 3; ++c) P[((size_t)c * vrows + v) * kTcK + p] = __float2half_rn(src[(size_t)v * 3 + c] * kTcPScale);
  }
  auto split = [](float x, __half* hi, __half* lo) {
    *hi = __float2half_rn(x);
    *lo = __float2half_rn(x - __half2float(*hi));
  };
  for (int v = 0; v < V; ++v)
    for (int c = 0; c < 3; ++c) {
      __half* row = &P[((size_t)c * vrows + v) * kTcK + kTcKPose];
      split(mh->v_template[(size_t)v * 3 + c] * kTcPScale, &row[0], &row[1]);
      for (int l = 0; l < 10 && l < d.NS; ++l) {
        __half hi, lo;
        split(mh->shapedirs[((size_t)v * 3 + c) * d.NS + l] * kTcPScale, &hi, &lo);
        row[2 + l] = hi; row[12 + l] = hi; row[22 + l] = lo;
      }
    }
  __half* dP; int* doff; float* dw; int* dcnt;
  if (device_upload(&dP, P.data(), P.size())) return 1;
  owned->push_back(dP);
  if (device_upload(&doff, off.data(), off.size())) return 1;
  owned->push_back(doff);
  if (device_upload(&dw, w.data(), w.size())) return 1;
  owned->push_back(dw);
  if (device_upload(&dcnt, cnt.data(), cnt.size())) return 1;
  owned->push_back(dcnt);
  tc->P = dP; tc->sk_off = doff; tc->sk_w = dw; tc->sk_cnt = dcnt;
  tc->KW = KW; tc->vtiles = vtiles;
  if (make_tmap_tiled_bf16(&tc->tmP, dP, (int64_t)3 * vrows, kTcK, kTcK, 128, 64)) return 1;   // 2-byte elements: fp16 rides the bf16 map
  if (make_tmap_tiled_bf16(&tc->tmPx, dP, (int64_t)3 * vrows, kTcK, kTcK, 128, kTcKShape, 64)) return 1;
  tc->ok = true;
  return 0;
}

// Mesh tiles per CTA: enough CTAs to fill the SMs, few enough that the 168 KB of P per CTA amortise.
static int pick_tiles_per_cta(int vtiles, int tiles_total) {
  const int sms = num_sms();
  double best = 1e30; int best_s = 1;
  for (int s = 1; s <= std::min(tiles_total, 32); ++s) {
    const int per = ceil_div(tiles_total, s);
    const int waves = ceil_div(vtiles * s, sms);
    const double cost = waves * (1.5 + per);         // P load ~ 1.5 tile times
    if (cost < best) { best = cost; best_s = s; }
  }
  return ceil_div(tiles_total, best_s);
}

int smplx_tc_forward(const SmplxDev& d, const SmplxTc& tc, const TcCall& c, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(smplx_vertex_tc_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    AP_CHECK_CUDA(cudaFuncSetAttribute(smplx_vertex_tc_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    AP_CHECK_CUDA(cudaFuncSetAttribute(smplx_vertex_tc_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    AP_CHECK_CUDA(cudaFuncSetAttribute(smplx_vertex_tc_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  CUtensorMap tmFh, tmFl, tmFx;
  if (make_tmap_tiled_bf16(&tmFh, c.fh, c.B, kTcK, kTcK, kTcMeshTile, 64)) return 1;
  if (make_tmap_tiled_bf16(&tmFl, c.fl, c.B, kTcK, kTcK, kTcMeshTile, 64)) return 1;
  if (make_tmap_tiled_bf16(&tmFx, c.fh, c.B, kTcK, kTcK, kTcMeshTile, kTcKShape, 64)) return 1;
  KArgs a{};
  a.V = d.V; a.B = c.B; a.nb = c.nb; a.NS = d.NS; a.KW = tc.KW; a.has_transl = c.has_transl;
  a.tiles_total = ceil_div(c.B, kTcMeshTile);
  a.tiles_per_cta = pick_tiles_per_cta(tc.vtiles, a.tiles_total);
  a.sk_off = tc.sk_off; a.sk_w = tc.sk_w; a.sk_cnt = tc.sk_cnt;
  a.rec = c.rec; a.out = c.out; a.out_cam = c.out_cam;
  a.vrows = tc.vtiles * 128;
  dim3 grid(tc.vtiles, ceil_div(a.tiles_total, a.tiles_per_cta));
  static const int un = getenv("AIRPOSE_SMPLX_UNROLL") ? atoi(getenv("AIRPOSE_SMPLX_UNROLL")) : 2;   // 4 spills at 96 registers (0.77 vs 0.64 ms)
  if (un == 2) {
    if (c.out_cam) AP_CHECK_CUDA(launch_chain_smem(smplx_vertex_tc_kernel<true, 2>, grid, dim3(kThreads), kSmemBytes, stream, tc.tmP, tc.tmPx, tmFh, tmFl, tmFx, a));
    else AP_CHECK_CUDA(launch_chain_smem(smplx_vertex_tc_kernel<false, 2>, grid, dim3(kThreads), kSmemBytes, stream, tc.tmP, tc.tmPx, tmFh, tmFl, tmFx, a));
  } else {
    if (c.out_cam) AP_CHECK_CUDA(launch_chain_smem(smplx_vertex_tc_kernel<true, 4>, grid, dim3(kThreads), kSmemBytes, stream, tc.tmP, tc.tmPx, tmFh, tmFl, tmFx, a));
    else AP_CHECK_CUDA(launch_chain_smem(smplx_vertex_tc_kernel<false, 4>, grid, dim3(kThreads), kSmemBytes, stream, tc.tmP, tc.tmPx, tmFh, tmFl, tmFx, a));
  }
  AP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace airpose
