// SMPL-X vertex kernel for LARGE batches ("mesh-lane" layout): the same math as smplx_tc.cu -- lbs.py:179 (blend_shapes),
// :197-203 (pose offsets), :209-220 (skinning), body_models.py:980-982 (+transl), utils/utils.py:237-239 (transform_smpl) of
// /root/reference/copenet/src/copenet[/smplx/smplx], call pattern of copenet_twoview.py:281-292 -- with the tcgen05 product
// TRANSPOSED:
//     D^T[b, (c, v)] = F[b, :] . P_c[v, :]^T       M = 128 meshes (TMEM lanes), N = 3 coordinates x 32 vertices, K = 192 + 32
// Why (profiles/r02y, r02z): with TMEM lane = vertex every thread needs, per vertex and mesh, the 12 floats of each skinning
// matrix A[b][j] out of shared memory -- 170 B per vertex.mesh through the LSU, most of it as warp-uniform broadcasts that use
// an eighth of a wavefront -- and the N = 32 MMAs re-read the 128-row P operand for every 32 meshes; the two together saturate
// the shared-memory data pipe (41 % + 39 %) at 22 % of the HBM roofline.  With TMEM lane = MESH:
//   * a thread's mesh is fixed for the life of the CTA and the vertex is warp-uniform, so the skinning matrices live in a
//     REGISTER cache of 4 slots x 12 floats per thread; a slot is reloaded (3 x 16-byte loads of the thread's own record, L2-
//     resident) only when the vertex stream moves to another joint -- consecutive vertices share their joints, the host
//     orders the slots to keep them (smplx_ml_create);
//   * skinning weights and slot ids are warp-uniform: two broadcast loads per vertex for all 32 meshes;
//   * F (hi + lo, 104 KB) is resident, P streams through a 2-stage TMA ring as 32-vertex tiles: 26 MMAs of 128 x 96 x 16
//     per 4096 vertex.mesh instead of 78 of 128 x 32 x 16;
//   * results go through a per-warp 32 x 12 transposition buffer so that global stores are 48-byte runs per mesh.
// Template, shape blend and pose-corrective offsets all come out of the contraction (split-fp16 products, smplx.cuh).
//
// CTA = (128-mesh tile, range of 32-vertex tiles), 14 warps:
//   warp 0  TMA: F once, then P tiles + the tile's slot table      warp 1  MMA issuer, TMEM owner
//   warps 2-13  epilogue: mesh quarter = warp % 4 (TMEM lanes), vertex pairs s, s + 3, ... of the tile with s = (warp - 2) / 4
//               (12 warps, not 16: 448 threads leave 128 registers per thread, which the 48-register matrix cache needs)
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"
#include "smplx.cuh"

namespace airpose {

namespace {

constexpr int kThreads = 448;
constexpr int kEpiWarp0 = 2;
constexpr int kEpiWarps = 12;                         // 3 per TMEM lane quarter: 128 registers per thread for the matrix cache
constexpr int kVT = kMlVertTile;                       // 32 vertices per tile
constexpr int kNCols = 3 * kVT;                        // 96 accumulator columns
constexpr int kFBlk = 128 * 128;                       // 128 meshes x 64 fp16
constexpr int kFxBlk = 128 * 64;                       // 128 meshes x 32 fp16 (template / shape block, hi only)
constexpr int kFhOff = 0;
constexpr int kFxOff = 3 * kFBlk;                      // 49152
constexpr int kFlOff = kFxOff + kFxBlk;                // 57344
constexpr int kFBytes = kFlOff + 3 * kFBlk;            // 106496
constexpr int kPBlk = kNCols * 128;                    // 96 rows x 64 fp16
constexpr int kPxBlk = kNCols * 64;                    // 96 rows x 32 fp16
constexpr int kPxOff = 3 * kPBlk;                      // 36864
constexpr int kTabOff = kPxOff + kPxBlk;               // 43008
constexpr int kTabBytes = kVT * 32;                    // per vertex: 4 slot ids + 4 weights
constexpr int kStageBytes = kTabOff + kTabBytes;       // 44032
constexpr int kStages = 2;
constexpr int kPOff = kFBytes;
constexpr int kTrWords = 32 * 7;                       // per-warp transposition buffer: 32 meshes x (2 vertices x 3 + 1 pad) floats
constexpr int kTrOff = kPOff + kStages * kStageBytes;  // 194560
constexpr int kBarOff = kTrOff + kEpiWarps * kTrWords * 4;
constexpr int kNumBars = 1 + 2 * kStages + 4;
constexpr int kSmemBytes = 1024 + kBarOff + kNumBars * 8 + 16;
constexpr int kTmemCols = 256;                         // 2 accumulators x 96 columns
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
static_assert(kStageBytes % 1024 == 0 && kFBytes % 1024 == 0 && kPBlk % 1024 == 0 && kPxBlk % 512 == 0, "swizzle atom alignment");

struct MlArgs {
  int V, B, has_transl;
  int vtiles, tiles_per_cta;           // 32-vertex tiles
  int vrows;                           // rows per coordinate plane of P
  const float* rec;                    // [Bpad][kTcRecFloats], plain (not pair-interleaved)
  const uint32_t* vtab;                // [vtiles * 32][8]
  float* out;
  float* out_cam;
};

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32xN(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {      // kind::f16, fp16 A/B, fp32 accumulator
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int OFF>
__device__ __forceinline__ float lds_f32_off(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1 + %2];" : "=f"(v) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// The warp's 32 x 6 transposition buffer (row stride 7 words) to global memory: lane (< 30) = (row r5 of five, column of six),
// pass i = 0..6 covers rows 5 i + r5.  `src` = this lane's element in pass 0, `dst` = its global element, five_rows = the
// global distance of five mesh rows; rows_left = valid rows from row r5 on (slow path only).
__device__ __forceinline__ void copy_out(uint32_t src, float* dst, size_t five_rows, bool active, bool tail_ok, bool fast, int rows_left,
                                         bool col_ok) {
  if (fast) {
    if (active) {
      const float v0 = lds_f32_off<0>(src), v1 = lds_f32_off<140>(src), v2 = lds_f32_off<280>(src), v3 = lds_f32_off<420>(src);
      dst[0] = v0; dst[five_rows] = v1; dst[2 * five_rows] = v2; dst[3 * five_rows] = v3;
      const float v4 = lds_f32_off<560>(src), v5 = lds_f32_off<700>(src);
      dst[4 * five_rows] = v4; dst[5 * five_rows] = v5;
      if (tail_ok) dst[6 * five_rows] = lds_f32_off<840>(src);         // rows 30, 31
    }
  } else if (active && col_ok) {
#pragma unroll 1
    for (int i = 0; i < 7; ++i) {
      if (5 * i < rows_left && (i < 6 || tail_ok)) dst[(size_t)i * five_rows] = lds_f32(src + i * 140);
    }
  }
}

template <bool kHasCam>
__global__ void __launch_bounds__(kThreads, 1)
smplx_vertex_ml_kernel(const __grid_constant__ CUtensorMap tmFh, const __grid_constant__ CUtensorMap tmFl,
                       const __grid_constant__ CUtensorMap tmFx, const __grid_constant__ CUtensorMap tmP,
                       const __grid_constant__ CUtensorMap tmPx, const MlArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* f_full = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* p_full = f_full + 1;
  uint64_t* p_empty = p_full + kStages;
  uint64_t* tfull = p_empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x;
  const int t0 = blockIdx.y * a.tiles_per_cta;
  const int t1 = min(a.vtiles, t0 + a.tiles_per_cta);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmFh); ptx::prefetch_tmap(&tmFl); ptx::prefetch_tmap(&tmFx);
    ptx::prefetch_tmap(&tmP); ptx::prefetch_tmap(&tmPx);
    ptx::mbar_init(f_full, 1);
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&p_full[s], 1); ptx::mbar_init(&p_empty[s], 1 + kEpiWarps); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], kEpiWarps); }
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

  if (warp == 0) {
    // ------------------------------------------------------------------ F (once) + P ring
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(f_full, kFBytes);
      for (int kb = 0; kb < 3; ++kb) {
        ptx::tma_load_2d(&tmFh, f_full, smem + kFhOff + kb * kFBlk, kb * 64, mt * 128);
        ptx::tma_load_2d(&tmFl, f_full, smem + kFlOff + kb * kFBlk, kb * 64, mt * 128);
      }
      ptx::tma_load_2d(&tmFx, f_full, smem + kFxOff, kTcKPose, mt * 128);
      int it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&p_empty[st], ph ^ 1, 100 + st);
        uint8_t* ps = smem + kPOff + st * kStageBytes;
        ptx::mbar_arrive_expect_tx(&p_full[st], kStageBytes);
        for (int c = 0; c < 3; ++c) {
          for (int kb = 0; kb < 3; ++kb)
            ptx::tma_load_2d(&tmP, &p_full[st], ps + kb * kPBlk + c * (kVT * 128), kb * 64, c * a.vrows + t * kVT);
          ptx::tma_load_2d(&tmPx, &p_full[st], ps + kPxOff + c * (kVT * 64), kTcKPose, c * a.vrows + t * kVT);
        }
        bulk_load(ps + kTabOff, a.vtab + (size_t)t * kVT * 8, kTabBytes, &p_full[st]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, kNCols);
      ptx::mbar_wait(f_full, 0, 200);
      const uint32_t fbase = ptx::smem_u32(smem);
      int it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&tempty[st], ph ^ 1, 210 + st);
        ptx::mbar_wait(&p_full[st], ph, 220 + st);
        ptx::tc_fence_after();
        const uint32_t ps = ptx::smem_u32(smem + kPOff + st * kStageBytes);
        const uint32_t d_tmem = tmem_base + st * kNCols;
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t ah = ptx::make_kmajor_sw128_desc(fbase + kFhOff + kb * kFBlk);
          const uint64_t al = ptx::make_kmajor_sw128_desc(fbase + kFlOff + kb * kFBlk);
          const uint64_t bd = ptx::make_kmajor_sw128_desc(ps + kb * kPBlk);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ptx::umma_bf16(d_tmem, ah + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
            ptx::umma_bf16(d_tmem, al + 2 * k, bd + 2 * k, idesc, 1);
          }
        }
        {                                               // K = 32: template + shape blend (64-byte swizzled rows, hi only)
          const uint64_t ax = ptx::make_kmajor_desc(fbase + kFxOff, 64);
          const uint64_t bx = ptx::make_kmajor_desc(ps + kPxOff, 64);
#pragma unroll
          for (int k = 0; k < 2; ++k) ptx::umma_bf16(d_tmem, ax + 2 * k, bx + 2 * k, idesc, 1);
        }
        ptx::umma_commit(&p_empty[st]);
        ptx::umma_commit(&tfull[st]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;                        // TMEM lane quarter this warp may read = which 32 meshes
    const int sub = (warp - kEpiWarp0) >> 2;          // 0..2: half-chunks (vertex pairs) sub, sub + 3, ... of every tile
    const int mesh0 = mt * 128 + quad * 32;
    const float* recb = a.rec + (size_t)min(mesh0 + lane, a.B - 1) * kTcRecFloats;
    const uint32_t tr = ptx::smem_u32(smem + kTrOff) + (uint32_t)((warp - kEpiWarp0) * kTrWords * 4);
    const uint32_t tr_w = tr + (uint32_t)lane * 7 * 4;                      // this mesh's row of the transposition buffer
    // copy-out pattern: lanes 0..29 = 5 mesh rows x 6 columns per pass, 7 passes (rows 5 i + r5)
    const int r5 = (lane * 43) >> 8, col = lane - 6 * r5;                   // lane / 6, lane % 6
    const uint32_t tr_r = tr + (uint32_t)(r5 * 7 + col) * 4;
    const int row_stride = a.V * 3;
    const int lane_off = (mesh0 + r5) * row_stride + col;                   // < 2^31 elements (checked on the host)
    const bool rows_full = mesh0 + 32 <= a.B;
    // register cache of skinning matrices: slot k = A (row-major 3 x 4) as float2 pairs; the rotation part is pre-multiplied
    // by 2^-10 (the accumulator's scale, exact) and transl is folded into the translation column (the weights of a vertex
    // sum to one, lbs.py:209) at reload time
    float2 cache[4][6];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int q = 0; q < 6; ++q) cache[k][q] = make_float2(0.f, 0.f);
    uint32_t force = 0xfu;                              // the first vertex of this CTA's range loads all four slots
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int st = it & 1; const uint32_t ph = (it >> 1) & 1;
      ptx::mbar_wait(&p_full[st], ph, 400 + st);       // the slot table of this tile (the MMA warp waits on the same barrier)
      ptx::mbar_wait(&tfull[st], ph, 410 + st);
      ptx::tc_fence_after();
      const uint32_t tab = ptx::smem_u32(smem + kPOff + st * kStageBytes + kTabOff);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + st * kNCols;
#pragma unroll 1
      for (int hc = sub; hc < kVT / 2; hc += 3) {       // half-chunks of two vertices
        uint32_t dx[2], dy[2], dz[2];
        tmem_ld_32xN(taddr + hc * 2, dx);
        tmem_ld_32xN(taddr + kVT + hc * 2, dy);
        tmem_ld_32xN(taddr + 2 * kVT + hc * 2, dz);
        ptx::tmem_ld_wait();
        if (hc + 3 >= kVT / 2) {                        // this warp's last accumulator read of the tile: MMA may refill it
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty[st]);
        }
        float cx[kHasCam ? 2 : 1], cy[kHasCam ? 2 : 1], cz[kHasCam ? 2 : 1];       // camera-frame pass (kHasCam)
#pragma unroll
        for (int vi = 0; vi < 2; ++vi) {
          const uint32_t te = tab + (uint32_t)(hc * 2 + vi) * 32;
          const float4 w = ptx::lds_f4(te + 16);
          const uint32_t flags = lds_u32(te) | force;   // bit k: slot k holds another joint than at the previous vertex (warp-uniform)
          if (flags != 0) {
            const uint32_t i01 = lds_u32(te + 4), i23 = lds_u32(te + 8);
            const uint32_t id[4] = {i01 & 0xffffu, i01 >> 16, i23 & 0xffffu, i23 >> 16};
            float ax = 0.f, ay = 0.f, az = 0.f;
            if (a.has_transl) { ax = __ldg(recb + kTcRecTransl); ay = __ldg(recb + kTcRecTransl + 1); az = __ldg(recb + kTcRecTransl + 2); }
            const float is = 1.f / kTcPScale;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if ((flags >> k) & 1u) {
                const float4* src = reinterpret_cast<const float4*>(recb + id[k]);
                const float4 r0 = __ldg(src), r1 = __ldg(src + 1), r2 = __ldg(src + 2);
                cache[k][0] = make_float2(r0.x * is, r0.y * is); cache[k][1] = make_float2(r0.z * is, r0.w + ax);
                cache[k][2] = make_float2(r1.x * is, r1.y * is); cache[k][3] = make_float2(r1.z * is, r1.w + ay);
                cache[k][4] = make_float2(r2.x * is, r2.y * is); cache[k][5] = make_float2(r2.z * is, r2.w + az);
              }
            force = 0u;
          }
          const float wk[4] = {w.x, w.y, w.z, w.w};
          float2 T[6];
#pragma unroll
          for (int q = 0; q < 6; ++q) T[q] = make_float2(wk[0] * cache[0][q].x, wk[0] * cache[0][q].y);
#pragma unroll
          for (int k = 1; k < 4; ++k) {
            const float2 ww = make_float2(wk[k], wk[k]);
#pragma unroll
            for (int q = 0; q < 6; ++q) T[q] = ptx::ffma2(ww, cache[k][q], T[q]);
          }
          // v = T [accumulator; 1]: the accumulator is 2^10 (v_template + shape blend + pose offsets) (lbs.py:179,203,215-220)
          const float x = __uint_as_float(dx[vi]), y = __uint_as_float(dy[vi]), z = __uint_as_float(dz[vi]);
          const float ox = fmaf(T[0].x, x, fmaf(T[0].y, y, fmaf(T[1].x, z, T[1].y)));
          const float oy = fmaf(T[2].x, x, fmaf(T[2].y, y, fmaf(T[3].x, z, T[3].y)));
          const float oz = fmaf(T[4].x, x, fmaf(T[4].y, y, fmaf(T[5].x, z, T[5].y)));
          ptx::sts_f32(tr_w + (uint32_t)(vi * 3) * 4, ox);
          ptx::sts_f32(tr_w + (uint32_t)(vi * 3 + 1) * 4, oy);
          ptx::sts_f32(tr_w + (uint32_t)(vi * 3 + 2) * 4, oz);
          if (kHasCam) { cx[vi] = ox; cy[vi] = oy; cz[vi] = oz; }
        }
        __syncwarp();
        // coalesced copy-out: 24-byte runs per mesh, completed to 96 bytes by the other half-chunks (scalar stores: V * 12
        // bytes is not a multiple of 8, so rows of different meshes have different alignments)
        const int vb = t * kVT + hc * 2;                // first vertex of this half-chunk
        const int ncol = min(6, (a.V - vb) * 3);        // the last tile is clipped at V (warp-uniform)
        const bool fast = rows_full && ncol == 6;
        const size_t e0 = (size_t)lane_off + (size_t)vb * 3;
        copy_out(tr_r, a.out + e0, (size_t)5 * row_stride, lane < 30, r5 < 2, fast, a.B - mesh0 - r5, col < ncol);
        if (kHasCam) {                                  // transform_smpl (utils.py:237-239): R v + t; record: R row-major (9), t (3)
          const float4* cs = reinterpret_cast<const float4*>(recb + kTcRecCam);
          const float4 c0 = __ldg(cs), c1 = __ldg(cs + 1), c2 = __ldg(cs + 2);       // R00 R01 R02 R10 | R11 R12 R20 R21 | R22 t0 t1 t2
          __syncwarp();
#pragma unroll
          for (int vi = 0; vi < 2; ++vi) {
            const float qx = fmaf(c0.x, cx[vi], fmaf(c0.y, cy[vi], fmaf(c0.z, cz[vi], c2.y)));
            const float qy = fmaf(c0.w, cx[vi], fmaf(c1.x, cy[vi], fmaf(c1.y, cz[vi], c2.z)));
            const float qz = fmaf(c1.z, cx[vi], fmaf(c1.w, cy[vi], fmaf(c2.x, cz[vi], c2.w)));
            ptx::sts_f32(tr_w + (uint32_t)(vi * 3) * 4, qx);
            ptx::sts_f32(tr_w + (uint32_t)(vi * 3 + 1) * 4, qy);
            ptx::sts_f32(tr_w + (uint32_t)(vi * 3 + 2) * 4, qz);
          }
          __syncwarp();
          copy_out(tr_r, a.out_cam + e0, (size_t)5 * row_stride, lane < 30, r5 < 2, fast, a.B - mesh0 - r5, col < ncol);
        }
        __syncwarp();                                   // the buffer is rewritten by the next half-chunk
      }
      if (lane == 0) ptx::mbar_arrive(&p_empty[st]);     // this warp is done with the tile's slot table
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

// Slot tables: per vertex 4 reload flags, 4 slot ids (float offset of the joint's A inside a record, j * 12) and 4 weights.  The order of the
// slots is chosen per epilogue warp's vertex stream (tile t, vertices 8 s .. 8 s + 7) so that a joint stays in its slot from
// one vertex to the next -- the kernel reloads a slot only when its id changes, so the choice affects speed, never results.
int smplx_ml_create(const airpose_smplx_model_host* mh, const SmplxDev& d, SmplxTc* tc, std::vector<void*>* owned) {
  tc->ml_ok = false;
  // Opt-in (AIRPOSE_SMPLX_ML=1): measured on B200 at B = 8192 this kernel ties with the vertex-lane kernel (0.70 vs 0.65 ms,
  // profiles/r02mo_ncu_full_lbs_b8192_meshlane.csv): 3x fewer shared-memory wavefronts per vertex.mesh, but 96 instructions per
  // vertex.mesh at 39 % issue utilisation on 12 epilogue warps (the 48-register matrix cache caps the CTA at 448 threads).
  if (!tc->ok || !getenv("AIRPOSE_SMPLX_ML") || atoi(getenv("AIRPOSE_SMPLX_ML")) == 0) return 0;
  const int V = d.V, J = d.J;
  std::vector<int> anc(J);
  for (int j = 0; j < J; ++j) anc[j] = j < kTcBodyJoints ? j : anc[(int)mh->parents[j]];
  std::vector<std::vector<std::pair<int, float>>> rows(V);
  for (int v = 0; v < V; ++v)
    for (int j = 0; j < J; ++j) {
      const float w = mh->lbs_weights[(size_t)v * J + j];
      if (w == 0.f) continue;
      bool merged = false;
      for (auto& e : rows[v])
        if (e.first == anc[j]) { e.second += w; merged = true; break; }
      if (!merged) rows[v].push_back({anc[j], w});
    }
  for (int v = 0; v < V; ++v)
    if ((int)rows[v].size() > 4) return 0;               // denser skinning: the vertex-lane kernel (8 slots) handles it
  const int vtiles = ceil_div(V, kMlVertTile);
  std::vector<uint32_t> tab((size_t)vtiles * kMlVertTile * 8, 0u);
  for (int s = 0; s < 3; ++s) {                          // the vertex stream of epilogue warps with (warp - 2) / 4 == s
    int slot[4] = {0, 0, 0, 0};                          // joint currently assigned to each slot (joint 0 to start with)
    int loaded[4] = {-1, -1, -1, -1};                    // joint the kernel's register cache holds in each slot at this point
    long last[4] = {-4, -3, -2, -1};                     // last use, for the replacement choice
    long tick = 0;
    for (int t = 0; t < vtiles; ++t)
      for (int hc = s; hc < kMlVertTile / 2; hc += 3)
        for (int i = 0; i < 2; ++i, ++tick) {
        const int v = t * kMlVertTile + hc * 2 + i;
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        if (v < V) {
          bool used[4] = {false, false, false, false};
          std::vector<std::pair<int, float>> miss;
          for (auto& e : rows[v]) {
            int k = -1;
            for (int q = 0; q < 4; ++q)
              if (!used[q] && slot[q] == e.first) { k = q; break; }
            if (k < 0) { miss.push_back(e); continue; }
            used[k] = true; w[k] = e.second; last[k] = tick;
          }
          for (auto& e : miss) {                         // least recently used free slot
            int k = -1;
            for (int q = 0; q < 4; ++q)
              if (!used[q] && (k < 0 || last[q] < last[k])) k = q;
            used[k] = true; slot[k] = e.first; w[k] = e.second; last[k] = tick;
          }
        }
        uint32_t* row = &tab[(size_t)v * 8];
        for (int k = 0; k < 4; ++k) {
          if (w[k] != 0.f && loaded[k] != slot[k]) {      // zero-weight slots are never loaded
            row[0] |= 1u << k;
            loaded[k] = slot[k];
          }
          row[1 + (k >> 1)] |= (uint32_t)(slot[k] * 12) << (16 * (k & 1));
          memcpy(&row[4 + k], &w[k], 4);
        }
      }
  }
  uint32_t* dtab;
  if (device_upload(&dtab, tab.data(), tab.size())) return 1;
  owned->push_back(dtab);
  tc->ml_vtab = dtab;
  tc->ml_vtiles = vtiles;
  const int vrows = tc->vtiles * 128;
  if (make_tmap_tiled_bf16(&tc->tmP32, tc->P, (int64_t)3 * vrows, kTcK, kTcK, kMlVertTile, 64)) return 1;
  if (make_tmap_tiled_bf16(&tc->tmPx32, tc->P, (int64_t)3 * vrows, kTcK, kTcK, kMlVertTile, kTcKShape, 64)) return 1;
  tc->ml_ok = true;
  return 0;
}

// Vertex tiles per CTA: whole waves over the SMs, few enough CTAs per mesh tile that the 104 KB of F per CTA amortise.
static int pick_vtiles_per_cta(int mesh_tiles, int vtiles) {
  const int sms = num_sms();
  double best = 1e30; int best_s = 1;
  for (int s = 1; s <= vtiles; ++s) {
    const int per = ceil_div(vtiles, s);
    const int waves = ceil_div(mesh_tiles * s, sms);
    const double cost = waves * (2.5 + per);           // F load ~ 2.5 tile times
    if (cost < best) { best = cost; best_s = s; }
  }
  return ceil_div(vtiles, best_s);
}

int smplx_ml_forward(const SmplxDev& d, const SmplxTc& tc, const TcCall& c, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    AP_CHECK_CUDA(cudaFuncSetAttribute(smplx_vertex_ml_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    AP_CHECK_CUDA(cudaFuncSetAttribute(smplx_vertex_ml_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  CUtensorMap tmFh, tmFl, tmFx;
  if (make_tmap_tiled_bf16(&tmFh, c.fh, c.B, kTcK, kTcK, 128, 64)) return 1;
  if (make_tmap_tiled_bf16(&tmFl, c.fl, c.B, kTcK, kTcK, 128, 64)) return 1;
  if (make_tmap_tiled_bf16(&tmFx, c.fh, c.B, kTcK, kTcK, 128, kTcKShape, 64)) return 1;
  MlArgs a{};
  a.V = d.V; a.B = c.B; a.has_transl = c.has_transl;
  a.vtiles = tc.ml_vtiles;
  const int mesh_tiles = ceil_div(c.B, 128);
  a.tiles_per_cta = pick_vtiles_per_cta(mesh_tiles, a.vtiles);
  a.vrows = tc.vtiles * 128;
  a.rec = c.rec; a.vtab = tc.ml_vtab; a.out = c.out; a.out_cam = c.out_cam;
  dim3 grid(mesh_tiles, ceil_div(a.vtiles, a.tiles_per_cta));
  if (c.out_cam) smplx_vertex_ml_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(tmFh, tmFl, tmFx, tc.tmP32, tc.tmPx32, a);
  else smplx_vertex_ml_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(tmFh, tmFl, tmFx, tc.tmP32, tc.tmPx32, a);
  AP_LAUNCH_CHECK();
  return 0;
}

}  // namespace airpose
