// Host-side interface of the tcgen05 GEMM / implicit-conv kernel (gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace airpose {

// Epilogue of one GEMM: v = acc*scale[n] + shift[n] (+ residual[m,n]); relu; store.
struct Epilogue {
  const float* scale = nullptr;           // [N] or null (=1)
  const float* shift = nullptr;           // [N] or null (=0)
  const void* residual = nullptr;         // [M,N] bf16 (or fp32 when residual_f32) or null
  int64_t ldr = 0;
  int residual_f32 = 0;
  int relu = 0;
  void* out_bf16 = nullptr; int64_t ldd = 0;     // bf16 [M,N]
  float* out_f32 = nullptr; int64_t ldf = 0;     // fp32 [M,N]
  void* out_split = nullptr; int64_t lds = 0;    // bf16 [M,3N]: hi | lo | hi  (split-bf16 operand of the next GEMM)
};

// Geometry of the A operand when it is an NHWC activation read through TMA im2col.
struct ConvGeom {
  int n = 0, H = 0, W = 0, Cin = 0;
  int Ho = 0, Wo = 0;
  int ksize = 1, stride = 1, pad = 0;
};

struct GemmLaunch {
  CUtensorMap tmA;     // tiled [M,K] map, or im2col map over (C,W,H,N)
  CUtensorMap tmB;     // tiled [N,K] map
  CUtensorMap tmD;     // tma_epi: tiled [M,N] bf16 output map, box 128 rows x 64 cols, 128B swizzle
  CUtensorMap tmR;     // tma_epi: same shape over the bf16 residual (valid when epi.residual != null)
  int M = 0, N = 0, K = 0;
  int block_n = 128;   // 64, 128 or 256
  int im2col = 0;
  // stem mode (gemm_tma.cu only): K blocks of 32 (64-byte swizzle); k-block r reads the A rows
  // img*stem_img_stride + stem_tap_off[r] + (m % stem_img_rows) of a plain [rows,32] matrix.
  int stem = 0;
  int stem_img_rows = 0, stem_img_stride = 0;
  int stem_tap_off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int tma_epi = 0;     // 1: bf16 output through smem staging + TMA store (gemm_tma.cu); 0: direct stores (gemm.cu)
  int pdl = 0;         // launch with programmatic stream serialization (prologue overlaps the previous kernel's tail)
  int pair_b_box = 0;  // 1: planned for the cta_group::2 pair kernel (gemm_sk2.cu): block_n = 256, tmB boxes of 128 rows
  // gemm_sk.cu only: the operand is given TRANSPOSED -- A as [K][M], B as [K][N], row-major (M / N contiguous) -- and read through
  // MN-major UMMA descriptors; its tensor map has boxes of 64 columns x 64 rows (make_tmap_tiled_bf16(rows = K, cols = M or N)).
  // The weight-gradient GEMMs contract over the pixels, which is the OUTER dimension of every NHWC tensor.
  int mt2 = 0;                 // gemm_sk.cu: 256 x block_n tiles (two 128-row sub-tiles share every B k-block; single-buffered accumulator)
  int mn_a = 0, mn_b = 0;      // mn_b == 2: B = im2col(x)^T, tmB an im2col map with boxes of 64 channels x 64 pixels, geom the conv
  ConvGeom geom;
  Epilogue epi;
};

// True when the epilogue can run through TMA (bf16 output only, whole 64-column chunks).
bool tma_epilogue_eligible(const GemmLaunch& L);
// Builds tmD / tmR from epi.out_bf16 / epi.residual and sets tma_epi (call after epi is filled in).
int enable_tma_epilogue(GemmLaunch* L);
int launch_gemm_tma(const GemmLaunch& L, cudaStream_t stream);
// Second-generation kernel (gemm_sk.cu): stream-K split, two epilogue warpgroups, resident weights.
// Chosen per problem by prefers_stream_k(); AIRPOSE_GEMM_V1=1 / AIRPOSE_GEMM_SK=1 force one kernel for A/B runs.
// With `sk` (ranges >= 2 on entry): plain split-K -- every tile is cut into `ranges` k-ranges, one CTA each, and EVERY range leaves
// its fp32 partial tile in the stream's workspace instead of anything being stored through tmD; the caller sums the partials in a
// fixed order with its own kernel, launched next on the same stream.  For GEMMs with a handful of tiles and a very long K (the
// weight gradients of the early layers: 1-6 tiles, thousands of k-blocks), where stream-K's one-owner gather caps a tile at 16 ranges.
// Partial of tile t (= m_blk * tiles_n + n_blk), range r: part + (t * ranges + r) * slot_floats; element (row, col) of the 128 x
// block_n tile at float ((col / 64 * 16 + (col % 64) / 4) * 128 + row) * 4 + col % 4.
struct SplitKInfo { int ranges; const float* part; int tiles_n, block_n, slot_floats; };
int splitk_ranges(int M, int N, int K, int block_n);      // 0: stream-K fills the GPU by itself
int launch_gemm_sk(const GemmLaunch& L, cudaStream_t stream, SplitKInfo* sk = nullptr);
bool use_stream_k();
// cta_group::2 pair kernel (gemm_sk2.cu): 256 x 256 tiles on CTA pairs.  prefers_pair() decides per problem.
int launch_gemm_sk2(const GemmLaunch& L, cudaStream_t stream);
bool prefers_pair(int M, int N, int K);
// rows of a tmB box for a launch of this shape (block_n, or 128 when the pair kernel will run it)
int b_box_rows(int M, int N, int K, int block_n);

// Tensor maps (cuTensorMapEncode* resolved through cudaGetDriverEntryPoint; no libcuda link).
int make_tmap_tiled_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld_elems,
                         int box_rows, int box_cols, int swizzle_bytes = 128);
int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, const ConvGeom& g, int channels_per_pixel,
                          int pixels_per_column);

// Tiled 4-d map over an NHWC bf16 tensor, dims (C, W, H, N), 128-byte swizzle, box (64 channels, box_w, box_h, 1 image).
// Boxes may overhang the tensor on every side (halo loads: zero fill; stores: clipped).
int make_tmap_nhwc4d_bf16(CUtensorMap* out, const void* base, int n, int H, int W, int C, int box_w, int box_h);

// Fused bottleneck tail (bneck.cu): conv2 3x3 + BN + ReLU -> conv3 1x1 + BN + residual + ReLU in one launch.
struct alignas(64) TailLaunch {
  unsigned char storage[832];      // five tensor maps + the kernel parameters (TailLaunchImpl in bneck.cu)
  int valid = 0;
  int pdl = 0;
};
bool bneck_tail_supported(int H, int W, int Cm, int Cout);
int build_bneck_tail(TailLaunch* L, const void* t1, const void* w2, const float* scale2, const float* shift2, const void* w3,
                     const float* scale3, const float* shift3, const void* residual, void* out, int n, int H, int W);
int launch_bneck_tail(const TailLaunch& L, cudaStream_t stream);

// 3x3 / stride 1 conv + BN + ReLU of the 128-channel stage with the input band resident in shared memory (conv3x3.cu)
struct alignas(64) SlabLaunch {
  unsigned char storage[512];      // three tensor maps + the kernel parameters (SlabLaunchImpl in conv3x3.cu)
  int valid = 0;
  int pdl = 0;
};
bool conv3x3_slab_supported(int H, int W, int Cin, int Cout, int ksize, int stride, int pad);
int build_conv3x3_slab(SlabLaunch* L, const void* x, const void* w, const float* scale, const float* shift, int relu, void* out, int n, int H,
                       int W);
int launch_conv3x3_slab(const SlabLaunch& L, cudaStream_t stream);

// Fused stem (stem.cu): conv 7x7 s2 + BN + ReLU + MaxPool 3x3 s2 in one launch (plus the operand pack kernel).
struct alignas(64) StemLaunch {
  unsigned char storage[320];      // two tensor maps + the kernel parameters (StemLaunchImpl in stem.cu)
  int valid = 0;
  int pdl = 0;
};
size_t stem_pairs_operand_elems(int n);      // bf16 elements of the packed operand of n images
size_t stem_pairs_weight_elems();
int stem_pack_pairs_weight(const float* w_f32, void* out_bf16, cudaStream_t st);
int build_stem_pool(StemLaunch* L, const void* x2p, const void* wp, const float* scale, const float* shift, void* out_nhwc, int n);
int launch_stem_pool(const StemLaunch& L, const float* x_nchw, void* x2p, int n, cudaStream_t stream);

int pick_block_n(int M, int N, int K);
bool prefers_stream_k(int M, int N, int K);
bool use_tma_epilogue();   // off with AIRPOSE_NO_TMA_EPI=1 (A/B runs)
bool use_pdl();            // off with AIRPOSE_NO_PDL=1
int launch_gemm(const GemmLaunch& L, cudaStream_t stream);
int num_sms();
// Stage A of a multi-chunk trunk call runs its chunks on two streams: a cap below the SM count lets kernels of the two streams
// be co-resident on disjoint SMs (their fixed per-launch latencies overlap).  0 = no cap.  Thread-local, set by trunk.cu.
void set_grid_cap(int cap);
int grid_limit();          // min(num_sms(), cap)

}  // namespace airpose
