// Parameter blocks shared by the tcgen05 implicit-GEMM kernels (tc_gemm.cu) and their host-side
// geometry builders (capi.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace b200 {

constexpr int kMaxTaps = 25;      // 5x5 filters
constexpr int kMaxPhases = 4;     // stride-2 dgrad output parities
constexpr int kTileM = 128;       // UMMA M (TMEM lanes)
constexpr int kBlockK = 64;       // bf16 elements per 128-byte swizzle row
constexpr int kTmemCols = 256;

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };

// "Tap GEMM": out[pixel, n] = epilogue( sum_taps sum_k A_tap[pixel, k] * B_tap[n, k] ).
// A tiles are TMA boxes of 128 pixels x 64 channels out of an NHWC activation tensor (rank-2..5
// tensor map, per-tap start offsets, zero fill = conv padding); B tiles are rows of a K-major
// weight matrix.  Covers conv fprop, strided dgrad / transposed-conv forward (one phase per
// blockIdx.z) and dense layers (one tap).
struct TapGemmParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmA_tail;               // same tensors with a narrow inner box for the last K chunk of a tap
  CUtensorMap tmB_tail;
  int tail_mode;                      // 0: last chunk uses the 64-wide maps; 1: 16-wide box, SWIZZLE_32B;
                                      // 2: 32-wide box, SWIZZLE_64B
  int merge_tail;                     // tail_mode 1 only: the tail rides in the stage of the last full chunk
  int a_rank;
  int kchunks;                        // ceil(K / 64) per tap
  int k_total;                        // K per tap (channels)
  int nphases;
  int phase_tap_begin[kMaxPhases + 1];
  int tap_a_off[kMaxTaps][4];         // start coordinate of A dims 1..4 for this tap
  int tap_b_row[kMaxTaps];            // first B row of this tap
  int a_mul[3][4];                    // [w|h|n] pixel index -> A dims 1..4 coordinate step
  int bw, bh, bn;                     // pixels per tile along w, h, n (product 128)
  int tiles_w, tiles_h, tiles_n;
  int phase_ext_w[kMaxPhases], phase_ext_h[kMaxPhases];
  int ext_n;
  long long phase_o_off[kMaxPhases];  // output element offset of the phase
  long long o_sw, o_sh, o_sn;         // output element strides per pixel index
  int ncols, bn_tile;                 // valid output columns, N tile (multiple of 16, <= 256)
  int stages;
  int epi_pipe;                       // epilogue: prefetch + pipelined mask loads
  int l2_prefetch;                    // K blocks the producer prefetches into L2 ahead of its loads (0 = off)
  int dual;                           // pixel tiles per CTA (1|2) sharing one B tile; 2 -> two TMEM accumulators
  int cluster;                        // 1, or 2: CTA pairs share B through TMA multicast (B box = bn_tile/2 rows)
  const void* b_base;                 // B operand (weights) base / row pitch / rows: L2 prefetch by the idle epilogue warps
  int b_pitch_bytes;
  int b_rows_total;
  int b_prefetch;                     // 1: CTAs of the first pixel tile pull their N tile's weight rows into L2 at kernel start
  int trace;                          // debug: CTA (0,0,0) prints clock stamps of its phases
  int quad;                           // 2-CTA kernel: 1 = cluster (2,2,1), the two pairs multicast their shared A tiles
  int cta2;                           // 1: 2-CTA kernel (cta_group::2): B split across the pair, M = 256 per MMA
  CUtensorMap tmOut[kMaxPhases];      // per phase: output viewed as [ext_n, ext_h, ext_w, ncols], box = the tile (tma_store)
  int tma_store;                      // 1: the epilogue stages its tiles in the (idle) pipeline smem and bulk-stores them
  int stage_pitch;                    // bytes per staged row = bn_tile * element size
  int cluster_y;                      // set by launch_tapgemm: 2 -> 2x2 clusters, the N-tile pair also shares A
  int splits;                         // > 1: split-K over blockIdx.z (1-CTA kernel only); the epilogue must be the
                                      // atomic fp32 accumulate (accumulate == 2) into a zeroed workspace
  // stream-K persistent form of the 2-CTA kernel (tapgemm2sm_sk_kernel), set by launch_tapgemm:
  int sk;                             // 1: this launch runs the persistent kernel
  int sk_snap;                        // 1: range boundaries snapped to item boundaries (whole items, no partial sums)
  int sk_groups;                      // pixel-tile groups (4 tiles = one CTA pair) per (phase, N tile)
  int sk_ny;                          // N tiles
  long long sk_total, sk_range;       // total (item, K-iteration) positions; positions per CTA pair
  long long sk_phase_base[kMaxPhases + 1];   // first position of each phase's items
  float* sk_partial;                  // [pair][cta][2 tiles][128 rows][bn_tile] fp32 partial accumulators
  int* sk_flags;                      // [pair][cta] 1 = that CTA's partial is published (reset by the consumer)
  long long sk_region;                // floats per CTA region = 2 * 128 * bn_tile
  void* out;
  int out_f32;                        // 0: bf16, 1: fp32
  int accumulate;                     // fp32 only: 1: out += result (one writer); 2: red.global.add (split-K partials)
  const float* bias;                  // [ncols] or null
  int act;
  float leak;
  const __nv_bfloat16* mask_src;      // same geometry as out; result *= act'(mask_src)
  int mask_kind;
  const uint16_t* mask_bits;          // sign bitmaps (see epilogue.cuh): consumer side / producer side
  uint16_t* bits_out;
  int bits_pitch, row_elems;
  float alpha;
};

// Weight gradient: out[tap][ca][cb] += alpha * sum_pixels A_tap[pixel, ca] * B[pixel, cb], both
// operands MN-major (channels contiguous), reduction over pixels, split-K over blockIdx.y with
// fp32 atomics.
struct WgradParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  int a_rank, b_rank;
  int ntaps;
  int tap_a_off[kMaxTaps][4];
  int a_mul[3][4];
  int bw, bh, bn;                     // pixels per K-chunk along w, h, n (product 64)
  int chunks_w, chunks_h, chunks_n;
  int total_chunks, chunks_per_split;
  int trace;                          // debug: CTA 0 prints where its producer / MMA threads wait
  int units, chunks_per_cta;          // set by launch_wgrad: (tap, M pair, N tile) units; linear chunk range per CTA
  int Ca, Cb;
  int m_tiles, n_tiles, bn_tile, nb_boxes;
  int bn_tile_t;                      // N tile of the transposed 2-CTA kernel (over Ca); 0 = not eligible
  int l2_prefetch;                    // K chunks prefetched into L2 ahead of the loads
  int dual;                           // 128-channel M tiles per CTA (1|2) sharing one B tile
  int stages;
  float* out;
  long long out_tap_stride;
  int ldo;
  float alpha;
};

int tapgemm_cluster_size(const TapGemmParams& p);
int tapgemm_dual(int m_tiles, int iters);
void set_dual_min_pct(int v);
void set_wgrad_min_chunks(int v);
int tapgemm_stage_bytes(int dual, int bn_tile, int merge_tail);
int tapgemm_2sm(int cluster, int dual, int tail_mode, int merge_tail, int bn_tile);
int tapgemm_stage_bytes_2sm(int bn_tile, int merge_tail);
int epilogue_pipelined();
int l2_prefetch_distance();
void launch_tapgemm(const TapGemmParams& p, cudaStream_t stream);

// Persistent small-K GEMM: out[M, ncols] = epi(A[M,K] * B[ncols,K]^T) with K <= 256 (image-side im2col GEMMs).
// The whole B stays resident in shared memory, A tiles (128 rows, full K) stream through a ring, and two TMEM
// accumulators alternate so the epilogue of tile i overlaps the MMAs of tile i+1.  grid = #SMs.
struct SmallKParams {
  CUtensorMap tmA;                    // [M, K], box 64 x 128
  CUtensorMap tmB;                    // [rows, K], box 64 x bn_tile
  int kchunks, k_total;
  int num_tiles;
  long long M;
  int ncols, bn_tile;
  int slots;                          // A ring depth
  int trace;                          // debug: print per-tile clock stamps of CTA 0
  CUtensorMap tmOut;                  // [M, ncols] output, box ncols x 128, no swizzle (tma_store)
  int tma_store;                      // 1: tiles are staged in shared memory and written with one bulk tensor store
  int stage_pitch;                    // bytes per staged row (ncols * element size)
  int bits_stage;                     // 1: the tile's sign words are staged and written with one bulk copy
  int l2_ahead;                       // tiles prefetched into L2 ahead of their loads (set by launch_smallk)
  int epi_pipe;
  long long ldo;
  void* out;
  int out_f32, accumulate;
  const float* bias;
  int act;
  float leak;
  const __nv_bfloat16* mask_src;
  int mask_kind;
  const uint16_t* mask_bits;          // sign bitmaps (see epilogue.cuh): consumer side / producer side
  uint16_t* bits_out;
  int bits_pitch, row_elems;
  float alpha;
};
bool smallk_fits(int kchunks, int bn_tile, int* slots, int stage_bytes = 0);
void launch_smallk(const SmallKParams& p, cudaStream_t stream);
int wgrad_dual(int m_tiles);
void launch_wgrad(const WgradParams& p, int splits, cudaStream_t stream);

}  // namespace b200
