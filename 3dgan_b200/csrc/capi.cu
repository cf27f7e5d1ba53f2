// extern "C" boundary (include/b200gan.h) and the host-side geometry builders that turn a TF-style
// SAME convolution into TMA tensor maps + tap tables for the tcgen05 kernels.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <string>

#include "../../include/b200gan.h"
#include "simt_kernels.cuh"
#include "tc_gemm.cuh"
#include "img_conv.cuh"

using namespace b200;

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }
static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptor encoding through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor, dims[0] contiguous; strides in ELEMENTS for dims 1..rank-1; 128-byte swizzle, zero OOB fill
static int make_tmap(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides,
                     const int* box, const int* estride, int swizzle_bytes = 128, int elem_bytes = 2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail("tensor map base not 16-byte aligned");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    b[i] = (cuuint32_t)box[i];
    es[i] = (cuuint32_t)estride[i];
    if (box[i] < 1 || box[i] > 256) return fail("TMA box dim out of range");
    if (i > 0) {
      gstr[i - 1] = (cuuint64_t)strides[i] * elem_bytes;
      if (gstr[i - 1] & 15) return fail("TMA stride not a multiple of 16 bytes (channels must be a multiple of 8)");
    }
  }
  CUresult r = enc(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                   : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                         : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return fail(buf);
  }
  return 0;
}

static long long align256(long long v);
static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static int cdiv(int a, int b) { return (a + b - 1) / b; }

static void pick_pixel_tile(int Wo, int Ho, int total, int* bw, int* bh, int* bn) {
  *bw = std::min(pow2ceil(Wo), total);
  *bh = std::min(pow2ceil(Ho), total / *bw);
  *bn = total / (*bw * *bh);
}
static int pick_bn_tile(int cols) {
  const int nt = cdiv(cols, 256);
  return std::max(16, cdiv(cdiv(cols, nt), 16) * 16);
}
// Run-time planner overrides (b200_set_tuning): N-tile cap (0 = the planner's own choice) and forced split-K factor
// (-1 = the planner's own choice); tools/tune_layers.py sweeps them per geometry.
static int g_bn_cap = 0, g_force_splits = -1;
static int pick_stages(int stage_bytes) {
  static int cap = -1;          // B200GAN_STAGES: cap the pipeline depth (2 leaves room for two CTAs per SM)
  if (cap < 0) { const char* e = getenv("B200GAN_STAGES"); cap = e ? atoi(e) : 8; }
  int s = (227 * 1024 - 3072) / stage_bytes;
  return std::max(2, std::min(s, std::min(cap, 8)));
}

static int tma_store_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200GAN_TMA_STORE"); v = e ? atoi(e) : 1; }
  return v;
}
static int weight_prefetch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200GAN_PREFETCH_B"); v = e ? atoi(e) : 1; }
  return v;
}

static bool is_small(int c) { return c <= 4; }

// Launch plan of a tensor-core conv tap GEMM (fprop: op 0, dgrad: op 1): pixel-tile layout, N tile and split-K
// factor.  Fitted to tools/tune_layers.py (every fprop / dgrad geometry of pix2pix, VAE and the cnn autoencoder under
// all combinations; gpurun_out/tune_layers.json): the fastest plan is the one that puts ~128 CTAs on the 148 SMs --
//   * two pixel tiles per CTA in CTA pairs (shared B tile, tcgen05 cta_group::2) whenever there are >= 4 pixel tiles;
//   * the widest N tile (<= 256 columns: one A tile feeds the most MACs) that still yields >= 96 CTAs: 128, or 64
//     when 64 alone gets there;
//   * and only then split-K over the taps (fp32 partial image + finalize pass; pix2pix's inner U-Net layers have
//     16..1024 output pixels and K = 16 taps x 512..1024 channels: a handful of CTAs would stream the whole weight
//     tensor), at most 8 (dgrad: 4) slices of >= 6 pipeline iterations.
// Layers that fill the machine anyway (the IWGAN c2 / c3 shapes: 128 .. 256 CTAs at N tile 208) are unaffected.
struct ConvPlan { int dual, bn_tile, splits, tiles, phases; };
static ConvPlan conv_plan(const b200_conv_geom* g, int op) {
  static int split_enabled = -1;
  if (split_enabled < 0) { const char* e = getenv("B200GAN_TAPSPLIT"); split_enabled = e ? atoi(e) : 1; }
  const int st = g->stride, kk = g->k * g->k;
  int bw, bh, bn;
  ConvPlan pl;
  int cols, iters, iters_min;
  if (op == 0) {
    pick_pixel_tile(g->Wo, g->Ho, kTileM, &bw, &bh, &bn);
    pl.tiles = cdiv(g->Wo, bw) * cdiv(g->Ho, bh) * cdiv(g->N, bn);
    pl.phases = 1;
    cols = g->Cout;
    iters = iters_min = kk * cdiv(g->Cin, kBlockK);
  } else {
    const int ew = cdiv(g->W, st), eh = cdiv(g->H, st);
    pick_pixel_tile(ew, eh, kTileM, &bw, &bh, &bn);
    pl.tiles = cdiv(ew, bw) * cdiv(eh, bh) * cdiv(g->N, bn);
    pl.phases = st * st;
    cols = g->Cin;
    const int kch = cdiv(g->Cout, kBlockK);
    iters = std::max(1, kk / pl.phases) * kch;                                  // per output parity
    iters_min = std::max(1, (g->k / st) * (g->k / st)) * kch;                   // lightest output parity
  }
  pl.dual = tapgemm_dual(pl.tiles, iters_min);
  // short-K layers stay single-tile (two co-resident CTAs per SM hide each other's latency) unless that grid needs a
  // second wave which two-tile CTAs avoid: the generator's fc1 as a 1x1 conv, 4 pixel tiles x 50 N tiles = 200 CTAs
  {
    const long long per_tile = (long long)pl.phases * cdiv(cols, 256);
    if (pl.dual == 1 && iters_min <= 4 && pl.tiles >= 4 && tapgemm_dual(pl.tiles, 5) == 2 &&
        pl.tiles * per_tile > 148 && 2LL * cdiv(pl.tiles, 4) * per_tile <= 148)
      pl.dual = 2;
  }
  const long long rows = (pl.dual == 2 ? 2LL * cdiv(pl.tiles, 4) : (long long)pl.tiles) * pl.phases;   // CTAs per N tile
  int cap = g_bn_cap > 0 ? g_bn_cap : 256;
  auto bn_of = [&](int cp) { const int nt = cdiv(cols, cp); return std::max(16, cdiv(cdiv(cols, nt), 16) * 16); };
  // (a single pixel tile -- <= 128 output pixels per parity -- is better served by 128 columns and one more split:
  //  4x4x512->1024 fprop 29.5 -> 19.9 us, dgrad 28.4 -> 18.4 us)
  //  64 columns only when that alone reaches 96 CTAs: otherwise 128 columns and split-K (16x16x512->1024 fprop:
  //  64 columns x 2 slices 41 us, 128 columns x 4 slices 35 us)
  //  -- for deep reductions (k*k*C >= 8192: pix2pix's inner layers); the shallower VAE / cnn layers (k*k*C <= 6400)
  //  are faster with 64 columns and fewer slices (cnn 7x7x128->256 fprop 16.5 vs 18.9 us)
  const bool deep = (long long)kk * (op == 0 ? g->Cin : g->Cout) >= 8192;
  if (g_bn_cap <= 0) {
    if (rows * cdiv(cols, bn_of(cap)) < 96) cap = 128;
    if (pl.tiles > 1 && rows * cdiv(cols, bn_of(128)) < 96 && (!deep || rows * cdiv(cols, bn_of(64)) >= 96)) cap = 64;
  }
  pl.bn_tile = bn_of(cap);
  const long long ctas = rows * cdiv(cols, pl.bn_tile);
  pl.splits = 1;
  if (g_force_splits >= 1) pl.splits = std::min(g_force_splits, std::max(1, iters / 2));
  else if (split_enabled && ctas < 96 && iters >= 16)
    while (pl.splits < (op == 1 ? 4 : 8) && ctas * pl.splits < 96 && iters / (2 * pl.splits) >= 6)
      pl.splits *= 2;       // 2, 4 or 8 (the measured factors); dgrad's output parities already multiply its CTAs
  return pl;
}
static int tc_tap_splits(const b200_conv_geom* g, int op) { return conv_plan(g, op).splits; }
static long long tc_out_elems(const b200_conv_geom* g, int op) {
  return op == 0 ? (long long)g->N * g->Ho * g->Wo * g->Cout : (long long)g->N * g->H * g->W * g->Cin;
}

extern "C" int b200_conv2d_route(const b200_conv_geom* g, int op) {
  const int kk = g->k * g->k;
  if (is_small(g->Cin)) {
    if (g->Cout % 8 == 0 || (kk * g->Cin <= 80 && g->Cout <= 1024)) return 2;
    return fail("small-channel conv: needs Cout % 8 == 0, or k*k*Cin <= 80 and Cout <= 1024");
  }
  if (g->Cin % 8) return fail("tensor-core conv needs Cin % 8 == 0");
  if (op != 0 && is_small(g->Cout)) return 3;     // <= 4 output channels: SIMT backward (PatchGAN head)
  if (op != 0 && g->Cout % 8) return fail("tensor-core dgrad/wgrad need Cout % 8 == 0");
  if (kk > kMaxTaps) return fail("filter larger than 5x5");
  if (op == 1 && g->stride * g->stride > kMaxPhases) return fail("dgrad stride > 2");
  return 1;
}

// 1 when the epilogue of this call honours b200_epilogue.bits_out / mask_bits (the tcgen05 GEMM epilogues do;
// the SIMT small-channel / small-output epilogues do not)
extern "C" int b200_conv2d_epilogue_bits(const b200_conv_geom* g, int op, int has_workspace) {
  const int r = b200_conv2d_route(g, op);
  if (r == 1 && op == 0 && g->Cout == 1) return 0;      // one output channel: the SIMT dot-product kernel (smallout_fprop)
  if (r == 1) return (op == 0 || op == 1) && !(has_workspace && tc_tap_splits(g, op) > 1);
  if (r == 2) return op == 0 && has_workspace && (g->Cout % 8 == 0);
  return 0;
}

static void fill_epilogue(TapGemmParams& p, const b200_epilogue* e) {
  p.bias = e ? e->bias : nullptr;
  p.act = e ? e->act : 0;
  p.leak = e ? e->leak : 0.f;
  p.mask_src = e ? (const __nv_bfloat16*)e->mask_src : nullptr;
  p.mask_kind = e ? e->mask_kind : 0;
  p.mask_bits = e ? (const uint16_t*)e->mask_bits : nullptr;
  p.bits_out = e ? (uint16_t*)e->bits_out : nullptr;
  p.bits_pitch = e ? e->bits_pitch : 0;
  p.out_f32 = e ? e->out_f32 : 0;
  p.accumulate = e ? e->accumulate : 0;
  p.alpha = 1.f;
  p.epi_pipe = epilogue_pipelined();
  p.l2_prefetch = l2_prefetch_distance();
}


// split-K factor from a small cost model: GEMM time / wave fill + fp32-atomic epilogue traffic per split
// (each CTA reduces a 128 x bn_tile fp32 tile into the gradient bucket with red.global.add)
static int pick_splits(int base, int total_chunks, int bn_tile, int dual = 1) {
  const int max_splits = std::max(1, std::min(total_chunks / 4, 148));
  const double t_gemm = 2.0 * 128 * dual * bn_tile * 64 * (double)total_chunks * base / 8e14;
  const double t_atomic = (double)base * 128 * dual * bn_tile * 4 / 2e12;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= max_splits; ++s) {
    const double waves = (double)base * s / 148.0;
    const double eff = waves / std::ceil(waves);
    const double cost = t_gemm / eff + s * t_atomic;
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}


// Output tensor maps for the bulk-store epilogue of the tap GEMM: per phase the output is the strided view
// [ext_n, ext_h, ext_w, ncols] the phase's pixels form; the box is the CTA tile.  Returns with p.tma_store = 0
// when a precondition fails (the epilogue then stores from registers).
static void setup_out_maps(TapGemmParams& p, const b200_epilogue* e) {
  p.tma_store = 0;
  if (!tma_store_enabled() || !epilogue_pipelined()) return;
  if (e && e->accumulate) return;
  const int elem = p.out_f32 ? 4 : 2;
  if (p.ncols % 8) return;
  const uintptr_t align = reinterpret_cast<uintptr_t>(p.out) | reinterpret_cast<uintptr_t>(p.mask_src);
  if (align & 15) return;
  if ((p.o_sw * elem) % 16 || (p.o_sh * elem) % 16 || (p.o_sn * elem) % 16 || (p.o_sw * 2) % 16) return;
  const long long need = (long long)p.dual * kTileM * p.bn_tile * elem;
  if (need > (long long)p.stages * (p.cta2 ? tapgemm_stage_bytes_2sm(p.bn_tile, p.merge_tail)
                                           : tapgemm_stage_bytes(p.dual, p.bn_tile, p.merge_tail))) return;
  for (int i = 0; i < p.nphases; ++i) {
    const char* base = reinterpret_cast<const char*>(p.out) + p.phase_o_off[i] * elem;
    if (reinterpret_cast<uintptr_t>(base) & 15) return;
    long long dims[4] = {p.ncols, p.phase_ext_w[i], p.phase_ext_h[i], p.ext_n};
    long long str[4] = {1, p.o_sw, std::max<long long>(p.o_sh, 1), std::max<long long>(p.o_sn, 1)};
    int box[4] = {p.bn_tile, p.bw, p.bh, p.bn};
    int es[4] = {1, 1, 1, 1};
    if (dims[1] < 1 || dims[2] < 1) return;
    if (make_tmap(&p.tmOut[i], base, 4, dims, str, box, es, 0, elem)) return;
  }
  p.stage_pitch = p.bn_tile * elem;
  p.tma_store = 1;
}

// ------------------------------------------------------------------------------------------------
// Plain GEMMs on the same kernels (rank-2 tensor maps): used by the small-channel im2col route.
// ------------------------------------------------------------------------------------------------
// out[m, n] = epi( sum_k A[m,k] * B[n,k] ),  A [M,K] row stride lda, B [Nrows,K] row stride ldb
static int dense_gemm(const void* A, long long M, int K, int lda, const void* B, int Nrows, int ldb, void* out,
                      long long ldo, int ncols, const b200_epilogue* e, cudaStream_t st) {
  {
    // small K, one N tile, many M tiles: persistent kernel with resident B and double-buffered accumulators
    int slots = 0;
    const int kch = cdiv(K, kBlockK), bn = pick_bn_tile(ncols);
    const long long tiles = (M + kTileM - 1) / kTileM;
    // bulk-store epilogue: possible when every 16-column chunk takes the epilogue's aligned fast path
    const int elem = (e && e->out_f32) ? 4 : 2;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(out) | (e ? reinterpret_cast<uintptr_t>(e->bias) : 0) |
                             (e ? reinterpret_cast<uintptr_t>(e->mask_src) : 0)) & 15) == 0;
    const bool tma_store = tma_store_enabled() && !(e && e->accumulate) && ncols % 8 == 0 && (ldo * elem) % 16 == 0 &&
                           (ldo * 2) % 16 == 0 && aligned16 && epilogue_pipelined();
    const int bits_pitch = e ? e->bits_pitch : 0;
    const bool bits_stage = tma_store && e && e->bits_out && M % kTileM == 0 &&
                            (reinterpret_cast<uintptr_t>(e->bits_out) & 15) == 0;
    const int stage_bytes = tma_store ? (((kTileM * ncols * elem + 127) & ~127) +
                                         (bits_stage ? ((kTileM * bits_pitch * 2 + 127) & ~127) : 0)) : 0;
    if (ncols <= 256 && tiles >= 296 && smallk_fits(kch, bn, &slots, stage_bytes)) {
      SmallKParams q;
      memset(&q, 0, sizeof q);
      if (tma_store) {
        long long dimsO[2] = {ncols, M}, strO[2] = {1, ldo};
        int boxO[2] = {ncols, kTileM}, esO[2] = {1, 1};
        if (make_tmap(&q.tmOut, out, 2, dimsO, strO, boxO, esO, 0, elem)) return -1;
        q.tma_store = 1; q.stage_pitch = ncols * elem; q.bits_stage = bits_stage ? 1 : 0;
      }
      long long dimsA[2] = {K, M}, strA[2] = {1, lda};
      int boxA[2] = {kBlockK, kTileM}, es[2] = {1, 1};
      if (make_tmap(&q.tmA, A, 2, dimsA, strA, boxA, es)) return -1;
      long long dimsB[2] = {K, Nrows}, strB[2] = {1, ldb};
      int boxB[2] = {kBlockK, bn};
      if (make_tmap(&q.tmB, B, 2, dimsB, strB, boxB, es)) return -1;
      q.kchunks = kch; q.k_total = K; q.num_tiles = (int)tiles; q.M = M; q.ncols = ncols; q.bn_tile = bn;
      q.slots = slots; q.ldo = ldo; q.out = out;
      q.out_f32 = e ? e->out_f32 : 0; q.accumulate = e ? e->accumulate : 0; q.bias = e ? e->bias : nullptr;
      q.act = e ? e->act : 0; q.leak = e ? e->leak : 0.f;
      q.mask_src = e ? (const __nv_bfloat16*)e->mask_src : nullptr; q.mask_kind = e ? e->mask_kind : 0;
      q.mask_bits = e ? (const uint16_t*)e->mask_bits : nullptr; q.bits_out = e ? (uint16_t*)e->bits_out : nullptr;
      q.bits_pitch = e ? e->bits_pitch : 0; q.row_elems = (int)ldo;
      q.alpha = 1.f;
      q.epi_pipe = epilogue_pipelined();
      launch_smallk(q, st);
      return 0;
    }
  }
  TapGemmParams p;
  memset(&p, 0, sizeof p);
  fill_epilogue(p, e);
  p.row_elems = (int)ldo;
  p.bw = kTileM; p.bh = 1; p.bn = 1;
  {
    long long dims[2] = {K, M};
    long long str[2] = {1, lda};
    int box[2] = {kBlockK, kTileM};
    int es[2] = {1, 1};
    if (make_tmap(&p.tmA, A, 2, dims, str, box, es)) return -1;
  }
  p.a_rank = 2;
  p.ncols = ncols;
  p.bn_tile = pick_bn_tile(ncols);
  p.cluster = tapgemm_cluster_size(p);
  {
    long long dims[2] = {K, Nrows};
    long long str[2] = {1, ldb};
    int box[2] = {kBlockK, p.bn_tile / p.cluster};
    int es[2] = {1, 1};
    if (make_tmap(&p.tmB, B, 2, dims, str, box, es)) return -1;
  }
  p.kchunks = cdiv(K, kBlockK);
  p.k_total = K;
  p.nphases = 1;
  p.phase_tap_begin[0] = 0; p.phase_tap_begin[1] = 1;
  p.a_mul[0][0] = 1;
  p.tiles_w = (int)((M + kTileM - 1) / kTileM); p.tiles_h = 1; p.tiles_n = 1;
  p.phase_ext_w[0] = (int)M; p.phase_ext_h[0] = 1; p.ext_n = 1;
  p.o_sw = ldo;
  p.dual = tapgemm_dual(p.tiles_w, p.kchunks);
  // short-K dense layers whose single-tile grid needs a second wave (the generator's fc1: 4 row tiles x 50 column
  // tiles = 200 CTAs on 148 SMs) take two row tiles per CTA instead
  if (p.dual == 1 && p.tiles_w >= 2 && (long long)p.tiles_w * cdiv(ncols, p.bn_tile) > 148 &&
      (long long)cdiv(p.tiles_w, 2) * cdiv(ncols, p.bn_tile) <= 148)
    p.dual = 2;
  p.stages = std::min(pick_stages(p.dual * kTileM * kBlockK * 2 + p.bn_tile * kBlockK * 2), std::max(2, p.kchunks));
  p.out = out;
  launch_tapgemm(p, st);
  return 0;
}

// out[ca, cb] += alpha * sum_m A[m,ca] * B[m,cb];  A [M, Ka] (Ka physical width, Ca logical), B [M, Cb]
static int dense_wgrad(const void* A, int Ca, int Ka, const void* B, int Cb, long long M, float* out, int ldo,
                       float alpha, cudaStream_t st) {
  WgradParams p;
  memset(&p, 0, sizeof p);
  p.bw = 64; p.bh = 1; p.bn = 1;
  {
    long long dims[2] = {Ka, M};
    long long str[2] = {1, Ka};
    int box[2] = {64, 64};
    int es[2] = {1, 1};
    if (make_tmap(&p.tmA, A, 2, dims, str, box, es)) return -1;
  }
  {
    long long dims[2] = {Cb, M};
    long long str[2] = {1, Cb};
    int box[2] = {64, 64};
    int es[2] = {1, 1};
    if (make_tmap(&p.tmB, B, 2, dims, str, box, es)) return -1;
  }
  p.a_rank = 2; p.b_rank = 2;
  p.ntaps = 1;
  p.a_mul[0][0] = 1;
  p.chunks_w = (int)((M + 63) / 64); p.chunks_h = 1; p.chunks_n = 1;
  p.total_chunks = p.chunks_w;
  p.Ca = Ca; p.Cb = Cb;
  p.m_tiles = cdiv(Ca, kTileM);
  p.bn_tile = pick_bn_tile(Cb);
  p.n_tiles = cdiv(Cb, p.bn_tile);
  p.nb_boxes = cdiv(p.bn_tile, 64);
  p.bn_tile_t = pick_bn_tile(p.Ca);       // N tile (over X channels) of the transposed 2-CTA kernel
  p.dual = wgrad_dual(p.m_tiles);
  p.l2_prefetch = l2_prefetch_distance();
  p.stages = pick_stages((2 * p.dual + p.nb_boxes) * 64 * 64 * 2);
  p.out = out;
  p.out_tap_stride = 0;
  p.ldo = ldo;
  p.alpha = alpha;
  const int base = cdiv(p.m_tiles, p.dual) * p.n_tiles;
  int splits = pick_splits(base, p.total_chunks, p.bn_tile, p.dual);
  p.chunks_per_split = cdiv(p.total_chunks, splits);
  splits = cdiv(p.total_chunks, p.chunks_per_split);
  launch_wgrad(p, splits, st);
  return 0;
}

// narrow-box variant for the last K chunk of a tap: 1 -> 16 channels (SWIZZLE_32B), 2 -> 32 (SWIZZLE_64B), 0 -> none
static int pick_tail_mode(int K) {
  if (getenv("B200GAN_NO_TAIL")) return 0;
  const int tail = K - (cdiv(K, kBlockK) - 1) * kBlockK;
  if (tail <= 16) return 1;
  if (tail <= 32) return 2;
  return 0;
}

static long long align256(long long v) { return (v + 255) / 256 * 256; }
static int small_kp(const b200_conv_geom* g) { return cdiv(g->k * g->k * g->Cin, 8) * 8; }
// the small-channel layers can run as im2col/col2im + tensor-core GEMM when the big side is TMA-aligned
static bool small_gemm_ok(const b200_conv_geom* g) { return g->Cout % 8 == 0; }

// Row-group K layout of the fused image-side kernels (img_conv.cuh): when the geometry qualifies, the fprop and wgrad
// workspaces are [gathered rows M x k*16 bf16 | weights Cout x k*16 bf16 | filter-gradient image k*16 x Cout fp32]
static ImgConvGeom img_geom(const b200_conv_geom* g) {
  return ImgConvGeom{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->k, g->stride, g->pad_t, g->pad_l};
}
static bool img_layout(const b200_conv_geom* g) {
  return is_small(g->Cin) && small_gemm_ok(g) && img_fprop_supported(img_geom(g), g->Cout, 1);
}
static long long img_ws_a(const b200_conv_geom* g) { return align256((long long)g->N * g->Ho * g->Wo * g->k * 16 * 2); }
static long long img_ws_w(const b200_conv_geom* g) { return align256((long long)g->Cout * g->k * 16 * 2); }
static long long img_ws_t(const b200_conv_geom* g) { return align256((long long)g->Cout * g->k * 16 * 4); }

extern "C" long long b200_conv2d_workspace_bytes(const b200_conv_geom* g, int op) {
  if (!is_small(g->Cin) && op != 2 && b200_conv2d_route(g, op) == 1)      // split-K partial sums: fp32 image of the output
    return tc_tap_splits(g, op) > 1 ? align256(tc_out_elems(g, op) * 4) : 0;
  if (!is_small(g->Cin) || !small_gemm_ok(g)) return 0;
  const long long M = (long long)g->N * g->Ho * g->Wo;
  if (img_layout(g) && op != 1) return img_ws_a(g) + img_ws_w(g) + img_ws_t(g);   // [gathered rows | weights | dW image]
  const int Kp = small_kp(g);
  if (op == 0) return align256(M * Kp * 2) + align256((long long)g->Cout * Kp * 2);
  if (op == 1) return align256(M * Kp * 4);
  return align256(M * Kp * 2);
}

// Split-K launch of a prepared tap GEMM (see tc_tap_splits): zero the fp32 workspace, let every K slice add its
// partial tile into it, then apply the caller's epilogue in one elementwise pass.  Returns false when the launch
// should go the normal way (no split for this geometry, or no workspace was passed).
static bool run_split_k(TapGemmParams& p, const b200_conv_geom* g, int op, const b200_epilogue* e, void* out,
                        void* workspace, long long workspace_bytes, cudaStream_t st) {
  const int splits = tc_tap_splits(g, op);
  const long long n = tc_out_elems(g, op);
  if (splits <= 1 || !workspace || workspace_bytes < n * 4 || (e && e->accumulate)) return false;
  cudaMemsetAsync(workspace, 0, n * 4, st);
  p.splits = splits;
  p.cta2 = 0;
  p.stages = pick_stages(tapgemm_stage_bytes(p.dual, p.bn_tile, p.merge_tail));
  p.out = workspace;
  p.out_f32 = 1; p.accumulate = 2;
  p.bias = nullptr; p.act = 0; p.mask_src = nullptr; p.mask_bits = nullptr; p.bits_out = nullptr;
  p.tma_store = 0;
  launch_tapgemm(p, st);
  splitk_finalize((const float*)workspace, out, e ? e->out_f32 : 0, n, p.ncols, e ? e->bias : nullptr, e ? e->act : 0,
                  e ? e->leak : 0.f, e ? e->mask_src : nullptr, e ? e->mask_kind : 0, st);
  return true;
}

extern "C" int b200_conv2d_fprop(const void* x, const void* w, const void* w_t, void* y, const b200_conv_geom* g,
                                 const b200_epilogue* e, void* workspace, long long workspace_bytes, b200_stream s) {
  cudaStream_t st = (cudaStream_t)s;
  const int route = b200_conv2d_route(g, 0);
  if (route < 0) return route;
  if (route == 2 && workspace && small_gemm_ok(g)) {
    // im2col (bf16 [M,Kp]) + padded transposed filter [Cout,Kp] -> tcgen05 GEMM with the fused epilogue
    if (workspace_bytes < b200_conv2d_workspace_bytes(g, 0)) return fail("conv2d_fprop: workspace too small");
    const long long M = (long long)g->N * g->Ho * g->Wo;
    // k = 4, Cin = 4 (pix2pix's discriminator input): the fused fprop in its virtual-fifth-row form; everything else
    // of that layer keeps the im2col route below
    const bool virt = !img_layout(g) && img_fprop_virtual_supported(img_geom(g), g->Cout);
    if (img_layout(g) || virt) {
      // fused gather + GEMM + epilogue (img_conv.cu); general epilogues (fp32 out, value masks, tanh / sigmoid,
      // accumulate) keep the GEMM route, fed in the same row-group layout so that the workspace always holds
      // what the filter gradient of this layer expects
      const ImgConvGeom ig = img_geom(g);
      const long long x_words = (long long)g->N * g->H * g->W * g->Cin / 2;
      const bool epi_ok = !e || (!e->out_f32 && !e->accumulate && e->act <= ACT_LRELU && (!e->mask_src || e->mask_bits));
      const uintptr_t al = reinterpret_cast<uintptr_t>(y) | (e ? reinterpret_cast<uintptr_t>(e->bits_out) : 0);
      if (epi_ok && (al & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        ImgFpropParams q;
        memset(&q, 0, sizeof q);
        q.g = ig; q.x = (const __nv_bfloat16*)x; q.x_words = x_words; q.w = (const __nv_bfloat16*)w; q.ldw = g->Cout;
        q.bias = e ? e->bias : nullptr; q.M = M; q.num_tiles = (int)((M + kTileM - 1) / kTileM); q.ncols = g->Cout;
        // staged rows padded so that the row pitch in words is 4 mod 8: the 16-byte row stores are conflict free
        int pitch = g->Cout * 2;
        if ((pitch / 4) % 8 != 4 && pitch + 16 <= 512) pitch += 16;
        q.stage_pitch = pitch;
        long long dimsO[2] = {g->Cout, M}, strO[2] = {1, g->Cout};
        int boxO[2] = {pitch / 2, kTileM}, esO[2] = {1, 1};
        if (make_tmap(&q.tmOut, y, 2, dimsO, strO, boxO, esO, 0, 2)) return -1;
        q.bits_out = e ? (uint16_t*)e->bits_out : nullptr;
        q.mask_bits = e ? (const uint16_t*)e->mask_bits : nullptr;
        q.bits_pitch = e ? e->bits_pitch : 0;
        q.bits_stage = (q.bits_out && M % kTileM == 0 && (q.bits_pitch * kTileM * 2) % 16 == 0) ? 1 : 0;
        q.act = e ? e->act : 0; q.leak = e ? e->leak : 0.f; q.mask_kind = e ? e->mask_kind : 0;
        q.im2col_out = nullptr;                       // (the filter gradient gathers for itself: img_wgrad_kernel)
        q.virt = virt ? 1 : 0;
        launch_img_fprop(q, st);
        return check_launch("conv2d_fprop(fused gather)");
      }
    }
    if (img_layout(g)) {
      const ImgConvGeom ig = img_geom(g);
      const int K16 = g->k * 16;
      __nv_bfloat16* A16 = (__nv_bfloat16*)workspace;
      const long long x_words = (long long)g->N * g->H * g->W * g->Cin / 2;
      __nv_bfloat16* Wt16 = (__nv_bfloat16*)((char*)workspace + img_ws_a(g));
      launch_img_im2col16((const __nv_bfloat16*)x, x_words, ig, A16, 1, st);
      launch_img_wpad16((const __nv_bfloat16*)w, g->Cout, g->Cout, g->k, g->k * g->Cin, Wt16, st);
      if (dense_gemm(A16, M, K16, K16, Wt16, g->Cout, K16, y, g->Cout, g->Cout, e, st)) return -1;
      return check_launch("conv2d_fprop(gather + GEMM)");
    }
    const int Kp = small_kp(g), kk = g->k * g->k * g->Cin;
    char* A = (char*)workspace;
    char* Wt = A + align256(M * Kp * 2);
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    nullptr, 0, 0.f, nullptr, 0, nullptr, 0};
    if (im2col_small(x, A, a, Kp, st)) return fail("im2col: k*k*Cin too large");
    wpad_transpose(w, Wt, kk, g->Cout, Kp, st);
    if (dense_gemm(A, M, Kp, Kp, Wt, g->Cout, Kp, y, g->Cout, g->Cout, e, st)) return -1;
    return check_launch("conv2d_fprop(im2col)");
  }
  if (route == 2) {
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    e ? e->bias : nullptr, e ? e->act : 0, e ? e->leak : 0.f, e ? e->mask_src : nullptr,
                    e ? e->mask_kind : 0, y, e ? e->out_f32 : 0};
    if (e && e->accumulate) return fail("small-channel fprop: accumulate unsupported");
    if (smallc_fprop(x, w, a, st)) return fail("smallc_fprop: unsupported shape");
    return check_launch("smallc_fprop");
  }
  if (g->Cout == 1 && !(e && (e->accumulate || e->mask_bits || e->bits_out))) {
    // one output channel (pix2pix PatchGAN head, hem/models/pix2pix.py:256): 1024 dot products, not a GEMM
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    e ? e->bias : nullptr, e ? e->act : 0, e ? e->leak : 0.f, e ? e->mask_src : nullptr,
                    e ? e->mask_kind : 0, y, e ? e->out_f32 : 0};
    if (smallout_fprop(x, w, a, st) == 0) return check_launch("smallout_fprop");
  }
  if (!w_t) return fail("conv2d_fprop: tensor-core path needs w_t");
  TapGemmParams p;
  memset(&p, 0, sizeof p);
  fill_epilogue(p, e);
  p.row_elems = g->Cout;
  pick_pixel_tile(g->Wo, g->Ho, kTileM, &p.bw, &p.bh, &p.bn);
  const int st_ = g->stride;
  if (p.bw * st_ > 256 || p.bh * st_ > 256) return fail("conv2d_fprop: tile exceeds TMA box limit");
  {
    long long dims[4] = {g->Cin, g->W, g->H, g->N};
    long long str[4] = {1, g->Cin, (long long)g->W * g->Cin, (long long)g->H * g->W * g->Cin};
    int box[4] = {kBlockK, p.bw * st_, p.bh * st_, p.bn};
    int es[4] = {1, st_, st_, 1};
    if (make_tmap(&p.tmA, x, 4, dims, str, box, es)) return -1;
  }
  p.a_rank = 4;
  p.ncols = g->Cout;
  const ConvPlan plan = conv_plan(g, 0);
  p.bn_tile = plan.bn_tile;
  p.cluster = tapgemm_cluster_size(p);
  {
    long long dims[2] = {g->Cin, (long long)g->k * g->k * g->Cout};
    long long str[2] = {1, g->Cin};
    int box[2] = {kBlockK, p.bn_tile / p.cluster};
    int es[2] = {1, 1};
    if (make_tmap(&p.tmB, w_t, 2, dims, str, box, es)) return -1;
    p.b_base = w_t; p.b_pitch_bytes = g->Cin * 2; p.b_rows_total = g->k * g->k * g->Cout; p.b_prefetch = weight_prefetch();
    p.tail_mode = pick_tail_mode(g->Cin);
    if (p.tail_mode) {
      const int tw = p.tail_mode == 1 ? 16 : 32;
      box[0] = tw;
      if (make_tmap(&p.tmB_tail, w_t, 2, dims, str, box, es, tw * 2)) return -1;
      long long adims[4] = {g->Cin, g->W, g->H, g->N};
      long long astr[4] = {1, g->Cin, (long long)g->W * g->Cin, (long long)g->H * g->W * g->Cin};
      int abox[4] = {tw, p.bw * st_, p.bh * st_, p.bn};
      int aes[4] = {1, st_, st_, 1};
      if (make_tmap(&p.tmA_tail, x, 4, adims, astr, abox, aes, tw * 2)) return -1;
    }
  }
  p.kchunks = cdiv(g->Cin, kBlockK);
  p.k_total = g->Cin;
  p.nphases = 1;
  p.phase_tap_begin[0] = 0;
  p.phase_tap_begin[1] = g->k * g->k;
  for (int r = 0; r < g->k; ++r)
    for (int c = 0; c < g->k; ++c) {
      const int t = r * g->k + c;
      p.tap_a_off[t][0] = c - g->pad_l;
      p.tap_a_off[t][1] = r - g->pad_t;
      p.tap_a_off[t][2] = 0;
      p.tap_b_row[t] = t * g->Cout;
    }
  p.a_mul[0][0] = st_; p.a_mul[1][1] = st_; p.a_mul[2][2] = 1;
  p.tiles_w = cdiv(g->Wo, p.bw); p.tiles_h = cdiv(g->Ho, p.bh); p.tiles_n = cdiv(g->N, p.bn);
  p.phase_ext_w[0] = g->Wo; p.phase_ext_h[0] = g->Ho; p.ext_n = g->N;
  p.phase_o_off[0] = 0;
  p.o_sw = g->Cout; p.o_sh = (long long)g->Wo * g->Cout; p.o_sn = (long long)g->Ho * g->Wo * g->Cout;
  p.dual = plan.dual;
  p.merge_tail = (p.tail_mode == 1 && p.kchunks >= 2 && !getenv("B200GAN_NO_MERGE_TAIL")) ? 1 : 0;
  p.cta2 = tapgemm_2sm(p.cluster, p.dual, p.tail_mode, p.merge_tail, p.bn_tile);
  p.stages = std::min(pick_stages(p.cta2 ? tapgemm_stage_bytes_2sm(p.bn_tile, p.merge_tail)
                                         : tapgemm_stage_bytes(p.dual, p.bn_tile, p.merge_tail)),
                      std::max(2, (p.kchunks - p.merge_tail) * g->k * g->k));
  p.out = y;
  if (run_split_k(p, g, 0, e, y, workspace, workspace_bytes, st)) return check_launch("conv2d_fprop(split-K)");
  setup_out_maps(p, e);
  launch_tapgemm(p, st);
  return check_launch("conv2d_fprop");
}

extern "C" int b200_conv2d_dgrad(const void* dy, const void* w, void* dx, const b200_conv_geom* g,
                                 const b200_epilogue* e, void* workspace, long long workspace_bytes, b200_stream s) {
  cudaStream_t st = (cudaStream_t)s;
  const int route = b200_conv2d_route(g, 1);
  if (route < 0) return route;
  if (route == 2 && workspace && small_gemm_ok(g)) {
    // T[M, k*k*Cin] = dy[M, Cout] . W[k*k*Cin, Cout]^T on tensor cores (fp32), then col2im + epilogue
    if (workspace_bytes < b200_conv2d_workspace_bytes(g, 1)) return fail("conv2d_dgrad: workspace too small");
    if (e && e->accumulate) return fail("small-channel dgrad: accumulate unsupported");
    const long long M = (long long)g->N * g->Ho * g->Wo;
    if (img_dgrad_supported(img_geom(g), g->Cout) && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
      // fused: tensor-core GEMM per image, the col2im gather reads its fp32 image from shared memory (img_conv.cu)
      ImgDgradParams q;
      memset(&q, 0, sizeof q);
      q.g = img_geom(g); q.cout = g->Cout; q.w = (const __nv_bfloat16*)w; q.ldw = g->Cout;
      long long dimsA[2] = {g->Cout, M}, strA[2] = {1, g->Cout};
      int boxA[2] = {64, kTileM}, boxT[2] = {16, kTileM}, es[2] = {1, 1};
      if (make_tmap(&q.tmA, dy, 2, dimsA, strA, boxA, es)) return -1;
      if (g->Cout % 64 && make_tmap(&q.tmA_tail, dy, 2, dimsA, strA, boxT, es, 32)) return -1;
      q.bias = e ? e->bias : nullptr; q.act = e ? e->act : 0; q.leak = e ? e->leak : 0.f;
      q.mask_src = e ? (const __nv_bfloat16*)e->mask_src : nullptr; q.mask_kind = e ? e->mask_kind : 0;
      q.out = dx; q.out_f32 = e ? e->out_f32 : 0;
      launch_img_dgrad(q, st);
      return check_launch("conv2d_dgrad(fused col2im)");
    }
    const int Kp = small_kp(g), kk = g->k * g->k * g->Cin;
    b200_epilogue te;
    memset(&te, 0, sizeof te);
    te.out_f32 = 1;
    // all Kp columns are produced (the weight rows past k*k*Cin are TMA zero fill): every 16-column chunk of the
    // epilogue is then full, the ragged 11-column tail used to cost more than the rest of the kernel
    if (dense_gemm(dy, M, g->Cout, g->Cout, w, kk, g->Cout, workspace, Kp, Kp, &te, st)) return -1;
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    e ? e->bias : nullptr, e ? e->act : 0, e ? e->leak : 0.f, e ? e->mask_src : nullptr,
                    e ? e->mask_kind : 0, dx, e ? e->out_f32 : 0};
    col2im_small((const float*)workspace, a, Kp, st);
    return check_launch("conv2d_dgrad(col2im)");
  }
  if (route == 2) {
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    e ? e->bias : nullptr, e ? e->act : 0, e ? e->leak : 0.f, e ? e->mask_src : nullptr,
                    e ? e->mask_kind : 0, dx, e ? e->out_f32 : 0};
    if (e && e->accumulate) return fail("small-channel dgrad: accumulate unsupported");
    if (smallc_dgrad(dy, w, a, st)) return fail("smallc_dgrad: unsupported shape");
    return check_launch("smallc_dgrad");
  }
  if (route == 3) {
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    e ? e->bias : nullptr, e ? e->act : 0, e ? e->leak : 0.f, e ? e->mask_src : nullptr,
                    e ? e->mask_kind : 0, dx, e ? e->out_f32 : 0};
    if (e && e->accumulate) return fail("small-output dgrad: accumulate unsupported");
    smallout_dgrad(dy, w, a, st);
    return check_launch("smallout_dgrad");
  }
  TapGemmParams p;
  memset(&p, 0, sizeof p);
  fill_epilogue(p, e);
  p.row_elems = g->Cin;
  const int st_ = g->stride;
  const int ext_w0 = cdiv(g->W, st_), ext_h0 = cdiv(g->H, st_);
  pick_pixel_tile(ext_w0, ext_h0, kTileM, &p.bw, &p.bh, &p.bn);
  {
    long long dims[4] = {g->Cout, g->Wo, g->Ho, g->N};
    long long str[4] = {1, g->Cout, (long long)g->Wo * g->Cout, (long long)g->Ho * g->Wo * g->Cout};
    int box[4] = {kBlockK, p.bw, p.bh, p.bn};
    int es[4] = {1, 1, 1, 1};
    if (make_tmap(&p.tmA, dy, 4, dims, str, box, es)) return -1;
  }
  p.a_rank = 4;
  p.ncols = g->Cin;
  const ConvPlan plan = conv_plan(g, 1);
  p.bn_tile = plan.bn_tile;
  p.cluster = tapgemm_cluster_size(p);
  {
    long long dims[2] = {g->Cout, (long long)g->k * g->k * g->Cin};
    long long str[2] = {1, g->Cout};
    int box[2] = {kBlockK, p.bn_tile / p.cluster};
    int es[2] = {1, 1};
    if (make_tmap(&p.tmB, w, 2, dims, str, box, es)) return -1;
    p.b_base = w; p.b_pitch_bytes = g->Cout * 2; p.b_rows_total = g->k * g->k * g->Cin; p.b_prefetch = weight_prefetch();
    p.tail_mode = pick_tail_mode(g->Cout);
    if (p.tail_mode) {
      const int tw = p.tail_mode == 1 ? 16 : 32;
      box[0] = tw;
      if (make_tmap(&p.tmB_tail, w, 2, dims, str, box, es, tw * 2)) return -1;
      long long adims[4] = {g->Cout, g->Wo, g->Ho, g->N};
      long long astr[4] = {1, g->Cout, (long long)g->Wo * g->Cout, (long long)g->Ho * g->Wo * g->Cout};
      int abox[4] = {tw, p.bw, p.bh, p.bn};
      int aes[4] = {1, 1, 1, 1};
      if (make_tmap(&p.tmA_tail, dy, 4, adims, astr, abox, aes, tw * 2)) return -1;
    }
  }
  p.kchunks = cdiv(g->Cout, kBlockK);
  p.k_total = g->Cout;
  // output parities, heaviest first so the tail of the grid is made of the cheap phases
  struct Ph { int ph, pw, nt; } phs[kMaxPhases];
  int np = 0;
  for (int ph = 0; ph < st_; ++ph)
    for (int pw = 0; pw < st_; ++pw) {
      int nr = 0, nc = 0;
      for (int r = 0; r < g->k; ++r) nr += ((ph + g->pad_t - r) % st_ == 0);
      for (int c = 0; c < g->k; ++c) nc += ((pw + g->pad_l - c) % st_ == 0);
      if (nr * nc == 0) return fail("conv2d_dgrad: output phase without taps (k < stride) unsupported");
      phs[np++] = {ph, pw, nr * nc};
    }
  std::stable_sort(phs, phs + np, [](const Ph& a, const Ph& b) { return a.nt > b.nt; });
  int t = 0;
  for (int i = 0; i < np; ++i) {
    const int ph = phs[i].ph, pw = phs[i].pw;
    p.phase_tap_begin[i] = t;
    for (int r = 0; r < g->k; ++r) {
      if ((ph + g->pad_t - r) % st_) continue;
      for (int c = 0; c < g->k; ++c) {
        if ((pw + g->pad_l - c) % st_) continue;
        p.tap_a_off[t][0] = (pw + g->pad_l - c) / st_;
        p.tap_a_off[t][1] = (ph + g->pad_t - r) / st_;
        p.tap_a_off[t][2] = 0;
        p.tap_b_row[t] = (r * g->k + c) * g->Cin;
        ++t;
      }
    }
    p.phase_ext_w[i] = cdiv(g->W - pw, st_);
    p.phase_ext_h[i] = cdiv(g->H - ph, st_);
    p.phase_o_off[i] = ((long long)ph * g->W + pw) * g->Cin;
  }
  p.phase_tap_begin[np] = t;
  p.nphases = np;
  p.a_mul[0][0] = 1; p.a_mul[1][1] = 1; p.a_mul[2][2] = 1;
  p.tiles_w = cdiv(ext_w0, p.bw); p.tiles_h = cdiv(ext_h0, p.bh); p.tiles_n = cdiv(g->N, p.bn);
  p.ext_n = g->N;
  p.o_sw = (long long)st_ * g->Cin; p.o_sh = (long long)st_ * g->W * g->Cin; p.o_sn = (long long)g->H * g->W * g->Cin;
  // two pixel tiles per CTA (and with them the 2-CTA kernels) whenever the lightest phase still has a real K loop
  p.dual = plan.dual;
  p.merge_tail = (p.tail_mode == 1 && p.kchunks >= 2 && !getenv("B200GAN_NO_MERGE_TAIL")) ? 1 : 0;
  p.cta2 = tapgemm_2sm(p.cluster, p.dual, p.tail_mode, p.merge_tail, p.bn_tile);
  p.stages = pick_stages(p.cta2 ? tapgemm_stage_bytes_2sm(p.bn_tile, p.merge_tail)
                                : tapgemm_stage_bytes(p.dual, p.bn_tile, p.merge_tail));
  p.out = dx;
  if (run_split_k(p, g, 1, e, dx, workspace, workspace_bytes, st)) return check_launch("conv2d_dgrad(split-K)");
  setup_out_maps(p, e);
  launch_tapgemm(p, st);
  return check_launch("conv2d_dgrad");
}

extern "C" int b200_conv2d_wgrad_folds_bias(const b200_conv_geom* g, int has_workspace) {
  return (b200_conv2d_route(g, 2) == 2 && has_workspace && img_layout(g)) ? 1 : 0;
}

extern "C" int b200_conv2d_wgrad(const void* x, const void* dy, float* dw, const b200_conv_geom* g, float alpha,
                                 void* workspace, long long workspace_bytes, int workspace_holds_im2col,
                                 b200_stream s) {
  return b200_conv2d_wgrad_bias(x, dy, dw, nullptr, g, alpha, workspace, workspace_bytes, workspace_holds_im2col, s);
}

extern "C" int b200_conv2d_wgrad_bias(const void* x, const void* dy, float* dw, float* dbias, const b200_conv_geom* g,
                                      float alpha, void* workspace, long long workspace_bytes,
                                      int workspace_holds_im2col, b200_stream s) {
  cudaStream_t st = (cudaStream_t)s;
  const int route = b200_conv2d_route(g, 2);
  if (route < 0) return route;
  if (dbias && !b200_conv2d_wgrad_folds_bias(g, workspace != nullptr))
    return fail("conv2d_wgrad_bias: this geometry does not fold the bias gradient (see b200_conv2d_wgrad_folds_bias)");
  if (route == 2 && workspace && small_gemm_ok(g)) {
    if (workspace_bytes < b200_conv2d_workspace_bytes(g, 2)) return fail("conv2d_wgrad: workspace too small");
    const long long M = (long long)g->N * g->Ho * g->Wo;
    if (img_layout(g)) {
      // rows gathered in the row-group layout (left by the fprop of the same input, or gathered here) x dy on the
      // tensor cores into a k*16 x Cout fp32 image, then folded into the TF layout; the ones column of the gather
      // makes row 15 the column sum of dy = the bias gradient
      if (img_wgrad_supported(img_geom(g), g->Cout) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) |
                                                        reinterpret_cast<uintptr_t>(dw)) & 15) == 0 &&
          (!dbias || (reinterpret_cast<uintptr_t>(dbias) & 15) == 0) && !getenv("B200GAN_NO_IMGWGRAD")) {
        // fused: the rows are gathered in shared memory inside the kernel (no im2col in HBM at all)
        ImgWgradParams q;
        memset(&q, 0, sizeof q);
        q.f.g = img_geom(g); q.f.x = (const __nv_bfloat16*)x;
        long long dimsD[2] = {g->Cout, M}, strD[2] = {1, g->Cout};
        int boxD[2] = {64, kTileM}, es[2] = {1, 1};
        if (make_tmap(&q.tmDy, dy, 2, dimsD, strD, boxD, es)) return -1;
        q.cout = g->Cout; q.dw = dw; q.ldo = g->Cout; q.dbias = dbias; q.alpha = alpha;
        launch_img_wgrad(q, st);
        return check_launch("conv2d_wgrad(fused gather)");
      }
      const int K16 = g->k * 16;
      __nv_bfloat16* A16 = (__nv_bfloat16*)workspace;
      float* T = (float*)((char*)workspace + img_ws_a(g) + img_ws_w(g));
      workspace_holds_im2col = 0;                      // (the fused fprop no longer leaves the gathered rows behind)
      if (!workspace_holds_im2col)
        launch_img_im2col16((const __nv_bfloat16*)x, (long long)g->N * g->H * g->W * g->Cin / 2, img_geom(g), A16, 1, st);
      if (cudaMemsetAsync(T, 0, (size_t)K16 * g->Cout * 4, st) != cudaSuccess) return check_launch("conv2d_wgrad(memset)");
      if (check_launch("conv2d_wgrad(gather)")) return -1;
      if (dense_wgrad(A16, K16, K16, dy, g->Cout, M, T, g->Cout, alpha, st)) return -1;
      if (check_launch("conv2d_wgrad(GEMM on gathered rows)")) return -1;
      launch_img_wgrad_fold(T, g->Cout, g->k, g->k * g->Cin, dw, g->Cout, dbias, st);
      return check_launch("conv2d_wgrad(gathered rows)");
    }
    const int Kp = small_kp(g), kk = g->k * g->k * g->Cin;
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    nullptr, 0, 0.f, nullptr, 0, nullptr, 0};
    // the fprop of the same input leaves im2col(x) at the start of its workspace: reuse it when told so
    // (not for the geometry whose fprop runs fused in the virtual-row form: it gathers in shared memory only)
    if (img_fprop_virtual_supported(img_geom(g), g->Cout)) workspace_holds_im2col = 0;
    if (!workspace_holds_im2col && im2col_small(x, workspace, a, Kp, st)) return fail("im2col: k*k*Cin too large");
    if (dense_wgrad(workspace, kk, Kp, dy, g->Cout, M, dw, g->Cout, alpha, st)) return -1;
    return check_launch("conv2d_wgrad(im2col)");
  }
  if (route == 2) {
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    nullptr, 0, 0.f, nullptr, 0, nullptr, 0};
    if (smallc_wgrad(x, dy, dw, a, alpha, st)) return fail("smallc_wgrad: unsupported shape");
    return check_launch("smallc_wgrad");
  }
  if (route == 3) {
    SmallConvArgs a{g->N, g->H, g->W, g->Cin, g->Ho, g->Wo, g->Cout, g->k, g->stride, g->pad_t, g->pad_l,
                    nullptr, 0, 0.f, nullptr, 0, nullptr, 0};
    if (smallout_wgrad(x, dy, dw, a, alpha, st)) return fail("smallout_wgrad: unsupported shape");
    return check_launch("smallout_wgrad");
  }
  WgradParams p;
  memset(&p, 0, sizeof p);
  const int st_ = g->stride;
  pick_pixel_tile(g->Wo, g->Ho, 64, &p.bw, &p.bh, &p.bn);
  if (p.bw * st_ > 256 || p.bh * st_ > 256) return fail("conv2d_wgrad: tile exceeds TMA box limit");
  {
    long long dims[4] = {g->Cin, g->W, g->H, g->N};
    long long str[4] = {1, g->Cin, (long long)g->W * g->Cin, (long long)g->H * g->W * g->Cin};
    int box[4] = {64, p.bw * st_, p.bh * st_, p.bn};
    int es[4] = {1, st_, st_, 1};
    if (make_tmap(&p.tmA, x, 4, dims, str, box, es)) return -1;
  }
  {
    long long dims[4] = {g->Cout, g->Wo, g->Ho, g->N};
    long long str[4] = {1, g->Cout, (long long)g->Wo * g->Cout, (long long)g->Ho * g->Wo * g->Cout};
    int box[4] = {64, p.bw, p.bh, p.bn};
    int es[4] = {1, 1, 1, 1};
    if (make_tmap(&p.tmB, dy, 4, dims, str, box, es)) return -1;
  }
  p.a_rank = 4; p.b_rank = 4;
  p.ntaps = g->k * g->k;
  for (int r = 0; r < g->k; ++r)
    for (int c = 0; c < g->k; ++c) {
      const int t = r * g->k + c;
      p.tap_a_off[t][0] = c - g->pad_l;
      p.tap_a_off[t][1] = r - g->pad_t;
      p.tap_a_off[t][2] = 0;
    }
  p.a_mul[0][0] = st_; p.a_mul[1][1] = st_; p.a_mul[2][2] = 1;
  p.chunks_w = cdiv(g->Wo, p.bw); p.chunks_h = cdiv(g->Ho, p.bh); p.chunks_n = cdiv(g->N, p.bn);
  p.total_chunks = p.chunks_w * p.chunks_h * p.chunks_n;
  p.Ca = g->Cin; p.Cb = g->Cout;
  p.m_tiles = cdiv(g->Cin, kTileM);
  p.bn_tile = pick_bn_tile(g->Cout);
  p.n_tiles = cdiv(g->Cout, p.bn_tile);
  p.nb_boxes = cdiv(p.bn_tile, 64);
  p.bn_tile_t = pick_bn_tile(p.Ca);       // N tile (over X channels) of the transposed 2-CTA kernel
  p.dual = wgrad_dual(p.m_tiles);
  p.l2_prefetch = l2_prefetch_distance();
  p.stages = pick_stages((2 * p.dual + p.nb_boxes) * 64 * 64 * 2);
  p.out = dw;
  p.out_tap_stride = (long long)g->Cin * g->Cout;
  p.ldo = g->Cout;
  p.alpha = alpha;
  const int base = cdiv(p.m_tiles, p.dual) * p.n_tiles * p.ntaps;
  int splits = pick_splits(base, p.total_chunks, p.bn_tile, p.dual);
  p.chunks_per_split = cdiv(p.total_chunks, splits);
  splits = cdiv(p.total_chunks, p.chunks_per_split);
  launch_wgrad(p, splits, st);
  return check_launch("conv2d_wgrad");
}

// ------------------------------------------------------------------------------------------------
// thin wrappers
// ------------------------------------------------------------------------------------------------
extern "C" const char* b200_last_error(void) { return g_err.c_str(); }

extern "C" int b200_set_tuning(const char* key, int value) {
  if (key && !strcmp(key, "dual_min_pct")) { set_dual_min_pct(value); return 0; }
  if (key && !strcmp(key, "bn_tile_cap")) { g_bn_cap = value < 0 ? 0 : value; return 0; }
  if (key && !strcmp(key, "tap_splits")) { g_force_splits = value; return 0; }
  if (key && !strcmp(key, "wgrad_min_chunks")) { set_wgrad_min_chunks(value); return 0; }
  return fail("set_tuning: unknown key");
}
extern "C" int b200_abi_version(void) { return 4; }
extern "C" int b200_device_check(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return fail("no CUDA device"); }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return fail("device is not compute capability 10.x (B200, sm_100a)");
  return 0;
}

#define WRAP(call, name) do { if ((call) != 0) return fail(name ": unsupported arguments"); return check_launch(name); } while (0)

extern "C" int b200_gemv_rows(const void* a, const void* w, const float* bias, float* out, int M, int K, int act,
                              float leak, b200_stream s) {
  WRAP(gemv_rows(a, w, bias, out, M, K, act, leak, (cudaStream_t)s), "gemv_rows");
}
extern "C" int b200_outer_mask(const float* g, const void* w, const void* mask, void* out, int M, int K, int kind,
                               float leak, b200_stream s) {
  WRAP(outer_mask(g, w, mask, out, M, K, kind, leak, (cudaStream_t)s), "outer_mask");
}
extern "C" int b200_bn_sums(const void* z, float* stats, long long R, int C, b200_stream s) {
  WRAP(bn_sums(z, stats, R, C, (cudaStream_t)s), "bn_sums");
}
extern "C" int b200_bn_apply(const void* z, const float* stats, const float* beta, void* out, long long R, int C,
                             float eps, int act, float leak, b200_stream s) {
  WRAP(bn_apply(z, stats, beta, out, R, C, eps, act, leak, (cudaStream_t)s), "bn_apply");
}
extern "C" int b200_bn_update_moving(const float* stats, long long R, int C, float* moving_mean, float* moving_var,
                                     float decay, int unbiased, b200_stream s) {
  WRAP(bn_update_moving(stats, R, C, moving_mean, moving_var, decay, unbiased, (cudaStream_t)s), "bn_update_moving");
}
extern "C" int b200_bn_bwd(const void* g, const void* z, const float* stats, float* bsum, void* dz, long long R, int C,
                           float eps, b200_stream s) {
  WRAP(bn_bwd(g, z, stats, bsum, dz, R, C, eps, (cudaStream_t)s), "bn_bwd");
}
extern "C" int b200_maskmul(const void* g, const void* a, void* out, long long n, int kind, float leak, b200_stream s) {
  WRAP(maskmul(g, a, out, n, kind, leak, (cudaStream_t)s), "maskmul");
}
extern "C" int b200_affine_act(const void* in, int in_type, void* out, int out_f32, long long n, float mul, float add,
                               int act, float leak, b200_stream s) {
  WRAP(affine_act(in, in_type, out, out_f32, n, mul, add, act, leak, (cudaStream_t)s), "affine_act");
}
extern "C" int b200_axpby(const void* a, int a_f32, float sa, const float* dev_sa, const void* b, int b_f32, float sb,
                          void* out, int out_f32, long long n, b200_stream s) {
  WRAP(axpby(a, a_f32, sa, dev_sa, b, b_f32, sb, out, out_f32, n, (cudaStream_t)s), "axpby");
}
extern "C" int b200_mul_add(const void* a, const void* b, const void* c, void* out, long long n, b200_stream s) {
  WRAP(mul_add(a, b, c, out, n, (cudaStream_t)s), "mul_add");
}
extern "C" int b200_fill_f32(float* out, long long n, float v, b200_stream s) {
  WRAP(fill_f32(out, n, v, (cudaStream_t)s), "fill_f32");
}
extern "C" int b200_interp(const void* x, const void* g, const float* alpha, void* out, int B, int D, b200_stream s) {
  WRAP(interp(x, g, alpha, out, B, D, (cudaStream_t)s), "interp");
}
extern "C" int b200_rowscale(const void* in, const float* s_row, float mul, float add, void* out, int B, int D,
                             b200_stream s) {
  WRAP(rowscale(in, s_row, mul, add, out, B, D, (cudaStream_t)s), "rowscale");
}
extern "C" int b200_dropout(const void* in, const float* u, void* out, long long n, float keep_prob, b200_stream s) {
  WRAP(dropout_apply(in, u, out, n, keep_prob, (cudaStream_t)s), "dropout");
}
extern "C" int b200_instnorm_fwd(const void* x, const float* scale, const float* shift, void* out, float* stats, int N,
                                 int HW, int C, float eps, b200_stream s) {
  WRAP(instnorm_fwd(x, scale, shift, out, stats, N, HW, C, eps, (cudaStream_t)s), "instnorm_fwd");
}
extern "C" int b200_instnorm_bwd(const void* g, const void* x, const float* stats, const float* scale, void* dx,
                                 float* dscale, float* dshift, int N, int HW, int C, b200_stream s) {
  WRAP(instnorm_bwd(g, x, stats, scale, dx, dscale, dshift, N, HW, C, (cudaStream_t)s), "instnorm_bwd");
}
extern "C" int b200_layout_convert(const void* in, int in_type, void* out, int to_nchw, int N, int C, int HW, float mul,
                                   float add, b200_stream s) {
  WRAP(layout_convert(in, in_type, out, to_nchw, N, C, HW, mul, add, (cudaStream_t)s), "layout_convert");
}
extern "C" int b200_summary_stats(const void* x, int x_type, long long n, float* out5, unsigned int* counts, int nb,
                                  b200_stream s) {
  WRAP(summary_stats(x, x_type, n, out5, counts, nb, (cudaStream_t)s), "summary_stats");
}
extern "C" int b200_montage(const void* x, int x_type, float* out, int m, int n, int H, int W, int C, float mul, float add,
                            b200_stream s) {
  WRAP(montage(x, x_type, out, m, n, H, W, C, mul, add, (cudaStream_t)s), "montage");
}
extern "C" int b200_slice_cols(const void* in, long long in_ld, int in_off, void* out, long long out_ld, int out_off,
                               long long rows, int cols, const void* mask, int mask_kind, float leak, b200_stream s) {
  WRAP(slice_cols(in, in_ld, in_off, out, out_ld, out_off, rows, cols, mask, mask_kind, leak, (cudaStream_t)s),
       "slice_cols");
}
extern "C" int b200_transpose_to_bf16(const void* in, int in_f32, void* out, int T, int A, int B, b200_stream s) {
  WRAP(transpose_to_bf16(in, in_f32, out, T, A, B, (cudaStream_t)s), "transpose_to_bf16");
}
extern "C" int b200_colsum(const void* x, const float* wrow, float* out, long long R, int C, float alpha,
                           b200_stream s) {
  WRAP(colsum(x, wrow, out, R, C, alpha, (cudaStream_t)s), "colsum");
}
extern "C" int b200_reduce_sum(const void* x, int x_f32, long long n, float* out, float alpha, int squared,
                               b200_stream s) {
  WRAP(reduce_sum(x, x_f32, n, out, alpha, squared, (cudaStream_t)s), "reduce_sum");
}
extern "C" int b200_wgan_loss(const float* sums, int B, int use_gp, float lambda, float* out4, b200_stream s) {
  WRAP(wgan_loss(sums, B, use_gp, lambda, out4, (cudaStream_t)s), "wgan_loss");
}
extern "C" int b200_eltloss(const void* a, int a_f32, const void* b, long long n, int kind, float label, float scale,
                            float gscale, float* out_sum, void* grad, int grad_f32, int mask_kind, float leak,
                            b200_stream s) {
  WRAP(eltloss(a, a_f32, b, n, kind, label, scale, gscale, out_sum, grad, grad_f32, mask_kind, leak, (cudaStream_t)s),
       "eltloss");
}
extern "C" int b200_philox(void* out, int out_f32, long long n, unsigned long long seed,
                           unsigned long long* dev_draw_counter, unsigned int stream_id, int normal, b200_stream s) {
  WRAP(philox_fill(out, out_f32, n, seed, dev_draw_counter, stream_id, normal, (cudaStream_t)s), "philox");
}
extern "C" int b200_optim_step(float* p, float* m, float* v, float* slot3, float* g, void* p_bf16, long long n, int kind,
                               float lr, float b1, float b2, float eps, float grad_scale, float clip, int zero_grad,
                               int* dev_step, b200_stream s) {
  WRAP(optim_step(p, m, v, slot3, g, p_bf16, n, kind, lr, b1, b2, eps, grad_scale, clip, zero_grad, dev_step,
                  (cudaStream_t)s), "optim_step");
}
extern "C" int b200_transpose_batch(const b200_transpose_entry* dev_table, int count, long long total_tiles,
                                    b200_stream s) {
  static_assert(sizeof(b200_transpose_entry) == sizeof(TransposeEntry), "table entry layout");
  WRAP(transpose_batch(dev_table, count, total_tiles, (cudaStream_t)s), "transpose_batch");
}
