// Fused image-side convolution kernels (see img_conv.cuh): the window gather of a <= 4-channel input happens in
// producer warps that write the swizzled A tiles directly, the bias rides in a spare K column, and the epilogue
// works on packed bf16 pairs.  Replaces im2col_k5c3 + wpad_transpose + smallk_kernel for the critic's first
// conv (models/gan.py:206 -> ops/layers.py:101) and for the input gradient of the generator's last deconv.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "img_conv.cuh"
#include "ptx.cuh"
#include "tc_gemm.cuh"
#include "epilogue.cuh"

namespace b200 {

namespace imgconv {

// 20 warps (5 per scheduler: 96 registers each): two producer groups of 4 warps (window staging, gather of one A row =
// output pixel per thread, thread 0 of a group issues its MMAs) taking alternate tiles, and 12 epilogue warps (they
// also build the weight tile at kernel start).  kImgProducers is the size of ONE group.
constexpr int kImgProducers = 128;
constexpr int kImgGroups = 2;
constexpr int kImgEpiWarps = 12;
constexpr int kImgCgs = kImgEpiWarps / 4;                // column groups: epilogue warps per TMEM lane quarter
constexpr int kImgEpiThreads = kImgEpiWarps * 32;
constexpr int kImgFirstEpiWarp = kImgGroups * kImgProducers / 32;
constexpr int kImgThreads = kImgGroups * kImgProducers + kImgEpiThreads;
constexpr int kAChunk0 = kTileM * 128;                   // K 0..63  (filter rows 0..3): 128 rows x 128 B, SWIZZLE_128B
constexpr int kATail = kTileM * 32;                      // K 64..79 (filter row 4):     128 rows x 32 B,  SWIZZLE_32B
constexpr int kASlot = kAChunk0 + kATail;
constexpr int kMaxSlots = 4;

struct ImgSmem {
  uint64_t b_ready;
  uint64_t a_empty[kMaxSlots];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__host__ __device__ inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// per output pixel: where its window starts and which of the 16 slots of a filter row are inside the image
struct RowCtx {
  int n, iy0, ix0;
  uint32_t wm[8];          // validity of slots (2i, 2i+1) as half-word masks
  bool ok;
};

__device__ __forceinline__ RowCtx make_row_ctx(const ImgConvGeom& g, long long pix, long long M) {
  RowCtx rc;
  rc.ok = pix < M;
  const int hw = g.Ho * g.Wo;
  const int p = rc.ok ? (int)pix : 0;
  rc.n = p / hw;
  const int rem = p - rc.n * hw;
  const int oy = rem / g.Wo;
  const int ox = rem - oy * g.Wo;
  rc.ix0 = ox * g.stride - g.pad_l;
  rc.iy0 = oy * g.stride - g.pad_t;
  const int j_lo = max(0, -rc.ix0) * g.Cin;
  const int j_hi = max(0, min(g.k, g.W - rc.ix0)) * g.Cin;
  const uint32_t em = rc.ok ? (((1u << j_hi) - 1u) & ~((1u << j_lo) - 1u)) : 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    rc.wm[i] = (((em >> (2 * i)) & 1u) ? 0xffffu : 0u) | (((em >> (2 * i + 1)) & 1u) ? 0xffff0000u : 0u);
  return rc;
}

// the 16 slots of filter row kh of one pixel: k*Cin contiguous input elements (2-byte aligned) fetched as aligned
// 32-bit words (gather_load) and funnel-shifted into place (gather_finish); slot 15 |= ones_bits.  Split in two so
// that a caller can put the loads of all filter rows in flight before it consumes the first.
struct GroupLoad {
  uint32_t wd[8];
  uint32_t sh;             // 0 | 16: the window starts on an even | odd element; 32: filter row outside the image
};
__device__ __forceinline__ GroupLoad gather_load(const ImgConvGeom& g, const uint32_t* __restrict__ xw, int x_words,
                                                 const RowCtx& rc, int kh) {
  GroupLoad L;
  const int iy = rc.iy0 + kh;
  if (rc.ok && iy >= 0 && iy < g.H) {
    const int e0 = ((rc.n * g.H + iy) * g.W + rc.ix0) * g.Cin;
    const int w0 = e0 >> 1;                     // floor (e0 is negative only for the first pixel's left padding)
    L.sh = (uint32_t)(e0 & 1) * 16u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = w0 + i;
      L.wd[i] = (idx >= 0 && idx < x_words) ? __ldg(xw + idx) : 0u;
    }
  } else {
    L.sh = 32u;
#pragma unroll
    for (int i = 0; i < 8; ++i) L.wd[i] = 0u;
  }
  return L;
}
__device__ __forceinline__ void gather_finish(const GroupLoad& L, const RowCtx& rc, uint32_t ones_bits, uint32_t* o) {
  const uint32_t sh = L.sh & 31u;               // (an outside row holds zeros: any shift will do)
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = __funnelshift_r(L.wd[i], i < 7 ? L.wd[i + 1] : 0u, sh) & rc.wm[i];
  o[7] |= ones_bits;
}

// ---- epilogue of one 16-column chunk: acc -> act -> sign word -> (mask) -> packed bf16 -> staged row.
// Every option is a template parameter: the epilogue warps are issue bound (measured: 2500 of the 3000 cycles a tile's
// epilogue takes are instruction issue, the rest TMEM read bandwidth), so nothing is decided at run time per chunk.
template <bool kMask, bool kAct, bool kWantBits>
__device__ __forceinline__ uint32_t img_chunk(const uint32_t* acc, float slope, float neg, uint32_t mbits, uint32_t dst) {
  uint32_t sign_word = 0;
  // two halves of 8 columns, each finished (packed and stored) before the next starts: keeps the live set small
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[8];
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[8 * h + j]);
    if (kAct) {                               // relu (slope 0) / lrelu: max(v, slope v)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], slope * v[j]);
    }
    if (kMask) {
      if (kWantBits) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sign_word |= (v[j] > 0.f ? 1u : 0u) << (8 * h + j);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ((mbits >> (8 * h + j)) & 1u) ? v[j] : v[j] * neg;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&b2);
    }
    if (kWantBits && !kMask) {
      // sign bits from the packed pairs: one compare per two elements (bf16(v) > 0 <=> v > 0 up to underflow);
      // element 2j -> bit 2j, element 2j+1 -> bit 2j+17 of w, folded below
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&pk[j]), zero2);
        sign_word |= m & ((1u << (8 * h + 2 * j)) | (1u << (8 * h + 2 * j + 17)));
      }
    }
    st_shared_v4(dst + 16 * h, pk[0], pk[1], pk[2], pk[3]);
  }
  if (kWantBits && !kMask) sign_word = (sign_word & 0xffffu) | (sign_word >> 16);
  return sign_word;
}

// this warp's chunks cg, cg+4, ... (my_n <= 4 of them) of one accumulator row, fully unrolled with two register buffers
// in ping-pong: the TMEM load (and mask word) of the next chunk is in flight while this one is processed.  All bases
// are for chunk cg; chunk i is 64 TMEM columns / 128 staged bytes / 4 sign words further.
// kBits: 0 no sign words, 1 staged in shared memory, 2 stored to global memory from registers (partial tiles)
template <bool kMask, bool kAct, int kBits>
__device__ __forceinline__ void img_row(uint32_t trow, int my_n, uint32_t srow, uint32_t sbits, const uint16_t* mrow,
                                        uint16_t* brow, bool row_ok, float slope, float neg) {
  // One accumulator buffer, no lookahead: the epilogue warps of a scheduler cover each other's TMEM-load latency,
  // and a second buffer in ping-pong made ptxas spill one of them in several of the option combinations
  // (a spilled buffer costs more than all the latency it could hide: local memory is an L2 round trip here).
  // All bases are for this warp's first chunk; its next chunk is kImgCgs chunks further.
#pragma unroll 1
  for (int i = 0; i < my_n; ++i) {
    uint32_t v[16];
    tmem_ld16(trow + 16 * kImgCgs * i, v);
    uint32_t m = 0;
    if (kMask) m = row_ok ? (uint32_t)__ldg(mrow + kImgCgs * i) : 0u;
    tmem_ld_wait16(v);
    const uint32_t sw = img_chunk<kMask, kAct, kBits != 0>(v, slope, neg, m, srow + 32 * kImgCgs * i);
    if (kBits == 1) st_shared_u16(sbits + 2 * kImgCgs * i, (uint16_t)sw);
    if (kBits == 2) { if (row_ok) brow[kImgCgs * i] = (uint16_t)sw; }
  }
}

// ---- input window in shared memory: the padded image rows a tile needs, zero padding included, so that the gather
// is mask free.  Row slot g holds padded row G_first + g, G = n * Hp + (iy + pad_t) with Hp = (Ho-1)*stride + k; the
// data of a row starts at element win_off (a multiple of 8: 16-byte aligned for cp.async), the columns before /
// after it are never written and stay zero.
// n / d for 0 <= n < 2^31 with a host-made reciprocal (ImgFpropParams::div_*): one wide multiply instead of the
// ~40-instruction (32-bit) or ~100-instruction (64-bit) division sequence, several of which sat on the producers'
// critical path per tile
__device__ __forceinline__ int fdiv(int n, ImgDiv d) { return (int)(((unsigned long long)(unsigned)n * d.mul) >> d.shr); }

struct TileSpan { int g_first, n_rows; };
__device__ __forceinline__ TileSpan tile_span(const ImgFpropParams& p, int tile) {
  const int hw = p.g.Ho * p.g.Wo;
  const int p0 = tile * kTileM;
  const int p1 = min((int)p.M, p0 + kTileM) - 1;
  const int n0 = fdiv(p0, p.div_hw), oy0 = fdiv(p0 - n0 * hw, p.div_wo);
  const int n1 = fdiv(p1, p.div_hw), oy1 = fdiv(p1 - n1 * hw, p.div_wo);
  TileSpan t;
  t.g_first = n0 * p.win_hp + oy0 * p.g.stride;
  t.n_rows = n1 * p.win_hp + oy1 * p.g.stride + p.g.k - t.g_first;
  return t;
}

// The 128 producer threads stage a tile's window: piece q = tid + 128*i covers bytes [j*piece, +piece) of window row
// g, (g, j) = divmod(q, pieces per row), piece = 16 (or 4) bytes.  Split in two so that the global loads of the NEXT
// tile's window are in flight while the current tile is gathered: stage_load (global -> registers; rows outside the
// image give zeros) and stage_store (registers -> window).  Host guarantees rows x pieces-per-row <= 128 * kWinRegs.
constexpr int kWinRegs = 6;
struct WinRegs { uint4 v[kWinRegs]; };

__device__ __forceinline__ const char* stage_src(const ImgFpropParams& p, const TileSpan& t, int q, int total,
                                                 int piece_bytes) {
  if (q >= total) return nullptr;
  const int g = fdiv(q, p.div_ppr), j = q - g * p.win_ppr;
  const int G = t.g_first + g;
  const int n = fdiv(G, p.div_hp);
  const int iy = G - n * p.win_hp - p.g.pad_t;
  if (n >= p.g.N || iy < 0 || iy >= p.g.H) return nullptr;
  return reinterpret_cast<const char*>(p.x) + (size_t)(n * p.g.H + iy) * (p.g.W * p.g.Cin * 2) + j * piece_bytes;
}
__device__ __forceinline__ void stage_load(const ImgFpropParams& p, const TileSpan& t, int tid, WinRegs& R) {
  const int total = t.n_rows * p.win_ppr;
#pragma unroll
  for (int i = 0; i < kWinRegs; ++i) {
    R.v[i] = make_uint4(0u, 0u, 0u, 0u);
    if (p.win_vec16) {
      const char* src = stage_src(p, t, tid + kImgProducers * i, total, 16);
      if (src) R.v[i] = __ldg(reinterpret_cast<const uint4*>(src));
    } else {                                              // 4-byte pieces, four per register slot
      const char* s0 = stage_src(p, t, tid + kImgProducers * (4 * i + 0), total, 4);
      const char* s1 = stage_src(p, t, tid + kImgProducers * (4 * i + 1), total, 4);
      const char* s2 = stage_src(p, t, tid + kImgProducers * (4 * i + 2), total, 4);
      const char* s3 = stage_src(p, t, tid + kImgProducers * (4 * i + 3), total, 4);
      if (s0) R.v[i].x = __ldg(reinterpret_cast<const uint32_t*>(s0));
      if (s1) R.v[i].y = __ldg(reinterpret_cast<const uint32_t*>(s1));
      if (s2) R.v[i].z = __ldg(reinterpret_cast<const uint32_t*>(s2));
      if (s3) R.v[i].w = __ldg(reinterpret_cast<const uint32_t*>(s3));
    }
  }
}
__device__ __forceinline__ void stage_store1(const ImgFpropParams& p, uint32_t win, int q, int total, uint32_t v) {
  if (q >= total) return;
  const int g = fdiv(q, p.div_ppr), j = q - g * p.win_ppr;
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(win + (uint32_t)(g * p.win_pitch + p.win_off) * 2 + j * 4), "r"(v) : "memory");
}
__device__ __forceinline__ void stage_store(const ImgFpropParams& p, const TileSpan& t, uint32_t win, int tid,
                                            const WinRegs& R) {
  const int total = t.n_rows * p.win_ppr;
#pragma unroll
  for (int i = 0; i < kWinRegs; ++i) {
    if (p.win_vec16) {
      const int q = tid + kImgProducers * i;
      if (q < total) {
        const int g = fdiv(q, p.div_ppr), j = q - g * p.win_ppr;
        st_shared_v4(win + (uint32_t)(g * p.win_pitch + p.win_off) * 2 + j * 16, R.v[i].x, R.v[i].y, R.v[i].z, R.v[i].w);
      }
    } else {
      stage_store1(p, win, tid + kImgProducers * (4 * i + 0), total, R.v[i].x);
      stage_store1(p, win, tid + kImgProducers * (4 * i + 1), total, R.v[i].y);
      stage_store1(p, win, tid + kImgProducers * (4 * i + 2), total, R.v[i].z);
      stage_store1(p, win, tid + kImgProducers * (4 * i + 3), total, R.v[i].w);
    }
  }
}
// the image rows of a window are one contiguous range of x: pull the lines of a tile two steps ahead into L2 (the
// input was usually evicted by the output stream of the previous launches, and DRAM latency under that store
// traffic is longer than one tile)
__device__ __forceinline__ void prefetch_window(const ImgFpropParams& p, const TileSpan& t, int tid) {
  const int row_bytes = p.g.W * p.g.Cin * 2;
  int n0 = fdiv(t.g_first, p.div_hp), iy0 = t.g_first - n0 * p.win_hp - p.g.pad_t;
  const int Gl = t.g_first + t.n_rows - 1;
  int n1 = fdiv(Gl, p.div_hp), iy1 = Gl - n1 * p.win_hp - p.g.pad_t;
  if (iy0 < 0) iy0 = 0;
  if (iy0 >= p.g.H) { iy0 = 0; ++n0; }
  if (iy1 >= p.g.H) iy1 = p.g.H - 1;
  if (iy1 < 0) { iy1 = p.g.H - 1; --n1; }
  if (n1 >= p.g.N) { n1 = p.g.N - 1; iy1 = p.g.H - 1; }
  const uint32_t lo = (uint32_t)(n0 * p.g.H + iy0) * row_bytes, hi = (uint32_t)(n1 * p.g.H + iy1 + 1) * row_bytes;
  const char* base = reinterpret_cast<const char*>(p.x);
  for (uint32_t o = (lo & ~127u) + (uint32_t)tid * 128u; o < hi; o += 32u * 128u)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
}

// One A row (output pixel r of `tile`): its kK filter rows are read from the window as aligned words, shifted into
// place and written as 16-slot groups into the swizzled operand tile(s): groups 0..3 -> the SWIZZLE_128B tile at a0,
// group 4 -> the SWIZZLE_32B tail tile at a1, or (kTail128) the first 32 bytes of the rows of a second SWIZZLE_128B
// tile at a1.  Slot 15 of groups 0 / 1 = 1.0 (bias pair of the fprop; column sum of dy in the filter gradient).
template <int kK, bool kTail128>
__device__ __forceinline__ void gather_row(const ImgFpropParams& p, const TileSpan& ts, uint32_t win, int tile, int r,
                                           uint32_t a0, uint32_t a1, const uint32_t* wm, bool full14) {
  const uint32_t ones = 0x3F800000u;
  const int hw = p.g.Ho * p.g.Wo;
  const int off0 = p.win_off - p.g.pad_l * p.g.Cin;
  const long long pix = (long long)tile * kTileM + r;
  const bool ok = pix < p.M;
  const int pi = ok ? (int)pix : 0;
  const int n = fdiv(pi, p.div_hw);
  const int rem = pi - n * hw;
  const int oy = fdiv(rem, p.div_wo);
  const int ox = rem - oy * p.g.Wo;
  const int e = (n * p.win_hp + oy * p.g.stride - ts.g_first) * p.win_pitch + off0 + ox * p.g.stride * p.g.Cin;
  const uint32_t src = win + (uint32_t)(e >> 1) * 4;
  const uint32_t sh = (uint32_t)(e & 1) * 16u;
  const uint32_t row_pitch = (uint32_t)p.win_pitch * 2;
#pragma unroll
  for (int kh = 0; kh < kK; ++kh) {
    uint32_t wd[9], o[8];
    if (kK == 5 && kh == 4 && p.virt) {             // virtual fifth row (k = 4, k*Cin = 16): zeros + the ones pair
#pragma unroll
      for (int i = 0; i < 7; ++i) o[i] = 0u;
      o[7] = 0x3F803F80u;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wd[i]) : "r"(src + kh * row_pitch + i * 4));
      wd[8] = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = __funnelshift_r(wd[i], wd[i + 1], sh);
        if (i == 7 || !full14) o[i] &= wm[i];         // k*Cin >= 14: only the last word holds empty slots
        if (!ok) o[i] = 0u;
      }
      if (kh < 2 && !p.virt) o[7] |= ones;
    }
    if (kh < 4) {
      const uint32_t base = a0 + r * 128;
      st_shared_v4(base + (((2 * kh) ^ (r & 7)) << 4), o[0], o[1], o[2], o[3]);
      st_shared_v4(base + (((2 * kh + 1) ^ (r & 7)) << 4), o[4], o[5], o[6], o[7]);
    } else if (kTail128) {
      const uint32_t base = a1 + r * 128;
      st_shared_v4(base + ((0 ^ (r & 7)) << 4), o[0], o[1], o[2], o[3]);
      st_shared_v4(base + ((1 ^ (r & 7)) << 4), o[4], o[5], o[6], o[7]);
    } else {
      const uint32_t base = a1 + r * 32;
      const uint32_t sw = (r >> 2) & 1;
      st_shared_v4(base + (sw << 4), o[0], o[1], o[2], o[3]);
      st_shared_v4(base + ((sw ^ 1) << 4), o[4], o[5], o[6], o[7]);
    }
    if (p.im2col_out && ok) {
      uint4* dst = reinterpret_cast<uint4*>(p.im2col_out + (size_t)pix * (kK * 16) + kh * 16);
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
  }
}

template <int kK>
__global__ void __launch_bounds__(kImgThreads, 1) img_fprop_kernel(const __grid_constant__ ImgFpropParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = warp >> 2;                              // 0 / 1: producer groups (alternate tiles), >= 2: epilogue
  const int gtid = threadIdx.x & (kImgProducers - 1);
  constexpr int kSteps0 = kK < 4 ? kK : 4;              // 16-wide MMA K steps in the SWIZZLE_128B chunk
  constexpr bool kTail = kK == 5;
  const int kcin = p.g.k * p.g.Cin;

  const int b0_bytes = p.ncols * 128;
  const int bt_bytes = kTail ? align_up(p.ncols * 32, 1024) : 0;
  const int out_bytes = align_up(kTileM * p.stage_pitch, 1024);
  const int bits_bytes = p.bits_stage ? align_up(kTileM * p.bits_pitch * 2, 128) : 0;
  const int win_bytes = align_up(p.win_rows * p.win_pitch * 2, 128);
  uint8_t* smem_b0 = smem;
  uint8_t* smem_bt = smem_b0 + b0_bytes;
  uint8_t* smem_a = smem_bt + bt_bytes;
  uint8_t* smem_out = smem_a + (size_t)p.slots * kASlot;
  uint8_t* smem_bits = smem_out + 2 * out_bytes;
  uint8_t* smem_win = smem_bits + 2 * bits_bytes;
  ImgSmem* ps = reinterpret_cast<ImgSmem*>(smem_win + 2 * kImgGroups * win_bytes);
  const int gstep = kImgGroups * gridDim.x;               // tile stride of one producer group
  const int tile0 = blockIdx.x + grp * gridDim.x;         // (producer groups) first tile

  // producers: the first window's global loads go out before any setup (DRAM latency under the prologue)
  TileSpan ts_next = {0, 0};
  WinRegs wr;
  if (grp < kImgGroups) {
    if (tile0 < p.num_tiles) {
      ts_next = tile_span(p, tile0);
      stage_load(p, ts_next, gtid, wr);
    }
    // all windows start out zero: the padding columns are never written again
    for (int i = threadIdx.x * 16; i < 2 * kImgGroups * win_bytes; i += kImgGroups * kImgProducers * 16)
      st_shared_v4(smem_u32(smem_win) + i, 0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmOut);
    mbar_init(smem_u32(&ps->b_ready), kImgEpiThreads);
    for (int s = 0; s < p.slots; ++s) mbar_init(smem_u32(&ps->a_empty[s]), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&ps->acc_full[a]), 1);
      mbar_init(smem_u32(&ps->acc_empty[a]), kImgEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<2 * kTmemCols>(smem_u32(&ps->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;

  if (grp < kImgGroups) {
    // ------------------------------------------------------------------ producers: window staging, gather, MMA issue.
    // Group g takes tiles g, g+2, ... of this CTA: its own window pair, A slots and accumulator (acc = g), so the two
    // groups only meet at the tensor pipe.  One group alone is a ~4000-cycle dependent chain per tile.
    const int r = gtid;                                    // the A row (output pixel of the tile) of this thread
    uint32_t wm[8];                                        // slots >= k*Cin are zero
#pragma unroll
    for (int i = 0; i < 8; ++i) wm[i] = (2 * i < kcin ? 0xffffu : 0u) | (2 * i + 1 < kcin ? 0xffff0000u : 0u);
    const bool full14 = kcin >= 14;
    const uint32_t win0 = smem_u32(smem_win) + grp * 2 * win_bytes;
    const int bar = 2 + grp;
    const int spg = p.slots / kImgGroups;                  // A slots per group
    // tile spans run two tiles ahead: ts_next (staged into registers now) and ts_next2 (its lines pulled into L2)
    TileSpan ts_next2 = ts_next;
    if (tile0 < p.num_tiles) {
      if (tile0 + gstep < p.num_tiles) {
        ts_next2 = tile_span(p, tile0 + gstep);
        if ((warp & 3) == 3) prefetch_window(p, ts_next2, lane);
      }
      stage_store(p, ts_next, win0, gtid, wr);            // (zeroed before the CTA barrier above)
    }
    named_barrier(bar, kImgProducers);

    const uint32_t idesc = make_idesc_bf16(kTileM, p.ncols, 0, 0);
    const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem_b0), 16, 1024);
    const uint64_t bdesct = make_smem_desc(smem_u32(smem_bt), 16, 256, 6);
    const uint32_t dacc = tmem + grp * kTmemCols;          // this group's accumulator
    int sl = 0, j = 0;
    uint32_t par = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += gstep, ++j) {
      const uint32_t win = win0 + (j & 1) * win_bytes;
      const TileSpan ts = ts_next;
      const int next = tile + gstep;
      if (next < p.num_tiles) {
        ts_next = ts_next2;
        stage_load(p, ts_next, gtid, wr);
        if (next + gstep < p.num_tiles) {
          ts_next2 = tile_span(p, next + gstep);
          if ((warp & 3) == 3) prefetch_window(p, ts_next2, lane);
        }
      }
      const int s = grp * spg + sl;
      mbar_wait(smem_u32(&ps->a_empty[s]), par ^ 1);
      const uint32_t a_addr = smem_u32(smem_a) + (uint32_t)s * kASlot;
      gather_row<kK, false>(p, ts, win, tile, r, a_addr, a_addr + kAChunk0, wm, full14);
      fence_proxy_async_smem();                            // A rows (generic proxy) -> visible to the tensor core
      if (next < p.num_tiles) stage_store(p, ts_next, win0 + ((j + 1) & 1) * win_bytes, gtid, wr);
      named_barrier(bar, kImgProducers);
      if (gtid == 0) {
        if (j == 0) mbar_wait(smem_u32(&ps->b_ready), 0);
        mbar_wait(smem_u32(&ps->acc_empty[grp]), (uint32_t)(j & 1) ^ 1);   // the epilogue has drained this accumulator
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
#pragma unroll
        for (int ks = 0; ks < kSteps0; ++ks) umma_bf16(dacc, adesc + 2 * ks, bdesc0 + 2 * ks, idesc, ks ? 1u : 0u);
        if (kTail) umma_bf16(dacc, make_smem_desc(a_addr + kAChunk0, 16, 256, 6), bdesct, idesc, 1u);
        umma_commit(smem_u32(&ps->a_empty[s]));
        umma_commit(smem_u32(&ps->acc_full[grp]));
      }
      __syncwarp();
      if (++sl == spg) { sl = 0; par ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int et = threadIdx.x - kImgGroups * kImgProducers;
    // first the weights: B[n][kh*16 + j] = w[(kh*kcin + j)][n]  (K-major, swizzled like a TMA load would leave it);
    // slot 15 of groups 0 / 1 = bias split into a bf16 high and low part (their A column is 1.0)
    {
      const int K16 = kK * 16;
      const int n8s = p.ncols >> 3;
      for (int item = et; item < n8s * K16; item += kImgEpiThreads) {
        const int kk = item % K16, n8 = item / K16;
        const int kh = kk >> 4, j = kk & 15;
        uint32_t v[4] = {0u, 0u, 0u, 0u};
        const bool vrow = p.virt && kh >= 4;                 // the virtual row group: bias pair in slots 14 / 15 only
        const bool bias_hi = p.virt ? (vrow && j == 14) : (j == 15 && kh == 0);
        const bool bias_lo = p.virt ? (vrow && j == 15) : (j == 15 && kh == 1);
        if (j < kcin && !vrow) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)(kh * kcin + j) * p.ldw + n8 * 8));
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else if (p.bias && (bias_hi || bias_lo)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t h2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float bv = __ldg(p.bias + n8 * 8 + 2 * i + h);
              const __nv_bfloat16 hi = __float2bfloat16(bv);
              const __nv_bfloat16 val = bias_hi ? hi : __float2bfloat16(bv - __bfloat162float(hi));
              h2[h] = (uint32_t)__bfloat16_as_ushort(val);
            }
            v[i] = h2[0] | (h2[1] << 16);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = n8 * 8 + i;
          const uint16_t val = (uint16_t)((i & 1) ? (v[i >> 1] >> 16) : (v[i >> 1] & 0xffffu));
          uint32_t dst;
          if (kh < 4) dst = smem_u32(smem_b0) + n * 128 + ((((kk >> 3)) ^ (n & 7)) << 4) + (kk & 7) * 2;
          else dst = smem_u32(smem_bt) + n * 32 + (((((kk - 64) >> 3)) ^ ((n >> 2) & 1)) << 4) + (kk & 7) * 2;
          st_shared_u16(dst, val);
        }
      }
      fence_proxy_async_smem();                             // generic-proxy writes -> visible to the tensor core's reads
      mbar_arrive(smem_u32(&ps->b_ready));
    }
    const int q = warp & 3;                                // TMEM lane quarter this warp may read
    const int cg = (warp - kImgFirstEpiWarp) >> 2;         // 0 .. kImgCgs-1: takes chunks cg, cg + kImgCgs, ...
    const int r = q * 32 + lane;
    const int nchunks = p.ncols >> 4;
    const bool issuer = et == 0;
    const int my_n = cg < nchunks ? (nchunks - cg + kImgCgs - 1) / kImgCgs : 0;
    const int bits_mode = p.bits_out ? (p.bits_stage ? 1 : 2) : 0;
    const float slope = p.act == ACT_NONE ? 1.f : (p.act == ACT_RELU ? 0.f : p.leak);
    const float neg = p.mask_kind == ACT_LRELU ? p.leak : 0.f;
    int acc = 0, so = 0;
    uint32_t accpar = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long row = (long long)tile * kTileM + r;
      const bool row_ok = row < p.M;
      mbar_wait(smem_u32(&ps->acc_full[acc]), accpar);
      tc_fence_after();
      if (issuer) bulk_wait_read1();                       // the store that last used this stage has read it
      named_barrier(1, kImgEpiThreads);
      // bases of this warp's first chunk (cg): TMEM column, staged row, staged / global sign words, mask words
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + acc * kTmemCols + cg * 16;
      const uint32_t srow = smem_u32(smem_out) + so * out_bytes + r * p.stage_pitch + cg * 32;
      const uint32_t sbits = smem_u32(smem_bits) + so * bits_bytes + r * p.bits_pitch * 2 + cg * 2;
      const uint16_t* mrow = p.mask_bits ? p.mask_bits + (row_ok ? row : 0) * p.bits_pitch + cg : nullptr;
      uint16_t* brow = p.bits_out ? p.bits_out + (row_ok ? row : 0) * p.bits_pitch + cg : nullptr;
#define IMG_ROW(M_, A_, B_) img_row<M_, A_, B_>(trow, my_n, srow, sbits, mrow, brow, row_ok, slope, neg)
      if (p.mask_bits) {
        if (bits_mode == 0) IMG_ROW(true, true, 0); else if (bits_mode == 1) IMG_ROW(true, true, 1); else IMG_ROW(true, true, 2);
      } else if (p.act != ACT_NONE) {
        if (bits_mode == 0) IMG_ROW(false, true, 0); else if (bits_mode == 1) IMG_ROW(false, true, 1); else IMG_ROW(false, true, 2);
      } else {
        if (bits_mode == 0) IMG_ROW(false, false, 0); else if (bits_mode == 1) IMG_ROW(false, false, 1); else IMG_ROW(false, false, 2);
      }
#undef IMG_ROW
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ps->acc_empty[acc]));
      fence_proxy_async_smem();                            // staged rows (generic proxy) -> visible to the bulk copy engine
      named_barrier(1, kImgEpiThreads);
      if (issuer) {
        tma_store_2d(&p.tmOut, smem_u32(smem_out) + so * out_bytes, 0, tile * kTileM);
        if (p.bits_stage)
          bulk_store_1d(p.bits_out + (size_t)tile * kTileM * p.bits_pitch, smem_u32(smem_bits) + so * bits_bytes,
                        (uint32_t)(kTileM * p.bits_pitch * 2));
        bulk_commit();
      }
      so ^= 1;
      acc ^= 1;
      if (acc == 0) accpar ^= 1;
    }
    if (issuer) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<2 * kTmemCols>(tmem);
}

// =============================================================================================
// Fused image-side wgrad (see ImgWgradParams)
// =============================================================================================
// Two producer groups of 4 warps take alternate tiles (a single group's gather is a ~4000-cycle dependent instruction
// chain per tile, longer than the dy traffic of the tile): warps 0-3 / 4-7 producers (thread 0 of each group issues its
// TMA loads and MMAs, into the group's own accumulator), warps 8-11 the final reduction.
constexpr int kWgGroups = 2;
constexpr int kWgThreads = kWgGroups * kImgProducers + 128;
struct WgSmem {
  uint64_t full[4], empty[4];
  uint64_t acc_full;
  uint32_t tmem_base;
};

template <int kK>
__global__ void __launch_bounds__(kWgThreads, 1) img_wgrad_kernel(const __grid_constant__ ImgWgradParams pw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const ImgFpropParams& p = pw.f;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = warp >> 2;                              // 0 / 1: producer groups, 2: reduction warps
  const int gtid = threadIdx.x & (kImgProducers - 1);
  const int kcin = p.g.k * p.g.Cin;
  const int a_bytes = 2 * kAChunk0;                       // two 64-slot blocks: filter rows 0..3 | row 4 (+ zeros)
  const int b_bytes = pw.nblocks * kAChunk0;              // dy: 64-channel boxes of 128 pixels
  const int stage_bytes = a_bytes + b_bytes;
  const int spg = pw.stages / kWgGroups;                  // stages per group
  const int win_bytes = align_up(p.win_rows * p.win_pitch * 2, 128);
  uint8_t* smem_win = smem + (size_t)pw.stages * stage_bytes;
  WgSmem* ps = reinterpret_cast<WgSmem*>(smem_win + 2 * kWgGroups * win_bytes);
  const int gstep = kWgGroups * gridDim.x;                // tile stride of one group
  const int tile0 = blockIdx.x + grp * gridDim.x;         // (producer groups) first tile

  TileSpan ts_next = {0, 0};
  WinRegs wr;
  if (grp < kWgGroups) {
    if (tile0 < p.num_tiles) {
      ts_next = tile_span(p, tile0);
      stage_load(p, ts_next, gtid, wr);
    }
    for (int i = threadIdx.x * 16; i < 2 * kWgGroups * win_bytes; i += kWgGroups * kImgProducers * 16)
      st_shared_v4(smem_u32(smem_win) + i, 0u, 0u, 0u, 0u);
    // the second A block only ever receives filter row 4 (its first 32 bytes per row): the rest stays zero
    for (int s = 0; s < pw.stages; ++s)
      for (int i = threadIdx.x * 16; i < kAChunk0; i += kWgGroups * kImgProducers * 16)
        st_shared_v4(smem_u32(smem) + s * stage_bytes + kAChunk0 + i, 0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&pw.tmDy);
    for (int s = 0; s < pw.stages; ++s) {
      mbar_init(smem_u32(&ps->full[s]), 1);
      mbar_init(smem_u32(&ps->empty[s]), 1);
    }
    mbar_init(smem_u32(&ps->acc_full), kWgGroups);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<2 * kTmemCols>(smem_u32(&ps->tmem_base));
  fence_proxy_async_smem();                                // (the zeroed A blocks, for the tensor core)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;

  if (grp < kWgGroups) {
    const int r = gtid;
    uint32_t wm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wm[i] = (2 * i < kcin ? 0xffffu : 0u) | (2 * i + 1 < kcin ? 0xffff0000u : 0u);
    const bool full14 = kcin >= 14;
    const uint32_t win0 = smem_u32(smem_win) + grp * 2 * win_bytes;
    const int bar = 2 + grp;
    TileSpan ts_next2 = ts_next;
    if (tile0 < p.num_tiles) {
      if (tile0 + gstep < p.num_tiles) {
        ts_next2 = tile_span(p, tile0 + gstep);
        if ((warp & 3) == 3) prefetch_window(p, ts_next2, lane);
      }
      stage_store(p, ts_next, win0, gtid, wr);
    }
    named_barrier(bar, kImgProducers);

    // MN-major operands: 64-element blocks kAChunk0 apart (LBO), 8-pixel groups 1024 B apart (SBO), 16 pixels per MMA
    const uint32_t idesc = make_idesc_bf16(kTileM, pw.cout, 1, 1);
    const uint32_t dacc = tmem + grp * kTmemCols;          // this group's accumulator
    int sl = 0, j = 0;
    uint32_t par = 0, accum = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += gstep, ++j) {
      const uint32_t win = win0 + (j & 1) * win_bytes;
      const TileSpan ts = ts_next;
      const int next = tile + gstep;
      if (next < p.num_tiles) {
        ts_next = ts_next2;
        stage_load(p, ts_next, gtid, wr);
        if (next + gstep < p.num_tiles) {
          ts_next2 = tile_span(p, next + gstep);
          if ((warp & 3) == 3) prefetch_window(p, ts_next2, lane);
        }
      }
      const int s = grp * spg + sl;
      mbar_wait(smem_u32(&ps->empty[s]), par ^ 1);
      const uint32_t a_addr = smem_u32(smem) + (uint32_t)s * stage_bytes;
      if (gtid == 0) {
        const uint32_t full = smem_u32(&ps->full[s]);
        mbar_arrive_expect_tx(full, b_bytes);
        for (int b = 0; b < pw.nblocks; ++b) tma_load_2d(a_addr + a_bytes + b * kAChunk0, &pw.tmDy, full, b * 64, tile * kTileM);
      }
      gather_row<kK, true>(p, ts, win, tile, r, a_addr, a_addr + kAChunk0, wm, full14);
      fence_proxy_async_smem();
      if (next < p.num_tiles) stage_store(p, ts_next, win0 + ((j + 1) & 1) * win_bytes, gtid, wr);
      named_barrier(bar, kImgProducers);
      if (gtid == 0) {
        mbar_wait(smem_u32(&ps->full[s]), par);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(a_addr, kAChunk0, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(a_addr + a_bytes, kAChunk0, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) { umma_bf16(dacc, adesc + 128 * k, bdesc + 128 * k, idesc, accum); accum = 1; }
        umma_commit(smem_u32(&ps->empty[s]));
      }
      __syncwarp();
      if (++sl == spg) { sl = 0; par ^= 1; }
    }
    if (gtid == 0) {
      if (tile0 < p.num_tiles) umma_commit(smem_u32(&ps->acc_full));
      else mbar_arrive(smem_u32(&ps->acc_full));
    }
  } else {
    // ------------------------------------------------------------------ one reduction at the end: rows = filter slots
    const int q = warp & 3;
    const int kk = q * 32 + lane;
    mbar_wait(smem_u32(&ps->acc_full), 0);
    tc_fence_after();
    const int kh = kk >> 4, jj = kk & 15;
    float* dst = nullptr;
    if (kk < kK * 16) {
      if (jj < kcin) dst = pw.dw + (size_t)(kh * kcin + jj) * pw.ldo;
      else if (kk == 15) dst = pw.dbias;
    }
    const bool two = (int)(blockIdx.x + gridDim.x) < p.num_tiles;      // the second group had tiles too
    for (int c = 0; c < pw.cout; c += 16) {
      uint32_t v[16], u[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c, v);
      if (two) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + kTmemCols + c, u);
      tmem_ld_wait16(v);
      if (two) {
        tmem_ld_wait16(u);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
      }
      if (dst) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c + i),
                       "f"(__uint_as_float(v[i]) * pw.alpha), "f"(__uint_as_float(v[i + 1]) * pw.alpha),
                       "f"(__uint_as_float(v[i + 2]) * pw.alpha), "f"(__uint_as_float(v[i + 3]) * pw.alpha) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<2 * kTmemCols>(tmem);
}

// =============================================================================================
// Fused image-side dgrad (see ImgDgradParams)
// =============================================================================================
constexpr int kDgThreads = 64 + 512;                     // warp 0 TMA producer, warp 1 MMA issuer, 16 epilogue warps
constexpr int kDgEpiThreads = 512;
constexpr int kTPitch = 84;                              // fp32 words per row of T in shared memory (4 x odd: the
                                                         // per-row 16-byte stores of a warp are conflict free)
struct DgSmem {
  uint64_t full[8], empty[8];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t b_ready;
  uint32_t tmem_base;
};

// col2im gather of one image from T (see img_dgrad_kernel): one input pixel (all its channels) per thread and step.
// Tap kh = kh0 + jh*stride with kh0 = (iy + pad_t) mod stride reads output row oy0 - jh (same along x): uniform loop
// bounds (ceil(k/stride) each way) and predicated loads instead of a divergent walk over all k*k taps.
// kS / kC: stride and Cin at compile time (0 = read them from the geometry): the hot shapes get fully unrolled code.
template <int kS, int kC>
__device__ __forceinline__ void dg_gather(const ImgDgradParams& p, const float* __restrict__ T, size_t obase, int et) {
  const int st = kS ? kS : p.g.stride;
  const int cin = kC ? kC : p.g.Cin;
  const int nj = (p.g.k + st - 1) / st;
  for (int pix = et; pix < p.g.H * p.g.W; pix += kDgEpiThreads) {
    const int iy = fdiv(pix, p.div_w);
    const int ix = pix - iy * p.g.W;
    const int ty = iy + p.g.pad_t, tx = ix + p.g.pad_l;
    int kh0, kw0, oy0, ox0;
    if (st == 2) { kh0 = ty & 1; kw0 = tx & 1; oy0 = ty >> 1; ox0 = tx >> 1; }
    else if (st == 1) { kh0 = 0; kw0 = 0; oy0 = ty; ox0 = tx; }
    else { oy0 = ty / st; kh0 = ty - oy0 * st; ox0 = tx / st; kw0 = tx - ox0 * st; }
    const int base = (oy0 * p.g.Wo + ox0) * kTPitch + kh0 * 16 + kw0 * cin;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int jh = 0; jh < 5; ++jh) {
      if (jh >= nj) break;
      const bool vh = kh0 + jh * st < p.g.k && oy0 - jh >= 0 && oy0 - jh < p.g.Ho;
      const int offh = jh * (st * 16 - p.g.Wo * kTPitch);
#pragma unroll
      for (int jw = 0; jw < 5; ++jw) {
        if (jw >= nj) break;
        const bool vv = vh && kw0 + jw * st < p.g.k && ox0 - jw >= 0 && ox0 - jw < p.g.Wo;
        const float* tp = T + (vv ? base + offh + jw * (st * cin - kTPitch) : 0);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < cin) acc[c] += vv ? tp[c] : 0.f;
      }
    }
    const size_t o = obase + (size_t)pix * cin;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= cin) break;
      float v = acc[c];
      if (p.bias) v += __ldg(p.bias + c);
      v = act_fwd(v, p.act, p.leak);
      if (p.mask_src) v *= act_grad_from_out(__bfloat162float(p.mask_src[o + c]), p.mask_kind, p.leak);
      if (p.out_f32) reinterpret_cast<float*>(p.out)[o + c] = v;
      else reinterpret_cast<__nv_bfloat16*>(p.out)[o + c] = __float2bfloat16(v);
    }
  }
}

__global__ void __launch_bounds__(kDgThreads, 1) img_dgrad_kernel(const __grid_constant__ ImgDgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int K16 = p.g.k * 16;                             // MMA N: the row-group layout of the filter taps
  const int kcin = p.g.k * p.g.Cin;
  const int nchunks = p.kfull + (p.ktail ? 1 : 0);        // K chunks (over Cout) per tile
  const int b_chunk = K16 * 128;
  uint8_t* smem_b = smem;                                 // kfull chunks [K16 rows][64 cout] SW128, then the tail [K16][16] SW32
  uint8_t* smem_bt = smem_b + (size_t)p.kfull * b_chunk;
  uint8_t* smem_a = smem_bt + (p.ktail ? align_up(K16 * 32, 1024) : 0);
  uint8_t* smem_t = smem_a + (size_t)p.stages * kAChunk0;
  DgSmem* ps = reinterpret_cast<DgSmem*>(smem_t + (size_t)p.tiles_per_image * kTileM * kTPitch * 4);
  const int n_images = p.g.N;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    if (p.ktail) tma_prefetch_desc(&p.tmA_tail);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&ps->full[s]), 1);
      mbar_init(smem_u32(&ps->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&ps->acc_full[a]), 1);
      mbar_init(smem_u32(&ps->acc_empty[a]), 16);
    }
    mbar_init(smem_u32(&ps->b_ready), kDgEpiThreads);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * kTmemCols>(smem_u32(&ps->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer: dy tiles, chunk by chunk
      int s = 0;
      uint32_t par = 0;
      for (int img = blockIdx.x; img < n_images; img += gridDim.x)
        for (int t = 0; t < p.tiles_per_image; ++t) {
          const int row0 = (img * p.tiles_per_image + t) * kTileM;
          for (int kc = 0; kc < nchunks; ++kc) {
            mbar_wait(smem_u32(&ps->empty[s]), par ^ 1);
            const uint32_t full = smem_u32(&ps->full[s]);
            const uint32_t dst = smem_u32(smem_a) + (uint32_t)s * kAChunk0;
            if (kc < p.kfull) {
              mbar_arrive_expect_tx(full, kAChunk0);
              tma_load_2d(dst, &p.tmA, full, kc * 64, row0);
            } else {
              mbar_arrive_expect_tx(full, kATail);
              tma_load_2d(dst, &p.tmA_tail, full, kc * 64, row0);
            }
            if (++s == p.stages) { s = 0; par ^= 1; }
          }
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      const uint32_t idesc = make_idesc_bf16(kTileM, K16, 0, 0);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem_b), 16, 1024);
      const uint64_t bdesct = make_smem_desc(smem_u32(smem_bt), 16, 256, 6);
      mbar_wait(smem_u32(&ps->b_ready), 0);
      int s = 0, ab = 0;
      uint32_t par = 0, abpar = 0;
      for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
        mbar_wait(smem_u32(&ps->acc_empty[ab]), abpar ^ 1);
        tc_fence_after();
        for (int t = 0; t < p.tiles_per_image; ++t) {
          const uint32_t d = tmem + ab * kTmemCols + t * 128;
          uint32_t accum = 0;
          for (int kc = 0; kc < nchunks; ++kc) {
            mbar_wait(smem_u32(&ps->full[s]), par);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem_a) + (uint32_t)s * kAChunk0;
            if (kc < p.kfull) {
              const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
              const uint64_t bdesc = bdesc0 + (uint64_t)(((uint32_t)b_chunk >> 4) * kc);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) { umma_bf16(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, accum); accum = 1; }
            } else {
              umma_bf16(d, make_smem_desc(a_addr, 16, 256, 6), bdesct, idesc, accum);
              accum = 1;
            }
            umma_commit(smem_u32(&ps->empty[s]));
            if (++s == p.stages) { s = 0; par ^= 1; }
          }
        }
        umma_commit(smem_u32(&ps->acc_full[ab]));
        ab ^= 1;
        if (ab == 0) abpar ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int et = threadIdx.x - 64;
    // the weights as the B operand: B[kh*16 + j][cout] = w[kh*kcin + j][cout] (rows of couts: K-major as they lie)
    {
      const int c8s = p.cout >> 3;
      for (int item = et; item < K16 * c8s; item += kDgEpiThreads) {
        const int c8 = item % c8s, kk = item / c8s;
        const int kh = kk >> 4, j = kk & 15;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (j < kcin) v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)(kh * kcin + j) * p.ldw + c8 * 8));
        const int kc = c8 >> 3, cc = c8 & 7;
        uint32_t dst;
        if (kc < p.kfull) dst = smem_u32(smem_b) + kc * b_chunk + kk * 128 + ((cc ^ (kk & 7)) << 4);
        else dst = smem_u32(smem_bt) + kk * 32 + ((cc ^ ((kk >> 2) & 1)) << 4);
        st_shared_v4(dst, v.x, v.y, v.z, v.w);
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&ps->b_ready));
    }
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int nq = K16 >> 4;                               // 16-column chunks of T
    const int hwc = p.g.H * p.g.W * p.g.Cin;
    const uint32_t t0 = smem_u32(smem_t);
    int ab = 0;
    uint32_t abpar = 0;
    for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
      mbar_wait(smem_u32(&ps->acc_full[ab]), abpar);
      tc_fence_after();
      // accumulators -> T in shared memory (fp32 rows of kTPitch words)
      for (int t = 0; t < p.tiles_per_image; ++t)
        for (int c = cg; c < nq; c += 4) {
          uint32_t v[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + ab * kTmemCols + t * 128 + c * 16, v);
          tmem_ld_wait16(v);
          const uint32_t dst = t0 + (uint32_t)((t * kTileM + r) * kTPitch + c * 16) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) st_shared_v4(dst + 16 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ps->acc_empty[ab]));
      named_barrier(1, kDgEpiThreads);
      // every output element gathers its taps: kh = kh0 + jh*stride with kh0 = (iy + pad_t) mod stride reads output row
      // oy0 - jh (same along x).  The loop bounds are uniform (ceil(k/stride) each way) and the loads predicated, so a
      // warp walks <= 9 (stride 2) combinations instead of diverging over all k*k
      const float* T = reinterpret_cast<const float*>(smem_t);
      const size_t obase = (size_t)img * hwc;
      if (p.g.stride == 2 && p.g.Cin == 3) dg_gather<2, 3>(p, T, obase, et);
      else if (p.g.stride == 2 && p.g.Cin == 1) dg_gather<2, 1>(p, T, obase, et);
      else if (p.g.stride == 1 && p.g.Cin == 3) dg_gather<1, 3>(p, T, obase, et);
      else dg_gather<0, 0>(p, T, obase, et);
      named_barrier(1, kDgEpiThreads);                     // T is free for the next image
      ab ^= 1;
      if (ab == 0) abpar ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * kTmemCols>(tmem);
}

// standalone gather (same K layout): one thread per output pixel
template <int kK>
__global__ void __launch_bounds__(256) img_im2col16_kernel(const __nv_bfloat16* __restrict__ x, int x_words,
                                                           ImgConvGeom g, long long M, __nv_bfloat16* __restrict__ out,
                                                           uint32_t ones) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= M) return;
  const RowCtx rc = make_row_ctx(g, pix, M);
  const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
#pragma unroll
  for (int kh = 0; kh < kK; ++kh) {
    uint32_t o[8];
    const GroupLoad L = gather_load(g, xw, x_words, rc, kh);
    gather_finish(L, rc, kh < 2 ? ones : 0u, o);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)pix * (kK * 16) + kh * 16);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}


// weights in the same K layout for the GEMM route (general epilogues): Wt[n][kh*16 + j] = w[kh*kcin + j][n], pads zero
__global__ void img_wpad16_kernel(const __nv_bfloat16* __restrict__ w, int ldw, int ncols, int k, int kcin,
                                  __nv_bfloat16* __restrict__ wt) {
  const int Kp = k * 16;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncols * Kp) return;
  const int n = i % ncols, kk = i / ncols;
  const int kh = kk >> 4, j = kk & 15;
  wt[(size_t)n * Kp + kk] = j < kcin ? w[(size_t)(kh * kcin + j) * ldw + n] : __float2bfloat16(0.f);
}

// filter gradient computed in the row-group layout -> the TF layout: dw[kh*kcin + j][n] += t[kh*16 + j][n];
// row 15 (the ones column of the gather) is the column sum of dy = the bias gradient
__global__ void img_wgrad_fold_kernel(const float* __restrict__ t, int ncols, int k, int kcin, float* __restrict__ dw,
                                      int ldo, float* __restrict__ dbias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncols * k * 16) return;
  const int n = i % ncols, kk = i / ncols;
  const int kh = kk >> 4, j = kk & 15;
  const float v = t[(size_t)kk * ncols + n];
  if (j < kcin) dw[(size_t)(kh * kcin + j) * ldo + n] += v;
  else if (kk == 15 && dbias) dbias[n] += v;
}

constexpr int kMaxDev = 64;
bool first_use(bool* flags) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) return true;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}
int dev_sms() {
  static int sms[kMaxDev] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) return 148;
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

}  // namespace imgconv
using namespace imgconv;

size_t img_fprop_smem(const ImgFpropParams& p) {
  const int bt = (p.g.k == 5 || p.virt) ? align_up(p.ncols * 32, 1024) : 0;
  return (size_t)p.ncols * 128 + bt + (size_t)p.slots * kASlot + 2 * (size_t)align_up(kTileM * p.stage_pitch, 1024) +
         2 * (size_t)(p.bits_stage ? align_up(kTileM * p.bits_pitch * 2, 128) : 0) +
         2 * kImgGroups * (size_t)align_up(p.win_rows * p.win_pitch * 2, 128) + sizeof(ImgSmem) + 1024;
}

static ImgDiv make_div(int d) {
  int L = 0;
  while ((1ll << L) < d) ++L;
  ImgDiv r;
  r.shr = 31 + L;
  r.mul = (uint32_t)(((1ull << r.shr) + d - 1) / d);       // ceil(2^(31+L) / d) < 2^32: exact quotients for n < 2^31
  return r;
}

// window layout for a geometry (see stage_window): padded-row pitch, data offset, padded rows per image and the
// largest number of padded rows one 128-pixel tile spans
static void img_window(const ImgConvGeom& g, int* pitch, int* off, int* hp, int* rows) {
  *off = align_up(g.pad_l * g.Cin, 8);
  const int reach = *off - g.pad_l * g.Cin + (g.Wo - 1) * g.stride * g.Cin + 18;   // last element a gather may touch
  *pitch = align_up(std::max(reach, *off + g.W * g.Cin), 8);
  *hp = (g.Ho - 1) * g.stride + g.k;
  static thread_local ImgConvGeom cg = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  static thread_local int crows = 0;
  if (memcmp(&cg, &g, sizeof g) != 0) {
    const long long M = (long long)g.N * g.Ho * g.Wo;
    const int hw = g.Ho * g.Wo;
    int best = 0;
    for (long long p0 = 0; p0 < M; p0 += kTileM) {
      const long long p1 = std::min(M, p0 + kTileM) - 1;
      const int n0 = (int)(p0 / hw), oy0 = (int)(p0 % hw) / g.Wo, n1 = (int)(p1 / hw), oy1 = (int)(p1 % hw) / g.Wo;
      best = std::max(best, (n1 - n0) * *hp + (oy1 - oy0) * g.stride + g.k);
      if (p0 / kTileM > 4096 && (p0 % hw) == 0) break;         // the pattern repeats once a tile starts an image
    }
    cg = g; crows = best;
  }
  *rows = crows;
}

bool img_fprop_supported(const ImgConvGeom& g, int ncols, int has_bias) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("B200GAN_NO_IMGFUSE"); off = e ? atoi(e) : 0; }
  if (off) return false;
  if (g.k < 3 || g.k > 5 || g.k * g.Cin > 15) return false;
  (void)has_bias;                                            // k >= 3: groups 0 and 1 both exist for the bias pair
  if (ncols % 16 || ncols < 16 || ncols > 256) return false;
  const long long numel = (long long)g.N * g.H * g.W * g.Cin;
  if ((g.W * g.Cin) & 1 || numel >= (1ll << 31)) return false;     // image rows are copied as 4-byte words
  if ((long long)g.N * g.Ho * g.Wo >= (1ll << 31) - 256) return false;
  // worst-case shared memory: two A slots, padded output stage, staged sign words, both input windows
  ImgFpropParams q;
  memset(&q, 0, sizeof q);
  q.g = g; q.ncols = ncols; q.slots = 2; q.stage_pitch = ncols * 2 + 16; q.bits_stage = 1; q.bits_pitch = ncols / 16;
  img_window(g, &q.win_pitch, &q.win_off, &q.win_hp, &q.win_rows);
  // the window is staged through registers (kWinRegs pieces per producer thread)
  const int rb = g.W * g.Cin * 2;
  if (q.win_rows * (rb / (rb % 16 == 0 ? 16 : 4)) > kImgProducers * kWinRegs * (rb % 16 == 0 ? 1 : 4)) return false;
  return img_fprop_smem(q) <= 227 * 1024;
}

// k = 4 with k*Cin = 16 (pix2pix's discriminator input: rgb + depth): every slot of the four filter rows holds data, so
// the bias pair moves into a fifth, otherwise empty row group (the kernel's k = 5 form; ImgFpropParams.virt).  Only the
// fused fprop takes this form: the layer's filter / input gradients keep the im2col route.
bool img_fprop_virtual_supported(const ImgConvGeom& g, int ncols) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("B200GAN_NO_IMGFUSE"); off = e ? atoi(e) : 0; }
  if (off) return false;
  if (g.k != 4 || g.k * g.Cin != 16) return false;          // (Cin = 4: every window starts on an even element)
  if (ncols % 16 || ncols < 16 || ncols > 256) return false;
  const long long numel = (long long)g.N * g.H * g.W * g.Cin;
  if (numel >= (1ll << 31) || (long long)g.N * g.Ho * g.Wo >= (1ll << 31) - 256) return false;
  ImgFpropParams q;
  memset(&q, 0, sizeof q);
  q.g = g; q.ncols = ncols; q.slots = 2; q.stage_pitch = ncols * 2 + 16; q.bits_stage = 1; q.bits_pitch = ncols / 16;
  q.virt = 1;
  img_window(g, &q.win_pitch, &q.win_off, &q.win_hp, &q.win_rows);
  const int rb = g.W * g.Cin * 2;
  if (q.win_rows * (rb / (rb % 16 == 0 ? 16 : 4)) > kImgProducers * kWinRegs * (rb % 16 == 0 ? 1 : 4)) return false;
  return img_fprop_smem(q) <= 227 * 1024;
}

void launch_img_fprop(const ImgFpropParams& p0, cudaStream_t stream) {
  ImgFpropParams p = p0;
  img_window(p.g, &p.win_pitch, &p.win_off, &p.win_hp, &p.win_rows);
  p.div_hw = make_div(p.g.Ho * p.g.Wo); p.div_wo = make_div(p.g.Wo); p.div_hp = make_div(p.win_hp);
  p.win_vec16 = ((p.g.W * p.g.Cin * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0) ? 1 : 0;
  p.win_ppr = p.g.W * p.g.Cin * 2 / (p.win_vec16 ? 16 : 4);
  p.div_ppr = make_div(p.win_ppr);
  p.slots = kMaxSlots;                                      // per producer group: 2 or 1
  if (img_fprop_smem(p) > 227 * 1024) p.slots = 2;
  const size_t smem = img_fprop_smem(p);
  static bool configured[kMaxDev] = {false};
  if (first_use(configured)) {
    cudaFuncSetAttribute(img_fprop_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(img_fprop_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(img_fprop_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int sms = dev_sms();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  if (p.g.k == 5 || p.virt) img_fprop_kernel<5><<<grid, kImgThreads, smem, stream>>>(p);
  else if (p.g.k == 4) img_fprop_kernel<4><<<grid, kImgThreads, smem, stream>>>(p);
  else img_fprop_kernel<3><<<grid, kImgThreads, smem, stream>>>(p);
}

static size_t img_wgrad_smem(const ImgFpropParams& f, int cout, int stages) {
  const int nblocks = (cout + 63) / 64;
  return (size_t)stages * (2 + nblocks) * kAChunk0 + 2 * kWgGroups * (size_t)align_up(f.win_rows * f.win_pitch * 2, 128) +
         sizeof(WgSmem) + 1024;
}

bool img_wgrad_supported(const ImgConvGeom& g, int cout) {
  if (!img_fprop_supported(g, cout, 1)) return false;
  ImgFpropParams q;
  memset(&q, 0, sizeof q);
  q.g = g;
  img_window(g, &q.win_pitch, &q.win_off, &q.win_hp, &q.win_rows);
  return img_wgrad_smem(q, cout, 2) <= 227 * 1024;
}

void launch_img_wgrad(const ImgWgradParams& p0, cudaStream_t stream) {
  ImgWgradParams p = p0;
  ImgFpropParams& f = p.f;
  f.M = (long long)f.g.N * f.g.Ho * f.g.Wo;
  f.num_tiles = (int)((f.M + kTileM - 1) / kTileM);
  f.x_words = (long long)f.g.N * f.g.H * f.g.W * f.g.Cin / 2;
  f.im2col_out = nullptr;
  img_window(f.g, &f.win_pitch, &f.win_off, &f.win_hp, &f.win_rows);
  f.div_hw = make_div(f.g.Ho * f.g.Wo); f.div_wo = make_div(f.g.Wo); f.div_hp = make_div(f.win_hp);
  f.win_vec16 = ((f.g.W * f.g.Cin * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(f.x) & 15) == 0) ? 1 : 0;
  f.win_ppr = f.g.W * f.g.Cin * 2 / (f.win_vec16 ? 16 : 4);
  f.div_ppr = make_div(f.win_ppr);
  p.nblocks = (p.cout + 63) / 64;
  p.stages = img_wgrad_smem(f, p.cout, 4) <= 227 * 1024 ? 4 : 2;      // per producer group: 2 or 1
  const size_t smem = img_wgrad_smem(f, p.cout, p.stages);
  static bool configured[kMaxDev] = {false};
  if (first_use(configured)) {
    cudaFuncSetAttribute(img_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(img_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(img_wgrad_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int sms = dev_sms();
  const int grid = f.num_tiles < sms ? f.num_tiles : sms;
  if (f.g.k == 5) img_wgrad_kernel<5><<<grid, kWgThreads, smem, stream>>>(p);
  else if (f.g.k == 4) img_wgrad_kernel<4><<<grid, kWgThreads, smem, stream>>>(p);
  else img_wgrad_kernel<3><<<grid, kWgThreads, smem, stream>>>(p);
}

static size_t img_dgrad_smem(const ImgConvGeom& g, int cout, int stages) {
  const int K16 = g.k * 16, kfull = cout / 64, ktail = cout % 64;
  return (size_t)kfull * K16 * 128 + (ktail ? align_up(K16 * 32, 1024) : 0) + (size_t)stages * kAChunk0 +
         (size_t)(g.Ho * g.Wo) * kTPitch * 4 + sizeof(DgSmem) + 1024;
}

bool img_dgrad_supported(const ImgConvGeom& g, int cout) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("B200GAN_NO_IMGFUSE"); off = e ? atoi(e) : 0; }
  if (off) return false;
  if (g.k < 1 || g.k > 5 || g.k * g.Cin > 15 || g.Cin > 4) return false;
  if (g.Ho * g.Wo != 128 && g.Ho * g.Wo != 256) return false;       // one image = 1 or 2 accumulator tiles
  if (cout % 64 != 0 && cout % 64 != 16) return false;              // K tail: none or one 16-wide SWIZZLE_32B box
  if (cout % 8 || cout > 1024) return false;
  if ((long long)g.N * g.H * g.W * g.Cin >= (1ll << 31) || (long long)g.N * g.Ho * g.Wo >= (1ll << 31)) return false;
  return img_dgrad_smem(g, cout, 2) <= 227 * 1024;
}

void launch_img_dgrad(const ImgDgradParams& p0, cudaStream_t stream) {
  ImgDgradParams p = p0;
  p.kfull = p.cout / 64; p.ktail = p.cout % 64;
  p.tiles_per_image = p.g.Ho * p.g.Wo / kTileM;
  p.div_w = make_div(p.g.W);
  p.stages = 8;
  while (p.stages > 2 && img_dgrad_smem(p.g, p.cout, p.stages) > 227 * 1024) --p.stages;
  const size_t smem = img_dgrad_smem(p.g, p.cout, p.stages);
  static bool configured[kMaxDev] = {false};
  if (first_use(configured))
    cudaFuncSetAttribute(img_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int sms = dev_sms();
  const int grid = p.g.N < sms ? p.g.N : sms;
  img_dgrad_kernel<<<grid, kDgThreads, smem, stream>>>(p);
}

void launch_img_im2col16(const __nv_bfloat16* x, long long x_words, const ImgConvGeom& g, __nv_bfloat16* out, int ones,
                         cudaStream_t stream) {
  const long long M = (long long)g.N * g.Ho * g.Wo;
  const int grid = (int)((M + 255) / 256);
  const uint32_t ob = ones ? 0x3F800000u : 0u;
  if (g.k == 5) img_im2col16_kernel<5><<<grid, 256, 0, stream>>>(x, (int)x_words, g, M, out, ob);
  else if (g.k == 4) img_im2col16_kernel<4><<<grid, 256, 0, stream>>>(x, (int)x_words, g, M, out, ob);
  else img_im2col16_kernel<3><<<grid, 256, 0, stream>>>(x, (int)x_words, g, M, out, ob);
}

void launch_img_wpad16(const __nv_bfloat16* w, int ldw, int ncols, int k, int kcin, __nv_bfloat16* wt,
                       cudaStream_t stream) {
  const int n = ncols * k * 16;
  img_wpad16_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, ldw, ncols, k, kcin, wt);
}

void launch_img_wgrad_fold(const float* t, int ncols, int k, int kcin, float* dw, int ldo, float* dbias,
                           cudaStream_t stream) {
  const int n = ncols * k * 16;
  img_wgrad_fold_kernel<<<(n + 255) / 256, 256, 0, stream>>>(t, ncols, k, kcin, dw, ldo, dbias);
}

}  // namespace b200
