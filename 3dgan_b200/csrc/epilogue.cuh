// Fused epilogue shared by the tensor-core and the small-channel kernels:
//   v = alpha*acc (+ bias[col]);  v = act(v);  v *= act'(mask_src[off+col]);  store bf16|fp32.
// Order follows the reference's layer definition conv -> +bias -> activation
// (ops/layers.py:101-105); the mask multiply is the activation-gradient of the consumer layer
// (derivatives are expressed through the stored post-activation value).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include "tc_gemm.cuh"
#include "ptx.cuh"

namespace b200 {

// tanh / sigmoid use the hardware MUFU.TANH (tanh.approx.f32, relative error about 2^-11): results are stored
// in bf16 (2^-9), and the libdevice tanhf expansion inlined into every epilogue bloats the hot loops
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float act_fwd(float v, int act, float leak) {
  if (act == ACT_NONE) return v;
  if (act == ACT_RELU || act == ACT_LRELU) return fmaxf(v, (act == ACT_RELU ? 0.f : leak) * v);
  if (act == ACT_TANH) return tanh_fast(v);
  return fmaf(0.5f, tanh_fast(0.5f * v), 0.5f);       // sigmoid(v) = (1 + tanh(v/2)) / 2
}
// derivative of act at the point whose OUTPUT is a (A.5: lrelu slope = leak for x <= 0)
__device__ __forceinline__ float act_grad_from_out(float a, int act, float leak) {
  switch (act) {
    case ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case ACT_LRELU: return a > 0.f ? 1.f : leak;
    case ACT_TANH: return 1.f - a * a;
    case ACT_SIGMOID: return a * (1.f - a);
    default: return 1.f;
  }
}

struct EpilogueArgs {
  const float* bias;
  int act;
  float leak;
  const __nv_bfloat16* mask_src;
  int mask_kind;
  float alpha;
  void* out;
  int out_f32;
  int accumulate;
  int ncols;
  int pipelined;          // prefetch mask rows to L2 + load them one chunk ahead (B200GAN_EPI_PIPE, default 1)
  // sign bitmaps (1 bit per element, 16-column words, `bits_pitch` words per output row): for relu / lrelu the
  // activation gradient only needs sign(out), so a consumer reads 2 bytes per chunk instead of 32
  const uint16_t* mask_bits;   // replaces mask_src when mask_kind is relu / lrelu
  uint16_t* bits_out;          // written next to the output: bit j of word (row, col/16) = out[row, col+j] > 0
  int bits_pitch;
  int row_elems;               // elements per output row: row index = element offset / row_elems
  // shared-memory staging (TMA-store epilogue): when stage_row != 0 the chunk is written to
  // stage_row + (col - stage_col0) * elem and its sign word to stage_bits + ((col - stage_col0) >> 4) * 2;
  // one thread later moves the whole tile to global memory with a bulk tensor store.  Row-strided
  // 16-byte global stores cost one L1 wavefront per lane; the bulk store writes full lines.
  uint32_t stage_row;
  uint32_t stage_bits;
  int stage_col0;
  // the tile's bias row staged in shared memory by the epilogue warps while they wait for the accumulators
  // (fp32, element i = bias[bias_col0 + i], zero past ncols): the global bias loads were the top stall of the
  // epilogue (L1 is almost all shared memory here, so they usually went to L2)
  uint32_t bias_smem;
  int bias_col0;
  // stream-K fixup (tapgemm2sm_sk_kernel): `partial_n` fp32 partial accumulator images of this row, `partial_stride`
  // elements apart, indexed by the column offset inside the tile; added to the accumulators before the epilogue
  const float* partial_row;
  int partial_n;
  long long partial_stride;
};

// Epilogue warps are idle during the main loop: pull the mask rows they will need into L2 meanwhile.
__device__ __forceinline__ void epilogue_prefetch_mask(const EpilogueArgs& e, long long off, int col0, int ncols_tile) {
  if (!e.mask_src || !e.pipelined) return;
  const char* p = reinterpret_cast<const char*>(e.mask_src + off + col0);
  const int bytes = min(ncols_tile, e.ncols - col0) * 2;
  for (int b = 0; b < bytes; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + b));
}
// raw 32 bytes (16 bf16) of the mask for one chunk, loaded ahead of use; valid only on the aligned fast path
struct MaskChunk {
  uint4 lo, hi;
  bool loaded;
};
__device__ __forceinline__ MaskChunk epilogue_load_mask(const EpilogueArgs& e, long long off, int col) {
  MaskChunk m;
  m.loaded = false;
  if (e.mask_src && e.pipelined && col + 8 <= e.ncols) {
    const __nv_bfloat16* p = e.mask_src + off + col;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      m.lo = __ldg(reinterpret_cast<const uint4*>(p));
      m.hi = (col + 16 <= e.ncols) ? __ldg(reinterpret_cast<const uint4*>(p) + 1) : make_uint4(0, 0, 0, 0);
      m.loaded = true;
    }
  }
  return m;
}

// fp32 bias for one chunk, loaded while the TMEM load of the chunk is in flight (aligned fast path only)
struct BiasChunk {
  float4 b[4];
  bool loaded;
};
__device__ __forceinline__ BiasChunk epilogue_load_bias(const EpilogueArgs& e, int col) {
  BiasChunk b;
  b.loaded = false;
  if (e.bias && col + 16 <= e.ncols) {
    const float* bp = e.bias + col;
    if ((reinterpret_cast<uintptr_t>(bp) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) b.b[j] = __ldg(reinterpret_cast<const float4*>(bp) + j);
      b.loaded = true;
    }
  }
  return b;
}

// everything a chunk needs besides its accumulators, fetched one chunk ahead
struct ChunkSide {
  MaskChunk m;
  uint32_t bits;
};
__device__ __forceinline__ ChunkSide epilogue_load_side(const EpilogueArgs& e, long long off, int col, bool row_ok,
                                                       long long bits_row) {
  ChunkSide s;
  s.m.loaded = false;
  s.bits = 0;
  if (row_ok) {
    if (e.mask_bits) s.bits = __ldg(e.mask_bits + bits_row + (col >> 4));
    else s.m = epilogue_load_mask(e, off, col);
  }
  return s;
}

// general (ragged / unaligned) path: a compact per-element loop over a local copy of the accumulators
__device__ __noinline__ static void epilogue_store16_slow(const EpilogueArgs& e, const float* acc, long long off, int col,
                                                          uint32_t mbits, long long bits_row) {
  const int nvalid = min(16, e.ncols - col);
  const long long o = off + col;
  uint32_t obits = 0;
#pragma unroll 1
  for (int j = 0; j < nvalid; ++j) {
    float v = acc[j] * e.alpha;
    if (e.bias) v += __ldg(e.bias + col + j);
    v = act_fwd(v, e.act, e.leak);
    obits |= (v > 0.f ? 1u : 0u) << j;
    if (e.mask_bits) v *= ((mbits >> j) & 1u) ? 1.f : (e.mask_kind == ACT_LRELU ? e.leak : 0.f);
    else if (e.mask_src) v *= act_grad_from_out(__bfloat162float(e.mask_src[o + j]), e.mask_kind, e.leak);
    if (e.out_f32) {
      float* dst = reinterpret_cast<float*>(e.out) + o + j;
      if (e.accumulate == 2) atomicAdd(dst, v);
      else *dst = e.accumulate ? *dst + v : v;
    } else {
      reinterpret_cast<__nv_bfloat16*>(e.out)[o + j] = __float2bfloat16(v);
    }
  }
  if (e.bits_out) e.bits_out[bits_row + (col >> 4)] = (uint16_t)obits;
}

// 16 consecutive columns [col, col+16) of one output row starting at element offset `off`.
// Fast path (full chunk, 16-byte aligned output, side data preloaded): straight-line code.
// kSimple kernels are launched when the epilogue is none/relu/lrelu with no value mask and no accumulate (bias,
// fp32 or bf16 output, sign bitmaps in and out allowed): the other variants are compiled out of their hot loop.
template <bool kSimple>
__device__ __forceinline__ void epilogue_store16(const EpilogueArgs& e, const uint32_t* acc, long long off,
                                                 int col, const ChunkSide* side = nullptr, long long bits_row = 0) {
  const long long o = off + col;
  // channel counts are multiples of 8 on the tensor-core route: a chunk holds 16 or (the last one) 8 columns
  const int nv = min(16, e.ncols - col);
  const bool full = nv == 16 || nv == 8;
  const bool f32 = e.out_f32 != 0;
  const uintptr_t oaddr = reinterpret_cast<uintptr_t>(e.out) + (uintptr_t)o * (f32 ? 4 : 2);
  const bool vmask = !kSimple && e.mask_src && !e.mask_bits;
  const bool fast = full && (e.stage_row || (oaddr & 15) == 0) &&
                    (!e.bias || e.bias_smem || (reinterpret_cast<uintptr_t>(e.bias + col) & 15) == 0) &&
                    (!vmask || (side && side->m.loaded));
  if (!fast) {
    if (e.stage_row) __trap();            // the host enables staging only when every chunk takes the fast path
    // cold: local copies, so that neither the accumulators nor the argument block of the hot path ever
    // have their address taken (they would be demoted to local memory, and L1 is almost all shared memory here)
    float tmp[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) tmp[j] = __uint_as_float(acc[j]);
    const EpilogueArgs ecopy = e;
    epilogue_store16_slow(ecopy, tmp, off, col, side ? side->bits : 0u, bits_row);
    return;
  }
  float v[16];
  const float a = e.alpha;
  if (e.bias) {
    const float4* bp = reinterpret_cast<const float4*>(e.bias + col);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b4 = e.bias_smem ? ld_shared_f4(e.bias_smem + (uint32_t)(col - e.bias_col0 + 4 * j) * 4)
                                    : ((4 * j < nv) ? __ldg(bp + j) : make_float4(0.f, 0.f, 0.f, 0.f));
      v[4 * j + 0] = fmaf(__uint_as_float(acc[4 * j + 0]), a, b4.x);
      v[4 * j + 1] = fmaf(__uint_as_float(acc[4 * j + 1]), a, b4.y);
      v[4 * j + 2] = fmaf(__uint_as_float(acc[4 * j + 2]), a, b4.z);
      v[4 * j + 3] = fmaf(__uint_as_float(acc[4 * j + 3]), a, b4.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]) * a;
  }
  if (!kSimple && (e.act == ACT_TANH || e.act == ACT_SIGMOID)) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_fwd(v[j], e.act, e.leak);
  } else {
    // none / relu / lrelu in one branch-free form: max(v, slope * v) with slope 1 / 0 / leak
    const float slope = e.act == ACT_NONE ? 1.f : (e.act == ACT_RELU ? 0.f : e.leak);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], slope * v[j]);
  }
  if (e.bits_out) {
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
    if (nv == 8) w &= 0xffu;
    if (e.stage_bits) st_shared_u16(e.stage_bits + (((col - e.stage_col0) >> 4) << 1), (uint16_t)w);
    else e.bits_out[bits_row + (col >> 4)] = (uint16_t)w;
  }
  if (e.mask_bits) {
    const float neg = e.mask_kind == ACT_LRELU ? e.leak : 0.f;
    const uint32_t w = side->bits;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ((w >> j) & 1u) ? v[j] : v[j] * neg;
  } else if (vmask) {
    const uint32_t mw[8] = {side->m.lo.x, side->m.lo.y, side->m.lo.z, side->m.lo.w,
                            side->m.hi.x, side->m.hi.y, side->m.hi.z, side->m.hi.w};
    if (e.mask_kind == ACT_LRELU || e.mask_kind == ACT_RELU) {
      // derivative from the sign of the stored activation: bf16 > 0  <=>  its 16 bits as a signed int > 0
      const float neg = e.mask_kind == ACT_LRELU ? e.leak : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool lo_pos = (short)(mw[j] & 0xffffu) > 0;
        const bool hi_pos = (int)mw[j] >= 0x10000;
        v[2 * j] = lo_pos ? v[2 * j] : v[2 * j] * neg;
        v[2 * j + 1] = hi_pos ? v[2 * j + 1] : v[2 * j + 1] * neg;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float lo = __uint_as_float(mw[j] << 16), hi = __uint_as_float(mw[j] & 0xffff0000u);
        v[2 * j] *= act_grad_from_out(lo, e.mask_kind, e.leak);
        v[2 * j + 1] *= act_grad_from_out(hi, e.mask_kind, e.leak);
      }
    }
  }
  if (e.stage_row) {
    if (f32) {
      const uint32_t dst = e.stage_row + (uint32_t)(col - e.stage_col0) * 4;
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        if (j < nv)
          st_shared_v4(dst + j * 4, __float_as_uint(v[j]), __float_as_uint(v[j + 1]), __float_as_uint(v[j + 2]),
                       __float_as_uint(v[j + 3]));
    } else {
      const uint32_t dst = e.stage_row + (uint32_t)(col - e.stage_col0) * 2;
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      st_shared_v4(dst, pk[0], pk[1], pk[2], pk[3]);
      if (nv == 16) st_shared_v4(dst + 16, pk[4], pk[5], pk[6], pk[7]);
    }
    return;
  }
  if (f32) {
    float* dst = reinterpret_cast<float*>(e.out) + o;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      if (j < nv) {
        float4 f = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (e.accumulate == 2) {                         // split-K partial tile: fire-and-forget vector reduction
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(f.x), "f"(f.y), "f"(f.z),
                       "f"(f.w) : "memory");
          continue;
        }
        if (e.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(dst + j);
          f.x += old.x; f.y += old.y; f.z += old.z; f.w += old.w;
        }
        *reinterpret_cast<float4*>(dst + j) = f;
      }
    }
  } else {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.out) + o;
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    if (nv == 16) *(reinterpret_cast<uint4*>(dst) + 1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// host side: can this epilogue run in a kSimple kernel?
inline bool epilogue_is_simple(int act, const void* mask_src, const void* mask_bits, int out_f32, int accumulate) {
  (void)out_f32;
  return !accumulate && (act == ACT_NONE || act == ACT_RELU || act == ACT_LRELU) &&
         (mask_src == nullptr || mask_bits != nullptr);
}

// One output row's chunks c0, c0+step, ... of an accumulator row in TMEM: the TMEM load and the side loads
// of chunk i+1 are issued before chunk i is processed.
// accumulators += the other contributors' partial sums of the same 16 columns (L2 loads: they were written by other SMs)
__device__ __forceinline__ void epilogue_add_partials(const EpilogueArgs& e, uint32_t* v, int c) {
  for (int k = 0; k < e.partial_n; ++k) {
    const float4* src = reinterpret_cast<const float4*>(e.partial_row + (long long)k * e.partial_stride + c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 f = __ldcg(src + j);
      v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + f.x);
      v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + f.y);
      v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + f.z);
      v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + f.w);
    }
  }
}

// stream-K contributor: the raw accumulator row goes to the partial-sum workspace (fp32, 64 bytes per chunk)
__device__ __forceinline__ void epilogue_dump_row(uint32_t trow, float* drow, int c_first, int c_step, int c_end) {
#pragma unroll 1
  for (int c = c_first; c < c_end; c += c_step) {
    uint32_t v[16];
    tmem_ld16(trow + c, v);
    tmem_ld_wait16(v);
    float4* dst = reinterpret_cast<float4*>(drow + c);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      __stcg(dst + j, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                  __uint_as_float(v[4 * j + 3])));
  }
}

// ---- lean path: the staged (TMA-store) bf16 epilogue of full 16-column chunks with every option resolved at
// compile time -- bias from shared memory, none / relu / lrelu, sign bitmap out, sign-bitmap mask in.  The general
// epilogue_store16 spends ~95 of its ~200 instructions per chunk on run-time option checks and register copies, and
// the epilogue warps are issue-bound (16 warps on 4 schedulers): this path is what the hot launches take.
template <bool kBias, bool kBitsOut, bool kMaskBits>
__device__ __forceinline__ void epilogue_chunk_lean(const EpilogueArgs& e, const uint32_t* acc, int col, float slope,
                                                    float neg, uint32_t mbits, long long bits_row) {
  float v[16];
  if (kBias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b4 = ld_shared_f4(e.bias_smem + (uint32_t)(col - e.bias_col0 + 4 * j) * 4);
      v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b4.x;
      v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b4.y;
      v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b4.z;
      v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b4.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  }
  if (slope != 1.f) {                       // relu (slope 0) / lrelu: max(v, slope v); uniform branch
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], slope * v[j]);
  }
  if (kBitsOut && kMaskBits) {                // (rare: sign word of the value BEFORE the mask multiply)
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
    if (e.stage_bits) st_shared_u16(e.stage_bits + (((col - e.stage_col0) >> 4) << 1), (uint16_t)w);
    else e.bits_out[bits_row + (col >> 4)] = (uint16_t)w;
  }
  if (kMaskBits) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ((mbits >> j) & 1u) ? v[j] : v[j] * neg;
  }
  const uint32_t dst = e.stage_row + (uint32_t)(col - e.stage_col0) * 2;
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    pk[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  if (kBitsOut && !kMaskBits) {
    // sign word from the packed pairs: one HSET2 per two elements instead of a compare + select + shift per element
    // (bf16(v) > 0 <=> v > 0 up to underflow); element 2j -> bit 2j, element 2j+1 -> bit 2j+17, folded at the end
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&pk[j]), zero2);
      w |= m & ((1u << (2 * j)) | (1u << (2 * j + 17)));
    }
    w = (w & 0xffffu) | (w >> 16);
    if (e.stage_bits) st_shared_u16(e.stage_bits + (((col - e.stage_col0) >> 4) << 1), (uint16_t)w);
    else e.bits_out[bits_row + (col >> 4)] = (uint16_t)w;
  }
  st_shared_v4(dst, pk[0], pk[1], pk[2], pk[3]);
  st_shared_v4(dst + 16, pk[4], pk[5], pk[6], pk[7]);
}

// chunks c_first, c_first + c_step, ... < c_end of one accumulator row, two register buffers in ping-pong: the TMEM
// load (and the mask word) of the next chunk is in flight while this one is processed, with no register copies
template <bool kBias, bool kBitsOut, bool kMaskBits>
__device__ __forceinline__ void epilogue_row_lean(const EpilogueArgs& e, uint32_t trow, bool row_ok, int n0, int c_first,
                                                  int c_step, int c_end, long long bits_row) {
  const float slope = e.act == ACT_NONE ? 1.f : (e.act == ACT_RELU ? 0.f : e.leak);
  const float neg = e.mask_kind == ACT_LRELU ? e.leak : 0.f;
  uint32_t va[16], vb[16], ma = 0, mb = 0;
  int c = c_first;
  tmem_ld16(trow + c, va);
  if (kMaskBits && row_ok) ma = __ldg(e.mask_bits + bits_row + ((n0 + c) >> 4));
  while (true) {
    const int c1 = c + c_step;
    tmem_ld_wait16(va);
    if (c1 < c_end) {
      tmem_ld16(trow + c1, vb);
      if (kMaskBits && row_ok) mb = __ldg(e.mask_bits + bits_row + ((n0 + c1) >> 4));
    }
    if (row_ok) epilogue_chunk_lean<kBias, kBitsOut, kMaskBits>(e, va, n0 + c, slope, neg, ma, bits_row);
    if (c1 >= c_end) break;
    const int c2 = c1 + c_step;
    tmem_ld_wait16(vb);
    if (c2 < c_end) {
      tmem_ld16(trow + c2, va);
      if (kMaskBits && row_ok) ma = __ldg(e.mask_bits + bits_row + ((n0 + c2) >> 4));
    }
    if (row_ok) epilogue_chunk_lean<kBias, kBitsOut, kMaskBits>(e, vb, n0 + c1, slope, neg, mb, bits_row);
    if (c2 >= c_end) break;
    c = c2;
  }
}

// kLean: bit `sel` set = instantiate the lean path for that option combination (sel = bias 4 | bits_out 2 | mask_bits 1);
// combinations left out take the general path (a kernel with a lot of live state of its own keeps only what it needs)
template <bool kSimple, int kLean = 0xff>
__device__ __forceinline__ void epilogue_row(const EpilogueArgs& e, uint32_t trow, long long off, bool row_ok,
                                             int n0, int c_first, int c_step, int c_end) {
  // c_end: first column offset (relative to the tile) that must not be processed
  if (c_first >= c_end) return;
  if (kSimple && kLean) {
    // all chunks full, staged bf16 output, bias (if any) in shared memory, masks (if any) as bitmaps: the lean path
    const bool lean = e.stage_row != 0 && !e.out_f32 && !e.accumulate && e.alpha == 1.f && (c_end & 15) == 0 &&
                      (!e.bias || e.bias_smem) && (!e.mask_src || e.mask_bits) && e.partial_n == 0 &&
                      (e.act == ACT_NONE || e.act == ACT_RELU || e.act == ACT_LRELU);
    if (lean) {
      const long long brow = (e.mask_bits || e.bits_out) ? (off / e.row_elems) * e.bits_pitch : 0;
      const int sel = (e.bias ? 4 : 0) | (e.bits_out ? 2 : 0) | (e.mask_bits ? 1 : 0);
#define B200_LEAN_CASE(S, A, B, C)                                                                   \
  if ((kLean >> S) & 1) {                                                                            \
    if (sel == S) { epilogue_row_lean<A, B, C>(e, trow, row_ok, n0, c_first, c_step, c_end, brow); return; } \
  }
      B200_LEAN_CASE(0, false, false, false)
      B200_LEAN_CASE(1, false, false, true)
      B200_LEAN_CASE(2, false, true, false)
      B200_LEAN_CASE(3, false, true, true)
      B200_LEAN_CASE(4, true, false, false)
      B200_LEAN_CASE(5, true, false, true)
      B200_LEAN_CASE(6, true, true, false)
      B200_LEAN_CASE(7, true, true, true)
#undef B200_LEAN_CASE
    }
  }
  // word offset of this output row in the sign bitmaps (the output is dense: row = element offset / row width)
  const long long bits_row = (e.mask_bits || e.bits_out) ? (off / e.row_elems) * e.bits_pitch : 0;
  if (kSimple) {
    uint32_t vn[16], vc[16];
    tmem_ld16(trow + c_first, vn);
    ChunkSide nxt = epilogue_load_side(e, off, n0 + c_first, row_ok, bits_row);
#pragma unroll 1
    for (int c = c_first; c < c_end; c += c_step) {
      tmem_ld_wait16(vn);
#pragma unroll
      for (int j = 0; j < 16; ++j) vc[j] = vn[j];
      const ChunkSide cur = nxt;
      const int c1 = c + c_step;
      if (c1 < c_end) {
        tmem_ld16(trow + c1, vn);
        nxt = epilogue_load_side(e, off, n0 + c1, row_ok, bits_row);
      }
      if (row_ok) {
        if (e.partial_n) epilogue_add_partials(e, vc, c);
        epilogue_store16<kSimple>(e, vc, off, n0 + c, &cur, bits_row);
      }
    }
  } else {
    // general epilogues (fp32 out, value masks, tanh/sigmoid) keep one accumulator buffer: fewer registers
#pragma unroll 1
    for (int c = c_first; c < c_end; c += c_step) {
      uint32_t v[16];
      tmem_ld16(trow + c, v);
      const ChunkSide cur = epilogue_load_side(e, off, n0 + c, row_ok, bits_row);
      tmem_ld_wait16(v);
      if (row_ok) {
        if (e.partial_n) epilogue_add_partials(e, v, c);
        epilogue_store16<kSimple>(e, v, off, n0 + c, &cur, bits_row);
      }
    }
  }
}

}  // namespace b200
