// Fused epilogue shared by the tensor-core and the small-channel kernels:
//   v = alpha*acc (+ bias[col]);  v = act(v);  v *= act'(mask_src[off+col]);  store bf16|fp32.
// Order follows the reference's layer definition conv -> +bias -> activation
// (ops/layers.py:101-105); the mask multiply is the activation-gradient of the consumer layer
// (derivatives are expressed through the stored post-activation value).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include "tc_gemm.cuh"

namespace b200 {

__device__ __forceinline__ float act_fwd(float v, int act, float leak) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LRELU: return fmaxf(leak * v, v);
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
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
  if (e.mask_src && e.pipelined && col + 16 <= e.ncols) {
    const __nv_bfloat16* p = e.mask_src + off + col;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      m.lo = __ldg(reinterpret_cast<const uint4*>(p));
      m.hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
      m.loaded = true;
    }
  }
  return m;
}

// 16 consecutive columns [col, col+16) of one output row starting at element offset `off`.
__device__ __forceinline__ void epilogue_store16(const EpilogueArgs& e, const uint32_t* acc, long long off,
                                                 int col, const MaskChunk* pre = nullptr) {
  float v[16];
  const int nvalid = min(16, e.ncols - col);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]) * e.alpha;
  if (e.bias) {
    const float* bp = e.bias + col;
    if (nvalid == 16 && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + j));
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) v[j] += __ldg(bp + j);
    }
  }
  if (e.act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_fwd(v[j], e.act, e.leak);
  }
  const long long o = off + col;
  if (e.mask_src && pre && pre->loaded) {
    uint4 raw[2] = {pre->lo, pre->hi};
    const __nv_bfloat16* mv = reinterpret_cast<const __nv_bfloat16*>(raw);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= act_grad_from_out(__bfloat162float(mv[j]), e.mask_kind, e.leak);
  } else if (e.mask_src) {
    const __nv_bfloat16* m = e.mask_src + o;
    if (nvalid == 16 && ((reinterpret_cast<uintptr_t>(m) & 15) == 0)) {
      uint4 raw[2];
      raw[0] = __ldg(reinterpret_cast<const uint4*>(m));
      raw[1] = __ldg(reinterpret_cast<const uint4*>(m) + 1);
      const __nv_bfloat16* mv = reinterpret_cast<const __nv_bfloat16*>(raw);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= act_grad_from_out(__bfloat162float(mv[j]), e.mask_kind, e.leak);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) v[j] *= act_grad_from_out(__bfloat162float(m[j]), e.mask_kind, e.leak);
    }
  }
  if (e.out_f32) {
    float* dst = reinterpret_cast<float*>(e.out) + o;
    if (nvalid == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 f = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (e.accumulate) {
          float4 old = *reinterpret_cast<float4*>(dst + j);
          f.x += old.x; f.y += old.y; f.z += old.z; f.w += old.w;
        }
        *reinterpret_cast<float4*>(dst + j) = f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) dst[j] = e.accumulate ? dst[j] + v[j] : v[j];
    }
  } else {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.out) + o;
    if (nvalid == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *(reinterpret_cast<uint4*>(dst) + 1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) dst[j] = __float2bfloat16(v[j]);
    }
  }
}

}  // namespace b200
