// HBM-bound kernels of the training step: small-channel convolutions (image <-> first feature map),
// batch-norm, activation gradients, GEMV for the critic's 1-unit dense layer, gradient-penalty
// helpers, loss reductions, Philox noise and the fused Adam/RMSProp update.  All coalesced /
// vectorised; grids are sized as multiples of the SM count where the kernel is a grid-stride loop.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "epilogue.cuh"
#include "simt_kernels.cuh"

namespace b200 {

typedef __nv_bfloat16 bf16;

static int num_sms() {                 // cached per device (a process may drive several)
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}
static int std_max(int a, int b) { return a > b ? a : b; }
static int stride_grid(long long n, int threads, int per_thread = 1) {
  long long need = (n + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
  long long cap = (long long)num_sms() * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// =============================================================================================
// Small-channel convolutions.  "small" side has Cs channels with k*k*Cs <= kSmallK taps
// (image side: 3, 1 or 4 channels); "big" side has Cb channels, contiguous, one thread per channel.
// Weight layout [k,k,Cs,Cb] == TF's [kh,kw,in,out] for a conv whose input is the small side, and
// TF's conv2d_transpose layout [kh,kw,out,in] for a deconv whose output is the small side.
// =============================================================================================
constexpr int kSmallK = 80;

struct ConvGeom {
  int N, H, W, Cs;      // small-channel tensor  [N,H,W,Cs]   (the strided conv's INPUT)
  int Ho, Wo, Cb;       // big-channel tensor    [N,Ho,Wo,Cb] (the strided conv's OUTPUT)
  int k, stride, pad_t, pad_l;
};

// big[n,oh,ow,cb] = epi( sum_{r,s,cs} small[n, oh*st+r-pt, ow*st+s-pl, cs] * w[r,s,cs,cb] )
// block = one thread per cb (rounded up to a warp multiple); each block walks pixels.
__global__ void smallc_fprop_kernel(const bf16* __restrict__ xs, const bf16* __restrict__ w, ConvGeom g,
                                    EpilogueArgs e, long long npix, int pix_per_block) {
  __shared__ float patch[kSmallK];
  const int kk = g.k * g.k * g.Cs;
  const int cb = threadIdx.x;
  float wreg[kSmallK];
#pragma unroll
  for (int j = 0; j < kSmallK; ++j)
    wreg[j] = (j < kk && cb < g.Cb) ? __bfloat162float(w[(long long)j * g.Cb + cb]) : 0.f;
  const float bias = (e.bias && cb < g.Cb) ? e.bias[cb] : 0.f;
  long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  for (long long p = p0; p < p1; ++p) {
    const int ow = (int)(p % g.Wo);
    const int oh = (int)((p / g.Wo) % g.Ho);
    const int n = (int)(p / ((long long)g.Wo * g.Ho));
    __syncthreads();
    for (int j = threadIdx.x; j < kk; j += blockDim.x) {
      const int cs = j % g.Cs;
      const int s = (j / g.Cs) % g.k;
      const int r = j / (g.Cs * g.k);
      const int ih = oh * g.stride + r - g.pad_t, iw = ow * g.stride + s - g.pad_l;
      float v = 0.f;
      if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
        v = __bfloat162float(xs[(((long long)n * g.H + ih) * g.W + iw) * g.Cs + cs]);
      patch[j] = v;
    }
    __syncthreads();
    if (cb < g.Cb) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < kSmallK; ++j)
        if (j < kk) acc = fmaf(patch[j], wreg[j], acc);
      float v = acc * e.alpha + bias;
      v = act_fwd(v, e.act, e.leak);
      const long long o = p * g.Cb + cb;
      if (e.mask_src) v *= act_grad_from_out(__bfloat162float(e.mask_src[o]), e.mask_kind, e.leak);
      if (e.out_f32) reinterpret_cast<float*>(e.out)[o] = v;
      else reinterpret_cast<bf16*>(e.out)[o] = __float2bfloat16(v);
    }
  }
}

// small[n,h,w,cs] = epi( sum_{r,s valid} sum_cb big[n,(h+pt-r)/st,(w+pl-s)/st,cb] * w[r,s,cs,cb] )
// one warp per small-side pixel; lanes stride over cb (coalesced), Cs (<=4) accumulators each.
__global__ void smallc_dgrad_kernel(const bf16* __restrict__ big, const bf16* __restrict__ w, ConvGeom g,
                                    EpilogueArgs e, long long npix) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp0; p < npix; p += nwarps) {
    const int x = (int)(p % g.W);
    const int y = (int)((p / g.W) % g.H);
    const int n = (int)(p / ((long long)g.W * g.H));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < g.k; ++r) {
      const int th = y + g.pad_t - r;
      if (th < 0 || th % g.stride) continue;
      const int oh = th / g.stride;
      if (oh >= g.Ho) continue;
      for (int s = 0; s < g.k; ++s) {
        const int tw = x + g.pad_l - s;
        if (tw < 0 || tw % g.stride) continue;
        const int ow = tw / g.stride;
        if (ow >= g.Wo) continue;
        const bf16* brow = big + (((long long)n * g.Ho + oh) * g.Wo + ow) * g.Cb;
        const bf16* wrow = w + (long long)((r * g.k + s) * g.Cs) * g.Cb;
        for (int c = lane; c < g.Cb; c += 32) {
          const float b = __bfloat162float(brow[c]);
#pragma unroll
          for (int cs = 0; cs < 4; ++cs)
            if (cs < g.Cs) acc[cs] = fmaf(b, __bfloat162float(wrow[(long long)cs * g.Cb + c]), acc[cs]);
        }
      }
    }
#pragma unroll
    for (int cs = 0; cs < 4; ++cs) acc[cs] = warp_sum(acc[cs]);
    if (lane < g.Cs) {
      float v = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) * e.alpha;
      if (e.bias) v += e.bias[lane];
      v = act_fwd(v, e.act, e.leak);
      const long long o = p * g.Cs + lane;
      if (e.mask_src) v *= act_grad_from_out(__bfloat162float(e.mask_src[o]), e.mask_kind, e.leak);
      if (e.out_f32) reinterpret_cast<float*>(e.out)[o] = v;
      else reinterpret_cast<bf16*>(e.out)[o] = __float2bfloat16(v);
    }
  }
}

// dw[r,s,cs,cb] += alpha * sum_pixels small[n, oh*st+r-pt, ow*st+s-pl, cs] * big[n,oh,ow,cb]
// thread per cb keeps all k*k*Cs accumulators in registers; the small-side patch is broadcast from
// shared memory; one atomicAdd per (tap, cb) per block.
__global__ void smallc_wgrad_kernel(const bf16* __restrict__ xs, const bf16* __restrict__ big, float* dw,
                                    ConvGeom g, float alpha, long long npix, int pix_per_block) {
  __shared__ float patch[kSmallK];
  const int kk = g.k * g.k * g.Cs;
  const int cb = threadIdx.x;
  float acc[kSmallK];
#pragma unroll
  for (int j = 0; j < kSmallK; ++j) acc[j] = 0.f;
  long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  for (long long p = p0; p < p1; ++p) {
    const int ow = (int)(p % g.Wo);
    const int oh = (int)((p / g.Wo) % g.Ho);
    const int n = (int)(p / ((long long)g.Wo * g.Ho));
    __syncthreads();
    for (int j = threadIdx.x; j < kk; j += blockDim.x) {
      const int cs = j % g.Cs;
      const int s = (j / g.Cs) % g.k;
      const int r = j / (g.Cs * g.k);
      const int ih = oh * g.stride + r - g.pad_t, iw = ow * g.stride + s - g.pad_l;
      float v = 0.f;
      if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
        v = __bfloat162float(xs[(((long long)n * g.H + ih) * g.W + iw) * g.Cs + cs]);
      patch[j] = v;
    }
    __syncthreads();
    if (cb < g.Cb) {
      const float b = __bfloat162float(big[p * g.Cb + cb]);
#pragma unroll
      for (int j = 0; j < kSmallK; ++j)
        if (j < kk) acc[j] = fmaf(patch[j], b, acc[j]);
    }
  }
  if (cb < g.Cb) {
#pragma unroll
    for (int j = 0; j < kSmallK; ++j)
      if (j < kk) atomicAdd(dw + (long long)j * g.Cb + cb, acc[j] * alpha);
  }
}


// ---------------------------------------------------------------------------------------------
// im2col / col2im for the small-channel layers: they turn the image-side conv into a plain GEMM
// that the tcgen05 kernels run (K = k*k*Cs padded to Kp, a multiple of 8 for TMA alignment).
// ---------------------------------------------------------------------------------------------
// A[m][j] = small[n, oh*st+r-pt, ow*st+s-pl, cs]  (j = (r*k+s)*Cs+cs; zero for padding and j >= k*k*Cs)
// one thread per 16-byte chunk of an output row (8 consecutive j): the (r, s, offset) decomposition of j
// comes from a shared-memory table built once per block, stores are 16-byte and fully coalesced.
__global__ void im2col_small_kernel(const bf16* __restrict__ xs, bf16* __restrict__ A, ConvGeom g, int Kp,
                                    long long total_chunks) {
  __shared__ short tr[128], ts[128];
  __shared__ int toff[128];
  const int kk = g.k * g.k * g.Cs;
  for (int j = threadIdx.x; j < Kp && j < 128; j += blockDim.x) {
    const int cs = j % g.Cs, s = (j / g.Cs) % g.k, r = j / (g.Cs * g.k);
    tr[j] = (short)(j < kk ? r : -1000);
    ts[j] = (short)s;
    toff[j] = (r * g.W + s) * g.Cs + cs;
  }
  __syncthreads();
  const int cpr = Kp / 8;                                   // chunks per row
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_chunks; i += stride) {
    const int ch = (int)(i % cpr);
    const long long m = i / cpr;
    const int ow = (int)(m % g.Wo);
    const int oh = (int)((m / g.Wo) % g.Ho);
    const int n = (int)(m / ((long long)g.Wo * g.Ho));
    const int ih0 = oh * g.stride - g.pad_t, iw0 = ow * g.stride - g.pad_l;
    const bf16* src = xs + (((long long)n * g.H + ih0) * g.W + iw0) * g.Cs;
    uint4 outv;
    bf16* o = reinterpret_cast<bf16*>(&outv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = ch * 8 + e;
      const int ih = ih0 + tr[j], iw = iw0 + ts[j];
      const bool ok = ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
      o[e] = ok ? src[toff[j]] : __float2bfloat16(0.f);
    }
    *reinterpret_cast<uint4*>(A + m * Kp + ch * 8) = outv;
  }
}
// specialisation for k = 5, Cs = 3, Kp = 80 (IWGAN c1 and the generator's last deconv): one thread builds a whole
// 160-byte row -- each of the 5 filter rows is 15 consecutive bf16 of the input -- and stores it with ten
// 16-byte writes (a quarter of the instructions of the generic kernel)
__global__ void im2col_k5c3_kernel(const bf16* __restrict__ xs, bf16* __restrict__ A, ConvGeom g, long long rows) {
  const unsigned short* x16 = reinterpret_cast<const unsigned short*>(xs);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < rows; m += stride) {
    const int ow = (int)(m % g.Wo);
    const int oh = (int)((m / g.Wo) % g.Ho);
    const int n = (int)(m / ((long long)g.Wo * g.Ho));
    const int ih0 = oh * g.stride - g.pad_t, iw0 = ow * g.stride - g.pad_l;
    const long long base = (((long long)n * g.H + ih0) * g.W + iw0) * 3;
    uint32_t w[40];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const bool rok = (unsigned)(ih0 + r) < (unsigned)g.H;
      const long long rb = base + (long long)r * g.W * 3;
#pragma unroll
      for (int e = 0; e < 15; ++e) {
        const bool ok = rok && (unsigned)(iw0 + e / 3) < (unsigned)g.W;
        const uint32_t v = ok ? (uint32_t)x16[rb + e] : 0u;
        const int j = r * 15 + e;
        if (j & 1) w[j >> 1] |= v << 16;
        else w[j >> 1] = v;
      }
    }
    w[38] = 0u; w[39] = 0u;
    uint4* dst = reinterpret_cast<uint4*>(A + m * 80);
#pragma unroll
    for (int q = 0; q < 10; ++q) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  }
}
int im2col_small(const void* xs, void* A, const SmallConvArgs& a, int Kp, cudaStream_t st) {
  ConvGeom g{a.N, a.H, a.W, a.Cs, a.Ho, a.Wo, a.Cb, a.k, a.stride, a.pad_t, a.pad_l};
  if (Kp > 128) return -1;
  if (a.k == 5 && a.Cs == 3 && Kp == 80 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
    const long long rows = (long long)a.N * a.Ho * a.Wo;
    im2col_k5c3_kernel<<<stride_grid(rows, 128, 1), 128, 0, st>>>((const bf16*)xs, (bf16*)A, g, rows);
    return 0;
  }
  const long long total = (long long)a.N * a.Ho * a.Wo * (Kp / 8);
  im2col_small_kernel<<<stride_grid(total, 256, 1), 256, 0, st>>>((const bf16*)xs, (bf16*)A, g, Kp, total);
  return 0;
}
// Wt[cb][j] = w[j][cb] for j < kk, 0 for kk <= j < Kp   (K-major B operand of the fprop GEMM)
__global__ void wpad_transpose_kernel(const bf16* __restrict__ w, bf16* __restrict__ wt, int kk, int Cb, int Kp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cb * Kp) return;
  const int j = i % Kp, cb = i / Kp;
  wt[i] = j < kk ? w[(long long)j * Cb + cb] : __float2bfloat16(0.f);
}
int wpad_transpose(const void* w, void* wt, int kk, int Cb, int Kp, cudaStream_t st) {
  wpad_transpose_kernel<<<(Cb * Kp + 255) / 256, 256, 0, st>>>((const bf16*)w, (bf16*)wt, kk, Cb, Kp);
  return 0;
}
// small[n,h,w,cs] = epi( sum_{r,s valid} T[(n,oh,ow)][(r*k+s)*Cs+cs] ),  T fp32 with row stride Kp
// one thread per small-side pixel: the valid taps are resolved once, the Cs channels are adjacent in T.
__global__ void col2im_small_kernel(const float* __restrict__ T, ConvGeom g, int Kp, EpilogueArgs e, long long npix) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
    const int x = (int)(p % g.W);
    const int y = (int)((p / g.W) % g.H);
    const int n = (int)(p / ((long long)g.W * g.H));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // taps r = r0 + jr*stride with r0 = (y + pad_t) mod stride read output row oh0 - jr (same along x): uniform loop
    // bounds and predicated loads instead of a divergent walk over all k*k taps (as in img_dgrad_kernel)
    const int st = g.stride, nj = (g.k + st - 1) / st;
    const int ty = y + g.pad_t, tx = x + g.pad_l;
    const int oh0 = ty / st, r0 = ty - oh0 * st, ow0 = tx / st, s0 = tx - ow0 * st;
    for (int jr = 0; jr < nj; ++jr) {
      const int r = r0 + jr * st, oh = oh0 - jr;
      const bool vr = r < g.k && oh >= 0 && oh < g.Ho;
      for (int js = 0; js < nj; ++js) {
        const int sx = s0 + js * st, ow = ow0 - js;
        const bool v = vr && sx < g.k && ow >= 0 && ow < g.Wo;
        const float* t = T + (v ? (((long long)n * g.Ho + oh) * g.Wo + ow) * Kp + (r * g.k + sx) * g.Cs : 0);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < g.Cs) acc[c] += v ? t[c] : 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= g.Cs) break;
      const long long i = p * g.Cs + c;
      float v = acc[c] * e.alpha;
      if (e.bias) v += e.bias[c];
      v = act_fwd(v, e.act, e.leak);
      if (e.mask_src) v *= act_grad_from_out(__bfloat162float(e.mask_src[i]), e.mask_kind, e.leak);
      if (e.out_f32) reinterpret_cast<float*>(e.out)[i] = v;
      else reinterpret_cast<bf16*>(e.out)[i] = __float2bfloat16(v);
    }
  }
}
int col2im_small(const float* T, const SmallConvArgs& a, int Kp, cudaStream_t st) {
  ConvGeom g{a.N, a.H, a.W, a.Cs, a.Ho, a.Wo, a.Cb, a.k, a.stride, a.pad_t, a.pad_l};
  EpilogueArgs e{a.bias, a.act, a.leak, (const bf16*)a.mask_src, a.mask_kind, 1.f, a.out, a.out_f32, 0, a.Cs, 0};
  const long long npix = (long long)a.N * a.H * a.W;
  col2im_small_kernel<<<stride_grid(npix, 128, 1), 128, 0, st>>>(T, g, Kp, e, npix);
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Convolutions whose OUTPUT has <= 4 channels (pix2pix PatchGAN head m5: 512 -> 1, hem/models/pix2pix.py:256).
// The forward runs on the tensor cores (N tile 16); TMA cannot address a 1-channel gradient tensor, so the
// two backward products are coalesced SIMT kernels (they are tiny: B*8*8 output pixels).
// ---------------------------------------------------------------------------------------------
// dx[n,h,w,ci] = epi( sum_{r,s valid} sum_co dy[n,(h+pt-r)/st,(w+pl-s)/st,co] * w[r,s,ci,co] )
__global__ void smallout_dgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ w, ConvGeom g,
                                      EpilogueArgs e, long long total) {
  // here g.Cs = Cin (big), g.Cb = Cout (small): reuse of the struct with swapped meaning is avoided by
  // passing Cin in g.Cs and Cout in g.Cb explicitly
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ci = (int)(i % g.Cs);
    const long long p = i / g.Cs;
    const int x = (int)(p % g.W);
    const int y = (int)((p / g.W) % g.H);
    const int n = (int)(p / ((long long)g.W * g.H));
    float acc = 0.f;
    for (int r = 0; r < g.k; ++r) {
      const int th = y + g.pad_t - r;
      if (th < 0 || th % g.stride) continue;
      const int oh = th / g.stride;
      if (oh >= g.Ho) continue;
      for (int s = 0; s < g.k; ++s) {
        const int tw = x + g.pad_l - s;
        if (tw < 0 || tw % g.stride) continue;
        const int ow = tw / g.stride;
        if (ow >= g.Wo) continue;
        const bf16* d = dy + (((long long)n * g.Ho + oh) * g.Wo + ow) * g.Cb;
        const bf16* wr = w + ((long long)(r * g.k + s) * g.Cs + ci) * g.Cb;
        for (int co = 0; co < g.Cb; ++co) acc = fmaf(__bfloat162float(d[co]), __bfloat162float(wr[co]), acc);
      }
    }
    float v = acc * e.alpha;
    if (e.bias) v += e.bias[ci];
    v = act_fwd(v, e.act, e.leak);
    if (e.mask_src) v *= act_grad_from_out(__bfloat162float(e.mask_src[i]), e.mask_kind, e.leak);
    if (e.out_f32) reinterpret_cast<float*>(e.out)[i] = v;
    else reinterpret_cast<bf16*>(e.out)[i] = __float2bfloat16(v);
  }
}
// dw[r,s,ci,co] += alpha * sum_{n,oh,ow} x[n,oh*st+r-pt,ow*st+s-pl,ci] * dy[n,oh,ow,co]
// block (tap, pixel range); threads over ci (coalesced reads of x rows); one atomic per (tap,ci,co) per block
__global__ void smallout_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* dw, ConvGeom g,
                                      float alpha, long long npix, int pix_per_block) {
  const int tap = blockIdx.y;
  const int r = tap / g.k, s = tap % g.k;
  long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  for (int ci = threadIdx.x; ci < g.Cs; ci += blockDim.x) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long p = p0; p < p1; ++p) {
      const int ow = (int)(p % g.Wo);
      const int oh = (int)((p / g.Wo) % g.Ho);
      const int n = (int)(p / ((long long)g.Wo * g.Ho));
      const int ih = oh * g.stride + r - g.pad_t, iw = ow * g.stride + s - g.pad_l;
      if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
      const float xv = __bfloat162float(x[(((long long)n * g.H + ih) * g.W + iw) * g.Cs + ci]);
#pragma unroll
      for (int co = 0; co < 4; ++co)
        if (co < g.Cb) acc[co] = fmaf(xv, __bfloat162float(dy[p * g.Cb + co]), acc[co]);
    }
#pragma unroll
    for (int co = 0; co < 4; ++co)
      if (co < g.Cb) atomicAdd(dw + ((long long)tap * g.Cs + ci) * g.Cb + co, acc[co] * alpha);
  }
}
// ---- one output channel (the PatchGAN head itself): the whole filter [k*k, Cin] sits in shared memory, a warp owns an
// input pixel at a time, gathers its <= ceil(k/stride)^2 (tap, dy) pairs and every lane produces 8 channels per step
// with 16-byte accesses (the generic kernel above walked all k*k taps per element with 2-byte loads: 68 us for the
// 4 MB gradient of pix2pix's m5 at batch 16)
__global__ void smallout1_dgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ w, ConvGeom g,
                                       EpilogueArgs e, int npix) {
  extern __shared__ uint4 sw4[];                       // [k*k][Cs] bf16
  const int C = g.Cs, kk = g.k * g.k;
  for (int i = threadIdx.x; i < kk * C / 8; i += blockDim.x) sw4[i] = reinterpret_cast<const uint4*>(w)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int st = g.stride, nj = (g.k + st - 1) / st;
  for (int p = blockIdx.x * nw + wid; p < npix; p += gridDim.x * nw) {
    const int x = p % g.W, y = (p / g.W) % g.H, n = p / (g.W * g.H);
    const int ty = y + g.pad_t, tx = x + g.pad_l;
    const int oh0 = ty / st, r0 = ty - oh0 * st, ow0 = tx / st, s0 = tx - ow0 * st;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int jr = 0; jr < nj; ++jr) {
        const int r = r0 + jr * st, oh = oh0 - jr;
        if (r >= g.k || oh < 0 || oh >= g.Ho) continue;
        for (int js = 0; js < nj; ++js) {
          const int sx = s0 + js * st, ow = ow0 - js;
          if (sx >= g.k || ow < 0 || ow >= g.Wo) continue;
          const float d = __bfloat162float(dy[(n * g.Ho + oh) * g.Wo + ow]);
          const uint4 wv = sw4[((r * g.k + sx) * C + c0) >> 3];
          const uint32_t w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[2 * j] = fmaf(d, __uint_as_float(w4[j] << 16), acc[2 * j]);
            acc[2 * j + 1] = fmaf(d, __uint_as_float(w4[j] & 0xffff0000u), acc[2 * j + 1]);
          }
        }
      }
      const long long o = (long long)p * C + c0;
      uint4 mv = make_uint4(0, 0, 0, 0);
      if (e.mask_src) mv = *reinterpret_cast<const uint4*>(e.mask_src + o);
      const uint32_t m4[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = acc[j] * e.alpha;
        if (e.bias) v += e.bias[c0 + j];
        v = act_fwd(v, e.act, e.leak);
        if (e.mask_src) {
          const float mval = (j & 1) ? __uint_as_float(m4[j >> 1] & 0xffff0000u) : __uint_as_float(m4[j >> 1] << 16);
          v *= act_grad_from_out(mval, e.mask_kind, e.leak);
        }
        acc[j] = v;
      }
      if (e.out_f32) {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + o);
        op[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        op[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      } else {
        uint32_t o4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
          o4[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.out) + o) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
}
// dw[tap, ci] += alpha * sum_p x[n, oh*st+r-pt, ow*st+s-pl, ci] * dy[p]: block = (8 output pixels, tap), a thread owns
// a channel pair; the 8 x-loads of a thread are independent (the generic kernel's were one dependent chain per pixel)
__global__ void smallout1_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* dw, ConvGeom g,
                                       float alpha, int npix) {
  constexpr int kPix = 8;
  const int tap = blockIdx.y, r = tap / g.k, s = tap % g.k;
  const int p0 = blockIdx.x * kPix;
  __shared__ float sd[kPix];
  __shared__ int soff[kPix];                            // element offset of the tap's input pixel, -1 = padding
  if (threadIdx.x < kPix) {
    const int p = p0 + threadIdx.x;
    int off = -1;
    float d = 0.f;
    if (p < npix) {
      const int ow = p % g.Wo, oh = (p / g.Wo) % g.Ho, n = p / (g.Wo * g.Ho);
      const int ih = oh * g.stride + r - g.pad_t, iw = ow * g.stride + s - g.pad_l;
      if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) { off = ((n * g.H + ih) * g.W + iw) * g.Cs; d = __bfloat162float(dy[p]); }
    }
    soff[threadIdx.x] = off; sd[threadIdx.x] = d;
  }
  __syncthreads();
  for (int c = threadIdx.x * 2; c < g.Cs; c += blockDim.x * 2) {
    uint32_t v[kPix];
#pragma unroll
    for (int i = 0; i < kPix; ++i) v[i] = soff[i] >= 0 ? *reinterpret_cast<const uint32_t*>(x + soff[i] + c) : 0u;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < kPix; ++i) {
      a0 = fmaf(__uint_as_float(v[i] << 16), sd[i], a0);
      a1 = fmaf(__uint_as_float(v[i] & 0xffff0000u), sd[i], a1);
    }
    atomicAdd(dw + (long long)tap * g.Cs + c, a0 * alpha);
    atomicAdd(dw + (long long)tap * g.Cs + c + 1, a1 * alpha);
  }
}
// y[p] = epi( sum_{r,s in bounds} x[n, oh*st+r-pt, ow*st+s-pl, :] . w[r,s,:] ): one block of 128 threads per output pixel,
// 16-byte loads of the x rows (the tcgen05 route pads the single column to an N tile of 16 and needs split-K: 20 us
// for 1024 dot products of 8192 elements)
__global__ void smallout1_fprop_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, ConvGeom g,
                                       EpilogueArgs e) {
  __shared__ float part[4];
  const int p = blockIdx.x;
  const int ow = p % g.Wo, oh = (p / g.Wo) % g.Ho, n = p / (g.Wo * g.Ho);
  const int C8 = g.Cs >> 3, kk = g.k * g.k;
  float acc = 0.f;
  for (int i = threadIdx.x; i < kk * C8; i += blockDim.x) {
    const int tap = i / C8, c0 = (i - tap * C8) << 3;
    const int r = tap / g.k, s = tap - r * g.k;
    const int ih = oh * g.stride + r - g.pad_t, iw = ow * g.stride + s - g.pad_l;
    if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
    const uint4 xv = *reinterpret_cast<const uint4*>(x + ((long long)(n * g.H + ih) * g.W + iw) * g.Cs + c0);
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + (long long)tap * g.Cs + c0));
    const uint32_t x4[4] = {xv.x, xv.y, xv.z, xv.w}, w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc = fmaf(__uint_as_float(x4[j] << 16), __uint_as_float(w4[j] << 16), acc);
      acc = fmaf(__uint_as_float(x4[j] & 0xffff0000u), __uint_as_float(w4[j] & 0xffff0000u), acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = (part[0] + part[1] + part[2] + part[3]) * e.alpha;
    if (e.bias) v += e.bias[0];
    v = act_fwd(v, e.act, e.leak);
    if (e.mask_src) v *= act_grad_from_out(__bfloat162float(e.mask_src[p]), e.mask_kind, e.leak);
    if (e.out_f32) reinterpret_cast<float*>(e.out)[p] = v;
    else reinterpret_cast<bf16*>(e.out)[p] = __float2bfloat16(v);
  }
}
// a.Cs = Cin, a.Cb = Cout (== 1); returns -1 when the shape is not taken
int smallout_fprop(const void* x, const void* w, const SmallConvArgs& a, cudaStream_t st) {
  const long long npix = (long long)a.N * a.Ho * a.Wo;
  if (a.Cb != 1 || a.Cs % 8 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) ||
      npix >= (1 << 30) || (long long)a.N * a.H * a.W >= (1 << 30))
    return -1;
  ConvGeom g{a.N, a.H, a.W, a.Cs, a.Ho, a.Wo, a.Cb, a.k, a.stride, a.pad_t, a.pad_l};
  EpilogueArgs e{a.bias, a.act, a.leak, (const bf16*)a.mask_src, a.mask_kind, 1.f, a.out, a.out_f32, 0, 1, 0};
  smallout1_fprop_kernel<<<(unsigned)npix, 128, 0, st>>>((const bf16*)x, (const bf16*)w, g, e);
  return 0;
}
int smallout_dgrad(const void* dy, const void* w, const SmallConvArgs& a, cudaStream_t st) {
  // a.Cs = Cin (many channels, the conv input), a.Cb = Cout (<= 4)
  ConvGeom g{a.N, a.H, a.W, a.Cs, a.Ho, a.Wo, a.Cb, a.k, a.stride, a.pad_t, a.pad_l};
  EpilogueArgs e{a.bias, a.act, a.leak, (const bf16*)a.mask_src, a.mask_kind, 1.f, a.out, a.out_f32, 0, a.Cs, 0};
  const long long total = (long long)a.N * a.H * a.W * a.Cs;
  const size_t wbytes = (size_t)a.k * a.k * a.Cs * 2;
  const uintptr_t al = reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.mask_src);
  if (a.Cb == 1 && a.Cs % 8 == 0 && wbytes <= 48 * 1024 && (al & 15) == 0 && total < (1ll << 31)) {
    const int npix = a.N * a.H * a.W;
    int blocks = (npix + 7) / 8;
    if (blocks > num_sms() * 4) blocks = num_sms() * 4;
    smallout1_dgrad_kernel<<<blocks, 256, wbytes, st>>>((const bf16*)dy, (const bf16*)w, g, e, npix);
    return 0;
  }
  smallout_dgrad_kernel<<<stride_grid(total, 256, 2), 256, 0, st>>>((const bf16*)dy, (const bf16*)w, g, e, total);
  return 0;
}
int smallout_wgrad(const void* x, const void* dy, float* dw, const SmallConvArgs& a, float alpha, cudaStream_t st) {
  if (a.Cb > 4) return -1;
  ConvGeom g{a.N, a.H, a.W, a.Cs, a.Ho, a.Wo, a.Cb, a.k, a.stride, a.pad_t, a.pad_l};
  const long long npix = (long long)a.N * a.Ho * a.Wo;
  if (a.Cb == 1 && a.Cs % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0 &&
      (long long)a.N * a.H * a.W * a.Cs < (1ll << 31) && npix < (1 << 30)) {
    smallout1_wgrad_kernel<<<dim3((unsigned)((npix + 7) / 8), a.k * a.k), 256, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, g,
                                                                                       alpha, (int)npix);
    return 0;
  }
  int blocks = num_sms() * 2 / (a.k * a.k) + 1;
  int ppb = (int)((npix + blocks - 1) / blocks);
  if (ppb < 1) ppb = 1;
  blocks = (int)((npix + ppb - 1) / ppb);
  smallout_wgrad_kernel<<<dim3(blocks, a.k * a.k), 256, 0, st>>>((const bf16*)x, (const bf16*)dy, dw, g, alpha, npix, ppb);
  return 0;
}

static int round_up32(int v) { return (v + 31) / 32 * 32; }

int smallc_fprop(const void* xs, const void* w, const SmallConvArgs& a, cudaStream_t st) {
  ConvGeom g{a.N, a.H, a.W, a.Cs, a.Ho, a.Wo, a.Cb, a.k, a.stride, a.pad_t, a.pad_l};
  if (a.k * a.k * a.Cs > kSmallK || a.Cb > 1024) return -1;
  EpilogueArgs e{a.bias, a.act, a.leak, (const bf16*)a.mask_src, a.mask_kind, 1.f, a.out, a.out_f32, 0, a.Cb, 0};
  const long long npix = (long long)a.N * a.Ho * a.Wo;
  const int threads = round_up32(a.Cb);
  long long blocks = (long long)num_sms() * (threads <= 256 ? 8 : 4);
  int ppb = (int)((npix + blocks - 1) / blocks);
  if (ppb < 1) ppb = 1;
  blocks = (npix + ppb - 1) / ppb;
  smallc_fprop_kernel<<<(int)blocks, threads, 0, st>>>((const bf16*)xs, (const bf16*)w, g, e, npix, ppb);
  return 0;
}

int smallc_dgrad(const void* big, const void* w, const SmallConvArgs& a, cudaStream_t st) {
  ConvGeom g{a.N, a.H, a.W, a.Cs, a.Ho, a.Wo, a.Cb, a.k, a.stride, a.pad_t, a.pad_l};
  if (a.Cs > 4) return -1;
  EpilogueArgs e{a.bias, a.act, a.leak, (const bf16*)a.mask_src, a.mask_kind, 1.f, a.out, a.out_f32, 0, a.Cs, 0};
  const long long npix = (long long)a.N * a.H * a.W;
  const int threads = 256;
  long long blocks = (npix * 32 + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  smallc_dgrad_kernel<<<(int)blocks, threads, 0, st>>>((const bf16*)big, (const bf16*)w, g, e, npix);
  return 0;
}

int smallc_wgrad(const void* xs, const void* big, float* dw, const SmallConvArgs& a, float alpha,
                 cudaStream_t st) {
  ConvGeom g{a.N, a.H, a.W, a.Cs, a.Ho, a.Wo, a.Cb, a.k, a.stride, a.pad_t, a.pad_l};
  if (a.k * a.k * a.Cs > kSmallK || a.Cb > 1024) return -1;
  const long long npix = (long long)a.N * a.Ho * a.Wo;
  const int threads = round_up32(a.Cb);
  long long blocks = (long long)num_sms() * 4;
  int ppb = (int)((npix + blocks - 1) / blocks);
  if (ppb < 1) ppb = 1;
  blocks = (npix + ppb - 1) / ppb;
  smallc_wgrad_kernel<<<(int)blocks, threads, 0, st>>>((const bf16*)xs, (const bf16*)big, dw, g, alpha, npix, ppb);
  return 0;
}

// =============================================================================================
// Elementwise
// =============================================================================================
__global__ void maskmul_kernel(const bf16* __restrict__ gsrc, const bf16* __restrict__ a, bf16* out,
                               long long n, int kind, float leak) {
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      uint4 gv = *reinterpret_cast<const uint4*>(gsrc + i);
      uint4 av = *reinterpret_cast<const uint4*>(a + i);
      const bf16* gp = reinterpret_cast<const bf16*>(&gv);
      const bf16* ap = reinterpret_cast<const bf16*>(&av);
      uint4 ov;
      bf16* op = reinterpret_cast<bf16*>(&ov);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        op[j] = __float2bfloat16(__bfloat162float(gp[j]) * act_grad_from_out(__bfloat162float(ap[j]), kind, leak));
      *reinterpret_cast<uint4*>(out + i) = ov;
    } else {
      for (long long j = i; j < n; ++j)
        out[j] = __float2bfloat16(__bfloat162float(gsrc[j]) * act_grad_from_out(__bfloat162float(a[j]), kind, leak));
    }
  }
}
int maskmul(const void* g, const void* a, void* out, long long n, int kind, float leak, cudaStream_t st) {
  maskmul_kernel<<<stride_grid(n, 256, 8), 256, 0, st>>>((const bf16*)g, (const bf16*)a, (bf16*)out, n, kind, leak);
  return 0;
}

// out = act(in * mul + add) elementwise; fp32, bf16 or uint8 in, bf16 or fp32 out.  With uint8 input this is the
// input stage of the reference's pipeline fused into the model's first op: data.py:21-22,29 casts the decoded
// image bytes to float32 and divides by 255, models/gan.py:50 / cnn.py:31 rescale to [-1,1] -- here the bytes go
// host -> device as they are (a quarter of the fp32 traffic) and are normalised in the same pass (the caller
// folds the 1/255 into mul).
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(uint8_t v) { return (float)v; }
__device__ __forceinline__ void from_f32(float& o, float v) { o = v; }
__device__ __forceinline__ void from_f32(bf16& o, float v) { o = __float2bfloat16(v); }
template <typename TI, typename TO>
__global__ void affine_act_kernel(const TI* __restrict__ in, TO* out, long long n, float mul, float add, int act,
                                  float leak) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    from_f32(out[i], act_fwd(fmaf(to_f32(in[i]), mul, add), act, leak));
}
int affine_act(const void* in, int in_type, void* out, int out_f32, long long n, float mul, float add, int act,
               float leak, cudaStream_t st) {
  const int grid = stride_grid(n, 256, 4);
#define AFF(TI, TO) affine_act_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)in, (TO*)out, n, mul, add, act, leak)
  if (in_type == 2) { if (out_f32) AFF(uint8_t, float); else AFF(uint8_t, bf16); }
  else if (in_type == 1) { if (out_f32) AFF(float, float); else AFF(float, bf16); }
  else if (in_type == 0) { if (out_f32) AFF(bf16, float); else AFF(bf16, bf16); }
  else return -1;
#undef AFF
  return 0;
}

// out[i] (+)= a[i]*sa*(*dev_a or 1) + b[i]*sb   (fp32/bf16 mixes through float conversion)
template <typename TA, typename TB, typename TO>
__global__ void axpby_kernel(const TA* __restrict__ a, float sa, const float* dev_sa, const TB* __restrict__ b,
                             float sb, TO* out, long long n) {
  const float s = dev_sa ? sa * (*dev_sa) : sa;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = (float)a[i] * s;
    if (b) v += (float)b[i] * sb;
    out[i] = (TO)v;
  }
}
int axpby(const void* a, int a_f32, float sa, const float* dev_sa, const void* b, int b_f32, float sb, void* out,
          int out_f32, long long n, cudaStream_t st) {
  const int grid = stride_grid(n, 256, 4);
#define AXPBY_CASE(TA, TB, TO) axpby_kernel<TA, TB, TO><<<grid, 256, 0, st>>>((const TA*)a, sa, dev_sa, (const TB*)b, sb, (TO*)out, n)
  const int key = (a_f32 ? 4 : 0) | (b_f32 ? 2 : 0) | (out_f32 ? 1 : 0);
  switch (key) {
    case 0: AXPBY_CASE(bf16, bf16, bf16); break;
    case 1: AXPBY_CASE(bf16, bf16, float); break;
    case 2: AXPBY_CASE(bf16, float, bf16); break;
    case 3: AXPBY_CASE(bf16, float, float); break;
    case 4: AXPBY_CASE(float, bf16, bf16); break;
    case 5: AXPBY_CASE(float, bf16, float); break;
    case 6: AXPBY_CASE(float, float, bf16); break;
    default: AXPBY_CASE(float, float, float); break;
  }
#undef AXPBY_CASE
  return 0;
}

// out = a + b*c (a optional): VAE reparameterisation z = mu + sigma*eps and its gradient g*eps (models/vae.py:127-128)
__global__ void mul_add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const bf16* __restrict__ c,
                               bf16* out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = __bfloat162float(b[i]) * __bfloat162float(c[i]);
    if (a) v += __bfloat162float(a[i]);
    out[i] = __float2bfloat16(v);
  }
}
int mul_add(const void* a, const void* b, const void* c, void* out, long long n, cudaStream_t st) {
  mul_add_kernel<<<stride_grid(n, 256, 4), 256, 0, st>>>((const bf16*)a, (const bf16*)b, (const bf16*)c, (bf16*)out, n);
  return 0;
}

__global__ void fill_kernel(float* out, long long n, float v) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}
int fill_f32(float* out, long long n, float v, cudaStream_t st) {
  fill_kernel<<<stride_grid(n, 256, 4), 256, 0, st>>>(out, n, v);
  return 0;
}

// x_hat[b,:] = x[b,:] + alpha[b]*(g[b,:]-x[b,:])      (models/gan.py:224-226)
__global__ void interp_kernel(const bf16* __restrict__ x, const bf16* __restrict__ g, const float* __restrict__ alpha,
                              bf16* out, int B, int D) {
  const long long n = (long long)B * D;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = alpha[i / D];
    const float xv = __bfloat162float(x[i]);
    out[i] = __float2bfloat16(xv + a * (__bfloat162float(g[i]) - xv));
  }
}
int interp(const void* x, const void* g, const float* alpha, void* out, int B, int D, cudaStream_t st) {
  interp_kernel<<<stride_grid((long long)B * D, 256, 4), 256, 0, st>>>((const bf16*)x, (const bf16*)g, alpha, (bf16*)out, B, D);
  return 0;
}
// rowscale: out[b,:] = in[b,:] * s[b] * mul   (backward of interp w.r.t. g and x)
__global__ void rowscale_kernel(const bf16* __restrict__ in, const float* __restrict__ s, float mul, float add,
                                bf16* out, int B, int D) {
  const long long n = (long long)B * D;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16(__bfloat162float(in[i]) * (s[i / D] * mul + add));
}
int rowscale(const void* in, const float* s, float mul, float add, void* out, int B, int D, cudaStream_t st) {
  rowscale_kernel<<<stride_grid((long long)B * D, 256, 4), 256, 0, st>>>((const bf16*)in, s, mul, add, (bf16*)out, B, D);
  return 0;
}

// Epilogue of a split-K tap GEMM: out[i] = act(ws[i] + bias[i % C]) * act'(mask[i]) over the fp32 workspace the K
// slices reduced into (same element order as the output tensor, channels innermost).
template <typename TO>
__global__ void splitk_finalize_kernel(const float* __restrict__ ws, TO* __restrict__ out, long long n, int C,
                                       const float* __restrict__ bias, int act, float leak,
                                       const bf16* __restrict__ mask, int mask_kind) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = ws[i];
    if (bias) v += __ldg(bias + (int)(i % C));
    v = act_fwd(v, act, leak);
    if (mask) v *= act_grad_from_out(__bfloat162float(mask[i]), mask_kind, leak);
    from_f32(out[i], v);
  }
}
int splitk_finalize(const float* ws, void* out, int out_f32, long long n, int C, const float* bias, int act, float leak,
                    const void* mask, int mask_kind, cudaStream_t st) {
  const int grid = stride_grid(n, 256, 2);
  if (out_f32) splitk_finalize_kernel<float><<<grid, 256, 0, st>>>(ws, (float*)out, n, C, bias, act, leak, (const bf16*)mask, mask_kind);
  else splitk_finalize_kernel<bf16><<<grid, 256, 0, st>>>(ws, (bf16*)out, n, C, bias, act, leak, (const bf16*)mask, mask_kind);
  return 0;
}

// out[r, out_off + c] = in[r, in_off + c] * act'(mask[r, c])  for c < cols: column-slice copy between row-major bf16
// matrices with different row strides.  Channel concatenation (pix2pix skip connections, hem/models/pix2pix.py:
// 210-222) writes each piece into its slice of the concat buffer; its backward reads the slice back and applies
// the piece's activation gradient in the same pass; also strips / restores zero channel padding.
template <int kVec>
__global__ void slice_cols_kernel(const bf16* __restrict__ in, long long in_ld, int in_off, bf16* __restrict__ out,
                                  long long out_ld, int out_off, long long rows, int cols,
                                  const bf16* __restrict__ mask, int mask_kind, float leak) {
  const int cpr = cols / kVec;                               // chunks per row
  const long long total = rows * cpr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / cpr;
    const int c = (int)(i - r * cpr) * kVec;
    const bf16* src = in + r * in_ld + in_off + c;
    bf16* dst = out + r * out_ld + out_off + c;
    if (kVec == 8) {
      uint4 v = *reinterpret_cast<const uint4*>(src);
      if (mask) {
        const uint4 mv = *reinterpret_cast<const uint4*>(mask + r * cols + c);
        bf16* vp = reinterpret_cast<bf16*>(&v);
        const bf16* mp = reinterpret_cast<const bf16*>(&mv);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          vp[j] = __float2bfloat16(__bfloat162float(vp[j]) * act_grad_from_out(__bfloat162float(mp[j]), mask_kind, leak));
      }
      *reinterpret_cast<uint4*>(dst) = v;
    } else {
      float v = __bfloat162float(*src);
      if (mask) v *= act_grad_from_out(__bfloat162float(mask[r * cols + c]), mask_kind, leak);
      *dst = __float2bfloat16(v);
    }
  }
}
int slice_cols(const void* in, long long in_ld, int in_off, void* out, long long out_ld, int out_off, long long rows,
               int cols, const void* mask, int mask_kind, float leak, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return -1;
  const uintptr_t al = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(mask);
  const bool vec = ((cols | in_off | out_off) & 7) == 0 && ((in_ld | out_ld) & 7) == 0 && (al & 15) == 0;
  if (vec)
    slice_cols_kernel<8><<<stride_grid(rows * (cols / 8), 256, 2), 256, 0, st>>>(
        (const bf16*)in, in_ld, in_off, (bf16*)out, out_ld, out_off, rows, cols, (const bf16*)mask, mask_kind, leak);
  else
    slice_cols_kernel<1><<<stride_grid(rows * cols, 256, 4), 256, 0, st>>>(
        (const bf16*)in, in_ld, in_off, (bf16*)out, out_ld, out_off, rows, cols, (const bf16*)mask, mask_kind, leak);
  return 0;
}

// [T][A][B] -> [T][B][A], fp32 or bf16 in, bf16 out (weight re-layout for the K-major GEMM operand)
template <typename TI>
__global__ void transpose_kernel(const TI* __restrict__ in, bf16* out, int A, int B) {
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * A * B;
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int a = a0 + i, b = b0 + threadIdx.x;
    if (a < A && b < B) tile[i][threadIdx.x] = (float)in[base + (long long)a * B + b];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int b = b0 + i, a = a0 + threadIdx.x;
    if (a < A && b < B) out[base + (long long)b * A + a] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}
int transpose_to_bf16(const void* in, int in_f32, void* out, int T, int A, int B, cudaStream_t st) {
  dim3 grid((B + 31) / 32, (A + 31) / 32, T), block(32, 8);
  if (in_f32) transpose_kernel<float><<<grid, block, 0, st>>>((const float*)in, (bf16*)out, A, B);
  else transpose_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)in, (bf16*)out, A, B);
  return 0;
}

// =============================================================================================
// Gen-2 layer options (hem/ops/layers.py): dropout, instance norm; input layout; summaries
// =============================================================================================
// tf.nn.dropout(h, keep_prob) (hem/ops/layers.py:64,132,208): out = in * [u >= 1 - keep] / keep with u ~ U[0,1) drawn by
// the caller (b200_philox); the backward applies the same kernel to the gradient with the same u.
__global__ void dropout_kernel(const bf16* __restrict__ in, const float* __restrict__ u, bf16* out, long long n, float keep) {
  const float inv = 1.f / keep, thr = 1.f - keep;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16(u[i] >= thr ? __bfloat162float(in[i]) * inv : 0.f);
}
int dropout_apply(const void* in, const float* u, void* out, long long n, float keep, cudaStream_t st) {
  if (!(keep > 0.f) || keep > 1.f) return -1;
  dropout_kernel<<<stride_grid(n, 256, 4), 256, 0, st>>>((const bf16*)in, u, (bf16*)out, n, keep);
  return 0;
}

// hem.instance_norm (hem/ops/images.py:73-89): per sample and channel, moments over H x W, eps 1e-3, then
// scale[c] * xhat + shift[c].  NHWC here.  One block per (sample, 32-channel slab): 32 x 8 threads, rows strided by 8.
// stats[n][2C] keeps (mean, rstd) for the backward.
__global__ void instnorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                    bf16* out, float* stats, int HW, int C, float eps) {
  __shared__ float red[2][8][33];
  const int n = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const bf16* xs = x + (long long)n * HW * C;
  float s = 0.f, q = 0.f;
  if (c < C)
    for (int r = ty; r < HW; r += 8) { const float v = __bfloat162float(xs[(long long)r * C + c]); s += v; q = fmaf(v, v, q); }
  red[0][ty][tx] = s; red[1][ty][tx] = q;
  __syncthreads();
  float mean = 0.f, rstd = 0.f;
  if (c < C) {
    s = q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += red[0][j][tx]; q += red[1][j][tx]; }
    mean = s / HW;
    rstd = rsqrtf(fmaxf(q / HW - mean * mean, 0.f) + eps);
    if (ty == 0) { stats[(long long)n * 2 * C + c] = mean; stats[(long long)n * 2 * C + C + c] = rstd; }
    const float a = scale[c] * rstd, b = shift[c] - mean * a;
    bf16* os = out + (long long)n * HW * C;
    for (int r = ty; r < HW; r += 8)
      os[(long long)r * C + c] = __float2bfloat16(fmaf(__bfloat162float(xs[(long long)r * C + c]), a, b));
  }
}
// dx = scale*rstd * (g - mean(g) - xhat * mean(g*xhat));  dscale[c] += sum g*xhat;  dshift[c] += sum g
__global__ void instnorm_bwd_kernel(const bf16* __restrict__ g, const bf16* __restrict__ x, const float* __restrict__ stats,
                                    const float* __restrict__ scale, bf16* dx, float* dscale, float* dshift, int HW, int C) {
  __shared__ float red[2][8][33];
  const int n = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long base = (long long)n * HW * C;
  float mean = 0.f, rstd = 0.f, s1 = 0.f, s2 = 0.f;
  if (c < C) {
    mean = stats[(long long)n * 2 * C + c]; rstd = stats[(long long)n * 2 * C + C + c];
    for (int r = ty; r < HW; r += 8) {
      const float gv = __bfloat162float(g[base + (long long)r * C + c]);
      const float xh = (__bfloat162float(x[base + (long long)r * C + c]) - mean) * rstd;
      s1 += gv; s2 = fmaf(gv, xh, s2);
    }
  }
  red[0][ty][tx] = s1; red[1][ty][tx] = s2;
  __syncthreads();
  if (c < C) {
    s1 = s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1 += red[0][j][tx]; s2 += red[1][j][tx]; }
    if (ty == 0) { atomicAdd(dshift + c, s1); atomicAdd(dscale + c, s2); }
    const float a = scale[c] * rstd, m1 = s1 / HW, m2 = s2 / HW;
    for (int r = ty; r < HW; r += 8) {
      const float gv = __bfloat162float(g[base + (long long)r * C + c]);
      const float xh = (__bfloat162float(x[base + (long long)r * C + c]) - mean) * rstd;
      dx[base + (long long)r * C + c] = __float2bfloat16(a * (gv - m1 - xh * m2));
    }
  }
}
int instnorm_fwd(const void* x, const float* scale, const float* shift, void* out, float* stats, int N, int HW, int C,
                 float eps, cudaStream_t st) {
  if (N < 1 || HW < 1 || C < 1) return -1;
  instnorm_fwd_kernel<<<dim3((C + 31) / 32, N), 256, 0, st>>>((const bf16*)x, scale, shift, (bf16*)out, stats, HW, C, eps);
  return 0;
}
int instnorm_bwd(const void* g, const void* x, const float* stats, const float* scale, void* dx, float* dscale,
                 float* dshift, int N, int HW, int C, cudaStream_t st) {
  if (N < 1 || HW < 1 || C < 1) return -1;
  instnorm_bwd_kernel<<<dim3((C + 31) / 32, N), 256, 0, st>>>((const bf16*)g, (const bf16*)x, stats, scale, (bf16*)dx,
                                                              dscale, dshift, HW, C);
  return 0;
}

// Input stage for NCHW sources (the Gen-2 pipeline yields NCHW, hem/ops/layers.py:117-119): out NHWC bf16 =
// in[n][c][h][w] * mul + add, in: uint8 (caller folds 1/255 into mul) | fp32 | bf16.  And the way back for samples.
template <typename TI>
__global__ void nchw_to_nhwc_kernel(const TI* __restrict__ in, bf16* __restrict__ out, int C, int HW, float mul, float add) {
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * C * HW;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = fmaf(to_f32(in[base + (long long)c * HW + p]), mul, add);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) out[base + (long long)p * C + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}
template <typename TI>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ in, float* __restrict__ out, int C, int HW, float mul, float add) {
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * C * HW;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = fmaf(to_f32(in[base + (long long)p * C + c]), mul, add);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) out[base + (long long)c * HW + p] = tile[threadIdx.x][i];
  }
}
int layout_convert(const void* in, int in_type, void* out, int to_nchw, int N, int C, int HW, float mul, float add,
                   cudaStream_t st) {
  if (N < 1 || C < 1 || HW < 1 || N > 65535) return -1;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  if (!to_nchw) {
    if (in_type == 2) nchw_to_nhwc_kernel<uint8_t><<<grid, block, 0, st>>>((const uint8_t*)in, (bf16*)out, C, HW, mul, add);
    else if (in_type == 1) nchw_to_nhwc_kernel<float><<<grid, block, 0, st>>>((const float*)in, (bf16*)out, C, HW, mul, add);
    else if (in_type == 0) nchw_to_nhwc_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)in, (bf16*)out, C, HW, mul, add);
    else return -1;
  } else {
    if (in_type == 1) nhwc_to_nchw_kernel<float><<<grid, block, 0, st>>>((const float*)in, (float*)out, C, HW, mul, add);
    else if (in_type == 0) nhwc_to_nchw_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)in, (float*)out, C, HW, mul, add);
    else return -1;
  }
  return 0;
}

// Summaries after the step (ops/summaries.py:13-40): what tf.summary.histogram / tf.nn.zero_fraction need of a tensor in
// one pass -- out5 = {min, max, sum, sum of squares, number of zeros}; counts[nb] (optional) = TensorBoard's default
// exponential buckets: index 0 holds v <= -1e-12*1.1^(nb/2-1) ... middle = [-1e-12, 1e-12), growth 1.1 on both sides.
template <typename TI>
__global__ void summary_kernel(const TI* __restrict__ x, long long n, float* out5, unsigned int* counts, int nb) {
  float mn = 3.4e38f, mx = -3.4e38f, s = 0.f, q = 0.f, z = 0.f;
  const int half = nb / 2;
  const float inv_log = 1.f / logf(1.1f);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = to_f32(x[i]);
    mn = fminf(mn, v); mx = fmaxf(mx, v); s += v; q = fmaf(v, v, q); z += (v == 0.f) ? 1.f : 0.f;
    if (counts) {
      const float a = fabsf(v);
      int k = a < 1e-12f ? 0 : min(half - 1, 1 + (int)floorf(logf(a * 1e12f) * inv_log));
      atomicAdd(counts + (v < 0.f ? half - 1 - k : half + k), 1u);
    }
  }
  __shared__ float red[5][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); z += __shfl_xor_sync(0xffffffffu, z, o);
  }
  if (lane == 0) { red[0][w] = mn; red[1][w] = mx; red[2][w] = s; red[3][w] = q; red[4][w] = z; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int j = 1; j < nw; ++j) {
      mn = fminf(mn, red[0][j]); mx = fmaxf(mx, red[1][j]); s += red[2][j]; q += red[3][j]; z += red[4][j];
    }
    // float atomic min / max through the ordered-int trick (values are finite)
    atomicMin(reinterpret_cast<int*>(out5), mn >= 0.f ? __float_as_int(mn) : (int)(0x80000000u - (unsigned)__float_as_int(mn)));
    atomicMax(reinterpret_cast<int*>(out5) + 1, mx >= 0.f ? __float_as_int(mx) : (int)(0x80000000u - (unsigned)__float_as_int(mx)));
    atomicAdd(out5 + 2, s); atomicAdd(out5 + 3, q); atomicAdd(out5 + 4, z);
  }
}
__global__ void summary_init_kernel(float* out5, unsigned int* counts, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    reinterpret_cast<int*>(out5)[0] = 0x7f7fffff;          // ordered-int encoding of +FLT_MAX
    reinterpret_cast<int*>(out5)[1] = (int)(0x80000000u - 0xff7fffffu);   // of -FLT_MAX
    out5[2] = out5[3] = out5[4] = 0.f;
  }
  if (counts && i < nb) counts[i] = 0u;
}
__global__ void summary_fix_kernel(float* out5) {          // decode the ordered ints back to floats
  for (int k = 0; k < 2; ++k) {
    const int e = reinterpret_cast<int*>(out5)[k];
    out5[k] = e >= 0 ? __int_as_float(e) : __int_as_float((int)(0x80000000u - (unsigned)e));
  }
}
int summary_stats(const void* x, int x_type, long long n, float* out5, unsigned int* counts, int nb, cudaStream_t st) {
  if (n < 1 || (counts && (nb < 4 || (nb & 1)))) return -1;
  summary_init_kernel<<<(std_max(nb, 1) + 255) / 256, 256, 0, st>>>(out5, counts, counts ? nb : 0);
  int grid = stride_grid(n, 256, 8);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  if (x_type == 1) summary_kernel<float><<<grid, 256, 0, st>>>((const float*)x, n, out5, counts, nb);
  else if (x_type == 0) summary_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, n, out5, counts, nb);
  else return -1;
  summary_fix_kernel<<<1, 1, 0, st>>>(out5);
  return 0;
}

// montage_summary (ops/summaries.py:95-124): images [B, H, W, C] -> one [m*H, n*W, C] grid image; image j goes to
// row block j % m, column block j / m (tf.split along the batch into n groups, concatenated side by side).
template <typename TI>
__global__ void montage_kernel(const TI* __restrict__ x, float* __restrict__ out, int m, int n, int H, int W, int C,
                               float mul, float add) {
  const long long total = (long long)m * n * H * W * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int gx = (int)(t % ((long long)n * W)); t /= (long long)n * W;
    const int gy = (int)t;
    const int j = (gx / W) * m + gy / H;
    out[i] = fmaf(to_f32(x[(((long long)j * H + gy % H) * W + gx % W) * C + c]), mul, add);
  }
}
int montage(const void* x, int x_type, float* out, int m, int n, int H, int W, int C, float mul, float add, cudaStream_t st) {
  if (m < 1 || n < 1 || H < 1 || W < 1 || C < 1) return -1;
  const int grid = stride_grid((long long)m * n * H * W * C, 256, 4);
  if (x_type == 1) montage_kernel<float><<<grid, 256, 0, st>>>((const float*)x, out, m, n, H, W, C, mul, add);
  else if (x_type == 0) montage_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, out, m, n, H, W, C, mul, add);
  else return -1;
  return 0;
}

// =============================================================================================
// Column reductions over a row-major [R, C] bf16 matrix (C contiguous)
// =============================================================================================
// out[c] += alpha * sum_r x[r,c] * (wrow ? wrow[r] : 1);   out2[c] += alpha * sum_r x[r,c]^2 (optional)
// Vector path (C % 8 == 0): block = 32 column-octets x 8 row lanes, one 16-byte load per thread and row, 4 rows
// in flight per thread; row lanes are combined through shared memory, one atomic per column.
template <bool kSquares>
__global__ void colsum_vec_kernel(const bf16* __restrict__ x, const float* __restrict__ wrow, float* out, float* out2,
                                  long long R, int C, float alpha, int rows_per_block) {
  __shared__ float red[kSquares ? 2 : 1][8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 8;
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (c < C) {
#pragma unroll 4
    for (long long r = r0 + ty; r < r1; r += 8) {
      const float wr = wrow ? wrow[r] : 1.f;
      const uint4 raw = *reinterpret_cast<const uint4*>(x + r * C + c);
      const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = __uint_as_float(w4[j] << 16), hi = __uint_as_float(w4[j] & 0xffff0000u);
        s[2 * j] = fmaf(lo, wr, s[2 * j]); s[2 * j + 1] = fmaf(hi, wr, s[2 * j + 1]);
        if (kSquares) { q[2 * j] = fmaf(lo, lo, q[2 * j]); q[2 * j + 1] = fmaf(hi, hi, q[2 * j + 1]); }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][ty][tx * 8 + j] = s[j];
    if (kSquares) red[kSquares ? 1 : 0][ty][tx * 8 + j] = q[j];
  }
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a += red[0][j][threadIdx.x];
      if (kSquares) b += red[kSquares ? 1 : 0][j][threadIdx.x];
    }
    atomicAdd(out + cc, a * alpha);
    if (kSquares) atomicAdd(out2 + cc, b * alpha);
  }
}
// block = 32 column-pairs x 8 row lanes (any C)
__global__ void colsum_kernel(const bf16* __restrict__ x, const float* __restrict__ wrow, float* out, float* out2,
                              long long R, int C, float alpha, int rows_per_block) {
  __shared__ float red[4][8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 2;
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const bool even = (C & 1) == 0;
  if (c < C) {
#pragma unroll 4
    for (long long r = r0 + ty; r < r1; r += 8) {
      const float wr = wrow ? wrow[r] : 1.f;
      float v0, v1 = 0.f;
      if (even) {
        __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(x + r * C + c);
        v0 = __low2float(h); v1 = __high2float(h);
      } else {
        v0 = __bfloat162float(x[r * C + c]);
        if (c + 1 < C) v1 = __bfloat162float(x[r * C + c + 1]);
      }
      s0 = fmaf(v0, wr, s0); s1 = fmaf(v1, wr, s1);
      q0 = fmaf(v0, v0, q0); q1 = fmaf(v1, v1, q1);
    }
  }
  red[0][ty][2 * tx] = s0; red[0][ty][2 * tx + 1] = s1;
  red[1][ty][2 * tx] = q0; red[1][ty][2 * tx + 1] = q1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < C) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { a += red[0][j][threadIdx.x]; b += red[1][j][threadIdx.x]; }
      atomicAdd(out + cc, a * alpha);
      if (out2) atomicAdd(out2 + cc, b * alpha);
    }
  }
}
static void colsum_launch(const bf16* x, const float* wrow, float* out, float* out2, long long R, int C, float alpha,
                          cudaStream_t st) {
  const bool vec = (C & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int gx = vec ? (C + 255) / 256 : (C + 63) / 64;
  static int per_sm = -1;       // row blocks per SM: fewer blocks = fewer same-address atomics at the end (B200GAN_COLSUM_BLOCKS)
  if (per_sm < 0) { const char* e = getenv("B200GAN_COLSUM_BLOCKS"); per_sm = e ? atoi(e) : 4; if (per_sm < 1) per_sm = 1; }
  long long want = ((long long)num_sms() * per_sm + gx - 1) / gx;
  if (want > (R + 31) / 32) want = (R + 31) / 32;
  if (want < 1) want = 1;
  const int rpb = (int)((R + want - 1) / want);
  const int gy = (int)((R + rpb - 1) / rpb);
  if (vec && out2) colsum_vec_kernel<true><<<dim3(gx, gy), 256, 0, st>>>(x, wrow, out, out2, R, C, alpha, rpb);
  else if (vec) colsum_vec_kernel<false><<<dim3(gx, gy), 256, 0, st>>>(x, wrow, out, out2, R, C, alpha, rpb);
  else colsum_kernel<<<dim3(gx, gy), 256, 0, st>>>(x, wrow, out, out2, R, C, alpha, rpb);
}
int colsum(const void* x, const float* wrow, float* out, long long R, int C, float alpha, cudaStream_t st) {
  colsum_launch((const bf16*)x, wrow, out, nullptr, R, C, alpha, st);
  return 0;
}

// ---- batch norm (tf.contrib.layers.batch_norm defaults: no gamma, eps 1e-3, batch statistics)
// stats[0:C] = sum, stats[C:2C] = sum of squares (zeroed by the caller)
int bn_sums(const void* z, float* stats, long long R, int C, cudaStream_t st) {
  colsum_launch((const bf16*)z, nullptr, stats, stats + C, R, C, 1.f, st);
  return 0;
}
// a = act((z-mean)*rstd + beta).  A block owns a slab of up to 1024 channels x a range of rows: the slab's
// rstd and shift are computed once into shared memory, then the rows are streamed with 16-byte accesses.
constexpr int kBnSlab = 1024;
__global__ void bn_apply_vec_kernel(const bf16* __restrict__ z, const float* __restrict__ stats,
                                    const float* __restrict__ beta, bf16* out, long long R, int C, float eps, int act,
                                    float leak, int rows_per_block) {
  __shared__ float sc[kBnSlab], sh[kBnSlab];
  const int cs0 = blockIdx.x * kBnSlab;
  const int cw = min(kBnSlab, C - cs0);
  const float invR = 1.f / (float)R;
  for (int c = threadIdx.x; c < cw; c += blockDim.x) {
    const float mean = stats[cs0 + c] * invR;
    const float rstd = rsqrtf(fmaxf(stats[C + cs0 + c] * invR - mean * mean, 0.f) + eps);
    sc[c] = rstd;
    sh[c] = beta[cs0 + c] - mean * rstd;
  }
  __syncthreads();
  const int cpr = cw >> 3;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  const int total = (int)(r1 - r0) * cpr;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int rr = t / cpr, ch = t - rr * cpr;
    const long long e0 = (r0 + rr) * C + cs0 + ch * 8;
    const uint4 zv = *reinterpret_cast<const uint4*>(z + e0);
    const uint32_t w4[4] = {zv.x, zv.y, zv.z, zv.w};
    uint32_t o4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = ch * 8 + 2 * j;
      const float lo = __uint_as_float(w4[j] << 16), hi = __uint_as_float(w4[j] & 0xffff0000u);
      const float a = act_fwd(fmaf(lo, sc[c], sh[c]), act, leak);
      const float b = act_fwd(fmaf(hi, sc[c + 1], sh[c + 1]), act, leak);
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      o4[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(out + e0) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
  }
}
__global__ void bn_apply_kernel(const bf16* __restrict__ z, const float* __restrict__ stats, const float* __restrict__ beta,
                                bf16* out, long long R, int C, float eps, int act, float leak) {
  const long long n = R * C;
  const float invR = 1.f / (float)R;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % C);
    const float mean = stats[c] * invR;
    const float var = fmaxf(stats[C + c] * invR - mean * mean, 0.f);
    const float v = (__bfloat162float(z[i]) - mean) * rsqrtf(var + eps) + beta[c];
    out[i] = __float2bfloat16(act_fwd(v, act, leak));
  }
}
int bn_apply(const void* z, const float* stats, const float* beta, void* out, long long R, int C, float eps, int act,
             float leak, cudaStream_t st) {
  if ((C & 7) == 0 && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int gx = (C + kBnSlab - 1) / kBnSlab;
    long long want = ((long long)num_sms() * 8 + gx - 1) / gx;
    if (want > R) want = R;
    const int rpb = (int)((R + want - 1) / want);
    const int gy = (int)((R + rpb - 1) / rpb);
    bn_apply_vec_kernel<<<dim3(gx, gy), 256, 0, st>>>((const bf16*)z, stats, beta, (bf16*)out, R, C, eps, act, leak, rpb);
  }
  else
    bn_apply_kernel<<<stride_grid(R * C, 256, 4), 256, 0, st>>>((const bf16*)z, stats, beta, (bf16*)out, R, C, eps, act, leak);
  return 0;
}
// moving_mean / moving_variance of tf.contrib.layers.batch_norm (decay 0.999, A.3): mv -= (mv - batch) * (1 - decay)
// from the sums bn_sums left in stats; `unbiased` feeds n/(n-1) * variance (the fused NCHW kernel of hem/ops/layers.py:124)
__global__ void bn_update_moving_kernel(const float* __restrict__ stats, float invR, float unbias, float* mm, float* mv,
                                        int C, float one_minus_decay) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = stats[c] * invR;
  const float var = fmaxf(stats[C + c] * invR - mean * mean, 0.f) * unbias;
  mm[c] -= (mm[c] - mean) * one_minus_decay;
  mv[c] -= (mv[c] - var) * one_minus_decay;
}
int bn_update_moving(const float* stats, long long R, int C, float* mm, float* mv, float decay, int unbiased, cudaStream_t st) {
  if (R < 1 || C < 1) return -1;
  const float unbias = (unbiased && R > 1) ? (float)R / (float)(R - 1) : 1.f;
  bn_update_moving_kernel<<<(C + 127) / 128, 128, 0, st>>>(stats, 1.f / (float)R, unbias, mm, mv, C, 1.f - decay);
  return 0;
}
// backward sums: bsum[0:C] += sum_r g ; bsum[C:2C] += sum_r g * xhat      (g already holds act')
__global__ void bn_bwd_sums_kernel(const bf16* __restrict__ g, const bf16* __restrict__ z, const float* __restrict__ stats,
                                   float* bsum, long long R, int C, float eps, int rows_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invR = 1.f / (float)R;
  const float mean = stats[c] * invR;
  const float rstd = rsqrtf(fmaxf(stats[C + c] * invR - mean * mean, 0.f) + eps);
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float s1 = 0.f, s2 = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float gv = __bfloat162float(g[r * C + c]);
    const float xh = (__bfloat162float(z[r * C + c]) - mean) * rstd;
    s1 += gv;
    s2 = fmaf(gv, xh, s2);
  }
  atomicAdd(bsum + c, s1);
  atomicAdd(bsum + C + c, s2);
}
// vector form (C % 8 == 0, 16-byte aligned): block = 32 column-octets x 8 row lanes as in colsum_vec_kernel, one
// 16-byte load of g and of z per thread and row; row lanes combined through shared memory, one atomic per column
__global__ void bn_bwd_sums_vec_kernel(const bf16* __restrict__ g, const bf16* __restrict__ z,
                                       const float* __restrict__ stats, float* bsum, long long R, int C, float eps,
                                       int rows_per_block) {
  __shared__ float red[2][8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 8;
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float s[8], q[8], mean[8], rstd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; mean[j] = 0.f; rstd[j] = 0.f; }
  if (c < C) {
    const float invR = 1.f / (float)R;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = stats[c + j] * invR;
      rstd[j] = rsqrtf(fmaxf(stats[C + c + j] * invR - mean[j] * mean[j], 0.f) + eps);
    }
#pragma unroll 2
    for (long long r = r0 + ty; r < r1; r += 8) {
      const uint4 gr = *reinterpret_cast<const uint4*>(g + r * C + c);
      const uint4 zr = *reinterpret_cast<const uint4*>(z + r * C + c);
      const uint32_t g4[4] = {gr.x, gr.y, gr.z, gr.w}, z4[4] = {zr.x, zr.y, zr.z, zr.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g0 = __uint_as_float(g4[j] << 16), g1 = __uint_as_float(g4[j] & 0xffff0000u);
        const float z0 = __uint_as_float(z4[j] << 16), z1 = __uint_as_float(z4[j] & 0xffff0000u);
        s[2 * j] += g0; s[2 * j + 1] += g1;
        q[2 * j] = fmaf(g0, z0 - mean[2 * j], q[2 * j]);
        q[2 * j + 1] = fmaf(g1, z1 - mean[2 * j + 1], q[2 * j + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][ty][tx * 8 + j] = s[j];
    red[1][ty][tx * 8 + j] = q[j] * rstd[j];           // sum g * (z - mean) * rstd
  }
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { a += red[0][j][threadIdx.x]; b += red[1][j][threadIdx.x]; }
    atomicAdd(bsum + cc, a);
    atomicAdd(bsum + C + cc, b);
  }
}
// dz = rstd * (g - s1/R - xhat*s2/R)
__global__ void bn_bwd_apply_kernel(const bf16* __restrict__ g, const bf16* __restrict__ z, const float* __restrict__ stats,
                                    const float* __restrict__ bsum, bf16* dz, long long R, int C, float eps) {
  const long long n = R * C;
  const float invR = 1.f / (float)R;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = (int)(i % C);
    const float mean = stats[c] * invR;
    const float rstd = rsqrtf(fmaxf(stats[C + c] * invR - mean * mean, 0.f) + eps);
    const float xh = (__bfloat162float(z[i]) - mean) * rstd;
    const float v = rstd * (__bfloat162float(g[i]) - bsum[c] * invR - xh * bsum[C + c] * invR);
    dz[i] = __float2bfloat16(v);
  }
}
// slab version (C % 8 == 0): per-channel coefficients once per block in shared memory, 16-byte accesses
//   dz = a*g + b*z + c   with  a = rstd,  b = -rstd^2*s2/R,  c = -rstd*s1/R + rstd^2*s2/R*mean
__global__ void bn_bwd_apply_vec_kernel(const bf16* __restrict__ g, const bf16* __restrict__ z, const float* __restrict__ stats,
                                        const float* __restrict__ bsum, bf16* dz, long long R, int C, float eps,
                                        int rows_per_block) {
  __shared__ float ca[kBnSlab], cb[kBnSlab], cc[kBnSlab];
  const int cs0 = blockIdx.x * kBnSlab;
  const int cw = min(kBnSlab, C - cs0);
  const float invR = 1.f / (float)R;
  for (int c = threadIdx.x; c < cw; c += blockDim.x) {
    const float mean = stats[cs0 + c] * invR;
    const float rstd = rsqrtf(fmaxf(stats[C + cs0 + c] * invR - mean * mean, 0.f) + eps);
    const float s1 = bsum[cs0 + c] * invR, s2 = bsum[C + cs0 + c] * invR;
    ca[c] = rstd;
    cb[c] = -rstd * rstd * s2;
    cc[c] = -rstd * s1 + rstd * rstd * s2 * mean;
  }
  __syncthreads();
  const int cpr = cw >> 3;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(R, r0 + rows_per_block);
  const int total = (int)(r1 - r0) * cpr;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int rr = t / cpr, ch = t - rr * cpr;
    const long long e0 = (r0 + rr) * C + cs0 + ch * 8;
    const uint4 gv = *reinterpret_cast<const uint4*>(g + e0);
    const uint4 zv = *reinterpret_cast<const uint4*>(z + e0);
    const uint32_t g4[4] = {gv.x, gv.y, gv.z, gv.w}, z4[4] = {zv.x, zv.y, zv.z, zv.w};
    uint32_t o4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = ch * 8 + 2 * j;
      const float glo = __uint_as_float(g4[j] << 16), ghi = __uint_as_float(g4[j] & 0xffff0000u);
      const float zlo = __uint_as_float(z4[j] << 16), zhi = __uint_as_float(z4[j] & 0xffff0000u);
      const float a = fmaf(ca[c], glo, fmaf(cb[c], zlo, cc[c]));
      const float b = fmaf(ca[c + 1], ghi, fmaf(cb[c + 1], zhi, cc[c + 1]));
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      o4[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dz + e0) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
  }
}
int bn_bwd(const void* g, const void* z, const float* stats, float* bsum, void* dz, long long R, int C, float eps,
           cudaStream_t st) {
  const int threads = 128;
  const int gx = (C + threads - 1) / threads;
  long long want = ((long long)num_sms() * 8 + gx - 1) / gx;
  if (want > R) want = R;
  if (want < 1) want = 1;
  const int rpb = (int)((R + want - 1) / want);
  const int gy = (int)((R + rpb - 1) / rpb);
  const uintptr_t al = reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(dz);
  if ((C & 7) == 0 && (al & 15) == 0) {
    const int vx = (C + 255) / 256;
    long long wv = ((long long)num_sms() * 4 + vx - 1) / vx;
    if (wv > (R + 31) / 32) wv = (R + 31) / 32;
    if (wv < 1) wv = 1;
    const int rpbv = (int)((R + wv - 1) / wv);
    const int vy = (int)((R + rpbv - 1) / rpbv);
    bn_bwd_sums_vec_kernel<<<dim3(vx, vy), 256, 0, st>>>((const bf16*)g, (const bf16*)z, stats, bsum, R, C, eps, rpbv);
  } else {
    bn_bwd_sums_kernel<<<dim3(gx, gy), threads, 0, st>>>((const bf16*)g, (const bf16*)z, stats, bsum, R, C, eps, rpb);
  }
  if ((C & 7) == 0 && (al & 15) == 0) {
    const int sx = (C + kBnSlab - 1) / kBnSlab;
    long long w2 = ((long long)num_sms() * 8 + sx - 1) / sx;
    if (w2 > R) w2 = R;
    const int rpb2 = (int)((R + w2 - 1) / w2);
    const int sy = (int)((R + rpb2 - 1) / rpb2);
    bn_bwd_apply_vec_kernel<<<dim3(sx, sy), 256, 0, st>>>((const bf16*)g, (const bf16*)z, stats, bsum, (bf16*)dz, R, C, eps, rpb2);
  } else {
    bn_bwd_apply_kernel<<<stride_grid(R * C, 256, 4), 256, 0, st>>>((const bf16*)g, (const bf16*)z, stats, bsum, (bf16*)dz, R, C, eps);
  }
  return 0;
}

// =============================================================================================
// GEMV family for 1-unit dense layers (critic fc2, models/gan.py:285)
// =============================================================================================
// out[m] = act( sum_k a[m,k]*w[k] + bias[0] )    one block (4 warps) per row: enough loads in flight for HBM
__global__ void gemv_rows_kernel(const bf16* __restrict__ a, const bf16* __restrict__ w, const float* bias, float* out,
                                 int M, int K, int act, float leak) {
  __shared__ float part[4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x;
  const bf16* ar = a + (long long)row * K;
  float s = 0.f;
  if ((K & 7) == 0) {
#pragma unroll 4
    for (int k = threadIdx.x * 8; k < K; k += 128 * 8) {
      uint4 av = *reinterpret_cast<const uint4*>(ar + k);
      uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + k));
      const bf16* ap = reinterpret_cast<const bf16*>(&av);
      const bf16* wp = reinterpret_cast<const bf16*>(&wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(__bfloat162float(ap[j]), __bfloat162float(wp[j]), s);
    }
  } else {
    for (int k = threadIdx.x; k < K; k += 128) s = fmaf(__bfloat162float(ar[k]), __bfloat162float(w[k]), s);
  }
  s = warp_sum(s);
  if (lane == 0) part[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) out[row] = act_fwd(part[0] + part[1] + part[2] + part[3] + (bias ? bias[0] : 0.f), act, leak);
}
int gemv_rows(const void* a, const void* w, const float* bias, float* out, int M, int K, int act, float leak,
              cudaStream_t st) {
  gemv_rows_kernel<<<M, 128, 0, st>>>((const bf16*)a, (const bf16*)w, bias, out, M, K, act, leak);
  return 0;
}
// out[m,k] = g[m] * w[k] * act'(mask[m,k])
__global__ void outer_mask_vec_kernel(const float* __restrict__ g, const bf16* __restrict__ w,
                                      const bf16* __restrict__ mask, bf16* out, int M, int K, int kind, float leak) {
  const long long n8 = (long long)M * K / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const long long e0 = i * 8;
    const float gv = g[e0 / K];
    uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + (e0 % K)));
    const bf16* wp = reinterpret_cast<const bf16*>(&wv);
    uint4 mv = make_uint4(0, 0, 0, 0);
    if (mask) mv = *reinterpret_cast<const uint4*>(mask + e0);
    const bf16* mp = reinterpret_cast<const bf16*>(&mv);
    uint4 ov;
    bf16* op = reinterpret_cast<bf16*>(&ov);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = gv * __bfloat162float(wp[e]);
      if (mask) v *= act_grad_from_out(__bfloat162float(mp[e]), kind, leak);
      op[e] = __float2bfloat16(v);
    }
    *reinterpret_cast<uint4*>(out + e0) = ov;
  }
}
__global__ void outer_mask_kernel(const float* __restrict__ g, const bf16* __restrict__ w, const bf16* __restrict__ mask,
                                  bf16* out, int M, int K, int kind, float leak) {
  const long long n = (long long)M * K;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = g[i / K] * __bfloat162float(w[i % K]);
    if (mask) v *= act_grad_from_out(__bfloat162float(mask[i]), kind, leak);
    out[i] = __float2bfloat16(v);
  }
}
int outer_mask(const float* g, const void* w, const void* mask, void* out, int M, int K, int kind, float leak,
               cudaStream_t st) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if ((K & 7) == 0 && aligned)
    outer_mask_vec_kernel<<<stride_grid((long long)M * K / 8, 256, 2), 256, 0, st>>>(g, (const bf16*)w, (const bf16*)mask, (bf16*)out, M, K, kind, leak);
  else
    outer_mask_kernel<<<stride_grid((long long)M * K, 256, 4), 256, 0, st>>>(g, (const bf16*)w, (const bf16*)mask, (bf16*)out, M, K, kind, leak);
  return 0;
}

// =============================================================================================
// Reductions to a device scalar
// =============================================================================================
// out[0] += alpha * sum x^2  (sq=1) or alpha * sum x (sq=0)
template <typename TI>
__global__ void reduce_kernel(const TI* __restrict__ x, long long n, float* out, float alpha, int sq) {
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = (float)x[i];
    s += sq ? v * v : v;
  }
  __shared__ float ws[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s * alpha);
  }
}
int reduce_sum(const void* x, int x_f32, long long n, float* out, float alpha, int sq, cudaStream_t st) {
  int grid = stride_grid(n, 256, 8);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  if (x_f32) reduce_kernel<float><<<grid, 256, 0, st>>>((const float*)x, n, out, alpha, sq);
  else reduce_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, n, out, alpha, sq);
  return 0;
}

// WGAN / IWGAN scalar losses (models/gan.py:194-205,229-230).  sums = {sum d_real, sum d_fake, sumsq grad}
// out = {g_loss, d_loss, slopes, dloss/d(sumsq)}
__global__ void wgan_loss_kernel(const float* sums, float inv_b, int use_gp, float lambda, float* out) {
  const float mr = sums[0] * inv_b, mf = sums[1] * inv_b;
  float d = mf - mr, slopes = 0.f, dss = 0.f;
  if (use_gp) {
    slopes = sqrtf(sums[2]);
    d += lambda * (slopes - 1.f) * (slopes - 1.f);
    dss = slopes > 0.f ? lambda * (slopes - 1.f) / slopes : 0.f;   // d/d(sumsq) of lambda*(sqrt(ss)-1)^2
  }
  out[0] = -mf; out[1] = d; out[2] = slopes; out[3] = dss;
}
int wgan_loss(const float* sums, int B, int use_gp, float lambda, float* out, cudaStream_t st) {
  wgan_loss_kernel<<<1, 1, 0, st>>>(sums, 1.f / (float)B, use_gp, lambda, out);
  return 0;
}

// Elementwise losses with fused gradient: kind 0: GAN d-loss/g-loss pieces etc. are assembled on the host
// side of the tape from these primitives:
//   kind 0  L1:            l = |a-b|                          dl/da = sign(a-b)
//   kind 1  -log(a+eps)                                       dl/da = -1/(a+eps)
//   kind 2  -log(1-a+eps)                                     dl/da =  1/(1-a+eps)
//   kind 3  Bernoulli recon: -(b*log(eps+a)+(1-b)*log(eps+1-a))   dl/da = -(b/(eps+a) - (1-b)/(eps+1-a))
//   kind 4  sigmoid-CE(logits a, label scalar `lab`)          dl/da = sigmoid(a)-lab
//   kind 5  squared error (a-b)^2                             dl/da = 2(a-b)
//   kind 6  0.5 a^2 ; kind 7  0.5 (a^2 - log(eps+a^2) - 1)    VAE latent loss pieces (models/vae.py:80-81)
// out_sum[0] += scale * sum l ;  grad[i] = gscale * dl/da * act'(a)  (bf16 or fp32), optional; act' is the
// derivative of the activation `mask_kind` that PRODUCED a, evaluated at a (0 = none): the gradient w.r.t. the
// pre-activation in one fp32 expression (Bernoulli loss behind a sigmoid: (1-b)/(1-a) * a(1-a) must not be
// rounded in between)
template <typename TA, typename TG>
__global__ void eltloss_kernel(const TA* __restrict__ a, const bf16* __restrict__ b, long long n, int kind, float lab,
                               float scale, float gscale, float* out_sum, TG* grad, int mask_kind, float leak) {
  float s = 0.f;
  const float eps = 1e-8f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float av = (float)a[i];
    const float bv = b ? __bfloat162float(b[i]) : 0.f;
    float l, d;
    switch (kind) {
      case 0: l = fabsf(av - bv); d = av > bv ? 1.f : (av < bv ? -1.f : 0.f); break;
      case 1: l = -logf(av + eps); d = -1.f / (av + eps); break;
      case 2: l = -logf(1.f - av + eps); d = 1.f / (1.f - av + eps); break;
      case 3: l = -(bv * logf(eps + av) + (1.f - bv) * logf(eps + 1.f - av));
              d = -(bv / (eps + av) - (1.f - bv) / (eps + 1.f - av)); break;
      case 4: l = fmaxf(av, 0.f) - av * lab + log1pf(__expf(-fabsf(av))); d = 1.f / (1.f + __expf(-av)) - lab; break;
      case 6: l = 0.5f * av * av; d = av; break;                                   // KL piece of the mean
      case 7: l = 0.5f * (av * av - logf(eps + av * av) - 1.f); d = av - av / (eps + av * av); break;  // KL piece of sigma
      default: l = (av - bv) * (av - bv); d = 2.f * (av - bv); break;
    }
    s += l;
    if (grad) grad[i] = (TG)(d * gscale * act_grad_from_out(av, mask_kind, leak));
  }
  __shared__ float ws[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0 && out_sum) atomicAdd(out_sum, s * scale);
  }
}
int eltloss(const void* a, int a_f32, const void* b, long long n, int kind, float lab, float scale, float gscale,
            float* out_sum, void* grad, int grad_f32, int mask_kind, float leak, cudaStream_t st) {
  int grid = stride_grid(n, 256, 8);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  if (a_f32 && grad_f32) eltloss_kernel<float, float><<<grid, 256, 0, st>>>((const float*)a, (const bf16*)b, n, kind, lab, scale, gscale, out_sum, (float*)grad, mask_kind, leak);
  else if (a_f32) eltloss_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)a, (const bf16*)b, n, kind, lab, scale, gscale, out_sum, (bf16*)grad, mask_kind, leak);
  else if (grad_f32) eltloss_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)a, (const bf16*)b, n, kind, lab, scale, gscale, out_sum, (float*)grad, mask_kind, leak);
  else eltloss_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)a, (const bf16*)b, n, kind, lab, scale, gscale, out_sum, (bf16*)grad, mask_kind, leak);
  return 0;
}

// =============================================================================================
// Philox4x32-10 noise (stands in for tf.random_normal / tf.random_uniform, models/gan.py:224,246)
// counter = (element/4, draw counter read from device memory), key = (seed lo, seed hi)
// =============================================================================================
__device__ __forceinline__ void philox_round(uint32_t* c, uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t* c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// normal=1: N(0,1) via Box-Muller; normal=0: U[0,1).  `draw` is advanced by the kernel (one block does it last
// is unnecessary: a separate 1-thread bump kernel follows) so CUDA-graph replays see fresh noise.
template <typename TO>
__global__ void philox_kernel(TO* out, long long n, unsigned long long seed, const unsigned long long* draw,
                              unsigned int stream_id, int normal) {
  const unsigned long long d = draw ? *draw : 0ull;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += stride) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)d, (uint32_t)(d >> 32) ^ stream_id};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float v[4];
    if (normal) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float u1 = ((float)c[2 * h] + 1.f) * 2.3283064365386963e-10f;   // (0,1]
        const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;
        const float rad = sqrtf(-2.f * __logf(u1));
        float sn, cs;
        __sincosf(6.283185307179586f * u2, &sn, &cs);
        v[2 * h] = rad * cs; v[2 * h + 1] = rad * sn;
      }
    } else {
#pragma unroll
      for (int h = 0; h < 4; ++h) v[h] = (float)(c[h] >> 8) * 5.9604644775390625e-08f;   // [0,1)
    }
    for (int h = 0; h < 4 && q * 4 + h < n; ++h) out[q * 4 + h] = (TO)v[h];
  }
}
__global__ void bump_kernel(unsigned long long* draw) { *draw += 1ull; }
int philox_fill(void* out, int out_f32, long long n, unsigned long long seed, unsigned long long* draw,
                unsigned int stream_id, int normal, cudaStream_t st) {
  const int grid = stride_grid((n + 3) / 4, 256, 1);
  if (out_f32) philox_kernel<float><<<grid, 256, 0, st>>>((float*)out, n, seed, draw, stream_id, normal);
  else philox_kernel<bf16><<<grid, 256, 0, st>>>((bf16*)out, n, seed, draw, stream_id, normal);
  if (draw) bump_kernel<<<1, 1, 0, st>>>(draw);
  return 0;
}

// =============================================================================================
// Fused optimizer update over a flat fp32 parameter bucket (util.py:150-183 -> tf.train.*Optimizer)
// One pass: read g,p,m,v (16-byte accesses), write p,m,v, the bf16 compute copy, and zeros into g (the next run's
// gradient accumulation starts from a clean bucket without a separate fill): 32 B/param + 2 B for the bf16 copy.
// =============================================================================================
// kind 0 Adam (TF epsilon-hat form, A.4), 1 RMSProp (ms init 1.0, A.4), 2 SGD, 3 Momentum,
//      4 Adagrad == ProximalAdagrad with l1 = l2 = 0 (accumulator init 0.1), 5 Adadelta (rho 0.95, eps 1e-8),
//      6 FTRL (lr_power -0.5, accumulator init 0.1, l1 = l2 = 0), 7 centered RMSProp (third slot = mean gradient)
struct OptimScalars { float lr, lr_t, b1, b2, eps, gscale, clip; int kind; };

__device__ __forceinline__ void optim_one(const OptimScalars& o, float gi, float& pi, float& mi, float& vi, float& si) {
  gi *= o.gscale;
  if (o.clip > 0.f) pi = fminf(fmaxf(pi, -o.clip), o.clip);     // WGAN: clip BEFORE the update (models/gan.py:142-148)
  switch (o.kind) {
    case 0: {
      mi = o.b1 * mi + (1.f - o.b1) * gi;
      vi = o.b2 * vi + (1.f - o.b2) * gi * gi;
      pi -= o.lr_t * mi / (sqrtf(vi) + o.eps);
    } break;
    case 1: {                                                   // b1 = decay, b2 = momentum
      vi = o.b1 * vi + (1.f - o.b1) * gi * gi;
      mi = o.b2 * mi + o.lr * gi * rsqrtf(vi + o.eps);
      pi -= mi;
    } break;
    case 2: pi -= o.lr * gi; break;
    case 3: {                                                   // b1 = momentum
      mi = o.b1 * mi + gi;
      pi -= o.lr * mi;
    } break;
    case 4: {                                                   // v = accumulator (init 0.1)
      vi += gi * gi;
      pi -= o.lr * gi * rsqrtf(vi);
    } break;
    case 5: {                                                   // v = accum, m = accum_update, b1 = rho
      vi = o.b1 * vi + (1.f - o.b1) * gi * gi;
      const float upd = sqrtf(mi + o.eps) * rsqrtf(vi + o.eps) * gi;
      mi = o.b1 * mi + (1.f - o.b1) * upd * upd;
      pi -= o.lr * upd;
    } break;
    case 6: {                                                   // v = accum (init 0.1), m = linear
      const float na = vi + gi * gi;
      mi += gi - (sqrtf(na) - sqrtf(vi)) / o.lr * pi;
      pi = -mi / (sqrtf(na) / o.lr);
      vi = na;
    } break;
    default: {                                                  // 7: b1 = decay, b2 = momentum, s = mean gradient
      si = o.b1 * si + (1.f - o.b1) * gi;
      vi = o.b1 * vi + (1.f - o.b1) * gi * gi;
      mi = o.b2 * mi + o.lr * gi * rsqrtf(vi - si * si + o.eps);
      pi -= mi;
    } break;
  }
}

__global__ void optim_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float* __restrict__ s3,
                             float* __restrict__ g, bf16* __restrict__ p16, long long n, OptimScalars o, int zero_grad,
                             const int* __restrict__ step) {
  if (o.kind == 0) {
    const float t = (float)(*step + 1);
    o.lr_t = o.lr * sqrtf(1.f - powf(o.b2, t)) / (1.f - powf(o.b1, t));
  }
  const bool slots = o.kind != 2, third = o.kind == 7;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<const float4*>(p)[i];
    float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = m4, s4 = m4;
    if (slots) { m4 = reinterpret_cast<const float4*>(m)[i]; v4 = reinterpret_cast<const float4*>(v)[i]; }
    if (third) s4 = reinterpret_cast<const float4*>(s3)[i];
    optim_one(o, g4.x, p4.x, m4.x, v4.x, s4.x);
    optim_one(o, g4.y, p4.y, m4.y, v4.y, s4.y);
    optim_one(o, g4.z, p4.z, m4.z, v4.z, s4.z);
    optim_one(o, g4.w, p4.w, m4.w, v4.w, s4.w);
    reinterpret_cast<float4*>(p)[i] = p4;
    if (slots) { reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4; }
    if (third) reinterpret_cast<float4*>(s3)[i] = s4;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p16) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
      reinterpret_cast<uint2*>(p16)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
  // tail (n not a multiple of 4; the engine's buckets are 64-element aligned, so this is for other callers)
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pi = p[i], mi = slots ? m[i] : 0.f, vi = slots ? v[i] : 0.f, si = third ? s3[i] : 0.f;
    optim_one(o, g[i], pi, mi, vi, si);
    p[i] = pi;
    if (slots) { m[i] = mi; v[i] = vi; }
    if (third) s3[i] = si;
    if (zero_grad) g[i] = 0.f;
    if (p16) p16[i] = __float2bfloat16(pi);
  }
}
__global__ void step_bump_kernel(int* step) { *step += 1; }
int optim_step(float* p, float* m, float* v, float* s3, float* g, void* p16, long long n, int kind, float lr, float b1,
               float b2, float eps, float gscale, float clip, int zero_grad, int* step, cudaStream_t st) {
  if (kind < 0 || kind > 7 || (kind == 7 && !s3)) return -1;
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                       reinterpret_cast<uintptr_t>(s3) | reinterpret_cast<uintptr_t>(g);
  if ((al & 15) || (reinterpret_cast<uintptr_t>(p16) & 7)) return -1;
  OptimScalars o{lr, lr, b1, b2, eps, gscale, clip, kind};
  optim_kernel<<<stride_grid(n >> 2, 256, 2), 256, 0, st>>>(p, m, v, s3, g, (bf16*)p16, n, o, zero_grad, step);
  step_bump_kernel<<<1, 1, 0, st>>>(step);
  return 0;
}

// Re-layout of every K-major weight copy of one optimizer group in ONE launch: entry e describes a [T][A][B] bf16
// block of the group's compute copy and the [T][B][A] destination (transposed per filter tap); the 32x32 tiles of
// all entries are numbered consecutively (tile_begin) and a block finds its entry by scanning the short table.
__global__ void __launch_bounds__(256) transpose_batch_kernel(const TransposeEntry* __restrict__ tab, int count) {
  // 64 x 64 tiles, 16-byte global accesses on both sides; the tile is transposed on its way into shared memory
  __shared__ bf16 sm[64][66];                 // sm[b][a]; 66: rows 4-byte aligned, bank = (b + a/2) % 32
  __shared__ TransposeEntry e;
  if (threadIdx.x == 0) {
    int k = 0;
    while (k + 1 < count && (long long)blockIdx.x >= tab[k + 1].tile_begin) ++k;
    e = tab[k];
  }
  __syncthreads();
  const int A = e.A, B = e.B;
  const int tb = (B + 63) / 64, ta = (A + 63) / 64;
  int t = (int)((long long)blockIdx.x - e.tile_begin);
  const int bx = t % tb; t /= tb;
  const int by = t % ta; t /= ta;
  const bf16* in = reinterpret_cast<const bf16*>(e.in) + (long long)t * A * B;
  bf16* out = reinterpret_cast<bf16*>(e.out) + (long long)t * A * B;
  const int a0 = by * 64, b0 = bx * 64;
  const int lane8 = (threadIdx.x & 7) * 8, row = threadIdx.x >> 3;          // 8 vectors per 64-wide row, 32 rows per pass
  const bool vin = (B & 7) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  const bool vout = (A & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int a = a0 + row + 32 * i, b = b0 + lane8;
    if (a < A) {
      if (vin && b + 8 <= B) {
        const uint4 v = *reinterpret_cast<const uint4*>(in + (long long)a * B + b);
        const bf16* vp = reinterpret_cast<const bf16*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[lane8 + j][row + 32 * i] = vp[j];
      } else {
        for (int j = 0; j < 8 && b + j < B; ++j) sm[lane8 + j][row + 32 * i] = in[(long long)a * B + b + j];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int b = b0 + row + 32 * i, a = a0 + lane8;
    if (b < B) {
      if (vout && a + 8 <= A) {
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(&sm[row + 32 * i][lane8]);
        *reinterpret_cast<uint4*>(out + (long long)b * A + a) = make_uint4(sp[0], sp[1], sp[2], sp[3]);
      } else {
        for (int j = 0; j < 8 && a + j < A; ++j) out[(long long)b * A + a + j] = sm[row + 32 * i][lane8 + j];
      }
    }
  }
}
int transpose_batch(const void* table, int count, long long total_tiles, cudaStream_t st) {
  if (count <= 0 || total_tiles <= 0 || total_tiles > 0x7fffffffLL) return -1;
  transpose_batch_kernel<<<(int)total_tiles, 256, 0, st>>>((const TransposeEntry*)table, count);
  return 0;
}

}  // namespace b200
