// Fused image-side convolutions (<= 4 input channels: the critic's c1, the generator's last deconv and their
// gradients; ops/layers.py:101,142 at Cin/Cout = 3).  One persistent tcgen05 kernel per direction does the window
// gather / scatter itself -- no im2col / col2im workspace round trip through HBM.
//
// K layout ("row groups of 16"): the k*k*Cin filter taps of a pixel are laid out as k groups (one per filter row kh)
// of 16 bf16 = 32 bytes: slots 0 .. k*Cin-1 hold (kw, c), the rest are zero.  k*Cin <= 15, so slot 15 of every group
// is spare: in the fprop it carries the bias (A = 1.0, B = bias split into a bf16 high and low part in groups 0 and 1),
// so the bias add rides in the GEMM; in the filter gradient the same 1.0 column yields the bias gradient.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace b200 {

struct ImgConvGeom {
  int N, H, W, Cin, Ho, Wo, k, stride, pad_t, pad_l;
};

struct ImgDiv { uint32_t mul; int shr; };     // n / d = (n * mul) >> shr for 0 <= n < 2^31

struct ImgFpropParams {
  ImgConvGeom g;
  const __nv_bfloat16* x;             // NHWC [N,H,W,Cin], 4-byte aligned, even element count
  long long x_words;                  // numel(x) / 2
  const __nv_bfloat16* w;             // TF layout [k*k*Cin][ldw] (bf16 compute copy)
  int ldw;
  const float* bias;                  // [ncols] or null
  long long M;                        // N*Ho*Wo output pixels
  int num_tiles;
  int ncols;                          // output channels (multiple of 16, <= 256) = N tile
  int slots;                          // A ring depth
  ImgDiv div_hw, div_wo, div_hp, div_ppr;    // reciprocals of Ho*Wo, Wo and the padded rows per image
  int win_pitch, win_off, win_hp, win_rows, win_vec16, win_ppr;   // input window in shared memory (set by launch_img_fprop)
  CUtensorMap tmOut;                  // [M, ncols] bf16, box (stage_pitch/2) x 128, no swizzle
  int stage_pitch;                    // bytes per staged output row (>= ncols*2; padded against bank conflicts)
  uint16_t* bits_out;                 // sign bitmap of the activated output (or null)
  const uint16_t* mask_bits;          // sign bitmap to multiply by (act' of the consumer side), or null
  int bits_pitch;                     // words per row of both bitmaps
  int bits_stage;                     // 1: the tile's sign words are staged and written with one bulk copy
  int act;
  float leak;
  int mask_kind;
  __nv_bfloat16* im2col_out;          // optional [M][k*16] copy of the gathered rows (for the later filter gradient)
  int virt;                           // 1: k == 4 and k*Cin == 16 (no free slot in any filter row): the kernel runs as
                                      // its k = 5 form whose fifth row group holds only the ones / bias pair (slots 14, 15)
};

// Fused input gradient / transposed-conv forward of an image-side layer: dx[N,H,W,Cin] = epi(conv^T(dy, W)).
// T[pixel, kh*16 + kw*Cin + c] = sum_cout dy[pixel, cout] * w[kh,kw,c,cout] runs on the tensor cores (dy tiles by TMA,
// the weights resident in shared memory), a whole image (1 or 2 tiles of 128 output pixels) at a time; the fp32 image
// of T stays in shared memory and every output element gathers its <= ceil(k/stride)^2 taps from it -- the 42 MB
// col2im workspace round trip through HBM is gone.
struct ImgDgradParams {
  ImgConvGeom g;                      // geometry of the forward conv: x [N,H,W,Cin], dy [N,Ho,Wo,Cout]
  CUtensorMap tmA;                    // dy as [M, Cout]: box 64 x 128, SWIZZLE_128B
  CUtensorMap tmA_tail;               // same, box 16 x 128, SWIZZLE_32B (Cout % 64 == 16)
  int cout, kfull, ktail;             // Cout = 64 * kfull + ktail, ktail in {0, 16}
  const __nv_bfloat16* w;             // [k*k*Cin][ldw]
  int ldw;
  int tiles_per_image;                // Ho*Wo / 128: 1 or 2
  ImgDiv div_w;                       // reciprocal of W
  const float* bias;                  // [Cin] or null
  int act;
  float leak;
  const __nv_bfloat16* mask_src;      // same shape as the output; result *= act'(mask_src)
  int mask_kind;
  void* out;
  int out_f32;
  int stages;
};
bool img_dgrad_supported(const ImgConvGeom& g, int cout);
void launch_img_dgrad(const ImgDgradParams& p, cudaStream_t stream);

// Fused filter (and bias) gradient of an image-side conv: dw[kh,kw,c,cout] += alpha * sum_pixels x_window * dy, and
// dbias[cout] += alpha * sum_pixels dy (the ones slot of the gather).  The gathered rows are built in shared memory by
// the same producer code as the fprop and used as the MN-major A operand (M = filter slots, K = pixels); dy tiles come
// by TMA as the MN-major B operand; every CTA accumulates its pixel tiles in TMEM and reduces once at the end.
struct ImgWgradParams {
  ImgFpropParams f;                   // gather side: g, x, M, num_tiles, window and reciprocals (set by the launcher)
  CUtensorMap tmDy;                   // dy as [M, Cout]: box 64 x 128, SWIZZLE_128B
  int cout, nblocks;                  // Cout (multiple of 16, <= 256), 64-channel boxes per tile
  int stages;
  float* dw;                          // [k*k*Cin][ldo] fp32, accumulated into
  int ldo;
  float* dbias;                       // [cout] or null
  float alpha;
};
bool img_wgrad_supported(const ImgConvGeom& g, int cout);
void launch_img_wgrad(const ImgWgradParams& p, cudaStream_t stream);

// 1 when the fused kernel takes this call (else the caller uses the im2col + GEMM route)
bool img_fprop_supported(const ImgConvGeom& g, int ncols, int has_bias);
bool img_fprop_virtual_supported(const ImgConvGeom& g, int ncols);
size_t img_fprop_smem(const ImgFpropParams& p);
void launch_img_fprop(const ImgFpropParams& p, cudaStream_t stream);

// standalone gather in the same K layout: out [M][k*16] bf16 (ones in slot 15 of groups 0 and 1 when `ones`)
void launch_img_im2col16(const __nv_bfloat16* x, long long x_words, const ImgConvGeom& g, __nv_bfloat16* out, int ones,
                         cudaStream_t stream);

// weights [k*k*Cin][ldw] -> [ncols][k*16] in the row-group layout (zero pads), for the GEMM route
void launch_img_wpad16(const __nv_bfloat16* w, int ldw, int ncols, int k, int kcin, __nv_bfloat16* wt,
                       cudaStream_t stream);
// t [k*16][ncols] fp32 (filter gradient in the row-group layout) -> dw [k*k*Cin][ldo] +=, dbias [ncols] += row 15
void launch_img_wgrad_fold(const float* t, int ncols, int k, int kcin, float* dw, int ldo, float* dbias,
                           cudaStream_t stream);

}  // namespace b200
