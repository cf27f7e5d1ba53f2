// Host-callable launchers of the HBM-bound kernels (simt_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace b200 {

struct SmallConvArgs {
  int N, H, W, Cs;          // small-channel tensor [N,H,W,Cs]
  int Ho, Wo, Cb;           // big-channel tensor   [N,Ho,Wo,Cb]
  int k, stride, pad_t, pad_l;
  const float* bias;
  int act;
  float leak;
  const void* mask_src;
  int mask_kind;
  void* out;
  int out_f32;
};

int smallc_fprop(const void* xs, const void* w, const SmallConvArgs& a, cudaStream_t st);
int smallc_dgrad(const void* big, const void* w, const SmallConvArgs& a, cudaStream_t st);
int smallc_wgrad(const void* xs, const void* big, float* dw, const SmallConvArgs& a, float alpha, cudaStream_t st);

int im2col_small(const void* xs, void* A, const SmallConvArgs& a, int Kp, cudaStream_t st);
int wpad_transpose(const void* w, void* wt, int kk, int Cb, int Kp, cudaStream_t st);
int col2im_small(const float* T, const SmallConvArgs& a, int Kp, cudaStream_t st);

int smallout_fprop(const void* x, const void* w, const SmallConvArgs& a, cudaStream_t st);
int smallout_dgrad(const void* dy, const void* w, const SmallConvArgs& a, cudaStream_t st);
int smallout_wgrad(const void* x, const void* dy, float* dw, const SmallConvArgs& a, float alpha, cudaStream_t st);

int maskmul(const void* g, const void* a, void* out, long long n, int kind, float leak, cudaStream_t st);
int affine_act(const void* in, int in_type, void* out, int out_f32, long long n, float mul, float add, int act,
               float leak, cudaStream_t st);
int axpby(const void* a, int a_f32, float sa, const float* dev_sa, const void* b, int b_f32, float sb, void* out,
          int out_f32, long long n, cudaStream_t st);
int mul_add(const void* a, const void* b, const void* c, void* out, long long n, cudaStream_t st);
int fill_f32(float* out, long long n, float v, cudaStream_t st);
int interp(const void* x, const void* g, const float* alpha, void* out, int B, int D, cudaStream_t st);
int rowscale(const void* in, const float* s, float mul, float add, void* out, int B, int D, cudaStream_t st);
int dropout_apply(const void* in, const float* u, void* out, long long n, float keep, cudaStream_t st);
int instnorm_fwd(const void* x, const float* scale, const float* shift, void* out, float* stats, int N, int HW, int C,
                 float eps, cudaStream_t st);
int instnorm_bwd(const void* g, const void* x, const float* stats, const float* scale, void* dx, float* dscale,
                 float* dshift, int N, int HW, int C, cudaStream_t st);
int layout_convert(const void* in, int in_type, void* out, int to_nchw, int N, int C, int HW, float mul, float add,
                   cudaStream_t st);
int summary_stats(const void* x, int x_type, long long n, float* out5, unsigned int* counts, int nb, cudaStream_t st);
int montage(const void* x, int x_type, float* out, int m, int n, int H, int W, int C, float mul, float add, cudaStream_t st);
int splitk_finalize(const float* ws, void* out, int out_f32, long long n, int C, const float* bias, int act, float leak,
                    const void* mask, int mask_kind, cudaStream_t st);
int slice_cols(const void* in, long long in_ld, int in_off, void* out, long long out_ld, int out_off, long long rows,
               int cols, const void* mask, int mask_kind, float leak, cudaStream_t st);
int transpose_to_bf16(const void* in, int in_f32, void* out, int T, int A, int B, cudaStream_t st);
int colsum(const void* x, const float* wrow, float* out, long long R, int C, float alpha, cudaStream_t st);
int bn_sums(const void* z, float* stats, long long R, int C, cudaStream_t st);
int bn_apply(const void* z, const float* stats, const float* beta, void* out, long long R, int C, float eps, int act,
             float leak, cudaStream_t st);
int bn_update_moving(const float* stats, long long R, int C, float* mm, float* mv, float decay, int unbiased, cudaStream_t st);
int bn_bwd(const void* g, const void* z, const float* stats, float* bsum, void* dz, long long R, int C, float eps,
           cudaStream_t st);
int gemv_rows(const void* a, const void* w, const float* bias, float* out, int M, int K, int act, float leak,
              cudaStream_t st);
int outer_mask(const float* g, const void* w, const void* mask, void* out, int M, int K, int kind, float leak,
               cudaStream_t st);
int reduce_sum(const void* x, int x_f32, long long n, float* out, float alpha, int sq, cudaStream_t st);
int wgan_loss(const float* sums, int B, int use_gp, float lambda, float* out, cudaStream_t st);
int eltloss(const void* a, int a_f32, const void* b, long long n, int kind, float lab, float scale, float gscale,
            float* out_sum, void* grad, int grad_f32, int mask_kind, float leak, cudaStream_t st);
int philox_fill(void* out, int out_f32, long long n, unsigned long long seed, unsigned long long* draw,
                unsigned int stream_id, int normal, cudaStream_t st);
int optim_step(float* p, float* m, float* v, float* s3, float* g, void* p16, long long n, int kind, float lr, float b1,
               float b2, float eps, float gscale, float clip, int zero_grad, int* step, cudaStream_t st);
// one entry of b200_transpose_batch's device table (5 x 8 bytes; mirrored by b200_transpose_entry in b200gan.h)
struct TransposeEntry { const void* in; void* out; long long tile_begin; int T, A; int B, pad_; };
int transpose_batch(const void* table, int count, long long total_tiles, cudaStream_t st);

}  // namespace b200
