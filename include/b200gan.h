/* b200gan C ABI — the drop-in boundary for the 3dgan training-step hot path on B200 (sm_100a).
 *
 * Each entry point replaces what TensorFlow executed for one call site of the reference layer API
 * (citations are into algoterranean/3dgan).  Conventions: every pointer is a DEVICE pointer owned by
 * the caller (the Python host keeps them in torch tensors used purely as buffers); activations are
 * NHWC bf16 unless a flag says fp32; parameters/gradients/optimizer state are fp32 with a bf16
 * compute copy; every call is asynchronous on the given CUDA stream and allocates nothing; return
 * value 0 = success, negative = error (b200_last_error() has the text).  Thread-compatible, not
 * thread-safe.  There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef B200GAN_H_
#define B200GAN_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200_stream;   /* cudaStream_t */

enum { B200_ACT_NONE = 0, B200_ACT_RELU = 1, B200_ACT_LRELU = 2, B200_ACT_TANH = 3, B200_ACT_SIGMOID = 4 };
enum { B200_OPT_ADAM = 0, B200_OPT_RMSPROP = 1, B200_OPT_SGD = 2, B200_OPT_MOMENTUM = 3 };

/* Geometry of one strided SAME convolution (TF semantics, ops/layers.py:101):
 * input [N,H,W,Cin] -> output [N,Ho,Wo,Cout], filter k x k, pad_t/pad_l = TF's pad_before. */
typedef struct {
  int N, H, W, Cin;
  int Ho, Wo, Cout;
  int k, stride, pad_t, pad_l;
} b200_conv_geom;

/* Fused epilogue: v = acc (+bias[c]); v = act(v); v *= act'(mask_src) ; store bf16 | fp32 (+=). */
typedef struct {
  const float* bias;      /* [channels] or NULL                       (tf.nn.bias_add, ops/layers.py:102,143) */
  int act;                /* B200_ACT_*                               (activation(h), ops/layers.py:104,145)   */
  float leak;             /* lrelu leak                               (ops/activations.py:11-29)               */
  const void* mask_src;   /* bf16, same shape as out, or NULL: multiply by the derivative of mask_kind at the
                             stored post-activation value (gradient of the consumer's activation)            */
  int mask_kind;
  int out_f32;            /* 0: bf16 output, 1: fp32 output */
  int accumulate;         /* fp32 only: out += result */
  /* Sign bitmaps (optional, honoured where b200_conv2d_epilogue_bits() returns 1): one bit per output element,
     uint16 words of 16 consecutive channels, bits_pitch words per output row (pixel).  relu / lrelu gradients
     only need sign(activation) (ops/activations.py:28: slope = leak for x <= 0), so the layer that produces an
     activation also writes its bitmap and the gradient of the layer above reads 2 bytes per 16 channels
     instead of the bf16 activation itself. */
  const void* mask_bits;  /* uint16 [rows, bits_pitch] of the tensor mask_src would name; used instead of it */
  void* bits_out;         /* uint16 [rows, bits_pitch]: bit j of word (row, c/16) = out[row, c + j] > 0      */
  int bits_pitch;
} b200_epilogue;

const char* b200_last_error(void);
int b200_abi_version(void);
/* 0 if the current CUDA device can run the library (compute capability 10.x), negative otherwise. */
int b200_device_check(void);
/* Launch-planner knobs a host may override at run time (the B200GAN_* environment variables of DESIGN.md 5.2 are
 * read once per process).  Keys: "dual_min_pct" - 0: the tap GEMMs use two pixel tiles per CTA (and the 2-CTA
 * kernels) whenever the geometry allows, which is how the tests reach those kernels at small batch; > 100: never;
 * anything else (default 65): the planner's rule (>= 4 pixel tiles).  "bn_tile_cap" - widest N tile of the conv tap
 * GEMMs (0 = planner's choice, else e.g. 64 / 128 / 256); "tap_splits" - split-K factor of the conv tap GEMMs
 * (-1 = planner's choice); "wgrad_min_chunks" - fewest 64-pixel chunks per CTA of the filter gradient's plain
 * stream-K cut (0 = the launcher's own cost model).  The last three exist for tools/tune_layers.py / tune_wgrad.py.
 * Unknown key: negative return. */
int b200_set_tuning(const char* key, int value);

/* ---- convolution family (tcgen05 implicit GEMM; small-channel image-side layers use coalesced SIMT kernels)
 * fprop : y  = epi(conv_SAME(x, W))            replaces tf.nn.conv2d            ops/layers.py:101, hem/ops/layers.py:118
 * dgrad : dx = epi(conv_SAME^T(dy, W))         replaces tf.nn.conv2d_transpose  ops/layers.py:142, hem/ops/layers.py:189
 *                                              and the input-gradient TF autodiff emits (models/gan.py:67-68,228)
 * wgrad : dW += alpha * d(conv)/dW             replaces the filter-gradient TF autodiff emits
 * w  : bf16 [k,k,Cin,Cout] (TF filter layout; for a deconv layer TF's [k,k,out,in] is the same thing)
 * w_t: bf16 [k,k,Cout,Cin] (per-tap transpose; only fprop on the tensor-core path reads it) */
int b200_conv2d_fprop(const void* x, const void* w, const void* w_t, void* y, const b200_conv_geom* g,
                      const b200_epilogue* e, void* workspace, long long workspace_bytes, b200_stream s);
int b200_conv2d_dgrad(const void* dy, const void* w, void* dx, const b200_conv_geom* g, const b200_epilogue* e,
                      void* workspace, long long workspace_bytes, b200_stream s);
int b200_conv2d_wgrad(const void* x, const void* dy, float* dw, const b200_conv_geom* g, float alpha,
                      void* workspace, long long workspace_bytes, int workspace_holds_im2col, b200_stream s);
/* workspace_holds_im2col: pass 1 with the workspace a previous b200_conv2d_fprop of the SAME x and geometry
 * used (its first bytes are im2col(x)); the small-channel wgrad then skips recomputing it. */
/* Same, and dbias[Cout] += column sums of dy (the bias gradient TF autodiff emits for ops/layers.py:101-105) in the
 * same pass: the gathered rows of the image-side route carry a ones column.  Only geometries for which
 * b200_conv2d_wgrad_folds_bias(g, workspace != NULL) returns 1 accept dbias != NULL. */
int b200_conv2d_wgrad_bias(const void* x, const void* dy, float* dw, float* dbias, const b200_conv_geom* g, float alpha,
                           void* workspace, long long workspace_bytes, int workspace_holds_im2col, b200_stream s);
int b200_conv2d_wgrad_folds_bias(const b200_conv_geom* g, int has_workspace);
/* scratch the call needs (0 for the direct tensor-core route).  Small-channel (image-side, Cin <= 4) layers
 * run on the tensor cores when a workspace of this size is passed: fused kernels that gather the filter window
 * (fprop, wgrad) or scatter-gather the col2im (dgrad) inside the kernel (csrc/img_conv.cu) where the geometry
 * allows, else im2col / col2im through the workspace + the same tcgen05 GEMM; with workspace == NULL they fall
 * back to the coalesced SIMT kernels. */
long long b200_conv2d_workspace_bytes(const b200_conv_geom* g, int op /*0 fprop,1 dgrad,2 wgrad*/);
/* which kernel family a geometry maps to: 1 tensor core, 2 small-channel input (image side), 3 small-channel
 * output backward (<= 4 output channels, SIMT), negative = unsupported */
int b200_conv2d_route(const b200_conv_geom* g, int op /*0 fprop,1 dgrad,2 wgrad*/);
/* 1 if that call honours b200_epilogue.bits_out / mask_bits (tensor-core epilogues), else 0: pass mask_src */
int b200_conv2d_epilogue_bits(const b200_conv_geom* g, int op, int has_workspace);

/* ---- dense with one output unit (critic fc2, models/gan.py:285; replaces tf.matmul ops/layers.py:57) */
int b200_gemv_rows(const void* a, const void* w, const float* bias, float* out, int M, int K, int act, float leak,
                   b200_stream s);                                   /* out[m] = act(a[m,:].w + bias[0])      */
int b200_outer_mask(const float* g, const void* w, const void* mask, void* out, int M, int K, int mask_kind,
                    float leak, b200_stream s);                      /* out[m,k] = g[m] w[k] act'(mask[m,k])  */

/* ---- batch norm, tf.contrib.layers.batch_norm defaults (ops/layers.py:58,103,144): beta only, eps, batch stats */
int b200_bn_sums(const void* z, float* stats /*[2C], zeroed*/, long long R, int C, b200_stream s);
int b200_bn_apply(const void* z, const float* stats, const float* beta, void* out, long long R, int C, float eps,
                  int act, float leak, b200_stream s);
/* UPDATE_OPS of batch_norm (run by the train ops that depend on `batchnorm_updates`, models/gan.py:69-70,130,165):
 * moving -= (moving - batch) * (1 - decay) from the sums of b200_bn_sums; unbiased != 0: n/(n-1) variance (fused NCHW kernel) */
int b200_bn_update_moving(const float* stats, long long R, int C, float* moving_mean, float* moving_var, float decay,
                          int unbiased, b200_stream s);
int b200_bn_bwd(const void* g, const void* z, const float* stats, float* bsum /*[2C], zeroed; bsum[0:C] = dbeta*/,
                void* dz, long long R, int C, float eps, b200_stream s);

/* ---- elementwise / reductions */
int b200_maskmul(const void* g, const void* a, void* out, long long n, int kind, float leak, b200_stream s);
int b200_affine_act(const void* in, int in_type /*0 bf16, 1 fp32, 2 uint8*/, void* out, int out_f32, long long n,
                    float mul, float add, int act, float leak, b200_stream s);
                                                                     /* out = act(in*mul+add); x -> 2(x-0.5), models/gan.py:50.  in_type 2 = the input stage:
                                                                        image bytes as decoded (data.py:14-22,29: cast to float32, / 255) are normalised here,
                                                                        the caller folds 1/255 into mul */
int b200_axpby(const void* a, int a_f32, float sa, const float* dev_sa, const void* b, int b_f32, float sb, void* out,
               int out_f32, long long n, b200_stream s);             /* out = a*sa*(*dev_sa) + b*sb */
int b200_mul_add(const void* a, const void* b, const void* c, void* out, long long n, b200_stream s);
                                                                     /* out = a + b*c (bf16): z = mu + sigma*eps, models/vae.py:127-128 */
int b200_fill_f32(float* out, long long n, float v, b200_stream s);
int b200_interp(const void* x, const void* g, const float* alpha, void* out, int B, int D, b200_stream s);
                                                                     /* x + alpha (g - x), models/gan.py:224-226 */
int b200_rowscale(const void* in, const float* s_row, float mul, float add, void* out, int B, int D, b200_stream s);
/* ---- Gen-2 layer options (hem/ops/layers.py) */
int b200_dropout(const void* in, const float* u, void* out, long long n, float keep_prob, b200_stream s);
                                                                     /* tf.nn.dropout(h, keep_prob) (hem/ops/layers.py:64,132,208): out = in [u >= 1-keep] / keep,
                                                                        u ~ U[0,1) from b200_philox; the gradient goes through the same call with the same u */
int b200_instnorm_fwd(const void* x, const float* scale, const float* shift, void* out, float* stats /*[N][2C]: mean, rstd*/,
                      int N, int HW, int C, float eps, b200_stream s);   /* hem.instance_norm (hem/ops/images.py:73-89), NHWC */
int b200_instnorm_bwd(const void* g, const void* x, const float* stats, const float* scale, void* dx, float* dscale,
                      float* dshift, int N, int HW, int C, b200_stream s);   /* dscale / dshift are accumulated (+=) */
/* ---- layout at the boundary: the Gen-2 pipeline and layers are NCHW (hem/ops/layers.py:117-119), the kernels NHWC.
 * to_nchw = 0: out (bf16 NHWC) = in (NCHW; in_type 0 bf16, 1 fp32, 2 uint8 image bytes) * mul + add  -- the input stage;
 * to_nchw = 1: out (fp32 NCHW) = in (NHWC bf16 | fp32) * mul + add                                   -- samples / summaries */
int b200_layout_convert(const void* in, int in_type, void* out, int to_nchw, int N, int C, int HW, float mul, float add,
                        b200_stream s);
/* ---- summaries after the step (ops/summaries.py:13-40,95-124) */
int b200_summary_stats(const void* x, int x_type, long long n, float* out5 /*min,max,sum,sumsq,zeros*/,
                       unsigned int* counts /*[nb] or NULL: TensorBoard-style exponential buckets, growth 1.1 from 1e-12, mirrored*/,
                       int nb, b200_stream s);                       /* tf.summary.histogram + tf.nn.zero_fraction in one pass */
int b200_montage(const void* x, int x_type, float* out, int m, int n, int H, int W, int C, float mul, float add,
                 b200_stream s);                                     /* montage_summary: [m*n,H,W,C] -> [m*H, n*W, C] fp32 */
int b200_slice_cols(const void* in, long long in_ld, int in_off, void* out, long long out_ld, int out_off,
                    long long rows, int cols, const void* mask, int mask_kind, float leak, b200_stream s);
                                                                     /* out[r, out_off+c] = in[r, in_off+c] * act'(mask[r,c]), bf16, row strides in elements:
                                                                        channel concat (tf.concat axis=1, hem/models/pix2pix.py:210-222) and its gradient */
int b200_transpose_to_bf16(const void* in, int in_f32, void* out, int T, int A, int B, b200_stream s);
int b200_colsum(const void* x, const float* wrow, float* out, long long R, int C, float alpha, b200_stream s);
                                                                     /* out[c] += alpha sum_r wrow[r] x[r,c]: bias / fc2 weight grads */
int b200_reduce_sum(const void* x, int x_f32, long long n, float* out, float alpha, int squared, b200_stream s);
int b200_wgan_loss(const float* sums, int B, int use_gp, float lambda, float* out4, b200_stream s);
                                                                     /* models/gan.py:194-205,229-230 */
int b200_eltloss(const void* a, int a_f32, const void* b, long long n, int kind, float label, float scale,
                 float gscale, float* out_sum, void* grad, int grad_f32, int mask_kind, float leak, b200_stream s);
                                                                     /* grad = gscale * dl/da * act'(a), act = mask_kind: the activation that produced a (0: none) */
                                                                     /* models/cnn.py:77, vae.py:76-83, gan.py:193-194, pix2pix.py:283-299 */
/* ---- noise (tf.random_normal / tf.random_uniform, models/gan.py:224,246; models/vae.py:127) */
int b200_philox(void* out, int out_f32, long long n, unsigned long long seed, unsigned long long* dev_draw_counter,
                unsigned int stream_id, int normal, b200_stream s);
/* ---- optimizer (util.py:150-183: every branch of init_optimizer), WGAN clip fused (gan.py:142).  One pass over the flat
 * bucket: p, slots and the bf16 compute copy are updated, and with zero_grad != 0 the gradient is reset to zero for the
 * next run.  kind / (b1, b2, eps) / slots (m, v, slot3):
 *   ADAM     (beta1, beta2, 1e-8)  m, v             tf.train.AdamOptimizer, lr_t = lr sqrt(1-b2^t)/(1-b1^t), t = *dev_step+1
 *   RMSPROP  (decay, momentum, 1e-10)  m = mom, v = ms (init 1)        CENTERED_RMSPROP: + slot3 = mean gradient
 *   SGD      -                      MOMENTUM (momentum)  m
 *   ADAGRAD  v = accumulator (init 0.1); also ProximalAdagrad with l1 = l2 = 0
 *   ADADELTA (rho 0.95, -, 1e-8)  v = accum, m = accum_update
 *   FTRL     lr_power -0.5, l1 = l2 = 0: v = accum (init 0.1), m = linear
 * *dev_step is incremented after the update. */
enum { B200_OPT_ADAGRAD = 4, B200_OPT_ADADELTA = 5, B200_OPT_FTRL = 6, B200_OPT_CENTERED_RMSPROP = 7 };
int b200_optim_step(float* p, float* m, float* v, float* slot3, float* g, void* p_bf16, long long n, int kind, float lr,
                    float b1, float b2, float eps, float grad_scale, float clip, int zero_grad, int* dev_step,
                    b200_stream s);
/* Per-tap transposed bf16 copies ([T][A][B] -> [T][B][A]) of several weights in one launch: the K-major fprop operand
 * of every conv of an optimizer group, refreshed after its update.  dev_table: `count` entries in device memory;
 * tile_begin = number of 64x64 tiles (T * ceil(A/64) * ceil(B/64)) of the entries before this one. */
typedef struct {
  const void* in;        /* bf16 [T][A][B] */
  void* out;             /* bf16 [T][B][A] */
  long long tile_begin;
  int T, A;
  int B, reserved;
} b200_transpose_entry;
int b200_transpose_batch(const b200_transpose_entry* dev_table, int count, long long total_tiles, b200_stream s);

/* ---- data-parallel exchange: average_gradients (util.py:118-147) as an NCCL all-reduce over NVLink.
 * One process per GPU = one tower (util.py:54-77).  libnccl is loaded at run time (b200_nccl_load: explicit path, else
 * "libnccl.so.2"); rank 0 creates the 128-byte unique id, the host exchanges it by any means (the Python host uses its
 * torch.distributed rendezvous), every rank calls b200_nccl_init.  The collectives are in place, asynchronous on the
 * given stream and may be captured in a CUDA graph; the 1/n of the average is folded into b200_optim_step's grad_scale.
 * Errors: negative return, text in b200_nccl_last_error(). */
const char* b200_nccl_last_error(void);
int b200_nccl_load(const char* libnccl_path /* may be NULL */);
int b200_nccl_version(void);
int b200_nccl_unique_id(void* out128);
int b200_nccl_init(const void* id128, int rank, int world, void** comm_out);
int b200_nccl_allreduce_f32(void* comm, float* buf, long long n, b200_stream s);          /* buf = sum over ranks */
int b200_nccl_broadcast_f32(void* comm, float* buf, long long n, int root, b200_stream s);
int b200_nccl_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* B200GAN_H_ */
