/* dd_b200.h — C ABI of libdd_b200.so, the B200 (sm_100a) replacement for the TensorFlow ops on the
 * DeepDenoiser hot path.  The reference has no FFI: its only seam is the Python method
 * Architecture.predict(features, mode) (TensorFlow/Architecture.py:537-617), which builds the graph out
 * of tf.layers.* / tf.nn.* calls.  Every entry point below replaces one family of those calls; the
 * reference call sites are cited on each declaration.  A maintainer binds them with ctypes
 * (INTEGRATION.md); deepdenoiser_b200/_lib.py is that binding.
 *
 * Conventions
 *   - all tensors are NHWC, described by dd_tensor: a device pointer to the START OF THE UNDERLYING
 *     BUFFER, logical dims n,h,w,c, the buffer's channel stride `cstride` (elements per pixel) and the
 *     view's first channel `coff`.  concat == writing at a channel offset of a wider buffer.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing,
 *     never synchronises, and returns 0 on success or a negative dd_status; dd_last_error() returns a
 *     thread-local message.
 *   - dtype: activations are DD_F16 (tensor-core path, fp32 accumulate) or DD_F32 (exact path);
 *     images (sources, predictions, weights of the per-pixel filter) are DD_F32.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef DD_B200_H_
#define DD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_B200_ABI_VERSION 1

typedef enum dd_status {
  DD_OK = 0,
  DD_ERR_INVALID = -1,     /* bad argument (shape / alignment / dtype) */
  DD_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  DD_ERR_UNSUPPORTED = -3, /* combination not implemented */
  DD_ERR_NO_DEVICE = -4
} dd_status;

typedef enum dd_dtype { DD_F32 = 0, DD_F16 = 1, DD_BF16 = 2 } dd_dtype;   /* 16-bit types: storage only, fp32 accumulation */
/* weight-packing code of the split-fp16 ("float16x2") tensor-core mode: W = W_hi + W_lo, both fp16 (dd_conv2d_fwd_split) */
#define DD_F16X2 3

typedef struct dd_tensor {
  void* ptr;
  int32_t dtype; /* dd_dtype */
  int32_t n, h, w, c;
  int32_t cstride;
  int32_t coff;
} dd_tensor;

typedef struct dd_ctx dd_ctx;

/* ---- context ------------------------------------------------------------------------------- */
int dd_abi_version(void);
const char* dd_last_error(void);
int dd_ctx_create(int device, dd_ctx** out);
int dd_ctx_destroy(dd_ctx* ctx);
int dd_ctx_sm_count(const dd_ctx* ctx);
/* Tuning / experiment knobs ("conv_shift_mode", "conv_rows", ...). Returns DD_ERR_INVALID if unknown. */
int dd_ctx_set_option(dd_ctx* ctx, const char* name, int value);
/* Debug: device buffer of 64*8 uint64 that CTA 0 of every subsequent tensor-core conv launch fills with clock64()
 * stamps of its pipeline roles (NULL disables).  Used by tools/probe_conv.py to see where tile time goes. */
int dd_ctx_set_trace_buffer(dd_ctx* ctx, void* device_buffer);
/* Number of kernels this library has launched through `ctx` (bench.py's gpu_launches). */
int64_t dd_ctx_launch_count(const dd_ctx* ctx);

/* ---- convolution weights -------------------------------------------------------------------- */
/* Packs TF-layout weights for dd_conv2d_fwd.
 *   w_hwio   host fp32, tf.layers.conv2d kernel layout [kh,kw,Cin,Cout], or for
 *            transposed convs tf.layers.conv2d_transpose layout [kh,kw,Cout,Cin] (A.5) with
 *            transposed != 0
 *   dtype    DD_F16: [tap][round16(Cout*groups)][round64(Cin)] fp16 (tcgen05 B operand)
 *            DD_F32: [tap][Cout][Cin] fp32 (exact SIMT path)
 * dd_conv2d_packed_bytes() gives the size of `packed` (device memory, written on `stream` from a
 * staging copy the function makes). */
size_t dd_conv2d_packed_bytes(int ksize, int cin, int cout, int dtype, int transposed);
int dd_conv2d_pack_weights(dd_ctx* ctx, const float* w_host, int ksize, int cin, int cout, int dtype,
                           int transposed, void* packed_dev, void* stream);

/* flags for dd_conv2d_fwd */
#define DD_CONV_RELU 1u        /* activation=tf.nn.relu on the output */
#define DD_CONV_RELU_COPY 2u   /* additionally write relu(output) to y_relu (dense-block inputs) */
#define DD_CONV_RESIDUAL_MASK 4u /* `residual` is a ReLU mask source: y = conv(x) * [residual > 0] (fp16 path; the input-gradient
                                   conv of the tensor-core training path fuses the previous layer's ReLU backward this way) */

/* y = conv2d(x, W) + b, stride 1, padding 'same', ksize 3 or 1; optional residual add and ReLU.
 * Replaces tf.layers.conv2d at UNet.py:29-31, Tiramisu.py:35-37,50-52,77-79, Architecture.py:238-243,
 * MultiScalePrediction.py:88-90.  x.dtype selects the path (F16: tcgen05, F32: SIMT); y may be F16 or
 * F32 on the F16 path.  bias: device fp32, round_up(cout,16) entries (zero padded), or NULL.
 * residual / y_relu may be NULL.  On the F16 path channels [cout, round_up(cout,8)) of y are written
 * as zeros, so y needs that much room after coff. */
int dd_conv2d_fwd(dd_ctx* ctx, const dd_tensor* x, const void* w_packed, const float* bias, int ksize,
                  uint32_t flags, const dd_tensor* residual, const dd_tensor* y, const dd_tensor* y_relu,
                  void* stream);

/* dd_conv2d_fwd on the 16-bit tensor-core path that also accumulates colsum_dev[c] += sum over all pixels of y[.., c] (fp32,
 * from the values before the 16-bit rounding).  With DD_CONV_RESIDUAL_MASK this is the input-gradient convolution of layer i
 * fused with the ReLU backward AND the BiasAddGrad of layer i-1 (Training.py:700-702 delegates both to TensorFlow's autodiff):
 * the separate read-only pass over the gradient tensor disappears. */
int dd_conv2d_fwd_colsum(dd_ctx* ctx, const dd_tensor* x, const void* w_packed, const float* bias, int ksize,
                         uint32_t flags, const dd_tensor* residual, const dd_tensor* y, float* colsum_dev, void* stream);

/* y = relu?(conv2d_transpose(x, W, k=2, stride=2, 'same') + b): y[2i+a,2j+b,o] = sum_c x[i,j,c] W[a,b,o,c].
 * Replaces tf.layers.conv2d_transpose at UNet.py:56-58. */
int dd_conv2d_transpose2x2_fwd(dd_ctx* ctx, const dd_tensor* x, const void* w_packed, const float* bias,
                               uint32_t flags, const dd_tensor* y, void* stream);

/* The high-accuracy tensor-core mode ("float16x2"): every activation x is an fp16 PAIR x = x_hi + x_lo (x_hi = fp16(x),
 * x_lo = fp16(x - x_hi): ~22 significant bits), weights likewise (packed with dtype DD_F16X2), and
 *   x.W ~= x_hi.W_hi + x_lo.W_hi + x_hi.W_lo
 * runs as three chunk passes of the same tcgen05 kernel over one fp32 accumulator (the dropped x_lo.W_lo term is 2^-22
 * relative).  The reference computes in fp32 (Training.py:518-524); this mode is what meets its 1e-4 parity bound on tensor
 * cores, at 3x the MMA work.  y is an fp16 pair (y_hi, y_lo) or one fp32 tensor (y_lo NULL); same shapes / flags as above
 * (no residual / relu copy). */
int dd_conv2d_fwd_split(dd_ctx* ctx, const dd_tensor* x_hi, const dd_tensor* x_lo, const void* w_packed, const float* bias,
                        int ksize, uint32_t flags, const dd_tensor* y_hi, const dd_tensor* y_lo, void* stream);
int dd_conv2d_transpose2x2_fwd_split(dd_ctx* ctx, const dd_tensor* x_hi, const dd_tensor* x_lo, const void* w_packed,
                                     const float* bias, uint32_t flags, const dd_tensor* y_hi, const dd_tensor* y_lo, void* stream);

/* y = relu?(conv2d_transpose(x, W, k=3, stride=2, 'same') + b) (Tiramisu.py:62-64; SURVEY A.5: full 2n+1
 * output cropped at the tail).  Computed as 4 output phases, each a stride-1 conv of x with the taps that
 * land on that phase: w_phase[py*2+px] is an ordinary packed 3x3 kernel (dd_conv2d_pack_weights,
 * transposed = 0) holding W[py-2dy, px-2dx]^T at tap (dy+1, dx+1), dy,dx in {0,-1}, zeros elsewhere.
 * y_relu (optional, with DD_CONV_RELU_COPY) receives relu(y). */
int dd_conv2d_transpose3x3_fwd(dd_ctx* ctx, const dd_tensor* x, const void* const* w_phase, const float* bias,
                               uint32_t flags, const dd_tensor* y, const dd_tensor* y_relu, void* stream);

/* ---- pooling / resampling ------------------------------------------------------------------- */
/* tf.layers.max_pooling2d(pool=ksize, strides=2, 'same'): ksize 3 (UNet.py:42-44, pad bottom/right only)
 * or 2 (Tiramisu.py:55-57).  y dims = ceil(x dims / 2). */
int dd_maxpool_s2_fwd(dd_ctx* ctx, const dd_tensor* x, int ksize, const dd_tensor* y, void* stream);
/* the same on fp16 (hi, lo) pairs of the float16x2 mode: max of hi + lo, re-split */
int dd_maxpool_s2_fwd_split(dd_ctx* ctx, const dd_tensor* x_hi, const dd_tensor* x_lo, int ksize, const dd_tensor* y_hi,
                            const dd_tensor* y_lo, void* stream);
/* tf.layers.average_pooling2d(factor, factor, 'same') (MultiScalePrediction.py:11-13); fp32 images. */
int dd_avgpool_fwd(dd_ctx* ctx, const dd_tensor* x, int factor, const dd_tensor* y, void* stream);

/* ---- source encoder ------------------------------------------------------------------------- */
/* FeatureStandardization.standardize + FeatureEngineering.variance for one render pass
 * (Architecture.py:39-46,114-132; FeatureEngineering.py:11-70; Utilities.py:3-4).
 *   src       fp32 [n,h,w,c] c in {1,3};      std_out fp32 [n,h,w,3] (1 channel replicated to 3,
 *             SourceEncoder.py:49-51) or NULL; var_out fp32 [n,h,w,1|3] or NULL
 *   variance_mode 0 'uniform' 1 'neighbor'; variance computed on the raw (before != 0) or standardised
 *   values; relative => / max(mean^2, epsilon); compress => mean over channels. */
typedef struct dd_standardize_params {
  int32_t use_log1p;
  float mean;
  float variance; /* divide by sqrt(variance) iff != 1 */
  int32_t use_variance;
  int32_t variance_mode;
  int32_t relative_variance;
  int32_t compute_before_standardization;
  int32_t compress_to_one_channel;
  float epsilon;
} dd_standardize_params;
int dd_standardize_variance(dd_ctx* ctx, const dd_tensor* src, const dd_standardize_params* prm,
                            const dd_tensor* std_out, const dd_tensor* var_out, void* stream);
/* The same for `count` passes of identical [n,h,w] in ONE launch (Architecture.py:549-555 loops over every pass): arrays of
 * `count` descriptor pointers (std_out[i] / var_out[i] may be NULL) and `count` parameter blocks.  table_dev: device scratch of
 * at least count * dd_standardize_variance_job_bytes() bytes (filled by this call on `stream`). */
int dd_standardize_variance_batch(dd_ctx* ctx, int count, const dd_tensor* const* src, const dd_standardize_params* prm,
                                  const dd_tensor* const* std_out, const dd_tensor* const* var_out, void* table_dev,
                                  size_t table_bytes, void* stream);
size_t dd_standardize_variance_job_bytes(void);

/* SourceEncoder.prepare_neural_network_input (SourceEncoder.py:29-79) for ALL tuples at once:
 * out[t*n + i, y, x, ch] = table[t][ch].ptr ? ptr[((i*h + y)*w + x)*cstride + cidx] : table[t][ch].constant
 * table: device array of tuples*out.c entries. */
typedef struct dd_gather_entry {
  const float* ptr;
  int32_t cstride;
  int32_t cidx;
  float constant;
  int32_t pad_;
} dd_gather_entry;
int dd_assemble_input(dd_ctx* ctx, const dd_gather_entry* table_dev, int tuples, int n, const dd_tensor* out,
                      void* stream);
/* float16x2 mode: the same gather written as an fp16 (hi, lo) pair */
int dd_assemble_input_split(dd_ctx* ctx, const dd_gather_entry* table_dev, int tuples, int n, const dd_tensor* out_hi,
                            const dd_tensor* out_lo, void* stream);

/* ---- kernel prediction ---------------------------------------------------------------------- */
/* KernelPrediction.kernel_prediction (KernelPrediction.py:11-63) with use_softmax=True, mode='symmetric',
 * including the tf.split of the post-processed tensor into one K*K slab per feature of the tuple
 * (Architecture.py:581-587) and the per-feature KernelPredictor loop (Architecture.py:260-289):
 *   out[o,y,x,c] = sum_{i,j} sym(src)[o,y+i-p,x+j-p,c] * softmax_k(logits[b,y,x,f*K*K + k])[i*K+j]
 * logits F16/F32 [B,h,w,features*K*K]; src/out fp32 [features*B,h,w,3].  Logits image b = tuple*ipt + n
 * (ipt = images_per_tuple) and feature f use src/out image o = (tuple*features + f)*ipt + n, i.e. the
 * images of the passes of one tuple are adjacent ("pass-major bank"). */
int dd_kernel_predict_fwd(dd_ctx* ctx, const dd_tensor* src, const dd_tensor* logits, int ksize, int features,
                          int images_per_tuple, const dd_tensor* out, void* stream);

/* ---- multi-scale composition ---------------------------------------------------------------- */
/* First layer of the compose weight net: relu(conv1x1_{6->C}(concat[up2(small), large]))
 * (MultiScalePrediction.py:16-33,57-66).  small fp32 [B,h/2,w/2,3], large fp32 [B,h,w,3],
 * w HOST fp32 [6][C] (TF [1,1,6,C]), b HOST fp32 [C] (copied into the launch parameters);
 * y F16/F32 [B,h,w,C..] (channels >= C zero-filled to y.c). */
int dd_compose_head_fwd(dd_ctx* ctx, const dd_tensor* small, const dd_tensor* large, const float* w,
                        const float* b, int c_mid, const dd_tensor* y, void* stream);
/* Tail: wgt = sigmoid(relu(conv1x1_{C->1}(t) + b)); out = large - wgt*up2(down2(large)) + wgt*up2(small)
 * (MultiScalePrediction.py:36-54,73-77), optionally followed by the inverse standardisation
 * (Architecture.py:48-55) when inv != NULL.  w HOST fp32 [C], b HOST fp32 [1]. */
typedef struct dd_invert_params {
  int32_t use_log1p;
  float mean;
  float variance;
} dd_invert_params;
int dd_compose_tail_fwd(dd_ctx* ctx, const dd_tensor* t, const float* w, const float* b, int c_mid,
                        const dd_tensor* small, const dd_tensor* large, const dd_invert_params* inv,
                        const dd_tensor* out, void* stream);
/* MultiScalePrediction.compose_scales (MultiScalePrediction.py:36-93) in one launch: head 1x1, two residual blocks of 3x3
 * 24->24 convolutions on tcgen05 (row-streaming, layer-pipelined; csrc/compose_rows.cuh), tail 1x1 + sigmoid and the blend
 * out = large - w*up2(down2(large)) + w*up2(small), optionally followed by the inverse standardisation.
 * Intermediates stay in shared memory / TMEM (16-bit between layers, fp32 accumulate; head / tail / blend in fp32).
 * packed_dev:  DEVICE blob of dd_compose_weights_bytes() bytes, 16-byte aligned, built by dd_compose_pack_weights (host) from
 *              the TF kernels [3,3,24,24] and biases [24] of the four convolutions in graph order (the biases and the residual
 *              connections are folded into the tensor-core operands); dtype DD_F16 / DD_BF16 = operand type.
 * params_host: HOST fp32 [dd_compose_params_floats()] = head_w[6][24], head_b[24], conv_b[4][24], tail_w[24], tail_b[1], pad
 *              (copied into the launch parameters). */
size_t dd_compose_weights_bytes(void);
size_t dd_compose_params_floats(void);
int dd_compose_pack_weights(const float* const* conv_w, const float* const* conv_b, int dtype, void* blob_host);
int dd_compose_scales_fwd(dd_ctx* ctx, const dd_tensor* small, const dd_tensor* large, const void* packed_dev,
                          const float* params_host, int dtype, const dd_invert_params* inv, const dd_tensor* out, void* stream);
/* FeatureStandardization.invert_standardization on an fp32 image (Architecture.py:48-55). */
int dd_invert_standardization(dd_ctx* ctx, const dd_tensor* x, const dd_invert_params* inv, const dd_tensor* y,
                              void* stream);

/* ---- fused output head ------------------------------------------------------------------------ */
/* AdjustNumberOfChannels (Architecture.py:230-244: conv1x1 C->O + ReLU, conv1x1 O->O, O = features*K*K) + the per-feature
 * split (Architecture.py:581-587) + KernelPrediction.kernel_prediction (KernelPrediction.py:11-63) of ONE scale in a single
 * kernel: the logits never reach HBM.  x: fp16 core output [B,h,w,C]; src / out: fp32 [B*features,h,w,3] banks laid out as
 * dd_kernel_predict_fwd expects.  Supported: K in {3,5}, features in {1,3} (dd_post_kp_supported); anything else runs the
 * unfused dd_conv2d_fwd x2 + dd_kernel_predict_fwd path.  blob_dev: dd_post_kp_pack_weights() copied to the device. */
int dd_post_kp_supported(int ksize, int features);
size_t dd_post_kp_weights_bytes(int cin, int ksize, int features);
int dd_post_kp_pack_weights(const float* w1, const float* b1, const float* w2, const float* b2, int cin, int ksize, int features,
                            void* blob_host);
int dd_post_kp_fwd(dd_ctx* ctx, const dd_tensor* x, const void* blob_dev, const dd_tensor* src, int ksize, int features,
                   int images_per_tuple, const dd_tensor* out, void* stream);

/* ---- training: loss, backward of every forward op, optimizer --------------------------------- */
/* Exact (fp32 accumulate, CUDA-core) training path; replaces what tf.train.AdamOptimizer.minimize derives by
 * autodiff (Training.py:700-702).  Gradient tensors are fp32. */

/* dz = dy * [y > 0]: backward of activation=tf.nn.relu (UNet.py:29-31 ...). */
int dd_relu_bwd(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* dz, void* stream);
/* dz_acc += dy * [y > 0]: the same through a concatenation (Tiramisu.py:40: every later layer of a dense block and the
 * transition read the same tensor, so their input gradients add up). */
int dd_relu_bwd_acc(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* dz_acc, void* stream);
/* out = a * (b + c): combined lighting pass color * (direct + indirect) (Training.py:420-433). */
int dd_muladd_fwd(dd_ctx* ctx, const dd_tensor* a, const dd_tensor* b, const dd_tensor* c, const dd_tensor* out, void* stream);
/* backward of the above for upstream gradient g: da_acc += g (b + c); dbc_inc = g a (the increment of BOTH db and dc). */
int dd_muladd_bwd(dd_ctx* ctx, const dd_tensor* a, const dd_tensor* b, const dd_tensor* c, const dd_tensor* g,
                  const dd_tensor* da_acc, const dd_tensor* dbc_inc, void* stream);
/* y += alpha * x;  y = value. */
int dd_axpy(dd_ctx* ctx, float alpha, const dd_tensor* x, const dd_tensor* y, void* stream);
int dd_fill(dd_ctx* ctx, float value, const dd_tensor* y, void* stream);
/* dx = dy * d(invert_standardization)/dx evaluated at x (Architecture.py:48-55). */
int dd_invert_standardization_bwd(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* x, const dd_invert_params* inv,
                                  const dd_tensor* dx, void* stream);
/* LossDifference.difference (LossDifference.py:15-36) + reduce_sum over channels + the weighted mean of
 * BaseFeatureTraining.loss (Training.py:126-129,210-243):  *loss_dev += weight * sum_pixels sum_c diff(pred, target),
 * dpred (=|+=) weight * d diff / d pred.  kind: 0 DIFFERENCE 1 ABSOLUTE 2 SMOOTH_ABSOLUTE 3 SQUARED 4 SMAPE;
 * the caller folds loss weight, scale weight and 1/(N*h*w) into `weight`.  dpred may be NULL (evaluation). */
int dd_loss_fwd_bwd(dd_ctx* ctx, const dd_tensor* pred, const dd_tensor* target, int kind, float weight, float epsilon,
                    float* loss_dev, const dd_tensor* dpred, int accumulate, void* stream);
/* BaseFeatureTraining.variation_mean (Training.py:139-186, 304-346): *loss_dev += weight * sum of LossDifference.difference
 * over the horizontal and vertical forward differences of pred / target; dpred_acc += the gradient (may be NULL).  The caller
 * folds loss weight, scale weight and 1 / (N * (h*(w-1) + (h-1)*w)) into `weight`. */
int dd_loss_variation_fwd_bwd(dd_ctx* ctx, const dd_tensor* pred, const dd_tensor* target, int kind, float weight, float epsilon,
                              float* loss_dev, const dd_tensor* dpred_acc, void* stream);
/* *sum_dev += number of pixels with sum_c |mask_src| > 0 (Conv2dUtilities.non_zero_mask, Conv2dUtilities.py:69-74). */
int dd_mask_sum(dd_ctx* ctx, const dd_tensor* mask_src, float* sum_dev, void* stream);
/* BaseFeatureTraining.masked_mean (Training.py:131-137): *loss_dev += weight * sum(difference * mask) / *mask_sum_dev
 * (nothing when the mask is empty), dpred_acc += its gradient. */
int dd_loss_masked_fwd_bwd(dd_ctx* ctx, const dd_tensor* pred, const dd_tensor* target, const dd_tensor* mask_src,
                           const float* mask_sum_dev, int kind, float weight, float epsilon, float* loss_dev,
                           const dd_tensor* dpred_acc, void* stream);
/* Multi-scale SSIM loss term (BaseFeatureTraining.ms_ssim, Training.py:188-204; tf.image.ssim_multiscale [external]).
 * dd_ssim_stats: stats[n, h-10, w-10, {mx, my, sxy, sxx+syy} x C] = the 11x11 Gaussian (sigma 1.5) VALID filter responses of
 * one level;  dd_ssim_reduce: sums_dev[n][c][2] += {sum cs, sum l*cs} over the filtered pixels;  dd_ssim_bwd: given
 * coef_dev[n][c][2] = {d objective / d mean(cs), d objective / d mean(l cs)} accumulates d objective / d x into dx_acc;
 * dd_avgpool2_adjoint: dfine_acc[2i+a, 2j+b] += dcoarse[i, j] / 4 (adjoint of the 2x2 average pooling between levels). */
int dd_ssim_stats(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, const dd_tensor* stats, void* stream);
int dd_ssim_reduce(dd_ctx* ctx, const dd_tensor* stats, int channels, float max_val, float* sums_dev, void* stream);
int dd_ssim_bwd(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, const dd_tensor* stats, const float* coef_dev, float max_val,
                const dd_tensor* dx_acc, void* stream);
int dd_avgpool2_adjoint(dd_ctx* ctx, const dd_tensor* dcoarse, const dd_tensor* dfine_acc, void* stream);
/* dW (TF layout [kh,kw,cin,cout]; transposed: [2,2,cout,cin]) += x^T * dz over all pixels, db += sum dz. */
int dd_conv2d_wgrad(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* dz, int ksize, int transposed, float* dw_dev,
                    float* db_dev, void* stream);
/* ---- mixed-precision (tensor-core) training: fp16 activations / activation gradients, fp32 master weights ---------- */
/* Tensor-core weight gradient (tcgen05, pixel axis = GEMM K, both operands MN-major): x, dz fp16 views of equal spatial
 * size, ksize 1 or 3 (stride 1, 'SAME'); dw_dev fp32 += scale * sum x (x) dz, layout 0 = TF [kh,kw,cin,cout],
 * 1 = [tap][cout][cin] (TF conv2d_transpose kernels).  Replaces Conv2DBackpropFilter of TF autodiff (Training.py:700-702). */
int dd_conv2d_wgrad_tc(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* dz, int ksize, int layout, float* dw_dev, float scale,
                       void* stream);
/* Device-side (re)pack of fp32 master weights into the fp16 layout of dd_conv2d_fwd after every optimizer step.
 * mode 0: forward, w_dev TF [k,k,cin,cout];  mode 1: the input-gradient convolution of that layer (flipped taps, swapped
 * channels; run it with dd_conv2d_fwd on dz);  mode 2: dd_conv2d_transpose2x2_fwd, w_dev TF [2,2,cout,cin].
 * mode | DD_PACK_BF16 writes bfloat16 instead of fp16.
 * packed_dev: dd_conv2d_packed_bytes() bytes (mode 1: of the swapped shape), zeroed once by the caller. */
#define DD_PACK_BF16 16
int dd_conv2d_pack_weights_dev(dd_ctx* ctx, const float* w_dev, int ksize, int cin, int cout, int mode, void* packed_dev,
                               void* stream);
/* out[n,i,j,sp*C+c] = dy[n,2i+ay,2j+ax,c] * [y[same] > 0] (y may be NULL), sp = 2*ay+ax: turns the backward of the
 * stride-2 2x2 transposed convolution (UNet.py:56-58) into 1x1 GEMMs on the coarse grid. */
int dd_space_to_depth2_mask(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* out, void* stream);
/* dz = dy * [y > 0] (y and dz may be NULL) and db_dev[c] += scale * sum_pixels dz[..,c] (db_dev may be NULL) in one pass. */
int dd_relu_bwd_bias(dd_ctx* ctx, const dd_tensor* dy, const dd_tensor* y, const dd_tensor* dz, float* db_dev, float scale,
                     void* stream);
/* input gradient of dd_conv2d_transpose2x2_fwd (exact path). */
int dd_conv2d_transpose2x2_dgrad(dd_ctx* ctx, const dd_tensor* dz, const float* w_dgrad, const dd_tensor* dx, void* stream);
/* Input gradient of the 3x3 stride-2 'SAME' transposed convolution (Tiramisu.py:62-64): dx[i,j,c] = sum dz[2i+r,2j+s,o] W[r,s,o,c].
 * w_dgrad: [tap][cin][cout] fp32 (dd_conv2d_repack_f32 with ksize = 3, transposed = 1). */
int dd_conv2d_transpose3x3_dgrad(dd_ctx* ctx, const dd_tensor* dz, const float* w_dgrad, const dd_tensor* dx, void* stream);
/* device-side repack of fp32 master weights after an optimizer step: forward layout and the layout of the
 * input-gradient convolution (spatially flipped, channels swapped); either output may be NULL. */
int dd_conv2d_repack_f32(dd_ctx* ctx, const float* w_dev, int ksize, int cin, int cout, int transposed, float* fwd_packed,
                         float* dgrad_packed, void* stream);
/* backward of dd_maxpool_s2_fwd: dx (fp32, pre-zeroed or accumulating) += dy routed to the first maximum of each window. */
int dd_maxpool_s2_bwd(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, const dd_tensor* dy, int ksize, const dd_tensor* dx,
                      void* stream);
/* Training pair for 16-bit tensors: the forward pass also records the first maximum of every window (index_dev: one byte per
 * (window, channel), [n, oh, ow, c] dense, value r * ksize + s), the backward pass adds dy to dx (in place, 16-bit) where a pixel
 * is the recorded maximum of a window - a pure gather.  tf.layers.max_pooling2d's gradient (UNet.py:42-44 under Training.py:700). */
int dd_maxpool_s2_fwd_index(dd_ctx* ctx, const dd_tensor* x, int ksize, const dd_tensor* y, uint8_t* index_dev, void* stream);
int dd_maxpool_s2_bwd_index(dd_ctx* ctx, const uint8_t* index_dev, const dd_tensor* dy, int ksize, const dd_tensor* dx, void* stream);
/* Same routing for fp16 / bf16 tensors in gather form (no atomics): dx += the gradient, in place, in the tensors' 16-bit type. */
int dd_maxpool_s2_bwd_acc(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, const dd_tensor* dy, int ksize, const dd_tensor* dx,
                          void* stream);
/* backward of dd_kernel_predict_fwd w.r.t. the logits (the source is data): softmax Jacobian included. */
int dd_kernel_predict_bwd(dd_ctx* ctx, const dd_tensor* src, const dd_tensor* logits, const dd_tensor* dout, int ksize,
                          int features, int images_per_tuple, const dd_tensor* dlogits, void* stream);
/* backward of dd_compose_tail_fwd (without fused inversion): dt written, dsmall / dlarge / dw_dev / db_dev accumulated. */
int dd_compose_tail_bwd(dd_ctx* ctx, const dd_tensor* t, const float* w, const float* b, int c_mid, const dd_tensor* small,
                        const dd_tensor* large, const dd_tensor* dout, const dd_tensor* dt, const dd_tensor* dsmall,
                        const dd_tensor* dlarge, float* dw_dev, float* db_dev, void* stream);
/* backward of dd_compose_head_fwd: y is the head's (ReLU'd) output, dy its gradient. */
int dd_compose_head_bwd(dd_ctx* ctx, const dd_tensor* small, const dd_tensor* large, const float* w, int c_mid,
                        const dd_tensor* y, const dd_tensor* dy, const dd_tensor* dsmall, const dd_tensor* dlarge,
                        float* dw_dev, float* db_dev, void* stream);
/* out_dev[g][c] += sum over the images of group g (x.n / groups each) and all pixels of x[..,c]: gradient of the
 * embedding row broadcast by the SourceEncoder (FeatureFlags.py:50-69). */
int dd_channel_sum(dd_ctx* ctx, const dd_tensor* x, int groups, float* out_dev, void* stream);
/* tf.train.AdamOptimizer (Training.py:701) on flat fp32 buffers; g is multiplied by grad_scale first. */
int dd_adam_step(dd_ctx* ctx, float* w, const float* g, float* m, float* v, size_t count, float lr, float beta1, float beta2,
                 float epsilon, int64_t step, float grad_scale, void* stream);

/* The same optimizer step guarded against fp16 overflow, without a host synchronisation: the step is skipped when the
 * gradient holds an inf / NaN (so a non-finite value never reaches w, m or v).  state_dev: device int32[4], zero-initialised by
 * the caller once: [0] steps skipped so far, [1] scratch flag, [2] steps applied (the `step` of the bias correction),
 * [3] scratch.  The caller reads [0] whenever convenient (e.g. when it logs) to lower its loss scale. */
int dd_adam_step_guarded(dd_ctx* ctx, float* w, const float* g, float* m, float* v, size_t count, float lr, float beta1,
                         float beta2, float epsilon, float grad_scale, int32_t* state_dev, void* stream);

/* ---- training input pipeline ----------------------------------------------------------------- */
/* CRC-32C (Castagnoli) of a host buffer: the checksum of the TFRecord framing written by TFRecordsCreator.py:221-230
 * (tf.python_io.TFRecordWriter); host code, needs no device. */
uint32_t dd_crc32c(const void* data, size_t size);
/* On-device data augmentation of one batch of square tiles (DataAugmentation.py:10-200, driver Training.py:794-821):
 * per example e: flip_left_right if flip[e] (:10-28), rot90 counter-clockwise rot[e] times (:47-61), channel permutation
 * perm[e] in 0..5 for 3-channel colour passes (:106-125; table :117-123), screen-space-normal sign fix-ups
 * (kind == DD_AUG_SCREEN_SPACE_NORMAL, :31-45,63-104) and the 3x3 rotation out = in . R[e] for world-space normals
 * (kind == DD_AUG_NORMAL, matrix [e][9] row major or NULL, :184-200).  x, y: [E,S,S,C] fp32, y != x; flip / rot / perm:
 * device int32 [E] (NULL = identity). */
#define DD_AUG_PLAIN 0
#define DD_AUG_COLOR 1
#define DD_AUG_SCREEN_SPACE_NORMAL 2
#define DD_AUG_NORMAL 3
int dd_augment_tiles(dd_ctx* ctx, const dd_tensor* x, int kind, const int32_t* flip_dev, const int32_t* rot_dev,
                     const int32_t* perm_dev, const float* rotation_dev, const dd_tensor* y, void* stream);

/* ---- inference tiling -------------------------------------------------------------------------- */
/* Prediction.py:282-310 / :384-441 on the device: tiles[t] = image[y_t : y_t+S, x_t : x_t+S] and the inverse paste of the
 * kept part [cy0,cy1) x [cx0,cx1) of every tile.  table_dev: int32 [T][6] = {y, x, cy0, cy1, cx0, cx1}; image [1,H,W,C],
 * tiles [T,S,S,C]; the caller must make sure every table entry lies inside the image. */
int dd_tiles_gather(dd_ctx* ctx, const dd_tensor* image, const int32_t* table_dev, const dd_tensor* tiles, void* stream);
int dd_tiles_scatter(dd_ctx* ctx, const dd_tensor* tiles, const int32_t* table_dev, const dd_tensor* image, void* stream);

/* ---- data-parallel exchange (NCCL over NVLink) --------------------------------------------------- */
/* One process per GPU; rank 0 calls dd_comm_unique_id and hands the 128 bytes to every rank (any side channel), every rank
 * calls dd_comm_init, then dd_comm_allreduce_sum_f32 sums the flat fp32 gradient buffer IN PLACE on `stream` (no host sync).
 * NCCL is loaded at run time (libnccl.so.2); without it these calls return DD_ERR_UNSUPPORTED and nothing else is affected. */
typedef struct dd_comm dd_comm;
int dd_comm_unique_id(void* id128);
int dd_comm_init(dd_ctx* ctx, const void* id128, int rank, int world, dd_comm** out);
int dd_comm_allreduce_sum_f32(dd_comm* comm, float* buf_dev, size_t count, void* stream);
int dd_comm_destroy(dd_comm* comm);

/* ---- utilities ------------------------------------------------------------------------------ */
/* dtype / channel-view conversion copy y = cast(x) (c channels). */
int dd_cast_copy(dd_ctx* ctx, const dd_tensor* x, const dd_tensor* y, void* stream);
/* Writes `bytes` of a scratch buffer (L2 flush between timed iterations in bench.py). */
int dd_l2_flush(dd_ctx* ctx, void* scratch, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DD_B200_H_ */
