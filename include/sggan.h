/* sggan.h -- C ABI of libsggan_sm100.so: the B200-native SG-GAN training step.
 *
 * The reference (fhfonsecaa/SG-GAN-TF2) has no FFI: its hot path sits behind Python callables
 * that dispatch TensorFlow eager kernels.  Each entry point below names the reference call it
 * replaces.  Conventions: plain C types only; every pointer is a raw DEVICE pointer unless the
 * name ends in _host; tensors are NHWC fp32 at the boundary (what the reference's numpy batches
 * are, model.py:246-256); every call returns 0 on success or a negative SGGAN_E_* code
 * (sggan_last_error() gives a string); launches are asynchronous on the given stream; the library
 * never allocates or frees device memory -- the caller supplies one workspace of
 * sggan_workspace_bytes() bytes; no CPU fallback exists.  A handle is bound to one device and is
 * not thread-safe.
 */
#ifndef SGGAN_H_
#define SGGAN_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGGAN_E_INVALID (-1)   /* bad argument / unsupported shape */
#define SGGAN_E_WORKSPACE (-2) /* workspace missing or too small */
#define SGGAN_E_CUDA (-3)      /* a CUDA runtime / driver call failed */
#define SGGAN_E_STATE (-4)     /* call order (e.g. step before weights) */

#define SGGAN_NET_G 0
#define SGGAN_NET_D 1
#define SGGAN_LOSS_P2P 0   /* as wired: model.py:190-191 */
#define SGGAN_LOSS_SGGAN 1 /* the defined-but-unwired SG-GAN losses, model.py:114-133 */

typedef struct sggan_config {
  int batch;            /* per-GPU batch B (args.batch_size x augmentation, model.py:234-244) */
  int image_height;     /* main.py:19 --img_height */
  int image_width;      /* main.py:20 --img_width  */
  int gf_dim;           /* module.py:221 (hard-coded 64 in the reference; must be 64 here) */
  int df_dim;           /* module.py:274 (64) */
  int segment_class;    /* module.py:275 / main.py:43 */
  int n_blocks;         /* 9 residual blocks, module.py:244-252 */
  int mask_height;      /* mask grid; must broadcast against the D logit grid */
  int mask_width;
  int loss_mode;        /* SGGAN_LOSS_* */
  int use_lsgan;        /* main.py:40, only read in SGGAN_LOSS_SGGAN mode */
  float lr;             /* model.py:82,205: 0.001 */
  float beta1;          /* main.py:28: 0.5 */
  float beta2;          /* Keras default 0.999 */
  float adam_eps;       /* Keras default 1e-7 */
  float in_eps;         /* tfa InstanceNormalization default 1e-3 */
  float p2p_lambda;     /* LAMBDA = 100, model.py:151 */
  float L1_lambda;      /* main.py:37 */
  float Lg_lambda;      /* main.py:38 */
  int world_size;       /* data-parallel ranks: gradients are scaled by 1/world_size before Adam */
} sggan_config;

typedef struct sggan_handle sggan_handle;

/* Fill `cfg` with the reference's effective defaults for an image size. */
void sggan_default_config(sggan_config* cfg, int batch, int height, int width);
/* D logit grid for an input size (module.py:284-311 shape arithmetic). */
void sggan_disc_logit_grid(int height, int width, int* hd, int* wd);
size_t sggan_workspace_bytes(const sggan_config* cfg);
int sggan_create(const sggan_config* cfg, void* workspace, size_t workspace_bytes, void* stream,
                 sggan_handle** out);
void sggan_destroy(sggan_handle* h);
const char* sggan_last_error(void);

/* Weights in Keras creation order (generator: 94 tensors, discriminator: 28). */
int sggan_num_tensors(const sggan_handle* h, int net);
int64_t sggan_tensor_numel(const sggan_handle* h, int net, int idx);
int sggan_tensor_rank(const sggan_handle* h, int net, int idx);
void sggan_tensor_shape(const sggan_handle* h, int net, int idx, int64_t shape[4]);
/* flat fp32 buffers, all tensors back to back in creation order: what=0 params, 1 grads, 2 adam m, 3 adam v */
float* sggan_flat_buffer(sggan_handle* h, int net, int what, int64_t* numel);
int64_t sggan_tensor_offset(const sggan_handle* h, int net, int idx);
/* Must be called after the params buffer was written (re-packs the bf16 GEMM operands). */
int sggan_weights_changed(sggan_handle* h);

/* generator(x): model.py:176,528-532.  real_A [B,H,W,3] fp32 -> fake_A [B,H,W,3] fp32. */
int sggan_gen_forward(sggan_handle* h, const float* real_A, float* fake_A);
/* discriminator([x, mask]): model.py:186-188.  x [B,H,W,3], mask [B,hm,wm,C] -> logits [B,Ho,Wo,1]. */
int sggan_disc_forward(sggan_handle* h, const float* x, const float* mask, float* logits);

/* sggan.train_step: model.py:169-200.  losses_out[0] = gen_loss, [1] = disc_loss (device floats).
 * sggan_train_step = forward_backward + adam; the split lets a data-parallel caller all-reduce the
 * flat gradient buffers in between (discriminator gradients are final after phase 1). */
int sggan_step_forward_backward_d(sggan_handle* h, const float* real_A, const float* seg_A, const float* mask,
                                  float* losses_out);
int sggan_step_backward_g(sggan_handle* h);
/* The same backward in two parts, for data parallelism: after part 0 the generator gradients from element
 * sggan_grad_split_offset(h) to the end of the flat buffer (sggan_flat_buffer(h, SGGAN_NET_G, 1): the middle residual block
 * and every layer above it) are final, so their all-reduce runs underneath part 1, which produces the rest. */
int sggan_step_backward_g_part(sggan_handle* h, int part);
int64_t sggan_grad_split_offset(const sggan_handle* h);
int sggan_step_adam(sggan_handle* h, int net);
/* Same update, issued on the handle's internal side stream so that it overlaps whatever the caller enqueues next
 * (the data-parallel trainer updates D this way while G's all-reduce is in flight).  The next sggan_step_adam /
 * sggan_step_* / sggan_train_step call on this handle joins it.  The step counter advances once both nets of a
 * step have been updated, in either order. */
int sggan_step_adam_async(sggan_handle* h, int net);
int sggan_train_step(sggan_handle* h, const float* real_A, const float* seg_A, const float* mask,
                     float* losses_out);
/* The same step as ONE CUDA graph: capture once per set of (stable) device pointers, then replay.  sggan_graph_capture
 * captures the step for these pointers -- or, if it already has (up to 4 sets are kept: double-buffered inputs alternate
 * between two), just selects that graph; sggan_graph_launch replays the selected one: a single cudaGraphLaunch on the
 * handle's stream that does exactly what sggan_train_step does (Adam's time step is read from a device-side counter, so
 * it advances from replay to replay).  Needs a non-default stream at sggan_create; run one eager sggan_train_step first
 * (lazy kernel attributes). */
int sggan_graph_capture(sggan_handle* h, const float* real_A, const float* seg_A, const float* mask, float* losses_out);
int sggan_graph_launch(sggan_handle* h);
int64_t sggan_step_count(const sggan_handle* h);
/* Resume support: set the number of completed steps (Adam's bias correction uses t = count + 1); the Adam slots
 * themselves are the flat buffers what = 2 / 3 of sggan_flat_buffer.  model.py:450-503 saves weights only; carrying
 * the optimizer over is what keeps a re-planned (new batch / image size) or restarted run on the same trajectory. */
int sggan_set_step_count(sggan_handle* h, int64_t completed_steps);
/* Re-bind the stream every later call of this handle launches on (the handle's internal side stream follows it
 * through events).  The caller orders the switch itself (e.g. synchronises the old stream first). */
int sggan_set_stream(sggan_handle* h, void* stream);
/* Data parallelism without torch: sum this rank's flat gradient buffer of `net` (SGGAN_NET_G, SGGAN_NET_D, or -1 for
 * both, D first) over the communicator with ncclAllReduce(ncclFloat32, ncclSum) in place, on `stream`, after the
 * handle's own stream has produced them (event dependency).  sggan_config.world_size makes Adam apply the 1/N.
 * `nccl_comm` is an ncclComm_t; libnccl.so.2 is resolved at the first call (no link-time dependency). */
int sggan_allreduce_grads(sggan_handle* h, int net, void* nccl_comm, void* stream);
int sggan_kernel_launches(const sggan_handle* h); /* kernels launched by the last train step */
/* fake_A of the last step / forward (device, [B,H,W,3] fp32). */
const float* sggan_last_fake(const sggan_handle* h);

/* Device-side timing of the dominant kernel: CUDA events on the step's stream around every launch of the
 * residual-block 3x3 convolution (forward) while profiling is on.  flops_per_launch = algorithmic
 * 2*B*H*W*Cin*Cout*9 of one launch (bytes for the norm passes, see sggan_profile_select).  The events sit between
 * the launches of the step's stream, so they perturb it slightly: profile in a separate pass, not in a timed one. */
int sggan_profile_begin(sggan_handle* h, int max_launches);
/* Which launches sggan_profile_begin/end bracket (default 0): 0 = the residual-block 3x3 convolutions (forward; work =
 * FLOPs), 1 = the instance-norm + ReLU apply pass behind the first convolution of every block (work = algorithmic bytes:
 * read Y, write the next frame), 2 = the instance-norm backward of the same layers (read Y and dX, write dY), 3 = the
 * generator-side loss kernels as one group per step (seg-edge weights and the gradient-sensitive loss in SG-GAN mode, the
 * L1 / GAN gradient seed, the loss finalize; work = every input read once, every output written once). */
int sggan_profile_select(sggan_handle* h, int kind);
int sggan_profile_end(sggan_handle* h, double* total_ms, int* launches, double* flops_per_launch);

/* Debug / test access to internal activations.  kind: 0 input frame X, 1 raw conv output Y,
 * 2 output-gradient frame dY, 3 input-gradient buffer dX, 4 forward stats.  Returns the device
 * pointer and describes the layout in `desc` (16 ints, see sggan_b200/_lib.py). */
void* sggan_debug_buffer(sggan_handle* h, int net, int layer, int kind, int64_t desc[16]);
int sggan_num_layers(const sggan_handle* h, int net);

/* ---- single operators (ops.py surface + criteria), NHWC fp32 in / out ------------------------- */
/* ops.conv2d (ops.py:24-28) / tf.keras.layers.Conv2D: kernel HWIO, padding 0 VALID, 1 SAME (TF),
 * 2 REFLECT((k-1)/2) then VALID.  bias may be null.  Cin, Cout multiples of 64, stride 1 or 2 (k=3). */
size_t sggan_conv2d_workspace(int B, int H, int W, int Cin, int Cout, int k, int stride, int padding);
int sggan_conv2d_fwd(const float* x, const float* kernel, const float* bias, float* y, int B, int H, int W, int Cin,
                     int Cout, int k, int stride, int padding, void* workspace, size_t workspace_bytes,
                     void* stream);
/* ops.deconv2d (ops.py:30-34) / Conv2DTranspose(3, strides 2, 'same'): kernel (kh,kw,Cout,Cin). */
int sggan_deconv2d_fwd(const float* x, const float* kernel, const float* bias, float* y, int B, int H, int W,
                       int Cin, int Cout, void* workspace, size_t workspace_bytes, void* stream);
/* The fp32-accurate tier of the same two operators (north_star: "rel 1e-4 tf32"): activations stay fp32, the operands
 * are rounded to tf32 and multiplied by tcgen05.mma.kind::tf32 with fp32 accumulation; any Cin / Cout (padded
 * internally).  The reference computes in fp32 (Keras defaults).  passes = 1: one tf32 product per term (operand
 * rounding 2^-11, ~3e-4 relative on the 2304-term dot products of the residual blocks); passes = 3: "3xTF32", the
 * operands are split hi + lo and hi*hi + lo*hi + hi*lo is accumulated (one convolution over 3x the input channels):
 * fp32-level accuracy, <= 1e-5 relative (tests/test_gpu_tf32.py).  bf16 storage (the training path): ~5e-3 per layer. */
size_t sggan_conv2d_tf32_workspace(int B, int H, int W, int Cin, int Cout, int k, int stride, int padding, int passes);
int sggan_conv2d_fwd_tf32(const float* x, const float* kernel, const float* bias, float* y, int B, int H, int W, int Cin,
                          int Cout, int k, int stride, int padding, int passes, void* workspace, size_t workspace_bytes,
                          void* stream);
int sggan_deconv2d_fwd_tf32(const float* x, const float* kernel, const float* bias, float* y, int B, int H, int W, int Cin,
                            int Cout, int passes, void* workspace, size_t workspace_bytes, void* stream);
/* InstanceNormalization (+ activation, + residual) on fp32 storage with double-precision statistics; any C.
 * workspace: B * C * 16 bytes. */
int sggan_instance_norm_fwd_f32(const float* x, const float* gamma, const float* beta, const float* residual, float* y, int B,
                                int H, int W, int C, float eps, int act, float alpha, void* workspace, size_t workspace_bytes,
                                void* stream);
/* Gradients of the same two operators (what gen_tape / disc_tape.gradient compute, model.py:196-197), through the
 * step's own dgrad / wgrad tensor-core kernels: dy [B,Ho,Wo,Cout] -> dx (shape of x), dw (shape of kernel), db [Cout]
 * (db may be null).  sggan_conv2d_bwd_workspace(..., stride = -2, ...) sizes the transposed-convolution case. */
size_t sggan_conv2d_bwd_workspace(int B, int H, int W, int Cin, int Cout, int k, int stride, int padding);
int sggan_conv2d_bwd(const float* x, const float* kernel, const float* dy, float* dx, float* dw, float* db, int B, int H,
                     int W, int Cin, int Cout, int k, int stride, int padding, void* workspace, size_t workspace_bytes,
                     void* stream);
int sggan_deconv2d_bwd(const float* x, const float* kernel, const float* dy, float* dx, float* dw, float* db, int B, int H,
                       int W, int Cin, int Cout, void* workspace, size_t workspace_bytes, void* stream);
/* Backward of InstanceNormalization (+ activation): dz = gradient w.r.t. act(norm(x)); outputs dx, dgamma, dbeta. */
int sggan_instance_norm_bwd(const float* x, const float* gamma, const float* beta, const float* dz, float* dx,
                            float* dgamma, float* dbeta, int B, int H, int W, int C, float eps, int act, float alpha,
                            void* workspace, size_t workspace_bytes, void* stream);
/* ops.instance_norm (ops.py:13-22) / tfa InstanceNormalization, optionally fused activation
 * (0 none, 1 relu, 2 leaky(alpha), 3 tanh) and residual add.  C multiple of 64. */
int sggan_instance_norm_fwd(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                            int B, int H, int W, int C, float eps, int act, float alpha, void* workspace,
                            size_t workspace_bytes, void* stream);
/* ops.lrelu (ops.py:36-37) */
int sggan_lrelu(const float* x, float* y, int64_t n, float leak, void* stream);
/* multiply([h4, mask]) + reduce_sum(axis=-1) (module.py:312-314) */
int sggan_mask_reduce(const float* h4, const float* mask, float* out, int B, int Hd, int Wd, int hm, int wm, int C,
                      void* stream);
/* abs_criterion / mae_criterion / sce_criterion (module.py:336-345): mode 0 / 1 / 2; out = device float */
int sggan_criterion(const float* a, const float* b, int64_t n, int mode, float* out, void* stream);
/* weighted_seg_A (model.py:115-119): seg [B,H,W,3] -> weight [B,H,W,1] */
int sggan_seg_edge_weight(const float* seg, float* weight, int B, int H, int W, void* stream);
/* gradloss_criterion (module.py:347-351): out = device float; d_in (may be null) = d out / d in */
int sggan_gradloss(const float* in, const float* target, const float* weight, float* out, float* d_in, int B, int H,
                   int W, void* stream);
/* tf_deriv (module.py:325-334): per-channel 3x3 Sobel x / y, x [B,H,W,C] -> out [B,Ho,Wo,2C] (channel c*2 + {x,y});
 * valid = 0: 'SAME' zero padding (Ho = H), 1: 'VALID' (Ho = H - 2) */
int sggan_tf_deriv(const float* x, float* out, int B, int H, int W, int C, int valid, void* stream);
/* Keras Adam apply (model.py:199-200): t = 1-based step index; the four buffers 16-byte aligned */
int sggan_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int64_t t, float lr, float beta1,
                    float beta2, float eps, void* stream);
/* one_hot + nearest resample of a class-id map to the D logit grid (utils.py:158-165,190):
 * ids [B,H,W] uint8 -> mask [B,hd,wd,C] fp32 */
int sggan_onehot_mask(const uint8_t* ids, float* mask, int B, int H, int W, int hd, int wd, int C, void* stream);
/* one_hot + scipy.ndimage.zoom(..., order 3, mode "nearest") of the class-id map -- the loader's mask construction
 * (utils.py:190,197-199) -- without the one-hot volume: mask[b,i,j,c] = round(sum_y sum_x wy[i][y] wx[j][x] [ids == c]).
 * wy [ho][wh] / wx [wo][ww]: windows of the per-axis spline weights (fp64, device), y0 [ho] / x0 [wo]: first input row /
 * column of each window (host helper: utils.zoom_weights).  ids [B,H,W] uint8 -> mask [B,ho,wo,C] fp32 (integers). */
int sggan_zoom_mask(const uint8_t* ids, const double* wy, const int* y0, const double* wx, const int* x0, int wh, int ww,
                    float* mask, int B, int H, int W, int ho, int wo, int C, void* stream);
/* RGB -> class id LUT (segment_class.py:60-70,95-97): rgb [n,3] uint8 -> ids [n] uint8 */
int sggan_rgb_to_class(const uint8_t* rgb, uint8_t* ids, int64_t n, void* stream);

/* ---- evaluation that follows the path each epoch (metric.py:18-47,71-77; model.py:307-378) ---- */
/* scores_seg_fake's label adapter: labels[b, w, h] = argmax_c uint8(255 * img[b, h, w, c]) (first maximum; note the
 * reference's (0,3,2,1) transpose).  img [B,H,W,3] fp32 -> labels [B,W,H] int32. */
int sggan_rgb_argmax_labels(const float* img, int32_t* labels, int B, int H, int W, void* stream);
/* _fast_hist: hist[t * n_class + p] += 1 over n label pairs with 0 <= t < n_class (hist: n_class^2 int64, accumulated
 * into -- zero it first for a fresh matrix).  n_class <= 96. */
int sggan_fast_hist(const int32_t* label_true, const int32_t* label_pred, int64_t n, int n_class, int64_t* hist,
                    void* stream);

/* ---- host helper: CRC-32C (Castagnoli) as used by the TF checkpoint format the reference saves (model.py:463-466);
 * crc = value to extend, 0 to start.  Runs on the CPU, touches no device. */
uint32_t sggan_crc32c(const void* data, size_t n, uint32_t crc);

#ifdef __cplusplus
}
#endif
#endif /* SGGAN_H_ */
