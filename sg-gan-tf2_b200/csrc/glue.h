// glue.h -- launchers of the bandwidth-bound kernels of the SG-GAN step (glue.cu): instance
// norm forward/backward fused with activation, residual add and frame (padding / phase-split)
// construction; image packing; losses and their gradients; Adam; weight packing.
#pragma once
#include <cuda_runtime.h>

#include "kparams.h"

namespace sggan {

// fp32 NHWC 3-channel image -> 8-channel bf16 frame (reflect border / phase split per dmap).
// The destination holds images [dst_b0, dst_b0 + B).
void launch_prep_image3(const float* src, int B, int H, int W, sg_bf16* dst, const FrameMap& dmap, int dst_b0,
                        cudaStream_t st);

// instance-norm passes as row streams (glue_rows.cu)
// All return 0 (launch_in_bwd_reduce: the number of partial-sum blocks per image, to be passed to the apply pass as
// InBwdParams::sums_nblk) or a negative cudaError.
int launch_in_apply(const InApplyParams& p, cudaStream_t st);
int launch_in_bwd_reduce(const InBwdParams& p, cudaStream_t st);
int launch_in_bwd_apply(const InBwdParams& p, cudaStream_t st);
int launch_in_bwd_fused(const InBwdParams& p, cudaStream_t st);  // reduce + apply in one launch (needs sync_ctr)
size_t in_bwd_partials_bytes(int C);  // size of InBwdParams::sums_part

// out[b,i,j,c] (plain bf16 [B][H][W][C]) = g1 + g2 (either may fold a reflected border back).
int launch_grad_gather(const GradSrc& g1, const GradSrc& g2, int B, int H, int W, int C, sg_bf16* out,
                       cudaStream_t st);

// Backward of an activation applied in a conv epilogue (discriminator h0: LeakyReLU without norm):
// dy = dz * act'(z), z read from the frame the epilogue wrote; also accumulates the bias gradient
// over the first nb_bias images.
// Deterministic cross-block sums: every block deposits its partial values in `scratch` ([blocks][K] floats), the block
// that draws the last ticket adds all deposits in block order (fixed assignment to threads, fixed combination tree) and
// accumulates the totals into the destination -- no floating-point atomics, so the result does not depend on the order in
// which blocks finish.  `ticket` wraps back to zero by itself (atomicInc).  scratch == nullptr: the kernels fall back to
// atomicAdd (stand-alone operator calls without a workspace).
struct OrderedSum {
  float* scratch;
  unsigned int* ticket;
};

struct ActBwdParams {
  GradSrc g;
  const sg_bf16* Z;
  FrameMap zmap;
  int B, H, W, C;
  int nb_act, act_wrap;
  float alpha;
  sg_bf16* dst;
  FrameMap dmap;
  float* dbias;
  int nb_bias;
  OrderedSum red;  // scratch: gridDim.x * nb_bias * C floats
};
void launch_act_bwd(const ActBwdParams& p, cudaStream_t st);

// Semantic-aware masking + GAN losses + their gradient w.r.t. the h4 logits (module.py:311-314,
// model.py:126-133,149-166).  h4: fp32 [2B][Hd][Wd][Cs] (real images first, then fake);
// mask: fp32 [B][hm][wm][Cs], broadcast against h4 the way tf.keras.layers.multiply does.
struct DiscLossParams {
  const float* h4;
  const float* mask;
  int B, Hd, Wd, hm, wm, Cs;
  int lsgan;        // 0: sigmoid cross-entropy (p2p / sce), 1: least squares (mae_criterion)
  float disc_scale;  // p2p: 1 (real + fake); sggan discriminator_loss: 0.5
  float* logits;     // [2B][Ho][Wo] out (may be null)
  float* loss;       // loss[0] += GAN part of the generator loss, loss[1] += discriminator loss
  sg_bf16* dst;      // dY frame of h4 for the 3B virtual images (real-D, fake-D, fake-G)
  FrameMap dmap;     // C = padded channel count
  float* dbias;      // [Cs] += over the first 2B images
  OrderedSum red;    // scratch: blocks * (2 + Cs) floats
};
void launch_disc_loss(const DiscLossParams& p, cudaStream_t st);

// L1 (+ optional gradient-sensitive) generator loss, tanh backward and the seed gradient of the
// generator: dpre = (l1_weight * sign(fake - target) / N + dD + dGrad) * (1 - fake^2).
struct FakeGradParams {
  const float* fake;    // [B][H][W][3] tanh output
  const float* target;  // L1 target (seg_A for p2p, real_A for the SG-GAN loss)
  const float* dD;      // gradient from the discriminator path, or null
  const float* dG;      // gradient of the gradient-sensitive loss, or null
  int B, H, W;
  float l1_weight;  // LAMBDA (100) or L1_lambda
  float* loss;      // loss[2] += sum |target - fake| (un-normalised)
  sg_bf16* dst;     // 8-channel dY frame of the output conv
  FrameMap dmap;
  float* dbias;  // [3]
  OrderedSum red;  // scratch: blocks * 4 floats
};
void launch_fake_grad(const FakeGradParams& p, cudaStream_t st);

// losses_out[0] = loss[0] + l1_weight * loss[2] / n_l1 + lg_weight * loss[3];  losses_out[1] = loss[1]
void launch_finalize_losses(const float* loss, float l1_weight, float n_l1, float lg_weight, float* out,
                            cudaStream_t st);

// Seg-edge weights (model.py:115-119) and gradient-sensitive loss (module.py:347-351) + its
// gradient w.r.t. `in`.  All fp32 NHWC 3-channel.
void launch_seg_edge_weight(const float* seg, int B, int H, int W, float* weight, cudaStream_t st);
void launch_gradloss(const float* in, const float* target, const float* weight, int B, int H, int W,
                     float scale, float* loss_slot, float* d_in, cudaStream_t st, OrderedSum red = OrderedSum{nullptr, nullptr});
// floats of OrderedSum scratch the four loss / seed kernels need for this problem size (the largest of them)
size_t ordered_sum_scratch_floats(int B, int H, int W, int Cs, int C_h0, int H_h0, int W_h0, int nb_bias);
// Plain criteria (module.py:336-345) as reductions: mode 0 abs, 1 squared, 2 sigmoid-CE(logits=a, labels=b)
void launch_criterion(const float* a, const float* b, int64_t n, int mode, float* out, cudaStream_t st);
// module.tf_deriv: Sobel x / y per channel, out [B][Ho][Wo][C * 2]; valid = 0: SAME zero padding, 1: VALID
void launch_sobel_deriv(const float* x, int B, int H, int W, int C, int valid, float* out, cudaStream_t st);

// Keras Adam (Appendix A.8) over a flat fp32 buffer; g is scaled by gscale first.
// step_dev != null: alpha_t is computed on the device from *step_dev (completed steps) and lr instead of being passed in.
void launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float alpha_t, float beta1, float beta2,
                 float eps, float gscale, cudaStream_t st, const long long* step_dev = nullptr, float lr = 0.f);
void launch_bump_step(long long* step_dev, cudaStream_t st);  // *step_dev += 1

// fp32 Keras-layout weights -> bf16 GEMM slabs [T][N][K]; see glue.cu for the modes.
struct PackParams {
  const float* src;
  sg_bf16* dst;
  int mode;
  int T, N, K;
  int KH, KW, Cin, Cout;
};
void launch_pack_weights(const PackParams& p, cudaStream_t st);
// all slabs of a net in ONE launch: jobs / starts (prefix sums of 256-thread blocks, njobs+1 entries) live on the device
void launch_pack_weights_batch(const PackParams* jobs, const int* starts, int njobs, int total_blocks, cudaStream_t st);
// window-wgrad scratch [pairs][128][ncol] -> Keras-layout gradient; mode 0 c1, 1 out conv, 2 h0.
void launch_unpack_wgrad(const float* scratch, float* dW, int mode, int KH, int KW, int Cin, int Cout, int ncol,
                         cudaStream_t st);
// dgamma[c] = sum_b sums[b][c][1], dbeta[c] = sum_b sums[b][c][0] over b < nb
void launch_in_param_grad(const float* sums, int nb, int C, float* dgamma, float* dbeta, cudaStream_t st);
// mask-multiply + channel sum as a standalone op (the K9 export): out[b,I,J] = sum_c x * mask
void launch_mask_reduce(const float* x, const float* mask, int B, int Hd, int Wd, int hm, int wm, int Cs,
                        float* out, cudaStream_t st);
// fp32 / tf32 operator tier (see glue.cu): fp32 frame with tf32-rounded values (Cs source channels, dmap.C >= Cs padded
// with zeros), tf32-rounded fp32 weight slabs, instance norm on fp32 storage with double-precision statistics
void launch_f32_to_frame_f32(const float* src, int B, int H, int W, int Cs, float* dst, const FrameMap& dmap, int Kp, cudaStream_t st);
void launch_pack_weights_f32(const PackParams& p, float* dst, int Kp, cudaStream_t st);
void launch_instance_norm_f32(const float* x, const float* gamma, const float* beta, const float* res, float* y, int B, int HW,
                              int C, float eps, int act, float alpha, double* stats, cudaStream_t st);
// elementwise helpers for ops.py
void launch_lrelu(const float* x, float* y, int64_t n, float leak, cudaStream_t st);
void launch_f32_to_frame(const float* src, int B, int H, int W, int C, sg_bf16* dst, const FrameMap& dmap,
                         cudaStream_t st);
void launch_bf16_to_f32(const sg_bf16* src, float* dst, int64_t n, cudaStream_t st);
void launch_f32_to_bf16(const float* src, sg_bf16* dst, int64_t n, cudaStream_t st);
// instance-norm statistics of a plain bf16 [B][HW][C] tensor (used by ops.instance_norm only; the
// step gets them from the conv epilogue)
void launch_in_stats(const sg_bf16* y, int B, int HW, int C, float* stats, cudaStream_t st);

}  // namespace sggan
