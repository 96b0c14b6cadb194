"""CPU oracle for the SG-GAN training step -- TEST INFRASTRUCTURE ONLY.

A PyTorch-CPU restatement (fp64 or fp32) of the reference's hot path.  Nothing in the product
package imports this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may, and there only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference ships no tests, golden vectors or weights for this path, and its
arithmetic lives in un-vendored TensorFlow 2.1.0 / tensorflow-addons 0.9.1
(requirements_VP_project.txt:89-90), which cannot be installed here.  Every TF/Keras semantic
this file relies on is listed in SURVEY.md Appendix A and restated next to the function that
uses it.  The frozen outputs of this oracle (tests/golden/*.npz, made by
tests/golden/make_golden.py) are the known-answer tests.

Reference lines followed (all under /root/reference):
  module.py:208-217   residule_block          -> residule_block()
  module.py:219-269   generator_resnet        -> generator_resnet()
  module.py:272-318   discriminator           -> discriminator()
  module.py:322-351   criteria / Sobel        -> tf_kernel_prep_3d, tf_deriv, *_criterion
  model.py:106-124    seg-edge kernel + generator_loss
  model.py:126-133    discriminator_loss
  model.py:149-166    gen_loss_p2p / disc_loss_p2p
  model.py:169-200    train_step              -> train_step()
  model.py:205-207    Adam hyper-parameters   -> keras_adam_update()
  utils.py:158-165    one_hot
  utils.py:190,197-204 mask zoom + flip       -> build_mask()
  segment_class.py:60-70,95-97  RGB->class LUT -> rgb_to_class()
Layout everywhere: NHWC activations, HWIO conv kernels, (kh,kw,Cout,Cin) transposed-conv
kernels -- the Keras variable layouts, in Keras creation order.
"""
from __future__ import annotations

import math
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# basic ops with TF semantics


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def tf_same_pads(size: int, k: int, s: int):
    """TF 'SAME': out=ceil(in/s), total=max((out-1)*s+k-in,0), before=total//2 (Appendix A.2)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d(x, kernel, bias=None, stride=1, padding="VALID"):
    """tf.keras.layers.Conv2D: cross-correlation, NHWC x HWIO (+bias) (Appendix A.1)."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    if padding.upper() == "SAME":
        pt, pb = tf_same_pads(x.shape[1], kh, stride)
        pl, pr = tf_same_pads(x.shape[2], kw, stride)
        x = F.pad(x, (0, 0, pl, pr, pt, pb))
    y = F.conv2d(_nchw(x), kernel.permute(3, 2, 0, 1), bias, stride=stride)
    return _nhwc(y)


def conv2d_transpose(x, kernel, bias=None, stride=2):
    """tf.keras.layers.Conv2DTranspose(k, strides=s, padding='same'), kernel (kh,kw,Cout,Cin).

    Output is s*in; it is the input-gradient of the SAME forward conv, i.e. the full transposed
    convolution cropped by the SAME pad_before at the top/left (Appendix A.3: for k=3, s=2 the
    crop is 0 at the top/left and 1 at the bottom/right)."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    H, W = x.shape[1], x.shape[2]
    full = F.conv_transpose2d(_nchw(x), kernel.permute(3, 2, 0, 1), None, stride=stride)
    pt, _ = tf_same_pads(H * stride, kh, stride)
    pl, _ = tf_same_pads(W * stride, kw, stride)
    y = full[:, :, pt:pt + H * stride, pl:pl + W * stride]
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return _nhwc(y)


def reflect_pad(x, p):
    """tf.pad(x, [[0,0],[p,p],[p,p],[0,0]], 'REFLECT') (Appendix A.4)."""
    return _nhwc(F.pad(_nchw(x), (p, p, p, p), mode="reflect"))


def instance_norm(x, gamma, beta, eps=1e-3):
    """tfa.layers.InstanceNormalization(): per-(n,c) biased moments over H,W; eps=1e-3 (A.5).
    ops.instance_norm (ops.py:13-22) is the same maths with eps=1e-5."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    inv = torch.rsqrt(var + eps) * gamma
    return x * inv + (beta - mean * inv)


def lrelu(x, leak=0.3):
    """tf.keras.layers.LeakyReLU() default alpha=0.3 (A.6); ops.lrelu uses leak=0.2 (ops.py:36)."""
    return torch.maximum(x, leak * x)


# --------------------------------------------------------------------------------------------
# bf16 storage emulation (test infrastructure for the CUDA path's numerics, not part of the reference)
#
# The CUDA engine stores activations, activation gradients and GEMM weights as bf16 and accumulates in
# fp32.  `BF16Emu` restates exactly WHERE it rounds, as straight-through autograd functions, so that
# tests can separate "the kernels compute something else" from "bf16 storage moved a ReLU / sign()":
#   fq(x)  forward: round to bf16, backward: identity        (a stored activation, a packed weight)
#   gq(x)  forward: identity,      backward: round to bf16   (a stored gradient)
#   q(x)   both                                              (raw conv output Y <-> its gradient frame dY)
# Rounding points of the engine (csrc/engine.cu): network inputs; every conv weight; every raw conv
# output Y in front of a norm (both directions); every post-norm activation X (forward) and the
# gradient w.r.t. every PADDED conv input (dgrad output, backward); the residual-stream sum (forward)
# and its gradient sum (backward); the gradients seeded into the two output convolutions.


class _FQ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _GQ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class BF16Emu:
    fq = staticmethod(_FQ.apply)
    gq = staticmethod(_GQ.apply)

    @staticmethod
    def q(x):
        return _GQ.apply(_FQ.apply(x))


class _NoEmu:
    fq = gq = q = staticmethod(lambda x: x)


# --------------------------------------------------------------------------------------------
# weights in Keras creation order


def _glorot(gen, shape, dtype):
    rf = shape[0] * shape[1]
    fan_in, fan_out = shape[2] * rf, shape[3] * rf
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(dtype)


def generator_spec(gf_dim=64, in_c=3, out_c=3, n_blocks=9):
    """(kind, kernel_shape, has_norm) per layer of generator_resnet in creation order."""
    spec = [("conv", (7, 7, in_c, gf_dim), True), ("conv", (3, 3, gf_dim, gf_dim * 2), True),
            ("conv", (3, 3, gf_dim * 2, gf_dim * 4), True)]
    for _ in range(n_blocks):
        spec += [("conv", (3, 3, gf_dim * 4, gf_dim * 4), True), ("conv", (3, 3, gf_dim * 4, gf_dim * 4), True)]
    spec += [("deconv", (3, 3, gf_dim * 2, gf_dim * 4), True), ("deconv", (3, 3, gf_dim, gf_dim * 2), True),
             ("conv", (7, 7, gf_dim, out_c), False)]
    return spec


def discriminator_spec(df_dim=64, in_c=3, segment_class=34):
    d = df_dim
    return [("conv", (3, 3, in_c, d), False), ("conv", (3, 3, d, d * 2), True), ("conv", (3, 3, d * 2, d * 4), True),
            ("conv", (3, 3, d * 4, d * 8), True), ("conv", (3, 3, d * 8, d * 8), True),
            ("conv", (3, 3, d * 8, d * 8), True), ("conv", (3, 3, d * 8, d * 8), True),
            ("conv", (3, 3, d * 8, segment_class), False)]


def init_weights(spec, seed, dtype=torch.float32, randomize_affine=False):
    """Keras defaults: glorot-uniform kernels, zero biases, gamma=1, beta=0 ([kernel,bias],[gamma,beta]).
    randomize_affine=True perturbs biases/gamma/beta so that parity tests exercise them."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for kind, shape, has_norm in spec:
        out.append(_glorot(gen, shape, dtype))
        cout = shape[2] if kind == "deconv" else shape[3]
        if randomize_affine:
            out.append((torch.rand(cout, generator=gen, dtype=torch.float64) * 0.2 - 0.1).to(dtype))
        else:
            out.append(torch.zeros(cout, dtype=dtype))
        if has_norm:
            if randomize_affine:
                out.append((1 + torch.rand(cout, generator=gen, dtype=torch.float64) * 0.4 - 0.2).to(dtype))
                out.append((torch.rand(cout, generator=gen, dtype=torch.float64) * 0.2 - 0.1).to(dtype))
            else:
                out.append(torch.ones(cout, dtype=dtype))
                out.append(torch.zeros(cout, dtype=dtype))
    return out


# --------------------------------------------------------------------------------------------
# networks


def residule_block(x, w, ks=3, s=1, emu=None):
    """module.py:208-217.  w = [k1,b1,g1,be1,k2,b2,g2,be2].  `emu`: see BF16Emu (None = exact restatement)."""
    E = emu or _NoEmu
    p = int((ks - 1) / 2)
    y = E.gq(reflect_pad(x, p))
    y = E.q(conv2d(y, E.fq(w[0]), w[1], stride=s, padding="VALID"))
    y = instance_norm(y, w[2], w[3])
    y = E.fq(torch.relu(y))
    y = E.gq(reflect_pad(y, p))
    y = E.q(conv2d(y, E.fq(w[4]), w[5], stride=s, padding="VALID"))
    y = instance_norm(y, w[6], w[7])
    return E.q(y + x)


def generator_resnet(x, w, n_blocks=None, taps=None, emu=None):
    """module.py:219-269.  `w` is the flat Keras-order list (94 tensors for 9 blocks; the block count
    is inferred from the list length when not given).
    taps, if a dict, receives named intermediates (used by layer-level parity tests).
    `emu`: see BF16Emu (None = exact restatement)."""
    E = emu or _NoEmu
    if n_blocks is None:
        n_blocks = (len(w) - 22) // 8
    idx = [0]

    def take(n):
        r = w[idx[0]:idx[0] + n]
        idx[0] += n
        return r

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    c0 = reflect_pad(E.fq(x), 3)
    k, b, g, be = take(4)
    c1 = tap("c1", E.q(torch.relu(instance_norm(E.q(conv2d(c0, E.fq(k), b, 1, "VALID")), g, be))))
    k, b, g, be = take(4)
    c2 = tap("c2", E.q(torch.relu(instance_norm(E.q(conv2d(c1, E.fq(k), b, 2, "SAME")), g, be))))
    k, b, g, be = take(4)
    c3 = tap("c3", E.q(torch.relu(instance_norm(E.q(conv2d(c2, E.fq(k), b, 2, "SAME")), g, be))))
    r = c3
    for i in range(n_blocks):
        r = tap("r%d" % (i + 1), residule_block(r, take(8), emu=emu))
    k, b, g, be = take(4)
    d1 = tap("d1", E.q(torch.relu(instance_norm(E.q(conv2d_transpose(r, E.fq(k), b, 2)), g, be))))
    k, b, g, be = take(4)
    d2 = tap("d2", E.fq(torch.relu(instance_norm(E.q(conv2d_transpose(d1, E.fq(k), b, 2)), g, be))))
    d2 = E.gq(reflect_pad(d2, 3))
    k, b = take(2)
    pred = torch.tanh(E.gq(conv2d(d2, E.fq(k), b, 1, "VALID")))
    assert idx[0] == len(w)
    return pred


def generator_unet_spec(gf_dim=64, in_c=3, out_c=3):
    """module.py:125-206 as (kind, kernel shape, has_norm) in Keras creation order: eight 3x3 'same' convolutions, then
    eight 3x3 'same' stride-1 transposed convolutions (kernel (kh, kw, Cout, Cin)); every layer but the last has a norm."""
    g = gf_dim
    enc = [in_c, g, 2 * g, 4 * g, 8 * g, 8 * g, 8 * g, 8 * g, 8 * g]
    dec = [8 * g, 8 * g, 8 * g, 8 * g, 8 * g, 4 * g, 2 * g, g, out_c]
    spec = [("conv", (3, 3, enc[i], enc[i + 1]), True) for i in range(8)]
    spec += [("deconv", (3, 3, dec[i + 1], dec[i]), i < 7) for i in range(8)]
    return spec


def generator_unet(x, w, training=False, drop_masks=None):
    """module.py:125-206.  All layers run at the input resolution (3x3, stride 1, 'same'); LeakyReLU() is Keras' default
    alpha 0.3; the skips are ADDS (d_k + e_{8-k}); Dropout(0.5) on d1..d3 only acts in training mode (inverted dropout:
    kept values are scaled by 2) -- `drop_masks` supplies the three keep masks then, since TF's generator cannot be matched."""
    idx = [0]

    def take(n):
        r = w[idx[0]:idx[0] + n]
        idx[0] += n
        return r

    e, h = [], x
    for i in range(8):
        k, b, g, be = take(4)
        h = instance_norm(conv2d(h, k, b, 1, "SAME"), g, be)
        h = torch.relu(h) if i == 7 else lrelu(h, 0.3)
        e.append(h)
    d = e[7]
    for i in range(7):
        k, b, g, be = take(4)
        d = conv2d_transpose(d, k, b, 1)
        if i < 3 and training:
            d = d * drop_masks[i] * 2.0
        d = instance_norm(d, g, be) + e[6 - i]
        if i in (2, 6):
            d = torch.relu(d)
    k, b = take(2)
    assert idx[0] == len(w)
    return torch.tanh(conv2d_transpose(d, k, b, 1))


def discriminator(x, mask, w, taps=None, emu=None):
    """module.py:272-318.  w = flat Keras-order list (28 tensors).  Returns (B,Hd,Wd,1).
    `emu`: see BF16Emu (None = exact restatement)."""
    E = emu or _NoEmu
    idx = [0]

    def take(n):
        r = w[idx[0]:idx[0] + n]
        idx[0] += n
        return r

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    k, b = take(2)
    h = tap("h0", E.q(lrelu(E.gq(conv2d(E.fq(x), E.fq(k), b, 2, "SAME")))))
    for name, stride, pad in (("h1", 2, "SAME"), ("h2", 2, "SAME"), ("h3", 1, "SAME"), ("h31", 2, "VALID"),
                              ("h32", 2, "VALID"), ("h33", 1, "VALID")):
        k, b, g, be = take(4)
        h = tap(name, E.q(lrelu(instance_norm(E.q(conv2d(h, E.fq(k), b, stride, pad)), g, be))))
    k, b = take(2)
    h4 = tap("h4", E.gq(conv2d(h, E.fq(k), b, 1, "SAME")))
    h4 = h4 * mask  # tf.keras.layers.multiply: numpy broadcasting (A.9)
    assert idx[0] == len(w)
    return h4.sum(dim=-1, keepdim=True)


def disc_logit_grid(H, W):
    """Spatial size of the discriminator output for an HxW input (Appendix B)."""
    def same(n):
        return -(-n // 2)

    def valid(n, s):
        return (n - 3) // s + 1

    h, w = same(same(same(H))), same(same(same(W)))
    h, w = valid(h, 2), valid(w, 2)
    h, w = valid(h, 2), valid(w, 2)
    return valid(h, 1), valid(w, 1)


# --------------------------------------------------------------------------------------------
# criteria (module.py:322-351) and losses (model.py:106-166)


def tf_kernel_prep_3d(kernel, n_channels):
    """module.py:322-323."""
    return np.tile(kernel, (n_channels, 1, 1)).swapaxes(0, 1).swapaxes(1, 2)


def _depthwise(x, kernel_np, padding):
    """tf.nn.depthwise_conv2d, filter (3,3,C,mult); output channel = c*mult + m (A.7)."""
    C, mult = kernel_np.shape[2], kernel_np.shape[3]
    k = torch.as_tensor(kernel_np, dtype=x.dtype)  # (3,3,C,mult)
    wt = k.permute(2, 3, 0, 1).reshape(C * mult, 1, 3, 3)
    xi = _nchw(x)
    if padding == "SAME":
        xi = F.pad(xi, (1, 1, 1, 1))
    return _nhwc(F.conv2d(xi, wt, None, groups=C))


def tf_deriv(batch, ksize=3, padding="SAME"):
    """module.py:325-334: Sobel x / y per channel, SAME zero padding."""
    n_ch = int(batch.shape[3])
    gx = tf_kernel_prep_3d(np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]), n_ch)
    gy = tf_kernel_prep_3d(np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]]), n_ch)
    kernel = np.stack([gx, gy], axis=-1).astype(np.float32)
    return _depthwise(batch, kernel, padding)


def abs_criterion(in_, target):
    return (in_ - target).abs().mean()


def mae_criterion(in_, target):
    return ((in_ - target) ** 2).mean()


def sce_criterion(logits, labels):
    """mean sigmoid_cross_entropy_with_logits = max(x,0) - x*z + log1p(exp(-|x|))."""
    x, z = logits, labels
    return (torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs()))).mean()


def gradloss_criterion(in_, target, weight):
    """module.py:347-351."""
    abs_deriv = (tf_deriv(in_).abs() - tf_deriv(target).abs()).abs()
    abs_deriv = abs_deriv.mean(dim=-1, keepdim=True)
    return (weight * abs_deriv).mean()


def seg_edge_weights(seg_A):
    """model.py:106-119: |sign(sum_c |dx seg| + |dy seg|)| on REFLECT-padded labels."""
    n_ch = int(seg_A.shape[3])
    k0 = tf_kernel_prep_3d(np.array([[0, 0, 0], [-1, 0, 1], [0, 0, 0]]), n_ch)
    k1 = tf_kernel_prep_3d(np.array([[0, -1, 0], [0, 0, 0], [0, 1, 0]]), n_ch)
    kernel = np.stack([k0, k1], axis=-1).astype(np.float32)
    segp = reflect_pad(seg_A, 1)
    conved = _depthwise(segp, kernel, "VALID").abs()
    return torch.sign(conved.sum(dim=-1, keepdim=True)).abs()


def bce_from_logits(labels, logits):
    """tf.keras.losses.BinaryCrossentropy(from_logits=True)(y_true, y_pred): global mean (A.7)."""
    return sce_criterion(logits, labels)


def gen_loss_p2p(DA_fake, fake_A, seg_A, LAMBDA=100):
    """model.py:149-158."""
    gan_loss = bce_from_logits(torch.ones_like(DA_fake), DA_fake)
    l1_loss = (seg_A - fake_A).abs().mean()
    return gan_loss + LAMBDA * l1_loss


def disc_loss_p2p(DA_real, DA_fake):
    """model.py:160-166."""
    return bce_from_logits(torch.ones_like(DA_real), DA_real) + bce_from_logits(torch.zeros_like(DA_fake), DA_fake)


def generator_loss(DA_fake, real_A, fake_A, seg_A, L1_lambda=10.0, use_lsgan=True, Lg_lambda=0.0):
    """model.py:114-124 (defined, never called by train_step).  Returns (g_loss, weighted_seg_A).
    With Lg_lambda > 0 the gradient-sensitive term the original SG-GAN adds is included:
    Lg_lambda * gradloss_criterion(real_A, fake_A, weighted_seg_A)."""
    crit = mae_criterion if use_lsgan else sce_criterion
    weighted = seg_edge_weights(seg_A)
    g_loss = crit(DA_fake, torch.ones_like(DA_fake)) + L1_lambda * abs_criterion(real_A, fake_A)
    if Lg_lambda:
        g_loss = g_loss + Lg_lambda * gradloss_criterion(real_A, fake_A, weighted)
    return g_loss, weighted


def discriminator_loss(DA_real, DA_fake_sample, use_lsgan=True):
    """model.py:126-133."""
    crit = mae_criterion if use_lsgan else sce_criterion
    return (crit(DA_real, torch.ones_like(DA_real)) + crit(DA_fake_sample, torch.zeros_like(DA_fake_sample))) / 2


# --------------------------------------------------------------------------------------------
# optimizer + step


def keras_adam_update(p, g, m, v, t, lr=1e-3, beta1=0.5, beta2=0.999, eps=1e-7):
    """Keras OptimizerV2 Adam, non-amsgrad (Appendix A.8); t counts from 1.  In place."""
    alpha = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m += (g - m) * (1.0 - beta1)
    v += (g * g - v) * (1.0 - beta2)
    p -= alpha * m / (torch.sqrt(v) + eps)


class StepState:
    """Weights + Adam slots of both nets (what model.py:81-89,205-207 hold)."""

    def __init__(self, g_weights, d_weights, lr=1e-3, beta1=0.5):
        self.g = [w.clone() for w in g_weights]
        self.d = [w.clone() for w in d_weights]
        self.gm = [torch.zeros_like(w) for w in self.g]
        self.gv = [torch.zeros_like(w) for w in self.g]
        self.dm = [torch.zeros_like(w) for w in self.d]
        self.dv = [torch.zeros_like(w) for w in self.d]
        self.t = 0
        self.lr, self.beta1 = lr, beta1


def step_grads(g_w, d_w, real_A, seg_A, mask_A, loss_mode="p2p", L1_lambda=10.0, Lg_lambda=5.0, use_lsgan=True,
               p2p_lambda=100, emu=None):
    """Forward + both gradients of model.py:169-197 ("fresh" fake_A branch).
    Returns dict(gen_loss, disc_loss, fake_A, da_real, da_fake, g_grads, d_grads).
    `emu=BF16Emu` adds the CUDA engine's bf16 storage rounding (see BF16Emu); p2p_lambda is LAMBDA of model.py:151."""
    g_w = [w.detach().clone().requires_grad_(True) for w in g_w]
    d_w = [w.detach().clone().requires_grad_(True) for w in d_w]
    fake_A = generator_resnet(real_A, g_w, emu=emu)
    da_real = discriminator(seg_A, mask_A, d_w, emu=emu)
    da_fake = discriminator(fake_A, mask_A, d_w, emu=emu)
    da_fake_sample = da_fake  # bit-identical duplicate forward in the reference (model.py:188)
    if loss_mode == "p2p":
        gen_loss = gen_loss_p2p(da_fake, fake_A, seg_A, LAMBDA=p2p_lambda)
        disc_loss = disc_loss_p2p(da_real, da_fake_sample)
    elif loss_mode == "sggan":
        gen_loss, _ = generator_loss(da_fake, real_A, fake_A, seg_A, L1_lambda, use_lsgan, Lg_lambda)
        disc_loss = discriminator_loss(da_real, da_fake_sample, use_lsgan)
    else:
        raise ValueError(loss_mode)
    g_grads = torch.autograd.grad(gen_loss, g_w, retain_graph=True)
    d_grads = torch.autograd.grad(disc_loss, d_w)
    return dict(gen_loss=gen_loss.detach(), disc_loss=disc_loss.detach(), fake_A=fake_A.detach(),
                da_real=da_real.detach(), da_fake=da_fake.detach(), g_grads=[g.detach() for g in g_grads],
                d_grads=[g.detach() for g in d_grads])


def train_step(state: StepState, real_A, seg_A, mask_A, **kw):
    """model.py:169-200: one simultaneous G+D update.  Returns the step_grads dict."""
    out = step_grads(state.g, state.d, real_A, seg_A, mask_A, **kw)
    state.t += 1
    for p, g, m, v in zip(state.g, out["g_grads"], state.gm, state.gv):
        keras_adam_update(p, g, m, v, state.t, state.lr, state.beta1)
    for p, g, m, v in zip(state.d, out["d_grads"], state.dm, state.dv):
        keras_adam_update(p, g, m, v, state.t, state.lr, state.beta1)
    return out


# --------------------------------------------------------------------------------------------
# mask construction -- integer work, bit-exact


def cityscape_lut():
    """segment_class.py:60-70: 21 RGB triples -> 8-class ids; unknown -> 0 (defaultdict(int))."""
    lut = defaultdict(int)
    maps = [((128, 64, 128), 4), ((244, 35, 232), 4), ((250, 170, 160), 4), ((230, 150, 140), 4), ((70, 70, 70), 5),
            ((102, 102, 156), 5), ((190, 153, 153), 5), ((180, 165, 180), 5), ((150, 100, 100), 5),
            ((150, 120, 90), 5), ((107, 142, 35), 7), ((70, 130, 180), 6), ((220, 20, 60), 2), ((255, 0, 0), 2),
            ((0, 0, 142), 1), ((0, 0, 70), 1), ((0, 60, 100), 1), ((0, 0, 90), 1), ((0, 0, 110), 1), ((0, 0, 230), 3),
            ((119, 11, 32), 3)]
    for k, v in maps:
        lut[k] = v
    return lut


def rgb_to_class(img_rgb):
    """segment_class.py:87-97 (the per-pixel loop), vectorised: (H,W,>=3) uint8 -> (H,W) int64."""
    lut = cityscape_lut()
    rgb = img_rgb[..., :3].astype(np.int64)
    key = (rgb[..., 0] << 16) | (rgb[..., 1] << 8) | rgb[..., 2]
    out = np.zeros(key.shape, dtype=np.int64)
    for (r, g, b), v in lut.items():
        out[key == ((r << 16) | (g << 8) | b)] = v
    return out


def one_hot(image_in, num_classes=8):
    """utils.py:158-165 (np.int -> np.int64)."""
    hot = np.zeros((image_in.shape[0], image_in.shape[1], num_classes))
    layer_idx = np.arange(image_in.shape[0]).reshape(image_in.shape[0], 1)
    component_idx = np.tile(np.arange(image_in.shape[1]), (image_in.shape[0], 1))
    hot[layer_idx, component_idx, image_in] = 1
    return hot.astype(np.int64)


def build_mask(seg_class, image_height, image_width, num_seg_masks, flip=False):
    """utils.py:190,197-204: one-hot then scipy cubic-spline zoom to (H/34, W/34), optional fliplr."""
    import scipy.ndimage
    m = one_hot(seg_class.astype(np.int64), num_seg_masks)
    m = scipy.ndimage.zoom(m, (image_height / 34.0 / m.shape[0], image_width / 34.0 / m.shape[1], 1), mode="nearest")
    if flip:
        m = np.fliplr(m)
    return m


def nearest_mask(seg_class, hd, wd, num_classes):
    """Documented deviation used for synthetic data: one-hot ids nearest-resampled to the D-logit
    grid: out[i,j] = onehot(seg[floor((i+0.5)*H/hd), floor((j+0.5)*W/wd)])."""
    H, W = seg_class.shape
    ii = np.minimum(((np.arange(hd) + 0.5) * H / hd).astype(np.int64), H - 1)
    jj = np.minimum(((np.arange(wd) + 0.5) * W / wd).astype(np.int64), W - 1)
    ids = seg_class[np.ix_(ii, jj)]
    return (ids[..., None] == np.arange(num_classes)).astype(np.int64)


def synthetic_batch(B, H, W, C, seed=19, dtype=torch.float32):
    """SURVEY 8(d) synthetic inputs: real_A, seg_A ~ U[0,1); piecewise-constant class-id map;
    mask on the D-logit grid."""
    rng = np.random.RandomState(seed)
    real_A = rng.rand(B, H, W, 3).astype(np.float32)
    seg_A = rng.rand(B, H, W, 3).astype(np.float32)
    ids = np.zeros((B, H, W), dtype=np.int64)
    for b in range(B):
        for _ in range(12):
            y0, x0 = rng.randint(0, H), rng.randint(0, W)
            y1, x1 = rng.randint(y0, H) + 1, rng.randint(x0, W) + 1
            ids[b, y0:y1, x0:x1] = rng.randint(0, C)
    hd, wd = disc_logit_grid(H, W)
    mask = np.stack([nearest_mask(ids[b], hd, wd, C) for b in range(B)]).astype(np.float32)
    return (torch.as_tensor(real_A).to(dtype), torch.as_tensor(seg_A).to(dtype), torch.as_tensor(mask).to(dtype), ids)


# --------------------------------------------------------------------------------------------
# evaluation scores -- integer work, bit-exact (metric.py:18-47,71-77)


def fast_hist(label_true, label_pred, n_class):
    """metric.py:18-24."""
    label_true, label_pred = np.asarray(label_true).reshape(-1), np.asarray(label_pred).reshape(-1)
    keep = (label_true >= 0) & (label_true < n_class)
    return np.bincount(n_class * label_true[keep].astype(int) + label_pred[keep], minlength=n_class ** 2).reshape(n_class, n_class)


def seg_fake_labels(seg_image, fake_img):
    """metric.py:71-77: argmax over the channels of the uint8-quantised images, in the (0,3,2,1) orientation."""
    a = np.argmax((255 * np.asarray(seg_image)).astype(np.uint8).transpose(0, 3, 2, 1), axis=1)
    b = np.argmax((255 * np.asarray(fake_img)).astype(np.uint8).transpose(0, 3, 2, 1), axis=1)
    return a, b
