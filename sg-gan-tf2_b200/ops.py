"""ops.py -- the reference's op wrappers (ops.py:10-49) over the sm_100a kernels.

The reference file is dead TF1 code (tf.variable_scope / slim); its *names and
signatures* are part of the surface, so they are re-created here: `conv2d`, `deconv2d`
(bias-free, truncated-normal(stddev) kernels, per the commented slim calls at ops.py:24-34),
`instance_norm` (eps 1e-5, scale ~ N(1, 0.02), offset 0, ops.py:13-22) and `lrelu` (leak 0.2).
Variables are created once per `name` (the tf.variable_scope behaviour) and kept in VARIABLES.
`batch_norm` and `linear` are never used anywhere in the reference and are out of scope.
"""
from __future__ import annotations

import ctypes as C
import zlib

import torch

from . import _lib as L

VARIABLES = {}  # name -> list of torch tensors, like a TF1 variable scope
_PAD = {"VALID": 0, "SAME": 1, "REFLECT": 2}
_ACT = {None: 0, "none": 0, "relu": 1, "lrelu": 2, "tanh": 3}
_PASSES = {"tf32": 1, "tf32x3": 3, "fp32": 3}  # precision tiers of the fp32-storage operators (include/sggan.h)
_ws = {}


def _workspace(nbytes, device):
    """Scratch for one operator call, keyed by (device, stream): calls on different streams never share a buffer."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    t = _ws.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _ws[key] = t
    return t


def _seed(name):
    """Deterministic per-name seed (Python's str hash is salted per process: replicas would start from different weights)."""
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def _trunc_normal(shape, stddev, seed):
    g = torch.Generator().manual_seed(seed)
    w = torch.empty(shape)
    torch.nn.init.trunc_normal_(w, 0.0, stddev, -2 * stddev, 2 * stddev, generator=g)
    return w


def conv2d_raw(x, kernel, bias=None, stride=1, padding="SAME", precision="bf16"):
    """Conv2D forward: x NHWC, kernel HWIO; stride 1 (odd k) or 2 (k=3).  precision "bf16" (the training path's storage;
    Cin and Cout multiples of 64) | "tf32" (fp32 storage, one tf32 product per term) | "tf32x3" / "fp32" (split operands,
    fp32-level accuracy); the fp32-storage tiers take any channel count."""
    x = L.as_cuda_f32(x)
    kernel = L.as_cuda_f32(kernel, x.device)
    bias = None if bias is None else L.as_cuda_f32(bias, x.device)
    B, H, W, Cin = x.shape
    k, Cout = kernel.shape[0], kernel.shape[3]
    pad = _PAD[padding.upper()]
    if pad == 0:
        Ho, Wo = (H - k) // stride + 1, (W - k) // stride + 1
    elif pad == 1:
        Ho, Wo = -(-H // stride), -(-W // stride)
    else:
        Ho, Wo = H, W
    if precision != "bf16":
        passes = _PASSES[precision]
        nbytes = L.lib().sggan_conv2d_tf32_workspace(B, H, W, Cin, Cout, k, stride, pad, passes)
        if nbytes == 0:
            raise L.SgganError("conv2d: unsupported shape (k %d, stride %d, %s)" % (k, stride, padding))
        y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
        ws = _workspace(nbytes, x.device)
        L.check(L.lib().sggan_conv2d_fwd_tf32(C.c_void_p(x.data_ptr()), C.c_void_p(kernel.data_ptr()),
                                              C.c_void_p(bias.data_ptr() if bias is not None else None),
                                              C.c_void_p(y.data_ptr()), B, H, W, Cin, Cout, k, stride, pad, passes,
                                              C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))
        return y
    nbytes = L.lib().sggan_conv2d_workspace(B, H, W, Cin, Cout, k, stride, pad)
    if nbytes == 0:
        raise L.SgganError("conv2d: unsupported shape (Cin %d, Cout %d, k %d, stride %d, %s)" % (Cin, Cout, k, stride, padding))
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
    ws = _workspace(nbytes, x.device)
    L.check(L.lib().sggan_conv2d_fwd(C.c_void_p(x.data_ptr()), C.c_void_p(kernel.data_ptr()),
                                     C.c_void_p(bias.data_ptr() if bias is not None else None),
                                     C.c_void_p(y.data_ptr()), B, H, W, Cin, Cout, k, stride, pad,
                                     C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))
    return y


def deconv2d_raw(x, kernel, bias=None, precision="bf16"):
    """Conv2DTranspose(3, strides 2, 'same') forward: kernel (kh, kw, Cout, Cin); precision as in conv2d_raw."""
    x = L.as_cuda_f32(x)
    kernel = L.as_cuda_f32(kernel, x.device)
    bias = None if bias is None else L.as_cuda_f32(bias, x.device)
    B, H, W, Cin = x.shape
    if kernel.shape[0] != 3 or kernel.shape[3] != Cin:
        raise L.SgganError("deconv2d: kernel must be (3, 3, Cout, Cin)")
    Cout = kernel.shape[2]
    if precision != "bf16":
        passes = _PASSES[precision]
        nbytes = L.lib().sggan_conv2d_tf32_workspace(B, H, W, Cin, Cout, 3, -2, 1, passes)
        if nbytes == 0:
            raise L.SgganError("deconv2d: unsupported shape")
        y = torch.empty((B, 2 * H, 2 * W, Cout), dtype=torch.float32, device=x.device)
        ws = _workspace(nbytes, x.device)
        L.check(L.lib().sggan_deconv2d_fwd_tf32(C.c_void_p(x.data_ptr()), C.c_void_p(kernel.data_ptr()),
                                                C.c_void_p(bias.data_ptr() if bias is not None else None),
                                                C.c_void_p(y.data_ptr()), B, H, W, Cin, Cout, passes,
                                                C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))
        return y
    nbytes = L.lib().sggan_conv2d_workspace(B, H, W, Cin, Cout, 3, -2, 1)
    if nbytes == 0:
        raise L.SgganError("deconv2d: unsupported shape")
    y = torch.empty((B, 2 * H, 2 * W, Cout), dtype=torch.float32, device=x.device)
    ws = _workspace(nbytes, x.device)
    L.check(L.lib().sggan_deconv2d_fwd(C.c_void_p(x.data_ptr()), C.c_void_p(kernel.data_ptr()),
                                       C.c_void_p(bias.data_ptr() if bias is not None else None),
                                       C.c_void_p(y.data_ptr()), B, H, W, Cin, Cout, C.c_void_p(ws.data_ptr()),
                                       ws.numel(), L.stream_ptr()))
    return y


def conv2d_bwd_raw(x, kernel, dy, stride=1, padding="SAME", transposed=False):
    """Gradients (dx, dw, db) of conv2d_raw / deconv2d_raw w.r.t. input, kernel and bias for an upstream dy."""
    x, dy = L.as_cuda_f32(x), L.as_cuda_f32(dy)
    kernel = L.as_cuda_f32(kernel, x.device)
    B, H, W, Cin = x.shape
    k = kernel.shape[0]
    Cout = kernel.shape[2] if transposed else kernel.shape[3]
    pad = _PAD[padding.upper()]
    nbytes = L.lib().sggan_conv2d_bwd_workspace(B, H, W, Cin, Cout, k, -2 if transposed else stride, pad)
    if nbytes == 0:
        raise L.SgganError("conv2d_bwd: unsupported shape")
    dx = torch.empty_like(x)
    dw = torch.empty(tuple(kernel.shape), dtype=torch.float32, device=x.device)
    db = torch.empty(Cout, dtype=torch.float32, device=x.device)
    ws = _workspace(nbytes, x.device)
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    if transposed:
        L.check(L.lib().sggan_deconv2d_bwd(P(x), P(kernel), P(dy), P(dx), P(dw), P(db), B, H, W, Cin, Cout, P(ws), ws.numel(),
                                           L.stream_ptr()))
    else:
        L.check(L.lib().sggan_conv2d_bwd(P(x), P(kernel), P(dy), P(dx), P(dw), P(db), B, H, W, Cin, Cout, k, stride, pad, P(ws),
                                         ws.numel(), L.stream_ptr()))
    return dx, dw, db


def instance_norm_bwd_raw(x, gamma, beta, dz, eps=1e-3, act=None, alpha=0.3):
    """Gradients (dx, dgamma, dbeta) of act(instance_norm(x)) for an upstream dz."""
    x, dz = L.as_cuda_f32(x), L.as_cuda_f32(dz)
    B, H, W, Cc = x.shape
    gamma, beta = L.as_cuda_f32(gamma, x.device), L.as_cuda_f32(beta, x.device)
    dx = torch.empty_like(x)
    dg = torch.empty(Cc, dtype=torch.float32, device=x.device)
    dbt = torch.empty(Cc, dtype=torch.float32, device=x.device)
    ws = _workspace(x.numel() * 6 + 2 * B * Cc * 8 + 160 * Cc * 8 + 16384, x.device)
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    L.check(L.lib().sggan_instance_norm_bwd(P(x), P(gamma), P(beta), P(dz), P(dx), P(dg), P(dbt), B, H, W, Cc, eps, _ACT[act],
                                            alpha, P(ws), ws.numel(), L.stream_ptr()))
    return dx, dg, dbt


def instance_norm_raw(x, gamma, beta, eps=1e-3, act=None, alpha=0.3, residual=None, precision="bf16"):
    """InstanceNormalization (+ activation, + residual).  precision "bf16": the training path's row-stream kernel (the
    input is rounded to bf16 storage, C a multiple of 64, <= 512); anything else: fp32 storage, double-precision statistics."""
    x = L.as_cuda_f32(x)
    B, H, W, Cc = x.shape
    gamma, beta = L.as_cuda_f32(gamma, x.device), L.as_cuda_f32(beta, x.device)
    residual = None if residual is None else L.as_cuda_f32(residual, x.device)
    y = torch.empty_like(x)
    if precision != "bf16":
        ws = _workspace(B * Cc * 16 + 256, x.device)
        L.check(L.lib().sggan_instance_norm_fwd_f32(C.c_void_p(x.data_ptr()), C.c_void_p(gamma.data_ptr()),
                                                    C.c_void_p(beta.data_ptr()),
                                                    C.c_void_p(residual.data_ptr() if residual is not None else None),
                                                    C.c_void_p(y.data_ptr()), B, H, W, Cc, eps, _ACT[act], alpha,
                                                    C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))
        return y
    ws = _workspace(x.numel() * 6 + B * Cc * 8 + 4096, x.device)
    L.check(L.lib().sggan_instance_norm_fwd(C.c_void_p(x.data_ptr()), C.c_void_p(gamma.data_ptr()),
                                            C.c_void_p(beta.data_ptr()),
                                            C.c_void_p(residual.data_ptr() if residual is not None else None),
                                            C.c_void_p(y.data_ptr()), B, H, W, Cc, eps, _ACT[act], alpha,
                                            C.c_void_p(ws.data_ptr()), ws.numel(), L.stream_ptr()))
    return y


# ---- the reference's signatures --------------------------------------------------------------------------

def conv2d(input_, output_dim, ks=4, s=2, stddev=0.02, padding="SAME", name="conv2d"):
    """ops.py:24-28 (slim.conv2d, activation_fn=None, biases_initializer=None)."""
    x = L.as_cuda_f32(input_)
    if name not in VARIABLES:
        VARIABLES[name] = [_trunc_normal((ks, ks, x.shape[-1], output_dim), stddev, _seed(name))]
    return conv2d_raw(x, VARIABLES[name][0], None, stride=s, padding=padding)


def deconv2d(input_, output_dim, ks=4, s=2, stddev=0.02, name="deconv2d"):
    """ops.py:30-34 (slim.conv2d_transpose, 'SAME').  Kernel layout (kh, kw, Cout, Cin); k=3, s=2 only."""
    x = L.as_cuda_f32(input_)
    if ks != 3 or s != 2:
        raise L.SgganError("deconv2d: only ks=3, s=2 (the generator's transposed convolutions) is implemented")
    if name not in VARIABLES:
        VARIABLES[name] = [_trunc_normal((ks, ks, output_dim, x.shape[-1]), stddev, _seed(name))]
    return deconv2d_raw(x, VARIABLES[name][0], None)


def instance_norm(input, name="instance_norm"):
    """ops.py:13-22: moments over axes [1,2], epsilon 1e-5, scale ~ N(1, 0.02), offset 0."""
    x = L.as_cuda_f32(input)
    depth = x.shape[3]
    if name not in VARIABLES:
        g = torch.Generator().manual_seed(_seed(name))
        VARIABLES[name] = [1.0 + 0.02 * torch.randn(depth, generator=g), torch.zeros(depth)]
    scale, offset = VARIABLES[name]
    return instance_norm_raw(x, scale, offset, eps=1e-5)


def lrelu(x, leak=0.2, name="lrelu"):
    """ops.py:36-37: tf.maximum(x, leak*x)."""
    x = L.as_cuda_f32(x)
    y = torch.empty_like(x)
    L.check(L.lib().sggan_lrelu(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), x.numel(), leak, L.stream_ptr()))
    return y


def mask_reduce(h4, mask):
    """multiply([h4, mask]) + reduce_sum(axis=-1, keepdims=True) (module.py:312-314)."""
    h4, mask = L.as_cuda_f32(h4), L.as_cuda_f32(mask)
    B, Hd, Wd, Cs = h4.shape
    hm, wm = mask.shape[1], mask.shape[2]
    out = torch.empty((B, max(Hd, hm), max(Wd, wm), 1), dtype=torch.float32, device=h4.device)
    L.check(L.lib().sggan_mask_reduce(C.c_void_p(h4.data_ptr()), C.c_void_p(mask.data_ptr()), C.c_void_p(out.data_ptr()),
                                      B, Hd, Wd, hm, wm, Cs, L.stream_ptr()))
    return out
