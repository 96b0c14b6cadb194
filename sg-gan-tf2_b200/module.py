"""module.py -- drop-in for the reference's module.py on the hot path (module.py:208-351).

Same names, argument meaning and defaults as the reference; the arithmetic runs in
libsggan_sm100.so (hand-written sm_100a kernels) through the ctypes binding in _lib.py.
`generator_resnet()` / `discriminator()` return callables with a Keras-like surface
(`model(x)`, `model([x, mask])`, `.trainable_variables` in Keras creation order,
`.save_weights / .load_weights`).  The builders take the sizes the reference hard-codes
(module.py:221,225,274-277) as keyword arguments defaulting to the reference's constants
(SURVEY D3); activations are fully convolutional, so a model is re-planned for whatever NHWC
shape it is called with.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from . import _lib as L

__all__ = ["generator_resnet", "generator_unet", "discriminator", "residule_block", "tf_kernel_prep_3d", "tf_deriv", "abs_criterion",
           "mae_criterion", "sce_criterion", "gradloss_criterion"]

_SEED = 19  # main.py:4 tf.random.set_seed(19)


def _glorot(gen, shape):
    rf = shape[0] * shape[1]
    lim = math.sqrt(6.0 / (shape[2] * rf + shape[3] * rf))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * lim


class Runtime:
    """One libsggan engine (generator + discriminator planned together for a fixed NHWC shape)."""
    _cache = {}

    def __init__(self, batch, height, width, segment_class=34, n_blocks=9, mask_hw=None, **cfg_kw):
        kw = dict(segment_class=segment_class, n_blocks=n_blocks)
        if mask_hw is not None:
            kw.update(mask_height=int(mask_hw[0]), mask_width=int(mask_hw[1]))
        kw.update(cfg_kw)
        self.cfg = L.default_config(batch, height, width, **kw)
        self.engine = L.Engine(self.cfg)
        self.bound = {L.NET_G: None, L.NET_D: None}

    @classmethod
    def get(cls, batch, height, width, **kw):
        """Cached plan for stand-alone forward calls; one at a time (workspaces are GBs).  A runtime that owns a
        network's weights stays alive through that network's reference even after it leaves the cache."""
        key = (batch, height, width, tuple(sorted(kw.items())))
        if key not in cls._cache:
            cls._cache.clear()
            cls._cache[key] = cls(batch, height, width, **kw)
        return cls._cache[key]


class _Net:
    """Keras-Model-like holder of one network's variables."""
    net_id = None

    def __init__(self, shapes, seed):
        gen = torch.Generator().manual_seed(seed)
        self._vars = []
        for s in shapes:
            if len(s) == 4:
                self._vars.append(_glorot(gen, s))
            else:
                self._vars.append(None)  # filled below: bias / beta zeros, gamma ones
        self._shapes = shapes
        self.runtime = None

    @property
    def trainable_variables(self):
        return self._vars

    def bind(self, runtime):
        """Make `runtime` the OWNER of this network's master weights: the variables move into its flat parameter
        buffer and become views of it.  If another runtime owned them (a re-plan for a new batch / image size), the
        Adam slots and the step count are carried over, so the optimizer continues instead of silently restarting."""
        if self.runtime is runtime:
            return
        old = self.runtime
        eng = runtime.engine
        eng.set_weights(self.net_id, self._vars)
        if old is not None:
            for what in (2, 3):  # Adam m, v
                eng.flat(self.net_id, what).copy_(old.engine.flat(self.net_id, what))
            L.check(L.lib().sggan_set_step_count(eng.h, max(L.lib().sggan_step_count(eng.h),
                                                             L.lib().sggan_step_count(old.engine.h))))
        self._vars = eng.tensors(self.net_id, 0)
        self.runtime = runtime
        runtime.bound[self.net_id] = self
        eng.weights_changed()
        self._restore_optimizer()

    def _infer_runtime(self, key_kw, batch, height, width):
        """A runtime for an off-plan forward call (e.g. sampling one image in the middle of training,
        model.py:528-532): it gets a COPY of the current master weights on every call; ownership stays with the
        training runtime, so training, save_weights and later samples keep seeing the trained weights."""
        rt = Runtime.get(batch, height, width, **key_kw)
        if rt is self.runtime:
            return rt
        rt.engine.set_weights(self.net_id, self._vars)
        rt.engine.weights_changed()
        return rt

    def get_weights(self):
        return [v.detach().cpu().numpy().copy() for v in self._vars]

    def set_weights(self, weights):
        if len(weights) != len(self._vars):
            raise ValueError("expected %d arrays, got %d" % (len(self._vars), len(weights)))
        for v, w in zip(self._vars, weights):
            w = torch.as_tensor(np.asarray(w), dtype=torch.float32)
            if tuple(w.shape) != tuple(v.shape):
                raise ValueError("shape mismatch %s vs %s" % (tuple(w.shape), tuple(v.shape)))
            v.copy_(w.to(v.device))
        if self.runtime is not None:
            self.runtime.engine.weights_changed()

    # Keras' Model.save_weights / load_weights (model.py:463-466,499-500): a path without an .npz / .h5 suffix is a TF
    # checkpoint prefix -> `<path>.index` + `<path>.data-00000-of-00001` + the directory's `checkpoint` state file, in
    # the tensor-bundle format with Keras' object-graph variable names (tf_checkpoint.py); `.npz` keeps the plain list.
    def layer_kinds(self):
        raise NotImplementedError

    def save_weights(self, path, optimizer=False):
        """optimizer=True also writes `<path>.opt.npz` with this network's Adam slots and the step count (the reference
        saves weights only, so a resumed run there restarts Adam; here the state can travel)."""
        if path.endswith(".npz"):
            np.savez(path, *self.get_weights())
        else:
            from . import tf_checkpoint
            tf_checkpoint.save(path, self.get_weights(), self.layer_kinds())
        if optimizer:
            if self.runtime is None:
                raise L.SgganError("save_weights(optimizer=True): the network has not been planned yet (no Adam state)")
            eng = self.runtime.engine
            np.savez(path + ".opt.npz", m=eng.flat(self.net_id, 2).cpu().numpy(), v=eng.flat(self.net_id, 3).cpu().numpy(),
                     step=np.int64(L.lib().sggan_step_count(eng.h)))

    def load_weights(self, path):
        if path.endswith(".npz") or (not os.path.exists(path + ".index") and os.path.exists(path + ".npz")):
            z = np.load(path if path.endswith(".npz") else path + ".npz")
            self.set_weights([z["arr_%d" % i] for i in range(len(z.files))])
        else:
            from . import tf_checkpoint
            self.set_weights(tf_checkpoint.load(path, self.layer_kinds()))
        self._pending_opt = path + ".opt.npz" if os.path.exists(path + ".opt.npz") else None
        self._restore_optimizer()

    def _restore_optimizer(self):
        """Adam slots + step count from the sidecar of the last load_weights, as soon as an engine owns the weights."""
        p = getattr(self, "_pending_opt", None)
        if p is None or self.runtime is None:
            return
        z = np.load(p)
        eng = self.runtime.engine
        if z["m"].size != eng.flat(self.net_id, 2).numel():
            raise L.SgganError("optimizer state %s does not match this network" % p)
        eng.flat(self.net_id, 2).copy_(torch.as_tensor(z["m"]))
        eng.flat(self.net_id, 3).copy_(torch.as_tensor(z["v"]))
        L.check(L.lib().sggan_set_step_count(eng.h, int(z["step"])))
        self._pending_opt = None


class GeneratorResnet(_Net):
    """generator_resnet (module.py:219-269): c7s1-64, d128, d256, 9 x R256, u128, u64, c7s1-3 + tanh."""
    net_id = L.NET_G

    def __init__(self, image_height=64, image_width=64, gf_dim=64, output_c_dim=3, n_blocks=9, seed=_SEED):
        if gf_dim != 64 or output_c_dim != 3:
            raise L.SgganError("generator_resnet: gf_dim=64, output_c_dim=3 only (module.py:221-222)")
        g = gf_dim
        shapes = [(7, 7, 3, g), (g,), (g,), (g,), (3, 3, g, 2 * g), (2 * g,), (2 * g,), (2 * g,),
                  (3, 3, 2 * g, 4 * g), (4 * g,), (4 * g,), (4 * g,)]
        for _ in range(2 * n_blocks):
            shapes += [(3, 3, 4 * g, 4 * g), (4 * g,), (4 * g,), (4 * g,)]
        shapes += [(3, 3, 2 * g, 4 * g), (2 * g,), (2 * g,), (2 * g,), (3, 3, g, 2 * g), (g,), (g,), (g,),
                   (7, 7, g, 3), (3,)]
        super().__init__(shapes, seed)
        # [kernel, bias, gamma, beta] per conv+norm; Keras defaults zeros / ones / zeros
        i = 0
        while i < len(shapes):
            self._vars[i + 1] = torch.zeros(shapes[i + 1])
            if i + 2 < len(shapes) and len(shapes[i + 2]) == 1:
                self._vars[i + 2] = torch.ones(shapes[i + 2])
                self._vars[i + 3] = torch.zeros(shapes[i + 3])
                i += 4
            else:
                i += 2
        self.n_blocks = n_blocks
        self.input_hw = (image_height, image_width)

    def layer_kinds(self):
        return ["conv", "norm"] * 3 + ["conv", "norm", "conv", "norm"] * self.n_blocks + ["deconv", "norm"] * 2 + ["conv"]

    def forward_fp32(self, x, precision="tf32x3"):
        """The same network on the fp32-storage operator tier (ops.conv2d_raw / deconv2d_raw / instance_norm_raw with
        precision "tf32" or "tf32x3"): what the reference computes in fp32 (module.py:219-269), for accuracy checks of
        the bf16 training path and for inference that must match the reference to 1e-4.  Any image size >= 8 px."""
        from . import ops
        v = [t if t.is_cuda else t.cuda() for t in self._vars]
        x = L.as_cuda_f32(x)

        def cna(h, i, act, stride=1, padding="SAME", deconv=False, residual=None):
            k, b, g, be = v[i:i + 4]
            h = ops.deconv2d_raw(h, k, b, precision=precision) if deconv else \
                ops.conv2d_raw(h, k, b, stride=stride, padding=padding, precision=precision)
            return ops.instance_norm_raw(h, g, be, eps=1e-3, act=act, residual=residual, precision=precision)

        h = cna(x, 0, "relu", padding="REFLECT")              # c7s1-64 on the 3-pixel reflect pad
        h = cna(h, 4, "relu", stride=2)
        h = cna(h, 8, "relu", stride=2)
        i = 12
        for _ in range(self.n_blocks):                        # residule_block (module.py:208-217)
            y = cna(h, i, "relu", padding="REFLECT")
            h = cna(y, i + 4, None, padding="REFLECT", residual=h)
            i += 8
        h = cna(h, i, "relu", deconv=True)
        h = cna(h, i + 4, "relu", deconv=True)
        k, b = v[i + 8], v[i + 9]
        return torch.tanh(ops.conv2d_raw(h, k, b, stride=1, padding="REFLECT", precision=precision))

    def __call__(self, x):
        x = L.as_cuda_f32(x)
        B, H, W, _ = x.shape
        rt = self.runtime
        if rt is None or (rt.cfg.batch, rt.cfg.image_height, rt.cfg.image_width) != (B, H, W):
            kw = dict(n_blocks=self.n_blocks)
            if H < 128 or W < 128:
                # the discriminator stack needs >= 128 px (Appendix B); plan it on a dummy grid instead
                raise L.SgganError("generator_resnet: images smaller than 128x128 are not supported by the joint plan")
            if rt is None:
                rt = Runtime.get(B, H, W, **kw)
                self.bind(rt)
            else:
                rt = self._infer_runtime(kw, B, H, W)
        return rt.engine.gen_forward(x)


class Discriminator(_Net):
    """discriminator (module.py:272-318): 8 convs, 6 instance norms, LeakyReLU(0.3), logits x mask, sum over C."""
    net_id = L.NET_D

    def __init__(self, image_height=128, image_width=128, df_dim=64, segment_class=34, seed=_SEED + 1):
        if df_dim != 64:
            raise L.SgganError("discriminator: df_dim=64 only (module.py:274)")
        d = df_dim
        shapes = [(3, 3, 3, d), (d,)]
        for cin, cout in ((d, 2 * d), (2 * d, 4 * d), (4 * d, 8 * d), (8 * d, 8 * d), (8 * d, 8 * d), (8 * d, 8 * d)):
            shapes += [(3, 3, cin, cout), (cout,), (cout,), (cout,)]
        shapes += [(3, 3, 8 * d, segment_class), (segment_class,)]
        super().__init__(shapes, seed)
        self._vars[1] = torch.zeros(shapes[1])
        for i in range(2, len(shapes) - 2, 4):
            self._vars[i + 1] = torch.zeros(shapes[i + 1])
            self._vars[i + 2] = torch.ones(shapes[i + 2])
            self._vars[i + 3] = torch.zeros(shapes[i + 3])
        self._vars[-1] = torch.zeros(shapes[-1])
        self.segment_class = segment_class

    def layer_kinds(self):
        return ["conv"] + ["conv", "norm"] * 6 + ["conv"]

    def __call__(self, inputs):
        x, mask = inputs
        x, mask = L.as_cuda_f32(x), L.as_cuda_f32(mask)
        B, H, W, _ = x.shape
        rt = self.runtime
        want = (B, H, W, int(mask.shape[1]), int(mask.shape[2]))
        if rt is None or (rt.cfg.batch, rt.cfg.image_height, rt.cfg.image_width, rt.cfg.mask_height,
                          rt.cfg.mask_width) != want:
            kw = dict(segment_class=self.segment_class, mask_hw=(int(mask.shape[1]), int(mask.shape[2])))
            if rt is None:
                rt = Runtime.get(B, H, W, **kw)
                self.bind(rt)
            else:
                rt = self._infer_runtime(kw, B, H, W)
        return rt.engine.disc_forward(x, mask)


class GeneratorUnet(_Net):
    """generator_unet (module.py:125-206), the reference CLI's default generator: eight 3x3 'same' convolutions + instance
    norm + LeakyReLU(0.3) at the input resolution, eight 3x3 'same' stride-1 transposed convolutions + instance norm with
    additive skips, Dropout(0.5) on the first three in training mode, tanh.

    FORWARD ONLY, on the operator tier (ops.conv2d_raw / instance_norm_raw; precision "bf16" = the training path's tcgen05
    kernels with bf16 storage, "tf32" / "tf32x3" = fp32 storage): sampling / testing with this generator works, training it
    does not (the fused step engine is the ResNet generator's, which is what BASELINE names).  A stride-1 'same' transposed
    convolution IS a convolution with the kernel flipped and its channel axes exchanged, so it runs on the same kernels."""
    net_id = None

    def __init__(self, gf_dim=64, output_c_dim=3, seed=_SEED):
        g = gf_dim
        enc = [3, g, 2 * g, 4 * g, 8 * g, 8 * g, 8 * g, 8 * g, 8 * g]
        dec = [8 * g, 8 * g, 8 * g, 8 * g, 8 * g, 4 * g, 2 * g, g, output_c_dim]
        shapes = []
        for i in range(8):
            shapes += [(3, 3, enc[i], enc[i + 1]), (enc[i + 1],), (enc[i + 1],), (enc[i + 1],)]
        for i in range(8):
            shapes += [(3, 3, dec[i + 1], dec[i]), (dec[i + 1],)]
            if i < 7:
                shapes += [(dec[i + 1],), (dec[i + 1],)]
        super().__init__(shapes, seed)
        i = 0
        while i < len(shapes):
            self._vars[i + 1] = torch.zeros(shapes[i + 1])
            if i + 2 < len(shapes) and len(shapes[i + 2]) == 1:
                self._vars[i + 2] = torch.ones(shapes[i + 2])
                self._vars[i + 3] = torch.zeros(shapes[i + 3])
                i += 4
            else:
                i += 2

    def layer_kinds(self):
        return ["conv", "norm"] * 8 + ["deconv", "norm"] * 7 + ["deconv"]

    def bind(self, runtime):
        raise L.SgganError("generator_unet runs on the operator tier only (no fused training engine)")

    def __call__(self, x, training=False, precision="bf16", drop_masks=None):
        from . import ops
        x = L.as_cuda_f32(x)
        v = [t if t.is_cuda else t.to(x.device) for t in self._vars]

        def prec(cin, cout):  # the bf16 tier takes channel counts in multiples of 64; the 3-channel ends go to the tf32 tier
            return precision if (precision != "bf16" or (cin % 64 == 0 and cout % 64 == 0)) else "tf32"

        e, h, i = [], x, 0
        for li in range(8):
            k, b, g, be = v[i:i + 4]
            i += 4
            h = ops.conv2d_raw(h, k, b, stride=1, padding="SAME", precision=prec(k.shape[2], k.shape[3]))
            h = ops.instance_norm_raw(h, g, be, eps=1e-3, act="relu" if li == 7 else "lrelu", alpha=0.3,
                                      precision=precision if h.shape[3] % 64 == 0 and h.shape[3] <= 512 else "fp32")
            e.append(h)
        d = e[7]
        for li in range(8):
            k, b = v[i], v[i + 1]
            # Conv2DTranspose(3x3, stride 1, 'same'), kernel (kh, kw, Cout, Cin)  ==  Conv2D with w'[kh, kw, ci, co] = w[2-kh, 2-kw, co, ci]
            kc = k.flip(0, 1).permute(0, 1, 3, 2).contiguous()
            d = ops.conv2d_raw(d, kc, b, stride=1, padding="SAME", precision=prec(kc.shape[2], kc.shape[3]))
            if li == 7:
                return torch.tanh(d)
            g, be = v[i + 2], v[i + 3]
            i += 4
            if li < 3 and training:  # Dropout(0.5), inverted: kept values are doubled
                m = drop_masks[li].to(d.device) if drop_masks is not None else (torch.rand_like(d) >= 0.5).float()
                d = d * m * 2.0
            d = ops.instance_norm_raw(d, g, be, eps=1e-3, act=None, residual=e[6 - li], precision=precision)
            if li in (2, 6):
                d = torch.relu(d)
        raise AssertionError("unreachable")


def generator_unet(gf_dim=64, output_c_dim=3):
    print("generator_unet")
    return GeneratorUnet(gf_dim, output_c_dim)


def generator_resnet(image_height=64, image_width=64, gf_dim=64, output_c_dim=3, n_blocks=9):
    print("generator_resnet")
    return GeneratorResnet(image_height, image_width, gf_dim, output_c_dim, n_blocks)


def discriminator(image_height=128, image_width=128, df_dim=64, segment_class=34):
    print("discriminator")
    return Discriminator(image_height, image_width, df_dim, segment_class)


def residule_block(x, dim, ks=3, s=1, weights=None):
    """module.py:208-217 as a standalone op: reflect-pad -> conv -> IN -> relu -> reflect-pad -> conv -> IN, + x.
    `weights` = [k1, b1, g1, be1, k2, b2, g2, be2] (Keras order); created glorot/zeros/ones if omitted."""
    from . import ops
    x = L.as_cuda_f32(x)
    if s != 1 or ks % 2 == 0:
        raise L.SgganError("residule_block: stride 1 and odd kernel only")
    cin = x.shape[-1]
    if weights is None:
        gen = torch.Generator().manual_seed(_SEED)
        weights = [_glorot(gen, (ks, ks, cin, dim)), torch.zeros(dim), torch.ones(dim), torch.zeros(dim),
                   _glorot(gen, (ks, ks, dim, dim)), torch.zeros(dim), torch.ones(dim), torch.zeros(dim)]
    y = ops.conv2d_raw(x, weights[0], weights[1], stride=1, padding="REFLECT")
    y = ops.instance_norm_raw(y, weights[2], weights[3], eps=1e-3, act="relu")
    y = ops.conv2d_raw(y, weights[4], weights[5], stride=1, padding="REFLECT")
    return ops.instance_norm_raw(y, weights[6], weights[7], eps=1e-3, act=None, residual=x)


# ---- criteria (module.py:322-351) ---------------------------------------------------------------------

def tf_kernel_prep_3d(kernel, n_channels):
    return np.tile(kernel, (n_channels, 1, 1)).swapaxes(0, 1).swapaxes(1, 2)


def tf_deriv(batch, ksize=3, padding="SAME"):
    """Sobel x / y per channel (module.py:325-334); returned as (B,H,W,2*C), channel = c*2 + {x,y} (depthwise_conv2d's
    channel order).  libsggan's sobel_deriv_kernel; the training step uses the fused gradloss kernel instead."""
    import ctypes as C
    x = L.as_cuda_f32(batch)
    if ksize != 3 or padding.upper() not in ("SAME", "VALID"):
        raise L.SgganError("tf_deriv: ksize 3, padding SAME or VALID (the reference's kernel is hard-coded 3x3)")
    B, H, W, n_ch = x.shape
    valid = padding.upper() == "VALID"
    out = torch.empty((B, H - 2 * valid, W - 2 * valid, 2 * n_ch), dtype=torch.float32, device=x.device)
    L.check(L.lib().sggan_tf_deriv(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, H, W, n_ch, int(valid), L.stream_ptr()))
    return out


def _criterion(a, b, mode):
    import ctypes as C
    a, b = L.as_cuda_f32(a), L.as_cuda_f32(b)
    if a.shape != b.shape:
        b = b.expand_as(a).contiguous()
    out = torch.zeros(1, dtype=torch.float32, device=a.device)
    L.check(L.lib().sggan_criterion(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), a.numel(), mode,
                                    C.c_void_p(out.data_ptr()), L.stream_ptr()))
    return out[0]


def abs_criterion(in_, target):
    return _criterion(in_, target, 0)


def mae_criterion(in_, target):
    return _criterion(in_, target, 1)


def sce_criterion(logits, labels):
    return _criterion(logits, labels, 2)


def gradloss_criterion(in_, target, weight, return_grad=False):
    import ctypes as C
    a, b, w = L.as_cuda_f32(in_), L.as_cuda_f32(target), L.as_cuda_f32(weight)
    B, H, W, ch = a.shape
    if ch != 3:
        raise L.SgganError("gradloss_criterion: 3-channel images only")
    out = torch.zeros(1, dtype=torch.float32, device=a.device)
    d_in = torch.empty_like(a) if return_grad else None
    L.check(L.lib().sggan_gradloss(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(w.data_ptr()),
                                   C.c_void_p(out.data_ptr()), C.c_void_p(d_in.data_ptr() if return_grad else None),
                                   B, H, W, L.stream_ptr()))
    return (out[0], d_in) if return_grad else out[0]
