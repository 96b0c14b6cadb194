"""ctypes binding of libsggan_sm100.so (include/sggan.h) + the `Engine` convenience wrapper.

Host code stays in Python and reaches CUDA only through this thin C ABI.  Tensors cross as raw
device pointers: anything that speaks DLPack (torch, tf.experimental.dlpack capsules, cupy) is
viewed zero-copy via torch.from_dlpack and its data_ptr() handed to the library.  There is no CPU
fallback: importing works without a GPU (so the CPU test-suite can check symbols and planning),
but creating an Engine without one raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGGAN_LIB", os.path.join(_HERE, "libsggan_sm100.so"))  # SGGAN_LIB: A/B-test a variant build

NET_G, NET_D = 0, 1
LOSS_P2P, LOSS_SGGAN = 0, 1


class SgganError(RuntimeError):
    pass


class Config(C.Structure):
    """Mirror of `struct sggan_config` (include/sggan.h)."""
    _fields_ = [("batch", C.c_int), ("image_height", C.c_int), ("image_width", C.c_int), ("gf_dim", C.c_int),
                ("df_dim", C.c_int), ("segment_class", C.c_int), ("n_blocks", C.c_int), ("mask_height", C.c_int),
                ("mask_width", C.c_int), ("loss_mode", C.c_int), ("use_lsgan", C.c_int), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("adam_eps", C.c_float), ("in_eps", C.c_float),
                ("p2p_lambda", C.c_float), ("L1_lambda", C.c_float), ("Lg_lambda", C.c_float),
                ("world_size", C.c_int)]


# every symbol include/sggan.h declares: name -> (restype, argtypes)
_P, _I, _I64, _F, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
SYMBOLS = {
    "sggan_default_config": (None, [C.POINTER(Config), _I, _I, _I]),
    "sggan_disc_logit_grid": (None, [_I, _I, C.POINTER(_I), C.POINTER(_I)]),
    "sggan_workspace_bytes": (_SZ, [C.POINTER(Config)]),
    "sggan_create": (_I, [C.POINTER(Config), _P, _SZ, _P, C.POINTER(_P)]),
    "sggan_destroy": (None, [_P]),
    "sggan_last_error": (C.c_char_p, []),
    "sggan_num_tensors": (_I, [_P, _I]),
    "sggan_tensor_numel": (_I64, [_P, _I, _I]),
    "sggan_tensor_rank": (_I, [_P, _I, _I]),
    "sggan_tensor_shape": (None, [_P, _I, _I, C.POINTER(_I64)]),
    "sggan_flat_buffer": (_P, [_P, _I, _I, C.POINTER(_I64)]),
    "sggan_tensor_offset": (_I64, [_P, _I, _I]),
    "sggan_weights_changed": (_I, [_P]),
    "sggan_gen_forward": (_I, [_P, _P, _P]),
    "sggan_disc_forward": (_I, [_P, _P, _P, _P]),
    "sggan_step_forward_backward_d": (_I, [_P, _P, _P, _P, _P]),
    "sggan_step_backward_g": (_I, [_P]),
    "sggan_step_backward_g_part": (_I, [_P, _I]),
    "sggan_grad_split_offset": (_I64, [_P]),
    "sggan_step_adam": (_I, [_P, _I]),
    "sggan_step_adam_async": (_I, [_P, _I]),
    "sggan_train_step": (_I, [_P, _P, _P, _P, _P]),
    "sggan_graph_capture": (_I, [_P, _P, _P, _P, _P]),
    "sggan_graph_launch": (_I, [_P]),
    "sggan_step_count": (_I64, [_P]),
    "sggan_set_step_count": (_I, [_P, _I64]),
    "sggan_set_stream": (_I, [_P, _P]),
    "sggan_allreduce_grads": (_I, [_P, _I, _P, _P]),
    "sggan_kernel_launches": (_I, [_P]),
    "sggan_last_fake": (_P, [_P]),
    "sggan_profile_begin": (_I, [_P, _I]),
    "sggan_profile_select": (_I, [_P, _I]),
    "sggan_profile_end": (_I, [_P, C.POINTER(C.c_double), C.POINTER(_I), C.POINTER(C.c_double)]),
    "sggan_debug_buffer": (_P, [_P, _I, _I, _I, C.POINTER(_I64)]),
    "sggan_num_layers": (_I, [_P, _I]),
    "sggan_conv2d_workspace": (_SZ, [_I] * 8),
    "sggan_conv2d_fwd": (_I, [_P, _P, _P, _P] + [_I] * 8 + [_P, _SZ, _P]),
    "sggan_deconv2d_fwd": (_I, [_P, _P, _P, _P] + [_I] * 5 + [_P, _SZ, _P]),
    "sggan_conv2d_tf32_workspace": (_SZ, [_I] * 9),
    "sggan_conv2d_fwd_tf32": (_I, [_P, _P, _P, _P] + [_I] * 9 + [_P, _SZ, _P]),
    "sggan_deconv2d_fwd_tf32": (_I, [_P, _P, _P, _P] + [_I] * 6 + [_P, _SZ, _P]),
    "sggan_instance_norm_fwd_f32": (_I, [_P] * 5 + [_I] * 4 + [_F, _I, _F, _P, _SZ, _P]),
    "sggan_conv2d_bwd_workspace": (_SZ, [_I] * 8),
    "sggan_conv2d_bwd": (_I, [_P] * 6 + [_I] * 8 + [_P, _SZ, _P]),
    "sggan_deconv2d_bwd": (_I, [_P] * 6 + [_I] * 5 + [_P, _SZ, _P]),
    "sggan_instance_norm_bwd": (_I, [_P] * 7 + [_I] * 4 + [_F, _I, _F, _P, _SZ, _P]),
    "sggan_instance_norm_fwd": (_I, [_P] * 5 + [_I] * 4 + [_F, _I, _F, _P, _SZ, _P]),
    "sggan_lrelu": (_I, [_P, _P, _I64, _F, _P]),
    "sggan_mask_reduce": (_I, [_P, _P, _P] + [_I] * 6 + [_P]),
    "sggan_criterion": (_I, [_P, _P, _I64, _I, _P, _P]),
    "sggan_seg_edge_weight": (_I, [_P, _P, _I, _I, _I, _P]),
    "sggan_gradloss": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "sggan_tf_deriv": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "sggan_crc32c": (C.c_uint32, [_P, _SZ, C.c_uint32]),
    "sggan_adam_step": (_I, [_P, _P, _P, _P, _I64, _I64, _F, _F, _F, _F, _P]),
    "sggan_onehot_mask": (_I, [_P, _P] + [_I] * 6 + [_P]),
    "sggan_rgb_to_class": (_I, [_P, _P, _I64, _P]),
    "sggan_rgb_argmax_labels": (_I, [_P, _P, _I, _I, _I, _P]),
    "sggan_fast_hist": (_I, [_P, _P, _I64, _I, _P, _P]),
    "sggan_zoom_mask": (_I, [_P, _P, _P, _P, _P, _I, _I, _P] + [_I] * 6 + [_P]),
}

_lib = None


def lib():
    """Load the shared library (built by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SgganError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU / PyTorch fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise SgganError("libsggan error %d: %s" % (rc, lib().sggan_last_error().decode()))


def as_cuda_f32(x, device=None):
    """Zero-copy view of a DLPack-capable CUDA tensor; host arrays are uploaded (pinned if possible)."""
    if not isinstance(x, torch.Tensor):
        if hasattr(x, "__dlpack__"):
            x = torch.from_dlpack(x)
        else:
            x = torch.as_tensor(x)
    if x.dtype != torch.float32:
        x = x.float()
    if not x.is_cuda:
        x = x.to(device or "cuda", non_blocking=True)
    return x.contiguous()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def default_config(batch, height, width, **kw):
    cfg = Config()
    lib().sggan_default_config(C.byref(cfg), batch, height, width)
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError("sggan_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def disc_logit_grid(height, width):
    a, b = C.c_int(), C.c_int()
    lib().sggan_disc_logit_grid(height, width, C.byref(a), C.byref(b))
    return a.value, b.value


def workspace_bytes(cfg):
    n = lib().sggan_workspace_bytes(C.byref(cfg))
    if n == 0:
        raise SgganError("invalid configuration: %s" % lib().sggan_last_error().decode())
    return n


class _EngineStream:
    """Every engine call runs on the engine's own (non-default, graph-capturable) stream, ordered after whatever the
    caller enqueued on its current stream and before whatever it enqueues next -- whichever stream is current at call
    time, so `with torch.cuda.stream(...)` around a call, NCCL work handles and DLPack producers all stay correct."""

    def __init__(self, eng):
        self.eng = eng

    def __enter__(self):
        self.cur = torch.cuda.current_stream(self.eng.device)
        self.eng.stream.wait_stream(self.cur)
        return self.eng.stream

    def __exit__(self, *exc):
        self.cur.wait_stream(self.eng.stream)
        return False


class Engine:
    """One SG-GAN step engine bound to a CUDA device (launches on its own stream, see _EngineStream)."""

    def __init__(self, cfg: Config, device=None):
        if not torch.cuda.is_available():
            raise SgganError("no CUDA device: the SG-GAN step has no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device or ("cuda:%d" % torch.cuda.current_device()))
        nbytes = workspace_bytes(cfg)
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.stream = torch.cuda.Stream(self.device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            with _EngineStream(self):  # the workspace is zero-filled on the engine's stream
                check(lib().sggan_create(C.byref(cfg), C.c_void_p(self.workspace.data_ptr()), nbytes,
                                         C.c_void_p(self.stream.cuda_stream), C.byref(h)))
        self.h = h
        self._graph_key = None      # device pointers of the step graph replayed last
        self._seen_keys = set()     # input pointer sets that have run eagerly once
        self._eager_steps = 0
        self.use_graph = os.environ.get("SGGAN_GRAPH", "1") != "0"
        self._flat = {}
        self.losses = torch.zeros(2, dtype=torch.float32, device=self.device)
        Ho = max(disc_logit_grid(cfg.image_height, cfg.image_width)[0], cfg.mask_height)
        Wo = max(disc_logit_grid(cfg.image_height, cfg.image_width)[1], cfg.mask_width)
        self.out_grid = (Ho, Wo)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().sggan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------
    def num_tensors(self, net):
        return lib().sggan_num_tensors(self.h, net)

    def tensor_shape(self, net, idx):
        s = (C.c_int64 * 4)()
        lib().sggan_tensor_shape(self.h, net, idx, s)
        return tuple(s[i] for i in range(lib().sggan_tensor_rank(self.h, net, idx)))

    def flat(self, net, what):
        """Flat fp32 view (torch, zero-copy) of params(0) / grads(1) / adam m(2) / adam v(3)."""
        key = (net, what)
        if key not in self._flat:
            n = C.c_int64()
            ptr = lib().sggan_flat_buffer(self.h, net, what, C.byref(n))
            off = ptr - self.workspace.data_ptr()
            self._flat[key] = self.workspace[off:off + 4 * n.value].view(torch.float32)
        return self._flat[key]

    def tensor_view(self, net, what, idx):
        off = lib().sggan_tensor_offset(self.h, net, idx)
        n = lib().sggan_tensor_numel(self.h, net, idx)
        return self.flat(net, what)[off:off + n].view(self.tensor_shape(net, idx))

    def tensors(self, net, what=0):
        return [self.tensor_view(net, what, i) for i in range(self.num_tensors(net))]

    def set_weights(self, net, weights):
        if len(weights) != self.num_tensors(net):
            raise SgganError("expected %d tensors, got %d" % (self.num_tensors(net), len(weights)))
        for i, w in enumerate(weights):
            dst = self.tensor_view(net, 0, i)
            w = torch.as_tensor(w)
            if tuple(w.shape) != tuple(dst.shape):
                raise SgganError("tensor %d: shape %s, expected %s" % (i, tuple(w.shape), tuple(dst.shape)))
            dst.copy_(w.to(self.device, torch.float32))

    def weights_changed(self):
        with _EngineStream(self):
            check(lib().sggan_weights_changed(self.h))

    # ---- forward / step ------------------------------------------------------------------------
    def gen_forward(self, real_A):
        x = as_cuda_f32(real_A, self.device)
        out = torch.empty_like(x)
        with _EngineStream(self):
            check(lib().sggan_gen_forward(self.h, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr())))
        return out

    def disc_forward(self, x, mask):
        x, mask = as_cuda_f32(x, self.device), as_cuda_f32(mask, self.device)
        out = torch.empty((x.shape[0],) + self.out_grid + (1,), dtype=torch.float32, device=self.device)
        with _EngineStream(self):
            check(lib().sggan_disc_forward(self.h, C.c_void_p(x.data_ptr()), C.c_void_p(mask.data_ptr()),
                                           C.c_void_p(out.data_ptr())))
        return out

    def step_forward_backward_d(self, real_A, seg_A, mask):
        self._keep = (as_cuda_f32(real_A, self.device), as_cuda_f32(seg_A, self.device), as_cuda_f32(mask, self.device))
        a, s, m = self._keep
        with _EngineStream(self):
            check(lib().sggan_step_forward_backward_d(self.h, C.c_void_p(a.data_ptr()), C.c_void_p(s.data_ptr()),
                                                      C.c_void_p(m.data_ptr()), C.c_void_p(self.losses.data_ptr())))

    def step_backward_g(self, part=None):
        """part None: the whole generator backward; 0 / 1: its two halves (after part 0 the gradients from
        grad_split_offset() on are final, see include/sggan.h)."""
        with _EngineStream(self):
            check(lib().sggan_step_backward_g(self.h) if part is None else lib().sggan_step_backward_g_part(self.h, part))

    def grad_split_offset(self):
        return int(lib().sggan_grad_split_offset(self.h))

    def step_adam(self, net, overlapped=False):
        """Adam + weight re-pack of one net; overlapped=True issues it on the engine's side stream (joined by the next
        engine call), so that it runs underneath whatever is enqueued next."""
        with _EngineStream(self):
            check((lib().sggan_step_adam_async if overlapped else lib().sggan_step_adam)(self.h, net))

    def allreduce_grads(self, net, nccl_comm, stream=None):
        """In-place NCCL sum of this rank's flat gradient buffer(s) through the C ABI (sggan_allreduce_grads)."""
        with _EngineStream(self):
            check(lib().sggan_allreduce_grads(self.h, net, C.c_void_p(nccl_comm),
                                              C.c_void_p(stream if stream is not None else self.stream.cuda_stream)))

    def train_step(self, real_A, seg_A, mask):
        """One full G+D step; returns the device tensor [gen_loss, disc_loss] (no host sync).  From the second call
        with the same device buffers on, the step is ONE CUDA-graph launch (SGGAN_GRAPH=0 keeps the 250 eager launches)."""
        self._keep = (as_cuda_f32(real_A, self.device), as_cuda_f32(seg_A, self.device), as_cuda_f32(mask, self.device))
        a, s, m = self._keep
        key = (a.data_ptr(), s.data_ptr(), m.data_ptr())
        with _EngineStream(self):
            if self.use_graph and self._eager_steps >= 1 and not getattr(self, "_profiling", False):
                # a set of buffers seen before is worth a graph (the library keeps up to four: double-buffered inputs
                # alternate between two); capture is also the "select" call for a set that is already captured
                if key in self._seen_keys:
                    rc = lib().sggan_graph_capture(self.h, C.c_void_p(key[0]), C.c_void_p(key[1]), C.c_void_p(key[2]),
                                                   C.c_void_p(self.losses.data_ptr()))
                    if rc == 0:
                        self._graph_key = key
                        check(lib().sggan_graph_launch(self.h))
                        return self.losses
                    self.use_graph = False  # capture refused (still the same kernels, launched one by one)
                    self._graph_key = None
            if len(self._seen_keys) > 64:
                self._seen_keys.clear()
            self._seen_keys.add(key)
            check(lib().sggan_train_step(self.h, C.c_void_p(a.data_ptr()), C.c_void_p(s.data_ptr()),
                                         C.c_void_p(m.data_ptr()), C.c_void_p(self.losses.data_ptr())))
            self._eager_steps += 1
        return self.losses

    def last_fake(self):
        ptr = lib().sggan_last_fake(self.h)
        off = ptr - self.workspace.data_ptr()
        c = self.cfg
        n = c.batch * c.image_height * c.image_width * 3
        return self.workspace[off:off + 4 * n].view(torch.float32).view(c.batch, c.image_height, c.image_width, 3)

    def profile_begin(self, max_launches=8192, kind=0):
        """kind 0: residual-block convolutions (work = FLOPs); 1: their norm-apply passes; 2: their norm backward; 3: the
        generator-side loss kernels, one group per step (work = algorithmic bytes)."""
        check(lib().sggan_profile_select(self.h, kind))
        check(lib().sggan_profile_begin(self.h, max_launches))
        self._profiling = True  # the events are recorded by eager launches: no graph replay while profiling

    def profile_end(self):
        """-> (total ms, launches, algorithmic flops per launch) of the residual-block conv kernel."""
        ms, n, fl = C.c_double(), C.c_int(), C.c_double()
        check(lib().sggan_profile_end(self.h, C.byref(ms), C.byref(n), C.byref(fl)))
        self._profiling = False
        return ms.value, n.value, fl.value

    @property
    def kernel_launches(self):
        return lib().sggan_kernel_launches(self.h)

    # ---- debug access (tests) ---------------------------------------------------------------------
    def debug_buffer(self, net, layer, kind, nimg=None):
        """Decode an internal buffer to a dense fp32 NHWC torch tensor.
        kind 0: input frame X, 1: raw conv output Y, 2: output-gradient frame dY, 3: input gradient dX
        (padded layout, as stored), 4: forward stats [nb, C, 2]."""
        d = (C.c_int64 * 16)()
        ptr = lib().sggan_debug_buffer(self.h, net, layer, kind, d)
        if not ptr:
            return None
        frame_pix, Cc, H, W, fkind, P, pt, pl, plane, refl, nb, nbv, f32 = [d[i] for i in range(13)]
        off = ptr - self.workspace.data_ptr()
        if kind == 4:
            return self.workspace[off:off + nb * Cc * 8].view(torch.float32).view(nb, Cc, 2).clone()
        n = nimg or (nb if kind in (0, 1) else nbv)
        esz = 4 if f32 else 2
        raw = self.workspace[off:off + n * frame_pix * Cc * esz].view(torch.float32 if f32 else torch.bfloat16)
        raw = raw.view(n, frame_pix, Cc).float()
        if kind == 3:
            return raw.view(n, H, W, Cc)
        ii = torch.arange(H, device=raw.device).view(H, 1).expand(H, W)
        jj = torch.arange(W, device=raw.device).view(1, W).expand(H, W)
        if fkind == 0:
            pix = (ii + pt) * P + (jj + pl)
        else:
            pix = ((ii & 1) * 2 + (jj & 1)) * plane + ((ii >> 1) + pt) * P + ((jj >> 1) + pl)
        return raw[:, pix.reshape(-1), :].view(n, H, W, Cc)


def keras_alpha_t(lr, beta1, beta2, t):
    return lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
