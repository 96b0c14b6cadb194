"""utils.py -- the loader's MASK CONSTRUCTION on the GPU (reference utils.py:158-165,190,197-204; segment_class.py:60-70).

Only the part of the reference's utils.py that feeds the training step's `mask_A` lives here: RGB -> class id (LUT),
one_hot, the scipy cubic-spline zoom to (H/34, W/34) and the flip.  PNG decoding, skimage `resize` and imgaug
augmentation (utils.py:27-156,167-189) stay host-side and out of scope.

`seg_mask` reproduces `scipy.ndimage.zoom(one_hot(ids), (H/34/h, W/34/w, 1), mode="nearest")` (order 3) exactly --
integer output equal on every shipped Cityscapes label map we tried (tests/golden/reference_fixtures.npz) -- without
ever building the (h, w, C) one-hot volume: spline prefilter + cubic B-spline evaluation are linear and separable, so
per-axis weight rows (a few dozen non-negligible taps each) are derived once per size on the host and one small
kernel sums them over the uint8 id map.
"""
from __future__ import annotations

import ctypes as C
import functools
import math

import numpy as np
import torch

from . import _lib as L

_Z = math.sqrt(3.0) - 2.0   # pole of the cubic B-spline prefilter
_NPAD = 12                  # scipy pre-pads by 12 edge samples for mode="nearest" (ndimage._prepad_for_spline_filter)


def _prefilter_columns(a):
    """Cubic B-spline prefilter with half-sample-symmetric ends (what scipy's spline_filter1d applies for
    mode="nearest") along axis 0 of a 2-D fp64 array, every column at once."""
    c = a.astype(np.float64).copy()
    n = c.shape[0]
    c *= (1.0 - _Z) * (1.0 - 1.0 / _Z)
    zn = math.pow(_Z, n)
    zi = _Z ** np.arange(1, n)
    c0 = c[0] + zn * c[n - 1] + (zi[:, None] * (c[1:] + zn * c[n - 2::-1][: n - 1])).sum(0)
    c[0] = c0 * (_Z / (1 - zn * zn)) + c[0]
    for i in range(1, n):
        c[i] += _Z * c[i - 1]
    c[n - 1] *= _Z / (_Z - 1)
    for i in range(n - 2, -1, -1):
        c[i] = _Z * (c[i + 1] - c[i])
    return c


@functools.lru_cache(maxsize=32)
def zoom_weights(n_in, n_out):
    """(n_out, n_in) fp64 matrix M with zoom(x, n_out / n_in, order=3, mode="nearest") == M @ x along one axis:
    edge-pad by 12, prefilter, evaluate the cubic B-spline at o * (n_in - 1) / (n_out - 1) (scipy's grid_mode=False)."""
    N = n_in + 2 * _NPAD
    E = np.zeros((N, n_in))
    E[np.arange(N), np.clip(np.arange(N) - _NPAD, 0, n_in - 1)] = 1.0
    P = _prefilter_columns(E)
    M = np.zeros((n_out, n_in))
    zoom = (n_in - 1) / (n_out - 1) if n_out > 1 else 1.0
    for o in range(n_out):
        cc = o * zoom + _NPAD
        fl = math.floor(cc)
        y = cc - fl
        z = 1.0 - y
        w = [z * z * z / 6.0, (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0, (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0]
        w.append(1.0 - w[0] - w[1] - w[2])
        for k in range(4):
            M[o] += w[k] * P[min(max(fl - 1 + k, 0), N - 1)]
    return M


def _windows(M, tol=1e-18):
    """Per output row: the input range that carries every weight above tol, padded to a common width."""
    n_out, n_in = M.shape
    lo = np.empty(n_out, dtype=np.int64)
    hi = np.empty(n_out, dtype=np.int64)
    for o in range(n_out):
        nz = np.nonzero(np.abs(M[o]) > tol)[0]
        lo[o], hi[o] = (nz[0], nz[-1] + 1) if len(nz) else (0, 1)
    width = int((hi - lo).max())
    start = np.minimum(lo, n_in - width).clip(min=0).astype(np.int32)
    win = np.stack([M[o, start[o]:start[o] + width] for o in range(n_out)])
    return np.ascontiguousarray(win), start, width


def one_hot(image_in, num_classes=8):
    """utils.py:158-165 on the device: (..., H, W) integer ids -> (..., H, W, num_classes) int64."""
    ids = torch.as_tensor(np.asarray(image_in)) if not isinstance(image_in, torch.Tensor) else image_in
    return (ids.to("cuda").long().unsqueeze(-1) == torch.arange(num_classes, device="cuda")).long()


def seg_mask(seg_class, image_height, image_width, num_seg_masks, flip=False):
    """utils.py:190,197-204: `zoom(one_hot(seg_class, C), (H/34/h, W/34/w, 1), mode="nearest")` (+ optional fliplr).
    seg_class: (h, w) or (B, h, w) class ids (uint8 range) -> float32 CUDA tensor (B, round(H/34), round(W/34), C)
    holding the integers scipy returns."""
    ids = torch.as_tensor(np.asarray(seg_class)) if not isinstance(seg_class, torch.Tensor) else seg_class
    if ids.dim() == 2:
        ids = ids.unsqueeze(0)
    ids = ids.to("cuda", torch.uint8).contiguous()
    B, h, w = ids.shape
    ho, wo = int(round(h * (image_height / 34.0 / h))), int(round(w * (image_width / 34.0 / w)))
    if ho < 1 or wo < 1:
        raise L.SgganError("seg_mask: image smaller than one 34-pixel cell")
    wy, y0, wh = _windows(zoom_weights(h, ho))
    wx, x0, ww = _windows(zoom_weights(w, wo))
    dev = ids.device
    t_wy, t_wx = torch.as_tensor(wy, device=dev), torch.as_tensor(wx, device=dev)
    t_y0, t_x0 = torch.as_tensor(y0, device=dev), torch.as_tensor(x0, device=dev)
    mask = torch.empty((B, ho, wo, num_seg_masks), dtype=torch.float32, device=dev)
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    L.check(L.lib().sggan_zoom_mask(P(ids), P(t_wy), P(t_y0), P(t_wx), P(t_x0), wh, ww, P(mask), B, h, w, ho, wo,
                                    num_seg_masks, L.stream_ptr()))
    return torch.flip(mask, dims=[2]) if flip else mask


def rgb_to_class(img_rgb):
    """segment_class.py:60-70,87-97 on the device: (..., 3) uint8 colours -> (...) uint8 class ids (unknown -> 0)."""
    rgb = torch.as_tensor(np.asarray(img_rgb)) if not isinstance(img_rgb, torch.Tensor) else img_rgb
    rgb = rgb[..., :3].to("cuda", torch.uint8).contiguous()
    out = torch.empty(rgb.shape[:-1], dtype=torch.uint8, device=rgb.device)
    L.check(L.lib().sggan_rgb_to_class(C.c_void_p(rgb.data_ptr()), C.c_void_p(out.data_ptr()), out.numel(), L.stream_ptr()))
    return out


class ImagePool(object):
    """utils.py:27-53: the history buffer of generated images, same call contract and the same use of numpy's global
    random state (np.random.rand twice per swap, so a seeded run replays the reference's choices).

    `image` is the 4-element list the reference passes; elements may be numpy arrays or CUDA tensors -- swapped-out
    entries are returned as they were stored, nothing is copied to the host.  The reference's trainer never calls the pool
    on the path that runs (model.py:169-200 feeds the fresh fake_A); it exists for callers that do."""

    def __init__(self, maxsize=50):
        self.maxsize = maxsize
        self.num_img = 0
        self.images = []

    def __call__(self, image):
        if self.maxsize <= 0:
            return image
        if self.num_img < self.maxsize:
            self.images.append(image)
            self.num_img += 1
            return image
        if np.random.rand() > 0.5:
            idx = int(np.random.rand() * self.maxsize)
            tmp1, tmp3 = self.images[idx][0], self.images[idx][2]
            self.images[idx][0], self.images[idx][2] = image[0], image[2]
            idx = int(np.random.rand() * self.maxsize)
            tmp2, tmp4 = self.images[idx][1], self.images[idx][3]
            self.images[idx][1], self.images[idx][3] = image[1], image[3]
            return [tmp1, tmp2, tmp3, tmp4]
        return image
