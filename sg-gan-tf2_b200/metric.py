"""metric.py -- the reference's evaluation scores (metric.py:18-47,71-77) with the pixel work on the GPU.

`_fast_hist` and the argmax label adapter `scores_seg_fake` are integer kernels in libsggan_sm100 (csrc/eval.cu); the
arithmetic on the n_class x n_class confusion matrix (`scores`) is the reference's numpy, line for line.  `dense_crf` and
the `scores_*_crf` adapters need pydensecrf and stay out of scope (SURVEY 8(f) row f3 covers the histogram scores).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _i32(x):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return t.to("cuda", torch.int32).contiguous()


def _fast_hist(label_true, label_pred, n_class, out=None):
    """metric.py:18-24 -> (n_class, n_class) int64 CUDA tensor (accumulated into `out` if given)."""
    lt, lp = _i32(label_true).reshape(-1), _i32(label_pred).reshape(-1)
    if lt.numel() != lp.numel():
        raise ValueError("label arrays differ in size")
    hist = out if out is not None else torch.zeros((n_class, n_class), dtype=torch.int64, device=lt.device)
    L.check(L.lib().sggan_fast_hist(C.c_void_p(lt.data_ptr()), C.c_void_p(lp.data_ptr()), lt.numel(), n_class,
                                    C.c_void_p(hist.data_ptr()), L.stream_ptr()))
    return hist


def scores(label_trues, label_preds, n_class):
    """metric.py:27-47: overall / mean / frequency-weighted accuracy, mean IoU, per-class IoU."""
    hist_d = torch.zeros((n_class, n_class), dtype=torch.int64, device="cuda")
    for lt, lp in zip(label_trues, label_preds):
        _fast_hist(lt, lp, n_class, out=hist_d)
    hist = hist_d.cpu().numpy().astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        acc_cls = np.diag(hist) / hist.sum(axis=1)
        acc_cls = np.nanmean(acc_cls)
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
        valid = hist.sum(axis=1) > 0
        mean_iu = np.nanmean(iu[valid])
        freq = hist.sum(axis=1) / hist.sum()
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
    cls_iu = dict(zip(range(n_class), iu))
    return {"Overall Acc": acc, "Mean Acc": acc_cls, "FreqW Acc": fwavacc, "Mean IoU": mean_iu, "Class IoU": cls_iu}


def _argmax_labels(img):
    x = L.as_cuda_f32(img)
    B, H, W, ch = x.shape
    if ch != 3:
        raise L.SgganError("label adapter: 3-channel images only")
    out = torch.empty((B, W, H), dtype=torch.int32, device=x.device)
    L.check(L.lib().sggan_rgb_argmax_labels(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, H, W, L.stream_ptr()))
    return out


def scores_seg_fake(seg_image, fake_img):
    """metric.py:71-77: true labels from seg_image, predicted labels from fake_img, both argmax over the RGB channels of the
    uint8-quantised image, in the reference's (B, W, H) orientation.  Returns CUDA int32 tensors."""
    return _argmax_labels(seg_image), _argmax_labels(fake_img)
