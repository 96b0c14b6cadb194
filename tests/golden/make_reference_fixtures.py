"""Fixtures taken from files the REFERENCE ITSELF ships (run in the build container, where /root/reference exists).

    python tests/golden/make_reference_fixtures.py        # writes tests/golden/reference_fixtures.npz

What is pinned here, and to what:
  * gta_rgb / gta_class -- every 4th pixel (both axes) of datasets/gta/trainA_seg/00005.png (palette PNG ->
    RGB) paired with the SAME pixels of datasets/gta/trainA_seg_class/00005.png.  The class PNG was written by
    the reference's own segment_class.py (lines 87-97), so this is a golden vector of the reference for the
    RGB -> class-id LUT (segment_class.py:60-70), not an output of our oracle.
  * city_ids_* -- full-resolution Cityscapes labelId maps (datasets/city/trainA_seg_class/aachen_*.png, ids 0..33)
    and city_mask_*_{256x512,512x1024}: utils.py:190,197-199 evaluated by calling numpy + scipy.ndimage.zoom
    exactly as the reference does (one_hot -> zoom(..., (H/34/h, W/34/w, 1), mode="nearest"), order 3).  scipy is
    the third-party code the reference calls for this step; the installed version is recorded (the reference pins
    1.4.1, whose spline boundary handling differs from >= 1.6).
"""
import os

import numpy as np
import scipy
import scipy.ndimage
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/datasets"


def one_hot(image_in, num_classes):  # utils.py:158-165 verbatim semantics (np.int -> np.int64)
    hot = np.zeros((image_in.shape[0], image_in.shape[1], num_classes))
    hot[np.arange(image_in.shape[0])[:, None], np.arange(image_in.shape[1])[None, :], image_in] = 1
    return hot.astype(np.int64)


def main():
    out = {"scipy_version": np.array(scipy.__version__)}
    rgb = np.array(Image.open(os.path.join(REF, "gta/trainA_seg/00005.png")).convert("RGB"))
    cls = np.array(Image.open(os.path.join(REF, "gta/trainA_seg_class/00005.png")))
    assert rgb.shape[:2] == cls.shape
    out["gta_rgb"] = np.ascontiguousarray(rgb[::4, ::4]).astype(np.uint8)
    out["gta_class"] = np.ascontiguousarray(cls[::4, ::4]).astype(np.uint8)
    names = ["aachen_000000", "aachen_000005", "aachen_000017", "aachen_000042"]
    kept = []
    for n in names:
        p = os.path.join(REF, "city/trainA_seg_class", n + ".png")
        if not os.path.exists(p):
            continue
        ids = np.array(Image.open(p)).astype(np.uint8)
        kept.append(n)
        out["city_ids_" + n] = ids
        hot = one_hot(ids.astype(np.int64), 34)
        for (H, W) in ((256, 512), (512, 1024)):
            m = scipy.ndimage.zoom(hot, (H / 34.0 / hot.shape[0], W / 34.0 / hot.shape[1], 1), mode="nearest")
            out["city_mask_%s_%dx%d" % (n, H, W)] = m.astype(np.int8)
    out["city_names"] = np.array(kept)
    path = os.path.join(HERE, "reference_fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
