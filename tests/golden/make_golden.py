"""Generates tests/golden/oracle_golden.npz -- frozen known-answer vectors of the CPU oracle.

The reference ships no golden vectors and cannot be imported here (TensorFlow 2.1), so these are the
oracle's own outputs on seeded inputs (fp64), frozen so that any later change to the oracle -- or a
different torch version computing something else -- is caught.  The mask fixture is a crop of a
class-id PNG shipped with the reference (datasets/city/trainA_seg_class/aachen_000000.png) when
/root/reference is present (data, not code).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sggan_oracle as O  # noqa: E402


def cases():
    g = torch.Generator().manual_seed(1234)
    r = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64) * 2 - 1  # noqa: E731
    out = {}
    x = r(2, 6, 8, 4)
    k3 = r(3, 3, 4, 5)
    b5 = r(5)
    out["x"], out["k3"], out["b5"] = x, k3, b5
    out["conv_same_s2"] = O.conv2d(x, k3, b5, 2, "SAME")
    out["conv_valid_s2"] = O.conv2d(x, k3, b5, 2, "VALID")
    out["conv_same_s1"] = O.conv2d(x, k3, b5, 1, "SAME")
    out["conv_reflect_s1"] = O.conv2d(O.reflect_pad(x, 1), k3, b5, 1, "VALID")
    kd = r(3, 3, 6, 4)
    out["kd"] = kd
    out["deconv"] = O.conv2d_transpose(x, kd, r(6) * 0, 2)
    gam, bet = r(4) + 1.5, r(4)
    out["gam"], out["bet"] = gam, bet
    out["inorm"] = O.instance_norm(x, gam, bet)
    out["inorm_eps5"] = O.instance_norm(x, gam, bet, eps=1e-5)
    out["lrelu03"] = O.lrelu(x)
    img_a, img_b = torch.rand(2, 9, 11, 3, generator=g, dtype=torch.float64), torch.rand(2, 9, 11, 3, generator=g, dtype=torch.float64)
    seg = (torch.rand(2, 9, 11, 3, generator=g, dtype=torch.float64) * 3).floor() / 3
    out["img_a"], out["img_b"], out["seg"] = img_a, img_b, seg
    w = O.seg_edge_weights(seg)
    out["edge_w"] = w
    out["tf_deriv"] = O.tf_deriv(img_a)
    out["gradloss"] = O.gradloss_criterion(img_a, img_b, w)
    logit = r(2, 3, 4, 1) * 3
    out["logit"] = logit
    out["sce_ones"] = O.sce_criterion(logit, torch.ones_like(logit))
    out["mae_ones"] = O.mae_criterion(logit, torch.ones_like(logit))
    out["gen_p2p"] = O.gen_loss_p2p(logit, img_a, img_b)
    out["disc_p2p"] = O.disc_loss_p2p(logit, -logit * 0.5)
    out["disc_lsgan"] = O.discriminator_loss(logit, -logit * 0.5, use_lsgan=True)
    # Keras Adam, 3 steps
    p, m, v = r(7), torch.zeros(7, dtype=torch.float64), torch.zeros(7, dtype=torch.float64)
    out["adam_p0"] = p.clone()
    grads = [r(7) * 0.1 for _ in range(3)]
    out["adam_grads"] = torch.stack(grads)
    for t, gr in enumerate(grads, 1):
        O.keras_adam_update(p, gr, m, v, t)
    out["adam_p3"], out["adam_m3"], out["adam_v3"] = p, m, v
    # tiny generator / discriminator forward, weights from the seeded initialiser
    gw = O.init_weights(O.generator_spec(n_blocks=1), 5, dtype=torch.float64, randomize_affine=True)
    xs = torch.rand(1, 16, 24, 3, generator=g, dtype=torch.float64)
    out["g_in"] = xs
    out["g_out"] = O.generator_resnet(xs, gw)
    dw = O.init_weights(O.discriminator_spec(segment_class=5), 6, dtype=torch.float64, randomize_affine=True)
    xd = torch.rand(1, 136, 144, 3, generator=g, dtype=torch.float64)
    hd, wd = O.disc_logit_grid(136, 144)
    mk = (torch.rand(1, hd, wd, 5, generator=g) > 0.5).double()
    out["d_in_sum"] = xd.sum()
    out["d_mask"] = mk
    out["d_out"] = O.discriminator(xd, mk, dw)
    # one full step on a 128x128 image (1 block): losses + a few gradient checksums
    gw = O.init_weights(O.generator_spec(n_blocks=1), 7, dtype=torch.float64, randomize_affine=True)
    dw = O.init_weights(O.discriminator_spec(segment_class=4), 8, dtype=torch.float64, randomize_affine=True)
    a, s, mk2, ids = O.synthetic_batch(1, 136, 136, 4, seed=3, dtype=torch.float64)
    st = O.step_grads(gw, dw, a, s, mk2)
    out["step_gen_loss"], out["step_disc_loss"] = st["gen_loss"], st["disc_loss"]
    out["step_g_gradnorms"] = torch.stack([x.norm() for x in st["g_grads"]])
    out["step_d_gradnorms"] = torch.stack([x.norm() for x in st["d_grads"]])
    out["synthetic_ids_sum"] = torch.tensor(float(ids.sum()))
    return {k: np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def mask_fixture():
    ref = "/root/reference/datasets/city/trainA_seg_class/aachen_000000.png"
    if not os.path.exists(ref):
        return {}
    from PIL import Image
    ids = np.array(Image.open(ref))[256:512:2, 512:1024:2].astype(np.uint8)  # 128 x 256 crop, ids in [0, 33]
    m = O.build_mask(ids, 256, 512, 34)  # utils.py:197-199 at 256x512 -> (8, 15, 34)
    rgb_ref = "/root/reference/datasets/gta/trainA_seg/00005.png"
    out = {"mask_ids": ids, "mask_zoom_256x512": m.astype(np.int8), "mask_onehot_sum": np.int64(O.one_hot(ids.astype(np.int64), 34).sum())}
    if os.path.exists(rgb_ref):
        rgb = np.array(Image.open(rgb_ref).convert("RGB"))[300:364, 600:728, :3].astype(np.uint8)
        out["lut_rgb"] = rgb
        out["lut_ids"] = O.rgb_to_class(rgb).astype(np.uint8)
    return out


if __name__ == "__main__":
    d = cases()
    d.update(mask_fixture())
    path = os.path.join(HERE, "oracle_golden.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes,", len(d), "arrays")
