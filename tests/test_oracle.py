"""The oracle against its frozen known-answer vectors and against independent formulations of the
TF/Keras semantics it restates.  CPU only."""
import math
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def t(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def close(a, b, tol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= tol * (1 + np.max(np.abs(b)))


def test_golden_regenerates_bit_for_bit(golden):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    fresh = make_golden.cases()
    for k, v in fresh.items():
        close(v, golden[k], 1e-10)


def test_conv_variants(O, golden):
    x, k3, b5 = t(golden["x"]), t(golden["k3"]), t(golden["b5"])
    close(O.conv2d(x, k3, b5, 2, "SAME"), golden["conv_same_s2"])
    close(O.conv2d(x, k3, b5, 2, "VALID"), golden["conv_valid_s2"])
    close(O.conv2d(x, k3, b5, 1, "SAME"), golden["conv_same_s1"])
    close(O.conv2d(O.reflect_pad(x, 1), k3, b5, 1, "VALID"), golden["conv_reflect_s1"])
    # A.2: TF SAME for k=3, s=2, even size pads (0 before, 1 after) -- not symmetric
    assert O.tf_same_pads(8, 3, 2) == (0, 1) and O.tf_same_pads(8, 3, 1) == (1, 1) and O.tf_same_pads(7, 3, 2) == (1, 1)
    manual = torch.nn.functional.conv2d(torch.nn.functional.pad(x.permute(0, 3, 1, 2), (0, 1, 0, 1)),
                                        k3.permute(3, 2, 0, 1), b5, stride=2).permute(0, 2, 3, 1)
    close(manual, golden["conv_same_s2"])
    sym = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), k3.permute(3, 2, 0, 1), b5, stride=2, padding=1).permute(0, 2, 3, 1)
    assert (sym - t(golden["conv_same_s2"])).abs().max() > 1e-2  # the usual torch idiom is NOT TF SAME


def test_deconv_is_gradient_of_same_conv(O, golden):
    # A.3: Conv2DTranspose(3, s=2, 'same') == input-gradient of the SAME stride-2 conv with the same kernel
    x, kd = t(golden["x"]), t(golden["kd"])
    close(O.conv2d_transpose(x, kd, None, 2), golden["deconv"])
    z = torch.zeros(2, 12, 16, 6, dtype=torch.float64, requires_grad=True)
    y = O.conv2d(z, kd.permute(0, 1, 2, 3), None, 2, "SAME")  # kernel (kh,kw,Cin=6,Cout=4) as HWIO
    (gz,) = torch.autograd.grad(y, z, x)
    close(gz, golden["deconv"], 1e-12)
    # out[2i] = x[i] w0 + x[i-1] w2 ; out[2i+1] = x[i] w1 along each axis (1-D check on a delta input)
    d = torch.zeros(1, 3, 3, 4, dtype=torch.float64)
    d[0, 1, 1, 0] = 1
    o = O.conv2d_transpose(d, kd, None, 2)
    close(o[0, 2, 2], kd[0, 0, :, 0])
    close(o[0, 3, 3], kd[1, 1, :, 0])
    close(o[0, 4, 4], kd[2, 2, :, 0])


def test_instance_norm(O, golden):
    x, g, b = t(golden["x"]), t(golden["gam"]), t(golden["bet"])
    close(O.instance_norm(x, g, b), golden["inorm"])
    close(O.instance_norm(x, g, b, eps=1e-5), golden["inorm_eps5"])
    ref = torch.nn.functional.group_norm(x.permute(0, 3, 1, 2), 4, g, b, eps=1e-3).permute(0, 2, 3, 1)
    close(ref, golden["inorm"], 1e-9)
    one = t(np.random.RandomState(0).rand(2, 1, 1, 4))
    close(O.instance_norm(one, g, b), b.view(1, 1, 1, 4).expand(2, 1, 1, 4), 1e-12)  # H*W == 1 -> exactly beta
    close(O.lrelu(x), golden["lrelu03"])
    close(O.lrelu(x, 0.2), np.maximum(golden["x"], 0.2 * golden["x"]))


def test_criteria_and_losses(O, golden):
    a, b, seg = t(golden["img_a"]), t(golden["img_b"]), t(golden["seg"])
    w = O.seg_edge_weights(seg)
    close(w, golden["edge_w"])
    assert set(np.unique(golden["edge_w"])) <= {0.0, 1.0}
    close(O.tf_deriv(a), golden["tf_deriv"])
    close(O.gradloss_criterion(a, b, w), golden["gradloss"])
    # independent Sobel: channel c*2+0 = x-derivative, c*2+1 = y-derivative, zero SAME padding
    pad = np.pad(golden["img_a"], ((0, 0), (1, 1), (1, 1), (0, 0)))
    gx = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=np.float64)
    man = np.zeros_like(golden["tf_deriv"])
    for c in range(3):
        for i in range(9):
            for j in range(11):
                win = pad[:, i:i + 3, j:j + 3, c]
                man[:, i, j, 2 * c] = (win * gx).sum(axis=(1, 2))
                man[:, i, j, 2 * c + 1] = (win * gx.T).sum(axis=(1, 2))
    close(man, golden["tf_deriv"], 1e-12)
    lg = t(golden["logit"])
    close(O.sce_criterion(lg, torch.ones_like(lg)), golden["sce_ones"])
    close(torch.nn.functional.binary_cross_entropy_with_logits(lg, torch.ones_like(lg)), golden["sce_ones"], 1e-12)
    close(O.mae_criterion(lg, torch.ones_like(lg)), golden["mae_ones"])
    close(O.gen_loss_p2p(lg, a, b), golden["gen_p2p"])
    close(O.disc_loss_p2p(lg, -lg * 0.5), golden["disc_p2p"])
    close(O.discriminator_loss(lg, -lg * 0.5, use_lsgan=True), golden["disc_lsgan"])
    close(golden["gen_p2p"], golden["sce_ones"] + 100 * np.abs(golden["img_b"] - golden["img_a"]).mean(), 1e-12)


def test_keras_adam(O, golden):
    p, m, v = t(golden["adam_p0"]).clone(), torch.zeros(7, dtype=torch.float64), torch.zeros(7, dtype=torch.float64)
    for step, g in enumerate(golden["adam_grads"], 1):
        O.keras_adam_update(p, t(g), m, v, step)
    close(p, golden["adam_p3"])
    close(m, golden["adam_m3"])
    close(v, golden["adam_v3"])
    # A.8: epsilon sits OUTSIDE the bias correction; torch.optim.Adam differs measurably at beta1 = 0.5
    q = torch.nn.Parameter(t(golden["adam_p0"]).clone())
    opt = torch.optim.Adam([q], lr=1e-3, betas=(0.5, 0.999), eps=1e-7)
    for g in golden["adam_grads"]:
        q.grad = t(g)
        opt.step()
    assert (q.detach() - t(golden["adam_p3"])).abs().max() > 0  # not identical ...
    assert (q.detach() - t(golden["adam_p3"])).abs().max() < 1e-5  # ... but the same algorithm
    # closed form of the first step: theta -= lr*sqrt(1-b2)/(1-b1) * (1-b1) g / (sqrt((1-b2) g^2) + eps)
    g0 = golden["adam_grads"][0]
    p1 = golden["adam_p0"] - 1e-3 * math.sqrt(1 - 0.999) / 0.5 * (0.5 * g0) / (np.sqrt(0.001 * g0 * g0) + 1e-7)
    pp, mm, vv = t(golden["adam_p0"]).clone(), torch.zeros(7, dtype=torch.float64), torch.zeros(7, dtype=torch.float64)
    O.keras_adam_update(pp, t(g0), mm, vv, 1)
    close(pp, p1, 1e-12)


def test_networks_and_step(O, golden):
    gw = O.init_weights(O.generator_spec(n_blocks=1), 5, dtype=torch.float64, randomize_affine=True)
    close(O.generator_resnet(t(golden["g_in"]), gw), golden["g_out"])
    assert np.abs(golden["g_out"]).max() <= 1.0  # tanh
    assert len(O.init_weights(O.generator_spec(), 1)) == 94 and len(O.init_weights(O.discriminator_spec(), 1)) == 28
    assert sum(w.numel() for w in O.init_weights(O.generator_spec(), 1)) == 11388675
    assert sum(w.numel() for w in O.init_weights(O.discriminator_spec(), 1)) == 8791970
    assert O.disc_logit_grid(256, 512) == (5, 13) and O.disc_logit_grid(512, 1024) == (13, 29)
    assert O.disc_logit_grid(128, 128) == (1, 1)
    gw = O.init_weights(O.generator_spec(n_blocks=1), 7, dtype=torch.float64, randomize_affine=True)
    dw = O.init_weights(O.discriminator_spec(segment_class=4), 8, dtype=torch.float64, randomize_affine=True)
    a, s, mk, ids = O.synthetic_batch(1, 136, 136, 4, seed=3, dtype=torch.float64)
    assert float(ids.sum()) == float(golden["synthetic_ids_sum"])
    st = O.step_grads(gw, dw, a, s, mk)
    close(st["gen_loss"], golden["step_gen_loss"], 1e-10)
    close(st["disc_loss"], golden["step_disc_loss"], 1e-10)
    close(torch.stack([g.norm() for g in st["g_grads"]]), golden["step_g_gradnorms"], 1e-8)
    close(torch.stack([g.norm() for g in st["d_grads"]]), golden["step_d_gradnorms"], 1e-8)
    # biases in front of an instance norm get (numerically) zero gradient
    assert golden["step_g_gradnorms"][1] < 1e-10 * golden["step_g_gradnorms"][0]


def test_discriminator_broadcast_128(O):
    # the reference-consistent case: 128x128 -> 1x1 logits against the loader's 4x4 mask
    torch.manual_seed(3)
    dw = O.init_weights(O.discriminator_spec(segment_class=3), 2, randomize_affine=True)
    x = torch.rand(2, 128, 128, 3)
    mask = (torch.rand(2, 4, 4, 3) > 0.5).float()
    out = O.discriminator(x, mask, dw)
    assert tuple(out.shape) == (2, 4, 4, 1)
    out2 = O.discriminator(torch.rand(2, 128, 128, 3), mask, dw)
    assert torch.allclose(out, out2, atol=1e-6)  # h33 has H*W == 1: the output does not depend on the image


def test_mask_construction_integer(O, golden):
    if "mask_ids" not in golden.files:
        pytest.skip("fixture made without /root/reference")
    ids = golden["mask_ids"]
    hot = O.one_hot(ids.astype(np.int64), 34)
    assert hot.dtype == np.int64 and hot.shape == (128, 256, 34) and int(hot.sum()) == int(golden["mask_onehot_sum"])
    assert np.array_equal(hot.argmax(-1), ids) and (hot.sum(-1) == 1).all()
    m = O.build_mask(ids, 256, 512, 34)
    assert m.shape == (8, 15, 34) and np.array_equal(m.astype(np.int8), golden["mask_zoom_256x512"])
    assert np.array_equal(np.fliplr(m), O.build_mask(ids, 256, 512, 34, flip=True))
    assert np.array_equal(O.rgb_to_class(golden["lut_rgb"]).astype(np.uint8), golden["lut_ids"])
    lut = O.cityscape_lut()
    assert len(lut) == 21 and lut[(1, 2, 3)] == 0 and lut[(107, 142, 35)] == 7
    nm = O.nearest_mask(ids.astype(np.int64), 5, 13, 34)
    assert nm.shape == (5, 13, 34) and (nm.sum(-1) == 1).all()


# ---------------------------------------------------------------------------------------------- reference-shipped fixtures
@pytest.fixture(scope="module")
def ref_fix():
    return np.load(os.path.join(HERE, "golden", "reference_fixtures.npz"))


def test_lut_against_the_class_png_the_reference_ships(O, ref_fix):
    """datasets/gta/trainA_seg_class/00005.png was written by the reference's own segment_class.py (lines 87-97) from
    datasets/gta/trainA_seg/00005.png: a golden vector OF THE REFERENCE for the RGB -> class-id LUT (every 4th pixel of
    both files, tests/golden/make_reference_fixtures.py).  This pins the oracle's restatement, not just itself."""
    rgb, cls = ref_fix["gta_rgb"], ref_fix["gta_class"]
    got = O.rgb_to_class(rgb)
    assert got.shape == cls.shape and np.array_equal(got.astype(np.uint8), cls)
    assert set(np.unique(cls)) == {0, 1, 2, 4, 5, 6, 7}  # the classes this image exercises


def test_mask_zoom_restatement_equals_scipy_on_shipped_label_maps(O, ref_fix):
    """utils.py:190,197-199 on full-resolution Cityscapes label maps the reference ships: the oracle (which calls
    scipy.ndimage.zoom exactly as the reference does) against the committed outputs, and the product's separable
    restatement of the same spline (sg-gan-tf2_b200/utils.py zoom_weights) evaluated with numpy -- integer for integer."""
    import importlib
    U = importlib.import_module("sg-gan-tf2_b200.utils")
    for name in ref_fix["city_names"]:
        ids = ref_fix["city_ids_" + str(name)].astype(np.int64)
        for (H, W) in ((256, 512), (512, 1024)):
            ref = ref_fix["city_mask_%s_%dx%d" % (name, H, W)]
            if str(name) == str(ref_fix["city_names"][0]) and (H, W) == (256, 512):
                assert np.array_equal(O.build_mask(ids, H, W, 34), ref)  # scipy here == scipy when the fixture was made
            Wy, Wx = U.zoom_weights(ids.shape[0], ref.shape[0]), U.zoom_weights(ids.shape[1], ref.shape[1])
            hot = (ids[..., None] == np.arange(34)).astype(np.float64)
            t = np.einsum("jx,ixc->ijc", Wx, np.einsum("iy,yxc->ixc", Wy, hot))
            got = np.where(t > 0, t + 0.5, t - 0.5).astype(np.int64)
            assert np.array_equal(got, ref), (name, H, W, int((got != ref).sum()))
            wy, y0, wh = U._windows(Wy)
            assert wh <= 96 and np.abs(Wy).sum(1).max() < 2.0  # the kernel's window carries the whole row of weights
            for o in range(Wy.shape[0]):
                full = np.zeros(Wy.shape[1])
                full[y0[o]:y0[o] + wh] = wy[o]
                assert np.abs(full - Wy[o]).max() < 1e-17


def test_stride1_transposed_conv_is_a_flipped_conv(O):
    """generator_unet's decoder (module.py:171-203) is built from Conv2DTranspose(3x3, stride 1, 'same').  Independent
    formulation: that is Conv2D with w'[kh, kw, ci, co] = w[2-kh, 2-kw, co, ci] -- the identity module.GeneratorUnet relies on
    to run the decoder on the convolution kernels -- and the whole oracle network keeps the input resolution."""
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 7, 9, 5, generator=g, dtype=torch.float64)
    k = torch.rand(3, 3, 4, 5, generator=g, dtype=torch.float64) - 0.5          # (kh, kw, Cout, Cin)
    b = torch.rand(4, generator=g, dtype=torch.float64)
    ref = O.conv2d_transpose(x, k, b, 1)
    alt = O.conv2d(x, k.flip(0, 1).permute(0, 1, 3, 2).contiguous(), b, 1, "SAME")
    assert ref.shape == (2, 7, 9, 4) and (ref - alt).abs().max() < 1e-12
    w = O.init_weights(O.generator_unet_spec(gf_dim=8), 5, dtype=torch.float64, randomize_affine=True)
    assert len(w) == 62
    img = torch.rand(1, 10, 12, 3, generator=g, dtype=torch.float64)
    y = O.generator_unet(img, w)
    assert y.shape == img.shape and float(y.abs().max()) <= 1.0
    masks = [(torch.rand(1, 10, 12, 64, generator=g) >= 0.5).double() for _ in range(3)]
    yt = O.generator_unet(img, w, training=True, drop_masks=masks)
    assert (yt - y).abs().max() > 1e-6     # dropout only acts in training mode
