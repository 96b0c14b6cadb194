"""The fp32-accurate operator tier (north_star: "rel 1e-2 bf16 / 1e-4 tf32") against the fp64 CPU oracle.

The reference computes in fp32 (Keras defaults, module.py:211-265).  Tolerances, relative L2 on the same inputs:
  * precision="tf32x3" (operands split hi + lo, hi*hi + lo*hi + hi*lo accumulated in fp32 by tcgen05.mma.kind::tf32):
    conv / deconv / instance norm < 1e-4 (measured 1e-6 .. 3e-5), generator output after 24 layers < 1e-3;
  * precision="tf32" (ONE tf32 product per term): < 2e-3 per operator -- operand rounding is 2^-11 = 4.9e-4 per factor,
    so a single pass cannot reach 1e-4 on these dot products whatever the kernel does (the test pins that figure too);
  * the bf16 training path of the same generator: < 3e-2, and it is CLOSER to the fp32-tier output than its bound, which
    is what says the 2e-2 of the training path is storage format, not kernel error.
"""
import importlib
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    return importlib.import_module("sg-gan-tf2_b200.ops")


@pytest.fixture(scope="module")
def mod():
    return importlib.import_module("sg-gan-tf2_b200.module")


CONV_CASES = [
    (2, 16, 24, 64, 64, 3, 1, "REFLECT"), (1, 20, 36, 128, 256, 3, 1, "SAME"), (2, 9, 13, 256, 128, 3, 1, "VALID"),
    (1, 12, 20, 3, 64, 7, 1, "REFLECT"), (2, 16, 32, 64, 128, 3, 2, "SAME"), (1, 15, 31, 128, 256, 3, 2, "VALID"),
    (1, 18, 30, 64, 3, 7, 1, "REFLECT"), (1, 9, 17, 512, 34, 3, 1, "SAME"), (1, 32, 64, 256, 256, 3, 1, "REFLECT"),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,pad", CONV_CASES)
def test_conv2d_fp32_tier(ops, O, B, H, W, Cin, Cout, k, stride, pad):
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + k)
    x = torch.rand(B, H, W, Cin, generator=g, dtype=torch.float64) * 2 - 1
    w = (torch.rand(k, k, Cin, Cout, generator=g, dtype=torch.float64) * 2 - 1) * (1.0 / (k * k * Cin) ** 0.5)
    b = torch.rand(Cout, generator=g, dtype=torch.float64) - 0.5
    xr = O.reflect_pad(x, (k - 1) // 2) if pad == "REFLECT" else x
    ref = O.conv2d(xr, w, b, stride, "VALID" if pad == "REFLECT" else pad)  # fp64
    y3 = ops.conv2d_raw(x.float(), w.float(), b.float(), stride=stride, padding=pad, precision="tf32x3")
    assert tuple(y3.shape) == tuple(ref.shape)
    assert rel(y3, ref) < 1e-4, rel(y3, ref)
    y1 = ops.conv2d_raw(x.float(), w.float(), b.float(), stride=stride, padding=pad, precision="tf32")
    r1 = rel(y1, ref)
    assert r1 < 2e-3, r1
    assert rel(y3, ref) < 0.25 * r1 + 1e-6  # the split removes most of the operand rounding; what stays is the tensor core's fp32 accumulation


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 12, 128, 64), (2, 16, 32, 256, 128), (1, 5, 7, 64, 3)])
def test_deconv2d_fp32_tier(ops, O, B, H, W, Cin, Cout):
    g = torch.Generator().manual_seed(B + H + Cin)
    x = torch.rand(B, H, W, Cin, generator=g, dtype=torch.float64) * 2 - 1
    w = (torch.rand(3, 3, Cout, Cin, generator=g, dtype=torch.float64) * 2 - 1) * (1.0 / (9 * Cin) ** 0.5)
    b = torch.rand(Cout, generator=g, dtype=torch.float64) - 0.5
    ref = O.conv2d_transpose(x, w, b, 2)
    y3 = ops.deconv2d_raw(x.float(), w.float(), b.float(), precision="tf32x3")
    assert tuple(y3.shape) == tuple(ref.shape)
    assert rel(y3, ref) < 1e-4, rel(y3, ref)
    assert rel(ops.deconv2d_raw(x.float(), w.float(), b.float(), precision="tf32"), ref) < 2e-3


@pytest.mark.parametrize("B,H,W,C,act,res", [(2, 16, 24, 64, "relu", False), (1, 9, 13, 256, None, True), (2, 7, 5, 34, "lrelu", False),
                                             (1, 1, 1, 64, None, False)])
def test_instance_norm_fp32_tier(ops, O, B, H, W, C, act, res):
    g = torch.Generator().manual_seed(H * W + C)
    x = torch.rand(B, H, W, C, generator=g, dtype=torch.float64) * 4 - 1
    gamma = 1 + 0.2 * (torch.rand(C, generator=g, dtype=torch.float64) - 0.5)
    beta = 0.2 * (torch.rand(C, generator=g, dtype=torch.float64) - 0.5)
    r = torch.rand(B, H, W, C, generator=g, dtype=torch.float64) if res else None
    ref = O.instance_norm(x, gamma, beta, 1e-3)
    ref = torch.relu(ref) if act == "relu" else (O.lrelu(ref, 0.3) if act == "lrelu" else ref)
    if res:
        ref = ref + r
    y = ops.instance_norm_raw(x.float(), gamma.float(), beta.float(), eps=1e-3, act=act, alpha=0.3,
                              residual=None if r is None else r.float(), precision="fp32")
    assert rel(y, ref) < 1e-4, rel(y, ref)  # measured ~1e-7: double-precision statistics, fp32 apply
    if H * W == 1:
        assert torch.equal(y.cpu().reshape(-1), beta.float())  # a single-pixel norm returns exactly beta


def test_generator_fp32_tier_matches_reference_to_1e3(mod, O):
    """The whole generator (24 conv / deconv + 23 norms) on the fp32 tier against the fp64 oracle: < 1e-3 (north_star's
    outputs bound for the accurate tier); single-pass tf32 and the bf16 training path on the same weights beside it."""
    B, H, W, nb = 1, 128, 128, 9
    gen = mod.generator_resnet(image_height=H, image_width=W, n_blocks=nb)
    gw = O.init_weights(O.generator_spec(n_blocks=nb), 1, randomize_affine=True)
    gen.set_weights([w.numpy() for w in gw])
    x = torch.rand(B, H, W, 3, generator=torch.Generator().manual_seed(5))
    ref = O.generator_resnet(x.double(), [w.double() for w in gw])
    y3 = gen.forward_fp32(x, precision="tf32x3")
    y1 = gen.forward_fp32(x, precision="tf32")
    yb = gen(x)  # bf16 training path (engine)
    r3, r1, rb = rel(y3, ref), rel(y1, ref), rel(yb, ref)
    print("generator output vs fp64 oracle: tf32x3 %.2e  tf32 %.2e  bf16 %.2e" % (r3, r1, rb))
    assert r3 < 1e-3, r3
    assert r1 < 1e-2, r1
    assert rb < 3e-2, rb
    assert r3 < r1 < rb


def test_generator_unet_forward(mod, O):
    """generator_unet (module.py:125-206, the reference CLI's default generator) on the operator tier: the bf16 path (the
    training kernels; 3-channel ends on the tf32 tier) and the fp32 tier against the fp64 oracle, inference mode and training
    mode with given dropout masks; weights round-trip through the reference's checkpoint format."""
    B, H, W = 1, 48, 64
    gen = mod.generator_unet()
    w = O.init_weights(O.generator_unet_spec(), 7, randomize_affine=True)
    gen.set_weights([t.numpy() for t in w])
    x = torch.rand(B, H, W, 3, generator=torch.Generator().manual_seed(9)) * 2 - 1
    w64 = [t.double() for t in w]
    ref = O.generator_unet(x.double(), w64)
    y3, yb = gen(x, precision="tf32x3"), gen(x)
    print("generator_unet vs fp64 oracle: tf32x3 %.2e  bf16 %.2e" % (rel(y3, ref), rel(yb, ref)))
    assert tuple(y3.shape) == (B, H, W, 3) and rel(y3, ref) < 1e-3
    assert rel(yb, ref) < 3e-2
    masks = [(torch.rand(B, H, W, 512, generator=torch.Generator().manual_seed(20 + i)) >= 0.5).float() for i in range(3)]
    reft = O.generator_unet(x.double(), w64, training=True, drop_masks=[m.double() for m in masks])
    yt = gen(x, training=True, precision="tf32x3", drop_masks=masks)
    assert rel(yt, reft) < 1e-3 and rel(reft, ref) > 1e-2  # dropout does something, and we do the same thing
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        gen.save_weights(os.path.join(d, "cp-0001.ckpt"))
        g2 = mod.generator_unet()
        g2.load_weights(os.path.join(d, "cp-0001.ckpt"))
        assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(g2.trainable_variables, gen.trainable_variables))
