"""Host-side logic without a GPU: the C-ABI library loads and exports every symbol include/sggan.h
declares, the planner sizes workspaces / rejects bad shapes loudly, and the Python surface mirrors the
reference's names.  No compute calls."""
import argparse
import ctypes as C
import importlib
import os
import re

import numpy as np

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(L):
    hdr = open(os.path.join(ROOT, "include", "sggan.h")).read()
    declared = set(re.findall(r"\b(sggan_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    lib = C.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(L.SYMBOLS), "ctypes table and header disagree: %s" % (declared ^ set(L.SYMBOLS))


def test_config_defaults_match_reference(L):
    cfg = L.default_config(8, 256, 512)
    assert (cfg.gf_dim, cfg.df_dim, cfg.segment_class, cfg.n_blocks) == (64, 64, 34, 9)  # module.py:221,274-275
    assert abs(cfg.lr - 1e-3) < 1e-9 and abs(cfg.beta1 - 0.5) < 1e-9  # model.py:82,205-207; main.py:28
    assert abs(cfg.beta2 - 0.999) < 1e-6 and abs(cfg.adam_eps - 1e-7) < 1e-12 and abs(cfg.in_eps - 1e-3) < 1e-9
    assert abs(cfg.p2p_lambda - 100.0) < 1e-6  # model.py:151
    assert (cfg.mask_height, cfg.mask_width) == (5, 13)
    assert L.disc_logit_grid(512, 1024) == (13, 29) and L.disc_logit_grid(128, 128) == (1, 1)
    assert C.sizeof(L.Config) == 20 * 4


def test_workspace_planning(L):
    small = L.workspace_bytes(L.default_config(1, 128, 128))
    mid = L.workspace_bytes(L.default_config(8, 256, 512))
    big = L.workspace_bytes(L.default_config(4, 512, 1024, segment_class=19))
    assert 0 < small < mid < big < 180 * 2 ** 30
    assert L.workspace_bytes(L.default_config(16, 256, 512)) > 1.8 * mid - 2 ** 30  # activations scale with batch
    # 128x128 with the loader's 4x4 mask broadcasts; 8x15 against 5x13 does not
    L.workspace_bytes(L.default_config(2, 128, 128, mask_height=4, mask_width=4))
    with pytest.raises(L.SgganError, match="broadcast"):
        L.workspace_bytes(L.default_config(1, 256, 512, mask_height=8, mask_width=15))
    with pytest.raises(L.SgganError, match="too small"):
        L.workspace_bytes(L.default_config(1, 64, 64))  # 64x64: negative dim at h33 (Appendix B)
    with pytest.raises(L.SgganError):
        L.workspace_bytes(L.default_config(1, 256, 512, gf_dim=32))
    with pytest.raises(AttributeError):
        L.default_config(1, 256, 512, not_a_field=1)


def test_no_cpu_fallback(L):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.SgganError, match="no CPU fallback"):
        L.Engine(L.default_config(1, 128, 128))
    h = C.c_void_p()
    cfg = L.default_config(1, 128, 128)
    buf = (C.c_char * 16)()
    rc = L.lib().sggan_create(C.byref(cfg), C.cast(buf, C.c_void_p), 16, None, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in L.lib().sggan_last_error()


def test_python_surface_names():
    mod = importlib.import_module("sg-gan-tf2_b200.module")
    ops = importlib.import_module("sg-gan-tf2_b200.ops")
    model = importlib.import_module("sg-gan-tf2_b200.model")
    for n in ("generator_resnet", "discriminator", "residule_block", "tf_kernel_prep_3d", "tf_deriv", "abs_criterion",
              "mae_criterion", "sce_criterion", "gradloss_criterion"):
        assert callable(getattr(mod, n))
    for n in ("conv2d", "deconv2d", "instance_norm", "lrelu"):
        assert callable(getattr(ops, n))
    import inspect
    assert str(inspect.signature(ops.conv2d)) == "(input_, output_dim, ks=4, s=2, stddev=0.02, padding='SAME', name='conv2d')"
    assert str(inspect.signature(ops.deconv2d)) == "(input_, output_dim, ks=4, s=2, stddev=0.02, name='deconv2d')"
    assert str(inspect.signature(ops.lrelu)) == "(x, leak=0.2, name='lrelu')"
    assert str(inspect.signature(mod.residule_block)).startswith("(x, dim, ks=3, s=1")
    for n in ("train_step", "train", "test", "save", "load", "generate_test_images", "gen_loss_p2p", "disc_loss_p2p",
              "generator_loss", "discriminator_loss"):
        assert callable(getattr(model.sggan, n))
    k = mod.tf_kernel_prep_3d(__import__("numpy").array([[0, 0, 0], [-1, 0, 1], [0, 0, 0]]), 3)
    assert k.shape == (3, 3, 3) and (k[1, 0] == -1).all() and (k[1, 2] == 1).all()


def test_keras_variable_order():
    mod = importlib.import_module("sg-gan-tf2_b200.module")
    g = mod.generator_resnet()
    d = mod.discriminator()
    gv, dv = g.trainable_variables, d.trainable_variables
    assert len(gv) == 94 and len(dv) == 28
    assert sum(v.numel() for v in gv) == 11388675 and sum(v.numel() for v in dv) == 8791970
    assert tuple(gv[0].shape) == (7, 7, 3, 64) and tuple(gv[-2].shape) == (7, 7, 64, 3)
    assert tuple(gv[84].shape) == (3, 3, 128, 256)  # Conv2DTranspose kernel is (kh, kw, Cout, Cin)
    assert tuple(dv[-2].shape) == (3, 3, 512, 34)
    assert float(gv[1].abs().sum()) == 0 and float(gv[2].min()) == 1.0 and float(gv[3].abs().sum()) == 0  # zeros/ones/zeros
    lim = (6.0 / (7 * 7 * 3 + 7 * 7 * 64)) ** 0.5
    assert float(gv[0].abs().max()) <= lim  # glorot-uniform bound
    model = importlib.import_module("sg-gan-tf2_b200.model")
    # use_resnet=False (the reference CLI's default) builds generator_unet: 62 variables in Keras order, forward only --
    # train_step says so loudly instead of training something else
    ns = argparse.Namespace(batch_size=1, image_width=512, image_height=256, use_resnet=False)
    m = model.sggan(ns)
    uv = m.generator.trainable_variables
    assert len(uv) == 62 and tuple(uv[0].shape) == (3, 3, 3, 64) and tuple(uv[-2].shape) == (3, 3, 3, 64)  # last: (kh, kw, Cout=3, Cin=64)
    assert tuple(uv[32].shape) == (3, 3, 512, 512) and m.generator.layer_kinds().count("deconv") == 8
    with pytest.raises(Exception, match="use_resnet"):
        m.train_step(ns)
    with pytest.raises(Exception, match="pix2pix"):
        model.sggan(argparse.Namespace(batch_size=1, image_width=512, image_height=256, use_pix2pix=True))


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores) prints ONE JSON line with the contract's
    keys; it must work without a GPU and without /root/reference."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--height", "128", "--width", "256"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "img/s"


def test_image_pool_follows_the_reference_contract():
    """utils.py:27-53: fills up to maxsize returning its input, then swaps elements 0/2 and 1/3 of two random slots half
    of the time; driven by numpy's global random state like the reference, so a seeded run is reproducible."""
    import importlib
    U = importlib.import_module("sg-gan-tf2_b200.utils")
    pool = U.ImagePool(maxsize=3)
    items = [[("a", i), ("b", i), ("c", i), ("d", i)] for i in range(8)]
    for i in range(3):
        assert pool(items[i]) is items[i] and pool.num_img == i + 1
    np.random.seed(4)
    outs = [pool(list(items[i])) for i in range(3, 8)]
    np.random.seed(4)
    # replay the reference's decisions with the same random draws
    mirror = [[("a", i), ("b", i), ("c", i), ("d", i)] for i in range(3)]
    for i, got in zip(range(3, 8), outs):
        img = items[i]
        if np.random.rand() > 0.5:
            j = int(np.random.rand() * 3)
            t1, t3 = mirror[j][0], mirror[j][2]
            mirror[j][0], mirror[j][2] = img[0], img[2]
            j = int(np.random.rand() * 3)
            t2, t4 = mirror[j][1], mirror[j][3]
            mirror[j][1], mirror[j][3] = img[1], img[3]
            assert got == [t1, t2, t3, t4]
        else:
            assert got == img
    assert U.ImagePool(0)(items[0]) is items[0]
