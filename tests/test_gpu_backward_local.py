"""Local consistency of EVERY backward kernel launch of a real training step.

Step-level gradient comparisons against the oracle are dominated by bf16 chaos (1e-2 forward noise flips ReLU masks
and the L1 sign gradient; tests/test_gpu_numerics.py), which could hide a systematic error in the custom backward
(a wrong reflect-fold row, a dropped stride-2 phase, a wrong virtual-image index).  Here each backward op of the step
is checked in isolation: its INPUTS are read back from the engine (input frame X, raw conv output Y, upstream
gradient, all bf16 as stored), the oracle's autograd computes what the op must produce from exactly those inputs, and
the engine's stored result (dX, dY, dW, dgamma, dbeta, dbias) has to match at single-op accuracy:
  * weight / affine / bias gradients (fp32 accumulation of identical bf16 operands): rel-L2 < 2e-3
  * dX, dY (stored as bf16):                                                        rel-L2 < 1e-2
Covers dgrad (incl. the padded-grid layout and the 4-phase stride-2 / transposed forms), wgrad (incl. the
sliding-window 3-channel layers), instance-norm backward (reduce + apply, dgamma / dbeta), the reflect-border fold,
the residual-stream gradient gather, the LeakyReLU backward of D's first layer and the 3B-virtual-image indexing
of the discriminator backward (reference: both GradientTape.gradient calls, model.py:196-197).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def bf(x):
    return x.to(torch.bfloat16).float()


# (kind, k, stride, pad, norm, act) per layer, in engine order (csrc/engine.cu build_net_g / build_net_d)
def g_layers(nb):
    ls = [("conv", 7, 1, "REFLECT", True, "relu"), ("conv", 3, 2, "SAME", True, "relu"), ("conv", 3, 2, "SAME", True, "relu")]
    for _ in range(nb):
        ls += [("conv", 3, 1, "REFLECT", True, "relu"), ("conv", 3, 1, "REFLECT", True, None)]
    ls += [("deconv", 3, 2, "SAME", True, "relu"), ("deconv", 3, 2, "SAME", True, "relu"), ("conv", 7, 1, "REFLECT", False, "tanh")]
    return ls


D_LAYERS = [("conv", 3, 2, "SAME", False, "lrelu"), ("conv", 3, 2, "SAME", True, "lrelu"), ("conv", 3, 2, "SAME", True, "lrelu"),
            ("conv", 3, 1, "SAME", True, "lrelu"), ("conv", 3, 2, "VALID", True, "lrelu"), ("conv", 3, 2, "VALID", True, "lrelu"),
            ("conv", 3, 1, "VALID", True, "lrelu"), ("conv", 3, 1, "SAME", False, None)]


def tensor_index(layers):
    """first tensor index (kernel) of each layer in the Keras-order list"""
    idx, out = 0, []
    for (_, _, _, _, norm, _) in layers:
        out.append(idx)
        idx += 4 if norm else 2
    return out


def conv_fwd(O, spec, x, w):
    """the layer's convolution on its (logical) input; returns (y, leaf) where leaf is the tensor the engine's dX is the
    gradient of (the padded grid for stride-1 layers with a border, the input itself otherwise)"""
    kind, k, stride, pad, _, _ = spec
    p = (k - 1) // 2
    if kind == "deconv":
        leaf = x.clone().requires_grad_(True)
        return O.conv2d_transpose(leaf, w, None, 2), leaf
    if stride == 1 and pad == "REFLECT":
        leaf = O.reflect_pad(x, p).clone().requires_grad_(True)
        return O.conv2d(leaf, w, None, 1, "VALID"), leaf
    if stride == 1 and pad == "SAME":
        leaf = F.pad(x, (0, 0, p, p, p, p)).clone().requires_grad_(True)
        return O.conv2d(leaf, w, None, 1, "VALID"), leaf
    leaf = x.clone().requires_grad_(True)
    return O.conv2d(leaf, w, None, stride, pad), leaf


def fold(O, spec, dx_raw, H, W):
    """engine dX buffer (padded-grid layout) -> gradient w.r.t. the layer's logical H x W input"""
    kind, k, stride, pad, _, _ = spec
    p = (k - 1) // 2
    if kind == "conv" and stride == 1 and pad == "REFLECT":
        z = torch.zeros(dx_raw.shape[0], H, W, dx_raw.shape[3], requires_grad=True)
        O.reflect_pad(z, p).backward(dx_raw[:, :H + 2 * p, :W + 2 * p])
        return z.grad
    if kind == "conv" and stride == 1 and pad == "SAME":
        return dx_raw[:, p:p + H, p:p + W]
    return dx_raw[:, :H, :W]


def act_fn(name):
    return {"relu": torch.relu, "lrelu": lambda v: torch.maximum(v, 0.3 * v), None: lambda v: v}[name]


def check_net(L, O, eng, net, layers, weights, nb_img, nbv_img, act_wrap, dz_of, tol_w=2e-3, tol_a=1e-2):
    """dz_of(li) -> gradient w.r.t. the logical OUTPUT (post norm / activation) of layer li, for the nbv virtual images"""
    tix = tensor_index(layers)
    grads = eng.tensors(net, 1)
    report = []
    for li, spec in enumerate(layers):
        kind, k, stride, pad, norm, act = spec
        wq = bf(weights[tix[li]])
        X = eng.debug_buffer(net, li, 0).cpu()                       # (nb, Hin, Win, Cx)
        cin = wq.shape[3] if kind == "deconv" else wq.shape[2]
        X = X[..., :cin]
        dY = eng.debug_buffer(net, li, 2).cpu()                      # (nbv, Hout, Wout, CoutK)
        cout = wq.shape[2] if kind == "deconv" else wq.shape[3]
        dY = dY[..., :cout]
        # ---- weight gradient: first nb images
        wl = wq.clone().requires_grad_(True)
        y, _ = conv_fwd(O, spec, X[:nb_img], wl)
        (dW,) = torch.autograd.grad(y, wl, dY[:nb_img])
        r = rel(grads[tix[li]], dW)
        report.append(("dW", li, r))
        assert r < tol_w, ("dW", net, li, r)
        if not norm:
            db = dY[:nb_img].sum(dim=(0, 1, 2))
            r = rel(grads[tix[li] + 1], db)
            assert r < 1e-2, ("dbias", net, li, r)  # accumulated from the fp32 seed, before its bf16 rounding into dY
        # ---- input gradient: all virtual images the engine computes it for
        has_dx = not (net == L.NET_G and li == 0)
        if has_dx:
            first = nb_img if (net == L.NET_D and li == 0) else 0  # D's first layer: only the generator-path images
            n = nbv_img - first
            dXe = eng.debug_buffer(net, li, 3, nimg=n).cpu()
            xin = torch.cat([X, X[nb_img - act_wrap:nb_img]])[first:first + n] if nbv_img > nb_img else X
            y, leaf = conv_fwd(O, spec, xin, wq)
            (dXo,) = torch.autograd.grad(y, leaf, dY[first:first + n])
            dXe = dXe[:, :dXo.shape[1], :dXo.shape[2]]
            r = rel(dXe, dXo)
            report.append(("dX", li, r))
            assert r < tol_a, ("dX", net, li, r)
        # ---- norm (+ activation) backward: dY from the upstream gradient
        if norm:
            Y = eng.debug_buffer(net, li, 1).cpu()
            g, be = weights[tix[li] + 2], weights[tix[li] + 3]
            dz = dz_of(li)
            Yv = torch.cat([Y, Y[nb_img - act_wrap:nb_img]]) if nbv_img > nb_img else Y
            Yl = Yv.clone().requires_grad_(True)
            gl, bl = g.clone().requires_grad_(True), be.clone().requires_grad_(True)
            z = act_fn(act)(O.instance_norm(Yl, gl, bl, eps=1e-3))
            (dYo,) = torch.autograd.grad(z, Yl, dz, retain_graph=True)
            r = rel(dY, dYo)
            report.append(("dY", li, r))
            assert r < tol_a, ("dY(norm bwd)", net, li, r)
            mask = torch.zeros_like(dz)
            mask[:nb_img] = 1
            dgo, dbo = torch.autograd.grad(z, (gl, bl), dz * mask)
            rg, rb = rel(grads[tix[li] + 2], dgo), rel(grads[tix[li] + 3], dbo)
            assert rg < 5e-3 and rb < 5e-3, ("dgamma/dbeta", net, li, rg, rb)
            assert float(grads[tix[li] + 1].abs().max()) == 0.0  # bias in front of a norm
    return report


@pytest.mark.parametrize("B,H,W", [(2, 128, 256), (1, 256, 512)])
def test_backward_local_consistency(L, O, B, H, W):
    nb, C = 2, 34
    cfg = L.default_config(B, H, W, n_blocks=nb, segment_class=C)
    eng = L.Engine(cfg)
    gw = O.init_weights(O.generator_spec(n_blocks=nb), 1, randomize_affine=True)
    dw = O.init_weights(O.discriminator_spec(segment_class=C), 2, randomize_affine=True)
    eng.set_weights(L.NET_G, gw)
    eng.set_weights(L.NET_D, dw)
    eng.weights_changed()
    real_A, seg_A, mask, _ = O.synthetic_batch(B, H, W, C, seed=11)
    eng.step_forward_backward_d(real_A, seg_A, mask)
    eng.step_backward_g()
    torch.cuda.synchronize()

    # ------------------------------------------------------------------ discriminator: 3B virtual images
    dl = D_LAYERS

    def d_dims(li):
        x = eng.debug_buffer(L.NET_D, li, 0)
        return x.shape[1], x.shape[2]

    def d_dz(li):  # gradient w.r.t. the output of D layer li = folded dX of layer li + 1
        Hn, Wn = d_dims(li + 1)
        return fold(O, dl[li + 1], eng.debug_buffer(L.NET_D, li + 1, 3).cpu(), Hn, Wn)

    rep_d = check_net(L, O, eng, L.NET_D, dl, dw, 2 * B, 3 * B, B, d_dz)
    # h0: LeakyReLU applied in the conv epilogue; dY0 = dz * lrelu'(z), z = h1's input frame (virtual-image wrap)
    z1 = eng.debug_buffer(L.NET_D, 1, 0).cpu()
    z1 = torch.cat([z1, z1[B:2 * B]])
    dy0 = d_dz(0) * torch.where(z1 > 0, torch.ones_like(z1), torch.full_like(z1, 0.3))
    assert rel(eng.debug_buffer(L.NET_D, 0, 2).cpu()[..., :64], dy0) < 1e-2

    # ------------------------------------------------------------------ generator
    gl = g_layers(nb)
    first_blk, n_l = 3, len(gl)

    def g_dims(li):
        x = eng.debug_buffer(L.NET_G, li, 0)
        return x.shape[1], x.shape[2]

    def g_fold(li):
        Hn, Wn = g_dims(li)
        return fold(O, gl[li], eng.debug_buffer(L.NET_G, li, 3).cpu(), Hn, Wn)

    # residual-stream gradients G_k (w.r.t. the input r_k of block k); the engine stores each sum as bf16
    Gk = {nb: g_fold(first_blk + 2 * nb)}  # dX of the first transposed convolution
    for kblk in range(nb - 1, -1, -1):
        Gk[kblk] = bf(Gk[kblk + 1] + g_fold(first_blk + 2 * kblk))

    def g_dz(li):
        in_blocks = first_blk <= li < first_blk + 2 * nb
        if in_blocks and (li - first_blk) % 2 == 1:  # conv_b of block k -> r_{k+1}
            return Gk[(li - first_blk) // 2 + 1]
        if li == first_blk - 1:                      # c3 -> r_0
            return Gk[0]
        return g_fold(li + 1)

    rep_g = check_net(L, O, eng, L.NET_G, gl, gw, B, B, 0, g_dz)
    # seed of the generator backward: dY(out) = (100 * sign(fake - seg) / N + dD) * (1 - fake^2), dD = D's fp32 input gradient
    fake = eng.last_fake().cpu()
    dD = eng.debug_buffer(L.NET_D, 0, 3, nimg=B).cpu()[..., :3]
    seed = (100.0 * torch.sign(fake - seg_A) / fake.numel() + dD) * (1 - fake * fake)
    assert rel(eng.debug_buffer(L.NET_G, n_l - 1, 2).cpu()[..., :3], seed) < 1e-2
    worst = max(r for (_, _, r) in rep_d + rep_g)
    print("backward local consistency: %d checks, worst rel-L2 %.3e" % (len(rep_d) + len(rep_g), worst))
