"""GPU debug harness (not a pytest): layer-by-layer parity of the CUDA engine against the oracle.
Usage: python tests/gpu/debug_step.py [B H W [n_blocks]]   -> prints relative errors per layer / tensor."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sggan_oracle as O  # noqa: E402

pkg = importlib.import_module("sg-gan-tf2_b200")
L = importlib.import_module("sg-gan-tf2_b200._lib")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), (a - b).abs().max().item(), b.abs().max().item()


def main():
    args = [int(a) for a in sys.argv[1:]]
    B, H, W = (args + [1, 128, 128])[:3] if len(args) < 3 else args[:3]
    nb = args[3] if len(args) > 3 else 9
    Cs = 34
    torch.manual_seed(0)
    cfg = L.default_config(B, H, W, n_blocks=nb, segment_class=Cs)
    print("config", B, H, W, "blocks", nb, "mask", cfg.mask_height, cfg.mask_width, "workspace MB",
          L.workspace_bytes(cfg) / 2 ** 20)
    eng = L.Engine(cfg)
    gw = O.init_weights(O.generator_spec(n_blocks=nb), 1, randomize_affine=True)
    dw = O.init_weights(O.discriminator_spec(segment_class=Cs), 2, randomize_affine=True)
    eng.set_weights(L.NET_G, gw)
    eng.set_weights(L.NET_D, dw)
    eng.weights_changed()
    real_A, seg_A, mask, _ = O.synthetic_batch(B, H, W, Cs, seed=19)
    torch.cuda.synchronize()

    # ---------------- generator forward
    taps = {}
    t0 = time.time()
    ref_fake = O.generator_resnet(real_A, gw, n_blocks=nb, taps=taps)
    print("oracle G fwd %.2fs" % (time.time() - t0))
    fake = eng.gen_forward(real_A)
    torch.cuda.synchronize()
    names = ["c1", "c2", "c3"] + [None if k % 2 == 0 else "r%d" % (k // 2 + 1) for k in range(2 * nb)] + ["d1", "d2"]
    for li, nm in enumerate(names):
        if nm is None:
            continue
        got = eng.debug_buffer(L.NET_G, li + 1, 0)
        print("G %-4s -> X[%2d]  rel %.3e  maxabs %.3e (ref max %.3e)" % ((nm, li + 1) + rel(got, taps[nm])))
    print("G fake            rel %.3e  maxabs %.3e (ref max %.3e)" % rel(fake, ref_fake))

    # ---------------- discriminator forward on seg_A
    dtaps = {}
    ref_logit = O.discriminator(seg_A, mask, dw, taps=dtaps)
    logit = eng.disc_forward(seg_A, mask)
    torch.cuda.synchronize()
    for li, nm in enumerate(["h0", "h1", "h2", "h3", "h31", "h32", "h33"]):
        got = eng.debug_buffer(L.NET_D, li + 1, 0, nimg=B)
        print("D %-4s -> X[%d]  rel %.3e  maxabs %.3e (ref max %.3e)" % ((nm, li + 1) + rel(got, dtaps[nm])))
    print("D logits          rel %.3e  maxabs %.3e (ref max %.3e)" % rel(logit, ref_logit))

    # ---------------- one training step
    t0 = time.time()
    ref = O.step_grads(gw, dw, real_A, seg_A, mask)
    print("oracle step %.2fs  gen_loss %.6f disc_loss %.6f" % (time.time() - t0, ref["gen_loss"], ref["disc_loss"]))
    eng.step_forward_backward_d(real_A, seg_A, mask)
    eng.step_backward_g()
    torch.cuda.synchronize()
    print("engine        gen_loss %.6f disc_loss %.6f  launches %d" % (eng.losses[0].item(), eng.losses[1].item(),
                                                                        eng.kernel_launches))
    print("fake (step)       rel %.3e  maxabs %.3e (ref max %.3e)" % rel(eng.last_fake(), ref["fake_A"]))
    for net, nm, rg in ((L.NET_D, "D", ref["d_grads"]), (L.NET_G, "G", ref["g_grads"])):
        gs = eng.tensors(net, 1)
        worst = 0.0
        for i, (g, r) in enumerate(zip(gs, rg)):
            e = rel(g, r)
            # biases in front of an instance norm have exactly zero gradient: compare absolutely
            flag = "" if (e[0] < 0.05 or e[2] < 1e-6) else "   <<<<"
            if e[2] >= 1e-6:
                worst = max(worst, e[0])
            print("%s grad[%2d] %-18s rel %.3e  maxabs %.3e (ref max %.3e)%s" % (nm, i, tuple(r.shape), e[0], e[1], e[2], flag))
        print("%s worst relative gradient error: %.3e" % (nm, worst))
    # ---------------- Adam
    st = O.StepState(gw, dw)
    O.train_step(st, real_A, seg_A, mask)
    eng.step_adam(L.NET_G)
    eng.step_adam(L.NET_D)
    torch.cuda.synchronize()
    for net, nm, rw, w0 in ((L.NET_G, "G", st.g, gw), (L.NET_D, "D", st.d, dw)):
        ws = eng.tensors(net, 0)
        num = sum(((a.cpu().double() - b.double()) ** 2).sum() for a, b in zip(ws, rw)) ** 0.5
        den = sum(((b.double() - c.double()) ** 2).sum() for b, c in zip(rw, w0)) ** 0.5
        print("%s post-Adam weights: |w - w_ref| / |w_ref - w0| = %.3e" % (nm, (num / den).item()))


if __name__ == "__main__":
    main()
