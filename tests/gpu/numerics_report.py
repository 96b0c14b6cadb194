"""Per-tensor numerics table of one engine step against the fp32 oracle and the bf16-rounding oracle
(run on a GPU box; the output is committed as profiles/r02_numerics_report.txt)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sggan_oracle as O  # noqa: E402

L = importlib.import_module("sg-gan-tf2_b200._lib")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def one(B, H, W, C, nb, **kw):
    cfg = L.default_config(B, H, W, n_blocks=nb, segment_class=C, **kw)
    eng = L.Engine(cfg, "cuda:0")
    gw = O.init_weights(O.generator_spec(n_blocks=nb), 1, randomize_affine=True)
    dw = O.init_weights(O.discriminator_spec(segment_class=C), 2, randomize_affine=True)
    eng.set_weights(L.NET_G, gw)
    eng.set_weights(L.NET_D, dw)
    eng.weights_changed()
    real_A, seg_A, mask, _ = O.synthetic_batch(B, H, W, C, seed=19)
    okw = {"p2p_lambda": kw["p2p_lambda"]} if "p2p_lambda" in kw else {}
    ref = O.step_grads(gw, dw, real_A, seg_A, mask, **okw)
    emu = O.step_grads(gw, dw, real_A, seg_A, mask, emu=O.BF16Emu, **okw)
    eng.step_forward_backward_d(real_A, seg_A, mask)
    eng.step_backward_g()
    torch.cuda.synchronize()
    gg, dg = eng.tensors(L.NET_G, 1), eng.tensors(L.NET_D, 1)
    print("== B%d %dx%d C%d blocks %d %s" % (B, H, W, C, nb, kw))
    print("   gen_loss %.6f (fp32 %.6f, emu %.6f)  disc_loss %.6f (fp32 %.6f, emu %.6f)" % (
        eng.losses[0].item(), ref["gen_loss"].item(), emu["gen_loss"].item(), eng.losses[1].item(),
        ref["disc_loss"].item(), emu["disc_loss"].item()))
    print("   fake_A rel-L2: vs fp32 %.3e, vs emu %.3e (emu vs fp32 %.3e)" % (
        rel(eng.last_fake(), ref["fake_A"]), rel(eng.last_fake(), emu["fake_A"]), rel(emu["fake_A"], ref["fake_A"])))
    for name, got, r32, re in (("G", gg, ref["g_grads"], emu["g_grads"]), ("D", dg, ref["d_grads"], emu["d_grads"])):
        worst32 = worste = 0.0
        for i, (a, b, c) in enumerate(zip(got, r32, re)):
            if float(b.abs().max()) < 1e-5:
                continue
            x, y, z = rel(a, b), rel(a, c), rel(c, b)
            worst32, worste = max(worst32, x), max(worste, y)
            if a.dim() == 4 or i >= len(got) - 4:
                print("   %s[%2d] %-18s engine vs fp32 %.3e | engine vs emu %.3e | emu vs fp32 %.3e" % (
                    name, i, tuple(a.shape), x, y, z))
        print("   worst %s: vs fp32 %.3e, vs emu %.3e" % (name, worst32, worste))
    del eng
    torch.cuda.empty_cache()


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    one(2, 256, 256, 34, 2 if quick else 9)
    if not quick:
        one(2, 256, 256, 34, 2, p2p_lambda=0.0)
        one(8, 256, 512, 34, 9)
        one(1, 512, 1024, 19, 9)
