"""Ad-hoc shape sweep (run on a GPU box): one forward/backward of the engine against the CPU oracle for batch
sizes, resolutions and class counts around the kernel-selection thresholds (transposed / staged / shift-sum paths
switch on tile counts).  Prints one line per configuration; exit code 1 if any deviates."""
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


def main():
    cfgs = [(1, 128, 256, 34, 2), (3, 128, 256, 19, 1), (1, 256, 512, 34, 1), (2, 192, 320, 8, 2), (5, 64, 128, 34, 1),
            (1, 136, 264, 34, 1), (2, 256, 256, 3, 2), (8, 128, 128, 34, 1)]
    bad = 0
    for (B, H, W, C, nb) in cfgs:
        cfg = L.default_config(B, H, W, n_blocks=nb, segment_class=C)
        try:
            eng = L.Engine(cfg, "cuda:0")
        except L.SgganError as e:
            print("B%d %dx%d C%d blocks%d: rejected loudly (%s)" % (B, H, W, C, nb, e), flush=True)
            continue
        gw = O.init_weights(O.generator_spec(n_blocks=nb), 1, randomize_affine=True)
        dw = O.init_weights(O.discriminator_spec(segment_class=C), 2, randomize_affine=True)
        eng.set_weights(L.NET_G, gw)
        eng.set_weights(L.NET_D, dw)
        eng.weights_changed()
        real_A, seg_A, mask, _ = O.synthetic_batch(B, H, W, C, seed=B + H)
        ref = O.step_grads(gw, dw, real_A, seg_A, mask)
        eng.step_forward_backward_d(real_A, seg_A, mask)
        eng.step_backward_g()
        torch.cuda.synchronize()
        lg = abs(eng.losses[0].item() - ref["gen_loss"].item()) / abs(ref["gen_loss"].item())
        ld = abs(eng.losses[1].item() - ref["disc_loss"].item()) / abs(ref["disc_loss"].item())
        rf = rel(eng.last_fake(), ref["fake_A"])
        gg = eng.tensors(L.NET_G, 1)
        r_out = rel(gg[-2], ref["g_grads"][-2])
        cos = torch.nn.functional.cosine_similarity(torch.cat([g.reshape(-1) for g in gg]).double().cpu(),
                                                    torch.cat([g.reshape(-1) for g in ref["g_grads"]]).double(), dim=0).item()
        ok = lg < 1e-2 and ld < 1e-2 and rf < 3e-2 and r_out < 3e-2 and cos > 0.9
        bad += not ok
        print("B%d %dx%d C%d blocks%d: gen %.2e disc %.2e fake %.2e out-conv dW %.2e cos(gradG) %.4f %s" %
              (B, H, W, C, nb, lg, ld, rf, r_out, cos, "ok" if ok else "DEVIATES"), flush=True)
        del eng
        torch.cuda.empty_cache()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
