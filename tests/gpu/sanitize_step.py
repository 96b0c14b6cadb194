"""One small training step + the fp32-tier operators, for `compute-sanitizer --tool memcheck python tests/gpu/sanitize_step.py`
(run it alone in a gpurun call; the sanitizer slows kernels by 10-50x, hence the tiny shapes)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sggan_oracle as O  # noqa: E402

L = importlib.import_module("sg-gan-tf2_b200._lib")
ops = importlib.import_module("sg-gan-tf2_b200.ops")
os.environ.setdefault("SGGAN_GRAPH", "0")
for (B, H, W, C, nb, mode) in ((2, 128, 256, 34, 1, L.LOSS_P2P), (1, 128, 128, 19, 1, L.LOSS_SGGAN), (3, 256, 256, 34, 2, L.LOSS_P2P)):
    cfg = L.default_config(B, H, W, n_blocks=nb, segment_class=C, loss_mode=mode)
    eng = L.Engine(cfg)
    eng.set_weights(L.NET_G, O.init_weights(O.generator_spec(n_blocks=nb), 1, randomize_affine=True))
    eng.set_weights(L.NET_D, O.init_weights(O.discriminator_spec(segment_class=C), 2, randomize_affine=True))
    eng.weights_changed()
    a, s, m, _ = O.synthetic_batch(B, H, W, C, seed=3)
    if H == 128 and W == 128:
        m = torch.rand(B, 4, 4, C)  # the loader's 4x4 mask against 1x1 logits
        del eng
        cfg = L.default_config(B, H, W, n_blocks=nb, segment_class=C, loss_mode=mode, mask_height=4, mask_width=4)
        eng = L.Engine(cfg)
        eng.set_weights(L.NET_G, O.init_weights(O.generator_spec(n_blocks=nb), 1, randomize_affine=True))
        eng.set_weights(L.NET_D, O.init_weights(O.discriminator_spec(segment_class=C), 2, randomize_affine=True))
        eng.weights_changed()
    for _ in range(2):
        losses = eng.train_step(a, s, m)
    torch.cuda.synchronize()
    print("step", (B, H, W, C, nb), [float(x) for x in losses.cpu()])
    del eng
x = torch.rand(1, 20, 36, 128)
w = torch.rand(3, 3, 128, 256) * 0.03
for prec in ("bf16", "tf32", "tf32x3"):
    y = ops.conv2d_raw(x, w, None, stride=1, padding="REFLECT", precision=prec)
    z = ops.instance_norm_raw(y, torch.ones(256), torch.zeros(256), act="relu", precision=prec)
torch.cuda.synchronize()
print("SANITIZE-DONE")
