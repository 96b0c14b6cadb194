"""Worker of the data-parallel equivalence test (tests/test_gpu_dp.py; also runnable by hand under torchrun):

    torchrun --nproc-per-node 2 tests/gpu/dp_worker.py --backend nccl            # one GPU per rank
    torchrun --nproc-per-node 2 tests/gpu/dp_worker.py --backend gloo --one-gpu  # both ranks on cuda:0

Every rank trains `--steps` steps on its shard of a global batch through model.sggan.train_step; rank 0 then repeats
the run alone on the concatenated batch and compares gradients (last step), losses and post-step weights."""
import argparse
import contextlib
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def run(M, L, O, ns, batches, gw, dw):
    with contextlib.redirect_stdout(sys.stderr):
        m = M.sggan(ns)
    losses = []
    for (a, s, k) in batches:
        m.real_A, m.seg_A, m.mask_A = a, s, k
        if m.runtime is None:  # plan, then load the common initial weights into the owner of the master weights
            B, H, W, _ = a.shape
            rt = m._ensure_runtime(B, H, W, (int(k.shape[1]), int(k.shape[2])))
            rt.engine.set_weights(L.NET_G, gw)
            rt.engine.set_weights(L.NET_D, dw)
            rt.engine.weights_changed()
        gl, dl = m.train_step(ns)
        losses.append((float(gl), float(dl)))
    eng = m.runtime.engine
    torch.cuda.synchronize()
    return m, losses, [eng.flat(n, w).clone() for n in (L.NET_G, L.NET_D) for w in (0, 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--one-gpu", action="store_true")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--height", type=int, default=128)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--per-rank", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(0 if args.one_gpu else int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(args.backend)
    import sggan_oracle as O
    L = importlib.import_module("sg-gan-tf2_b200._lib")
    M = importlib.import_module("sg-gan-tf2_b200.model")
    H, W, C, nb, b = args.height, args.width, 34, 2, args.per_rank
    ns = argparse.Namespace(batch_size=b, image_width=W, image_height=H, segment_class=C, use_resnet=True)
    gw = O.init_weights(O.generator_spec(n_blocks=9), 1, randomize_affine=True)
    dw = O.init_weights(O.discriminator_spec(segment_class=C), 2, randomize_affine=True)
    glob = [O.synthetic_batch(world * b, H, W, C, seed=40 + s)[:3] for s in range(args.steps)]
    mine = [tuple(t[rank * b:(rank + 1) * b].contiguous() for t in g) for g in glob]
    m, losses, (gp, gg, dp, dg) = run(M, L, O, ns, mine, gw, dw)
    assert m.world_size == world
    # mean of the per-rank losses = loss of the global batch
    lt = torch.tensor(losses[-1], dtype=torch.float64)
    dist.all_reduce(lt)
    lt /= world
    # every rank must hold identical weights after the step (same all-reduced gradients, same Adam)
    chk = torch.stack([gp.double().sum(), dp.double().sum()]).cpu()
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "replicas diverged: %s vs %s" % (lo, hi)
    dist.barrier()
    ok = True
    if rank == 0:
        dist_was = dist.is_initialized
        # single-rank reference on the concatenated batch: hide the process group from model.sggan
        dist.is_initialized = lambda: False
        try:
            ns1 = argparse.Namespace(batch_size=world * b, image_width=W, image_height=H, segment_class=C, use_resnet=True)
            m1, losses1, (gp1, gg1, dp1, dg1) = run(M, L, O, ns1, glob, gw, dw)
        finally:
            dist.is_initialized = dist_was
        assert m1.world_size == 1
        # gradient buffers of the last step: the DP ranks hold the SUM over ranks (Adam applies 1/world)
        rg, rd = rel(gg / world, gg1), rel(dg / world, dg1)
        wg, wd = (gp - gp1).abs().max().item(), (dp - dp1).abs().max().item()
        fg = ((gp - gp1).abs() > 1e-4).float().mean().item()
        fd = ((dp - dp1).abs() > 1e-4).float().mean().item()
        lg = abs(lt[0].item() - losses1[-1][0]) / abs(losses1[-1][0])
        ld = abs(lt[1].item() - losses1[-1][1]) / abs(losses1[-1][1])
        print("world %d backend %s: grad rel-L2 G %.3e D %.3e | weights max|diff| G %.2e D %.2e, fraction > 1e-4: G %.4f D %.4f | "
              "loss rel G %.2e D %.2e" % (world, args.backend, rg, rd, wg, wd, fg, fd, lg, ld), flush=True)
        # bounds: a missing 1/world or a dropped rank is an O(1) error in the gradients and a 2x error in the losses;
        # the per-sample math is identical, what differs is the summation order of the weight-gradient GEMMs (and the
        # bf16-noise-sized differences that follow from it after the first weight update)
        ok = rg < 3e-2 and rd < 8e-2 and lg < 5e-3 and ld < 5e-3 and wg < 4.1e-3 and wd < 4.1e-3 and fg < 0.05 and fd < 0.05
        print("DP-OK" if ok else "DP-MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
