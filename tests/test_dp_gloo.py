"""Data-parallel host logic on CPU (gloo, world_size 2): rank r takes samples [r*b, (r+1)*b), gradients
are summed across ranks and divided by the world size; because instance norm is per-sample and the losses
are batch means, the result equals the single-process gradient on the concatenated batch.
The compute here is the oracle (no GPU); the communication pattern is the one model.sggan.train_step uses
(flat gradient buffer per net, SUM all-reduce, 1/world scaling, then Adam)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sggan_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    gw = O.init_weights(O.generator_spec(n_blocks=1), 11, randomize_affine=True)
    dw = O.init_weights(O.discriminator_spec(segment_class=4), 12, randomize_affine=True)
    for w in gw + dw:  # identical replicas: broadcast rank 0's weights
        dist.broadcast(w, src=0)
    a, s, m, _ = O.synthetic_batch(world, 136, 136, 4, seed=5)
    out = O.step_grads(gw, dw, a[rank:rank + 1], s[rank:rank + 1], m[rank:rank + 1])
    flat_g = torch.cat([g.reshape(-1) for g in out["g_grads"]])
    flat_d = torch.cat([g.reshape(-1) for g in out["d_grads"]])
    hd = dist.all_reduce(flat_d, async_op=True)  # D first (ready first), then G -- as in model.train_step
    hg = dist.all_reduce(flat_g, async_op=True)
    hd.wait(); hg.wait()
    flat_g /= world
    flat_d /= world
    losses = torch.stack([out["gen_loss"], out["disc_loss"]])
    dist.all_reduce(losses)
    losses /= world
    if rank == 0:
        full = O.step_grads(gw, dw, a, s, m)
        ref_g = torch.cat([g.reshape(-1) for g in full["g_grads"]])
        ref_d = torch.cat([g.reshape(-1) for g in full["d_grads"]])
        ret["g"] = ((flat_g - ref_g).norm() / ref_g.norm()).item()
        ret["d"] = ((flat_d - ref_d).norm() / ref_d.norm()).item()
        ret["lg"] = abs(losses[0].item() - full["gen_loss"].item())
        ret["ld"] = abs(losses[1].item() - full["disc_loss"].item())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_dp_gradients_equal_big_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    print("dp-vs-big-batch deviations:", dict(ret))
    # fp32 CPU math with a thread-count dependent summation order: a wrong reduction (sum instead of mean,
    # a dropped rank) is an O(1) error, rounding is ~1e-6; the bound sits between the two
    assert ret["g"] < 1e-3 and ret["d"] < 1e-3, dict(ret)
    assert ret["lg"] < 1e-4 and ret["ld"] < 1e-4, dict(ret)
