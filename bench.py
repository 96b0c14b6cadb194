#!/usr/bin/env python
"""bench.py -- G+D train img/s of the SG-GAN step (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (libsggan_sm100)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement, host cores

N > 1 is launched by the driver under torchrun (one rank per GPU, NCCL); per-GPU batch is fixed
(weak scaling) and gradients are all-reduced between backward and Adam.  One JSON line on rank 0.

What is timed: K full steps (G fwd, D fwd on real+fake, D backward, G backward, Adam on both nets,
weight re-pack) with inputs resident in HBM -> `value`; the same K steps through model.sggan.train_step
with HOST numpy batches (pinned H2D + D2H of the two losses inside the timed region) -> `e2e`.
Inputs are synthetic; the working set (>5 GB at batch 8) is far larger than the 126 MB L2.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "G+D train img/s at 256x512"
UNIT = "img/s"
GFLOP_PER_IMG = {(256, 512, 34): 667.3, (512, 1024, 19): 2677.4}  # 3*G + 7*D algorithmic


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sggan_b200", choices=["sggan_b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch (config 3: 8)")
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--classes", type=int, default=34)
    ap.add_argument("--config", default=None, choices=["c2", "c3", "c5"],
                    help="BASELINE.json configs: c2 generator-only inference 256x512 batch 1; c3 (default) full step 256x512 "
                         "batch 8 C=34; c5 full step 512x1024 batch 4 C=19")
    ap.add_argument("--loss-mode", default="p2p", choices=["p2p", "sggan"])
    ap.add_argument("--sustained-seconds", type=float, default=5.0, help="length of the extra power-capped-regime loop (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    if a.config == "c2":
        a.batch, a.height, a.width, a.classes = 1, 256, 512, 34
    elif a.config == "c5":
        a.batch, a.height, a.width, a.classes = 4, 512, 1024, 19
    elif a.config == "c3":
        a.batch, a.height, a.width, a.classes = 8, 256, 512, 34
    return a


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clock + throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def cpu_baseline(H, W, C, steps, warmup, max_seconds=30.0):
    """The reference's CPU path: the PyTorch-CPU fp32 restatement of train_step (oracle/, 'port';
    TensorFlow 2.1 is not installable here), batch 1, all host threads."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sggan_oracle as O
    ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)
    gw = O.init_weights(O.generator_spec(), 1)
    dw = O.init_weights(O.discriminator_spec(segment_class=C), 2)
    st = O.StepState(gw, dw)
    a, s, m, _ = O.synthetic_batch(1, H, W, C, seed=19)
    for _ in range(warmup):
        O.train_step(st, a, s, m)
    times = []
    t_start = time.time()
    for _ in range(steps):
        t0 = time.time()
        O.train_step(st, a, s, m)
        times.append(time.time() - t0)
        if time.time() - t_start > max_seconds:
            break
    sec = statistics.median(times)
    return {"value": 1.0 / sec, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": "%d timed steps (+%d warm-up) of the fp32 CPU restatement of train_step at %dx%d, batch 1, C=%d; "
                      "median %.2f s/step" % (len(times), warmup, H, W, C, sec)}, sec, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = min(args.warmup, 1)
    k = max(1, min(args.steps, 5))
    cb, sec, n = cpu_baseline(args.height, args.width, args.classes, k, w, max_seconds=120.0)
    H, W = args.height, args.width
    line = {"impl": "reference", "metric": METRIC if (H, W) == (256, 512) else "G+D train img/s at %dx%d" % (H, W), "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": w, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SG-GAN train_step %dx%d C=%d, CPU restatement of the reference (TF2 unavailable), "
                                   "batch 1 per step" % (args.height, args.width, args.classes)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = importlib.import_module("sg-gan-tf2_b200._lib")
    M = importlib.import_module("sg-gan-tf2_b200.model")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    B, H, W, C = args.batch, args.height, args.width, args.classes

    # ---- synthetic inputs (seed 19 + rank), generated once, resident on the device
    rng = np.random.RandomState(19 + rank)
    real_h = rng.rand(B, H, W, 3).astype(np.float32)
    seg_h = rng.rand(B, H, W, 3).astype(np.float32)
    hd, wd = L.disc_logit_grid(H, W)
    ids = np.zeros((B, hd, wd), dtype=np.int64)
    for b in range(B):
        for _ in range(6):
            y0, x0 = rng.randint(0, hd), rng.randint(0, wd)
            ids[b, y0:rng.randint(y0, hd) + 1, x0:rng.randint(x0, wd) + 1] = rng.randint(0, C)
    mask_h = (ids[..., None] == np.arange(C)).astype(np.float32)
    real_d, seg_d, mask_d = (torch.as_tensor(x).cuda() for x in (real_h, seg_h, mask_h))

    ns = argparse.Namespace(batch_size=B, image_width=W, image_height=H, segment_class=C, use_resnet=True,
                            loss_mode=args.loss_mode)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):  # the builders print their names like the reference (module.py:220,273)
        model = M.sggan(ns)
    infer = args.config == "c2"
    if infer:
        # BASELINE config 2: generator-only inference (model.py:528-532); the plan is made by the first call
        def step():
            return model.generate_test_images(real_d)
    else:
        model.real_A, model.seg_A, model.mask_A = real_d, seg_d, mask_d

        def step():
            model.train_step(ns)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    for _ in range(max(args.warmup, 3)):
        step()
    eng = (model.generator.runtime if infer else model.runtime).engine
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: inputs resident in HBM, nothing but the step's own launches on the stream
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.kernel_launches
    losses = None if infer else [float(model.gen_loss), float(model.disc_loss)]
    if rank == 0 and not sampler.rows:
        # the timed region was shorter than one nvidia-smi poll (inference: 50 x 0.8 ms): keep the same load running until a
        # sample has been taken, so that the line still carries clocks under THIS load
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 3.0:
            for _ in range(20):
                step()
            torch.cuda.synchronize()
    clocks_timed = sampler.summary() if rank == 0 else None
    if clocks_timed is not None and args.steps * (ms / args.steps) < 150.0:
        clocks_timed["note"] = "timed region < 150 ms: sampled while the same step kept running right after it"
    # ---- timed region 2: end to end through the reference-facing API with host batches
    e2e = None
    if not args.no_e2e:
        # host batches in page-locked memory (made once, outside the timed region); every timed step copies them
        # host -> device and reads the step's result back
        real_p, seg_p, mask_p = (torch.from_numpy(x).pin_memory() for x in (real_h, seg_h, mask_h))
        if infer:
            out_host = torch.empty((B, H, W, 3), dtype=torch.float32).pin_memory()

            def step_e2e():
                out_host.copy_(model.generate_test_images(real_p), non_blocking=True)
                torch.cuda.current_stream().synchronize()  # the caller reads the image (model.py:561-567 saves it)
            h2d, d2h = int(real_h.nbytes), int(out_host.numel() * 4)
        else:
            model.real_A, model.seg_A, model.mask_A = real_p, seg_p, mask_p

            def step_e2e():
                model.train_step(ns)
                # the reference prints both losses every step (model.py:260): every step's two floats come back to the
                # host through an asynchronous copy and are READ one step late, so the next batch's H2D copy (copy
                # stream, double-buffered device inputs) runs underneath this step's kernels
                return model.losses_host(lag=1)
            h2d, d2h = int(real_h.nbytes + seg_h.nbytes + mask_h.nbytes), 8
        for _ in range(2):
            step_e2e()
        sync_all()
        e0.record()
        for _ in range(args.steps):
            step_e2e()
        if not infer:
            last = model.losses_host(lag=0)  # the last step's losses are read inside the timed region too
            assert last is not None and all(x == x for x in last)
        e1.record()
        sync_all()
        ms2 = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": B * world * args.steps / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms2 / args.steps}
        if not infer:
            model.real_A, model.seg_A, model.mask_A = real_d, seg_d, mask_d
    # ---- profiling passes (separate from the timed regions: the events sit between the step's launches): the
    #      residual-block convolution (tensor roofline) and the two instance-norm passes around it (HBM roofline)
    prof = {}
    if not infer:
        for kind in (0, 1, 2, 3):
            eng.profile_begin(18 * max(4, min(args.steps, 10)) + 8, kind=kind)
            for _ in range(max(4, min(args.steps, 10))):
                step()
            sync_all()
            prof[kind] = eng.profile_end()  # (ms, launches, work per launch)
    else:
        eng.profile_begin(18 * 12 + 8, kind=0)
        for _ in range(10):
            step()
        sync_all()
        prof[0] = eng.profile_end()
        eng.profile_begin(18 * 12 + 8, kind=1)
        for _ in range(10):
            step()
        sync_all()
        prof[1] = eng.profile_end()
    # ---- sustained regime: the same step for several seconds (the 20-step region above is a 0.2 s burst that never
    #      reaches the 1 kW power cap); its own clock summary
    sustained = None
    if args.sustained_seconds > 0 and not infer:
        n_s = max(args.steps, int(args.sustained_seconds * 1e3 / (ms / args.steps)))
        if rank == 0:
            sampler.rows = []
        sync_all()
        e0.record()
        for _ in range(n_s):
            step()
        e1.record()
        sync_all()
        ms3 = max_over_ranks(e0.elapsed_time(e1))
        eng.profile_begin(18 * 12 + 8, kind=0)
        for _ in range(10):
            step()
        sync_all()
        c_ms, c_n, c_fl = eng.profile_end()
        sustained = {"steps": n_s, "seconds": ms3 * 1e-3, "value": B * world * n_s / (ms3 * 1e-3), "unit": UNIT,
                     "ms_per_step": ms3 / n_s, "clocks": sampler.summary() if rank == 0 else None,
                     "conv_tflops_after": (c_fl * c_n / (c_ms * 1e-3) / 1e12) if c_ms > 0 else None}
    sampler.stop_flag = True
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_burst = peaks.get("bf16_tflops", 1590.0)
    peak_hbm = peaks.get("hbm_gbs", 6650.0)
    src = "MEASURED_PEAKS.json (of measured)" if peaks else "B200_PROFILING.md fallback (of fallback)"
    conv_ms, conv_n, conv_flops = prof[0]
    achieved = conv_flops * conv_n / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
    ms_step = ms / args.steps
    value = B * world * args.steps / (ms * 1e-3)
    gf = GFLOP_PER_IMG.get((H, W, C))
    # ncu --set full capture of the same kernel (dram__bytes_read.sum + dram__bytes_write.sum per launch); committed
    # under profiles/, valid for the c3 workload only -- not measured by this run
    traffic_file = os.path.join(ROOT, "profiles", "r02_ncu_conv_res.json")
    traffic, traffic_src = None, None
    if (H, W, B) == (256, 512, 8) and os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/r02_ncu_conv_res.json (ncu --set full, not measured in this run)"
        except Exception:
            pass

    def hbm_obj(kind, name):
        if kind not in prof or prof[kind][0] <= 0:
            return None
        t_ms, n_l, bytes_l = prof[kind]
        gbs = bytes_l * n_l / (t_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm, "traffic": None,
                "kernel": "%s (%d launches timed with CUDA events in a separate pass; algorithmic %.1f MB per launch, "
                          "%.1f us per launch)" % (name, n_l, bytes_l / 1e6, t_ms * 1e3 / n_l), "peak_source": src}

    if infer:
        metric = "G inference img/s at %dx%d" % (H, W)
        workload = ("SG-GAN generator_resnet (9 blocks) inference A->B, %dx%d, batch %d (BASELINE config 2: forward conv / "
                    "deconv + instance_norm path)" % (H, W, B))
    else:
        metric = METRIC if (H, W) == (256, 512) else "G+D train img/s at %dx%d" % (H, W)
        workload = ("SG-GAN full G+D train step (generator_resnet 9 blocks + semantic-aware D, fwd+bwd+Adam), "
                    "%dx%d, batch %d per GPU, C=%d, loss %s" % (H, W, B, C, args.loss_mode))
    line = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "working set %.1f GB per step >> 126 MB L2 (no flush needed)" % (L.workspace_bytes(eng.cfg) / 2 ** 30)},
        "clocks": clocks_timed, "e2e": e2e, "gpu_launches": launches * args.steps,
        "losses_last_step": losses,
        "step_tflops": (gf * B / ms_step) if (gf and not infer) else None,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved / peak_tf) if achieved else None,
                     "frac_of_burst_peak": (achieved / peak_burst) if achieved else None, "peak_burst": peak_burst,
                     "traffic": traffic, "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
                     "kernel": "conv_gemm_pair_kernel (cta_group::2), residual-block 3x3 256->256 forward (%d launches timed with CUDA events "
                               "in a separate pass after the timed region; algorithmic %.2f GFLOP per launch)" % (conv_n, conv_flops / 1e9),
                     "peak_source": "bf16_tflops_sustained / bf16_tflops of " + src},
        "roofline_hbm": [x for x in (hbm_obj(1, "row_stream_kernel<APPLY>: instance norm + ReLU + reflect-pad frame behind the first "
                                                 "conv of each residual block, read Y + write X"),
                                     hbm_obj(2, "row_stream_kernel<BWD_REDUCE> + <BWD_APPLY>: instance-norm backward of the same layers (two "
                                                "launches timed together, side stream joined first); algorithmic = read Y + read dX + "
                                                "write dY, the reduce pass re-reads Y and dX on top of that"),
                                     hbm_obj(3, "loss kernels of the generator side, one group per step: " +
                                             ("seg_edge_weight + gradloss (tiled Sobel loss and its gradient) + " if args.loss_mode == "sggan" else "") +
                                             "fake_grad (L1 sign + GAN gradient, tanh') + finalize_losses; algorithmic = every input "
                                             "read once, every output written once")) if x],
        "sustained": sustained,
    }
    if not args.no_cpu_baseline and world == 1 and not infer:
        line["cpu_baseline"] = cpu_baseline(H, W, C, 3, 1)[0]
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
