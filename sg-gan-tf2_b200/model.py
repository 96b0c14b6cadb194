"""model.py -- drop-in for the reference trainer's hot path (model.py:39-200,450-567).

`sggan(args)` keeps the reference's attribute and method names: `.generator`, `.discriminator`,
`.train_step(args)` (reads `.real_A/.seg_A/.mask_A`, sets `.fake_A/.gen_loss/.disc_loss`),
`.gen_loss_p2p`, `.disc_loss_p2p`, `.generator_loss`, `.discriminator_loss`,
`.generate_test_images`, `.save`, `.load`, `.train`, `.test`.  One call of `train_step` is ONE call
into libsggan_sm100.so (forward, both backward passes, Adam) -- no TensorFlow, no autograd tape.

Deliberate deviations: D5 -- `fake_A = G(real_A)` every step (the reference's
concat-with-previous-fake branch only survives step 2 at batch 10); D6 -- `lr` is live but
defaults to the reference-effective 0.001; D9 -- training uses the `use_resnet` generator; `generator_unet` (the CLI default)
runs forward only.
Data loading / eval / TensorBoard (model.py:202-448) are host orchestration outside the path:
`train()` consumes an iterable of numpy batches instead of globbing PNGs.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from . import module
from .module import (abs_criterion, discriminator, generator_resnet, generator_unet, gradloss_criterion,  # noqa: F401
                     mae_criterion, sce_criterion, tf_kernel_prep_3d)


class sggan(object):
    def __init__(self, args):
        self.batch_size = args.batch_size
        self.image_width = args.image_width
        self.image_height = args.image_height
        self.input_c_dim = getattr(args, "input_nc", 3)
        self.output_c_dim = getattr(args, "output_nc", 3)
        self.L1_lambda = getattr(args, "L1_lambda", 10.0)
        self.Lg_lambda = getattr(args, "Lg_lambda", 5.0)
        self.dataset_dir = getattr(args, "dataset_dir", "city")
        self.segment_class = getattr(args, "segment_class", 34)
        r = getattr(args, "ratio_gan2seg", 10)
        self.alpha_recip = 1. / r if r > 0 else 0
        self.use_pix2pix = getattr(args, "use_pix2pix", False)
        self.use_resnet = bool(getattr(args, "use_resnet", True))
        if self.use_pix2pix:
            raise L.SgganError("the pix2pix generator / discriminator pair (BatchNorm, 4x4 stride-2, Dropout) is not built "
                               "(SURVEY D9 / 8(f) row f4)")
        self.use_lsgan = bool(getattr(args, "use_lsgan", True))
        self.criterionGAN = mae_criterion if self.use_lsgan else sce_criterion
        # loss_mode "p2p" = what train_step actually calls (model.py:190-191); "sggan" = the defined-but-unwired
        # SG-GAN losses (model.py:114-133) + gradient-sensitive term
        self.loss_mode = getattr(args, "loss_mode", "p2p")
        self.lr = getattr(args, "lr_effective", 0.001)  # model.py:82,205: hard-coded 0.001 (args.lr unused)
        self.beta1 = getattr(args, "beta1", 0.5)
        self.discriminator = discriminator(self.image_height, self.image_width, getattr(args, "ndf", 64), self.segment_class)
        # use_resnet=False is the reference CLI's default (main.py:39): generator_unet, forward only here -- sampling and
        # testing work, train_step needs the ResNet generator (the fused engine, which is what BASELINE names)
        self.generator = generator_resnet(self.image_height, self.image_width, getattr(args, "ngf", 64), self.output_c_dim) \
            if self.use_resnet else module.generator_unet(getattr(args, "ngf", 64), self.output_c_dim)
        self.kernels = [tf_kernel_prep_3d(np.array([[0, 0, 0], [-1, 0, 1], [0, 0, 0]]), self.input_c_dim),
                        tf_kernel_prep_3d(np.array([[0, -1, 0], [0, 0, 0], [0, 1, 0]]), self.input_c_dim)]
        self.kernel = np.stack(self.kernels, axis=-1).astype(np.float32)  # "DerivKernel_seg" (model.py:109-112)
        self.weighted_seg_A = []
        self.real_A = self.seg_A = self.mask_A = self.fake_A = None
        self.gen_loss = self.disc_loss = None
        self.runtime = None
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._pinned = {}
        self._dev = {}
        self._copy_stream = None
        self._slot = 1
        self._uploaded = False
        self._h2d_done = [None, None]
        self._consumed = [None, None]
        self._loss_ev = [None, None]
        self._loss_host = None
        self._steps_done = 0

    # ---- plan -----------------------------------------------------------------------------------------
    def _ensure_runtime(self, B, H, W, mask_hw):
        rt = self.runtime
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world != self.world_size and rt is not None:
            raise L.SgganError("the process group changed after the first train_step (world size %d -> %d)" %
                               (self.world_size, world))
        self.world_size = world  # read lazily: init_process_group may follow the constructor
        if rt is not None and (rt.cfg.batch, rt.cfg.image_height, rt.cfg.image_width, rt.cfg.mask_height,
                               rt.cfg.mask_width) == (B, H, W, mask_hw[0], mask_hw[1]):
            # the training runtime must own both networks' master weights (a stand-alone G(x) / D([x, m]) call at
            # another shape never takes ownership, but a first call before training may have bound another plan)
            if self.generator.runtime is not rt:
                self.generator.bind(rt)
            if self.discriminator.runtime is not rt:
                self.discriminator.bind(rt)
            return rt
        rt = module.Runtime(B, H, W, segment_class=self.segment_class, n_blocks=self.generator.n_blocks,
                            mask_hw=mask_hw, loss_mode=L.LOSS_P2P if self.loss_mode == "p2p" else L.LOSS_SGGAN,
                            use_lsgan=int(self.use_lsgan), lr=self.lr, beta1=self.beta1, L1_lambda=self.L1_lambda,
                            Lg_lambda=self.Lg_lambda, world_size=self.world_size)
        self.generator.bind(rt)       # carries Adam m / v and the step count over from a previous plan
        self.discriminator.bind(rt)
        if self.world_size > 1:  # identical replicas: broadcast rank 0's weights
            for net in (L.NET_G, L.NET_D):
                dist.broadcast(rt.engine.flat(net, 0), src=0)
            rt.engine.weights_changed()
        self.runtime = rt
        return rt

    # ---- host batches -> device: double-buffered, on a copy stream -------------------------------------------
    # Step k's three host arrays are copied into device slot k % 2 by a copy stream that only waits for step k-2 (the
    # previous user of that slot), so the copy of step k+1 runs underneath the kernels of step k whenever the caller
    # does not block in between -- read the losses with losses_host(lag=1) for that.  Each slot has stable device
    # pointers, i.e. its own captured step graph (the library keeps up to four).
    def _begin_uploads(self):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        self._slot = (self._slot + 1) & 1
        self._uploaded = False

    def _upload(self, name, x):
        """One input of the step: device tensors pass through; host data goes to this step's device slot."""
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.float().contiguous()
        slot = self._slot
        pinned = isinstance(x, torch.Tensor) and x.dtype == torch.float32 and x.is_contiguous() and x.is_pinned()
        if not pinned:
            # pageable source: stage it in a persistent page-locked buffer of this slot (multi-threaded host copy)
            x = torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32))) if not isinstance(x, torch.Tensor) \
                else x.float().contiguous()
            buf = self._pinned.get((name, slot))
            if buf is None or buf.shape != x.shape:
                buf = torch.empty(tuple(x.shape), dtype=torch.float32).pin_memory()
                self._pinned[(name, slot)] = buf
            ev = self._h2d_done[slot]
            if ev is not None:
                ev.synchronize()  # the copy out of this staging buffer two steps ago has finished
            buf.copy_(x)
            x = buf
        dev = self._dev.get((name, slot))
        if dev is None or dev.shape != x.shape:
            dev = torch.empty(tuple(x.shape), dtype=torch.float32, device="cuda")
            self._dev[(name, slot)] = dev
        cs = self._copy_stream
        if not self._uploaded:
            self._uploaded = True
            if self._consumed[slot] is not None:
                cs.wait_event(self._consumed[slot])  # the step that last read this slot
            else:
                cs.wait_stream(torch.cuda.current_stream())  # first use: order after the allocation
        with torch.cuda.stream(cs):
            dev.copy_(x, non_blocking=True)
        return dev

    def _end_uploads(self):
        if self._uploaded:
            ev = self._h2d_done[self._slot] or torch.cuda.Event()
            ev.record(self._copy_stream)
            self._h2d_done[self._slot] = ev
            torch.cuda.current_stream().wait_event(ev)

    def _step_enqueued(self, eng):
        """After the step's launches: mark the input slot as consumed and start the losses' way back to the host."""
        slot = self._slot
        cur = torch.cuda.current_stream()
        if self._uploaded:
            ev = self._consumed[slot] or torch.cuda.Event()
            ev.record(cur)
            self._consumed[slot] = ev
        if self._loss_host is None:
            self._loss_host = torch.zeros((2, 2), dtype=torch.float32).pin_memory()
        if self._loss_ev[slot] is not None:
            self._loss_ev[slot].synchronize()  # nobody may still be waiting for the values of two steps ago
        self._loss_host[slot].copy_(eng.losses, non_blocking=True)
        ev = self._loss_ev[slot] or torch.cuda.Event()
        ev.record(cur)
        self._loss_ev[slot] = ev
        self._steps_done += 1

    def losses_host(self, lag=0):
        """(gen_loss, disc_loss) of the step `lag` steps back as Python floats, through an asynchronous copy into
        page-locked memory that was enqueued right behind that step.  lag=0 waits for the step just enqueued (what
        float(self.gen_loss) does); lag=1 returns the previous step's values while the current one runs -- the train
        loop's per-step print (model.py:260) then costs no pipeline bubble.  None if that step does not exist."""
        if lag not in (0, 1) or self._steps_done <= lag:
            return None
        slot = (self._slot - lag) & 1
        self._loss_ev[slot].synchronize()
        return float(self._loss_host[slot, 0]), float(self._loss_host[slot, 1])

    # ---- the hot path ---------------------------------------------------------------------------------
    def train_step(self, args=None):
        """model.py:169-200.  Inputs come from self.real_A / self.seg_A / self.mask_A exactly as in the
        reference's train loop (model.py:249-256)."""
        if not self.use_resnet:
            raise L.SgganError("train_step: the fused training engine is built for generator_resnet (use_resnet=True); "
                               "generator_unet runs forward only (module.GeneratorUnet)")
        self._begin_uploads()
        real_A = self._upload("real_A", self.real_A)
        seg_A = self._upload("seg_A", self.seg_A)
        mask_A = self._upload("mask_A", self.mask_A)
        self._end_uploads()
        B, H, W, _ = real_A.shape
        rt = self._ensure_runtime(B, H, W, (int(mask_A.shape[1]), int(mask_A.shape[2])))
        eng = rt.engine
        if self.world_size == 1:
            eng.train_step(real_A, seg_A, mask_A)
        else:
            # data parallel: D gradients are final after phase 1 and are all-reduced on NCCL's stream while
            # the generator backward runs; G gradients follow; Adam applies 1/world_size.
            eng.step_forward_backward_d(real_A, seg_A, mask_A)
            hd = dist.all_reduce(eng.flat(L.NET_D, 1), async_op=True)
            gg = eng.flat(L.NET_G, 1)
            if os.environ.get("SGGAN_DP_BUCKETS", "0") == "1":
                # opt-in: G's gradients in two buckets -- the upper half of the network (final after the first half of the
                # backward: a contiguous tail of the flat buffer) is reduced underneath the second half of the backward.
                # Correct (tests) and measured once on 2 GPUs (1871 img/s); the single-bucket form below is the one every
                # multi-GPU number in profiles/ was taken with, so it stays the default.
                off = eng.grad_split_offset()
                eng.step_backward_g(part=0)
                hs = [dist.all_reduce(gg[off:], async_op=True)]
                eng.step_backward_g(part=1)
                hs.append(dist.all_reduce(gg[:off], async_op=True))
            else:
                eng.step_backward_g()
                hs = [dist.all_reduce(gg, async_op=True)]
            hd.wait()                                  # long finished: it ran underneath the generator backward
            eng.step_adam(L.NET_D, overlapped=True)    # side stream: runs while G's all-reduce is in flight
            for hnd in hs:
                hnd.wait()
            eng.step_adam(L.NET_G)                     # joins the side stream
        self._step_enqueued(eng)
        # views into buffers the NEXT step overwrites (clone to keep); float(self.gen_loss) synchronises, losses_host() is
        # the asynchronous way
        self.fake_A = eng.last_fake()
        self.gen_loss, self.disc_loss = eng.losses[0], eng.losses[1]
        return self.gen_loss, self.disc_loss

    def generate_test_images(self, sample_imgA):
        """model.py:528-532."""
        return self.generator(sample_imgA)

    # ---- losses as callables (model.py:114-166); train_step fuses them, these serve API parity ---------------
    def gen_loss_p2p(self, DA_fake, fake_A, seg_A):
        LAMBDA = 100
        DA_fake = L.as_cuda_f32(DA_fake)
        gan_loss = sce_criterion(DA_fake, torch.ones_like(DA_fake))
        return gan_loss + LAMBDA * abs_criterion(seg_A, fake_A)

    def disc_loss_p2p(self, DA_real, DA_fake):
        DA_real, DA_fake = L.as_cuda_f32(DA_real), L.as_cuda_f32(DA_fake)
        return sce_criterion(DA_real, torch.ones_like(DA_real)) + sce_criterion(DA_fake, torch.zeros_like(DA_fake))

    def generator_loss(self, DA_fake, args):
        import ctypes as C
        seg = L.as_cuda_f32(self.seg_A)
        B, H, W, _ = seg.shape
        w = torch.empty((B, H, W, 1), dtype=torch.float32, device=seg.device)
        L.check(L.lib().sggan_seg_edge_weight(C.c_void_p(seg.data_ptr()), C.c_void_p(w.data_ptr()), B, H, W, L.stream_ptr()))
        self.weighted_seg_A = w
        DA_fake = L.as_cuda_f32(DA_fake)
        return self.criterionGAN(DA_fake, torch.ones_like(DA_fake)) + args.L1_lambda * abs_criterion(self.real_A, self.fake_A)

    def discriminator_loss(self, DA_real, DA_fake_sample):
        DA_real, DA_fake_sample = L.as_cuda_f32(DA_real), L.as_cuda_f32(DA_fake_sample)
        return (self.criterionGAN(DA_real, torch.ones_like(DA_real)) +
                self.criterionGAN(DA_fake_sample, torch.zeros_like(DA_fake_sample))) / 2

    # ---- orchestration around the path ---------------------------------------------------------------------
    def train(self, args, batches=None):
        """model.py:202-275 with the PNG loader replaced by `batches`: an iterable (or a callable taking the
        epoch) of (real_A, seg_A, mask_A) numpy/torch batches."""
        if batches is None:
            raise L.SgganError("train(): pass `batches` -- the PNG/skimage loader (utils.py:167-233) is outside the "
                               "accelerated path (SURVEY 8(f) row f1)")
        start_time = time.time()
        if getattr(args, "continue_train", False):
            print(" [*] Load SUCCESS" if self.load(args.checkpoint_dir) else " [!] Load failed...")
        epoch = 0
        try:
            for epoch in range(args.epoch):
                it = batches(epoch) if callable(batches) else batches
                for idx, (a, s, m) in enumerate(it):
                    self.real_A, self.seg_A, self.mask_A = a, s, m
                    self.train_step(args)
                    # the reference prints both losses every step (model.py:260); here the line is one step late so that
                    # the host never waits for the step it has just enqueued
                    got = self.losses_host(lag=1)
                    if got is not None:
                        print("Epoch: [%2d] [%4d] time: %4.4f Gen_Loss: %f Disc_Loss: %f " % (
                            epoch, idx - 1, time.time() - start_time, got[0], got[1]))
        finally:
            if getattr(args, "checkpoint_dir", None):
                self.save(args.checkpoint_dir, epoch)

    def save(self, checkpoint_dir, ep, optimizer=True):
        """model.py:450-468: `<dir>/<dataset>/gen/cp-NNNN.ckpt` and `…/disc/cp-NNNN.ckpt` as TF checkpoints (the format
        Keras' save_weights writes there, tf_checkpoint.py); plus, unlike the reference, the Adam state beside them."""
        path = "%s/%s" % (checkpoint_dir, self.dataset_dir)
        for sub in ("gen", "disc"):
            os.makedirs(os.path.join(path, sub), exist_ok=True)
        opt = optimizer and self.runtime is not None
        self.generator.save_weights(os.path.join(path, "gen/cp-%04d.ckpt" % ep), optimizer=opt)
        self.discriminator.save_weights(os.path.join(path, "disc/cp-%04d.ckpt" % ep), optimizer=opt)

    def load(self, checkpoint_dir):
        """model.py:471-503: the latest checkpoint of both nets (tf.train.latest_checkpoint = the directory's `checkpoint`
        state file; .npz files of earlier versions of this package are still found), False if either is missing."""
        from . import tf_checkpoint
        path = "%s/%s" % (checkpoint_dir, self.dataset_dir)

        def latest(sub):
            d = os.path.join(path, sub)
            p = tf_checkpoint.latest_checkpoint(d) if os.path.isdir(d) else None
            if p:
                return p
            fs = sorted(f for f in os.listdir(d) if f.endswith(".ckpt.npz")) if os.path.isdir(d) else []
            return os.path.join(d, fs[-1]) if fs else None

        g, d = latest("gen"), latest("disc")
        if g and d:
            self.generator.load_weights(g)
            self.discriminator.load_weights(d)
            return True
        return False

    def test(self, args, images=None):
        """model.py:535-567 without file IO: returns G(x) for each image (x in [0,255] like model.py:555-561)."""
        if not self.load(args.checkpoint_dir):
            print(" [!] Load failed...")
        return [self.generator(np.asarray(im, dtype=np.float32)[None]) for im in (images or [])]
