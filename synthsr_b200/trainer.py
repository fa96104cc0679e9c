"""One training step of SynthSR.training.training() (SynthSR/training.py:330-453) on the B200 engine:
on-the-fly generator -> U-Net forward/backward -> (data-parallel gradient all-reduce) -> Adam.

Data parallelism (one process per GPU, torch.distributed / NCCL): every rank generates and trains on its own
mini-batch shard; the ONLY exchange per step is one all-reduce of the flat gradient buffer (the BN moving statistics
ride at its tail so all replicas keep identical state).  BN batch statistics stay rank-local (equals the reference's
batchsize-per-GPU semantics per shard; documented deviation from a single big batch).
"""
import contextlib
import os

import numpy as np
import torch

from .draws import sample_draws
from .generator import SynthGenerator
from .unet import UNet3D


# when a list, every engine appends a timing CUDA event after each trained step (bench.py measures the throughput of the
# drop-in SynthSR.training.training() call from them, without touching its code path)
STEP_EVENTS = None
_NVTX = os.environ.get('SSR_NVTX') == '1'


@contextlib.contextmanager
def _range(name):
    """NVTX range (SSR_NVTX=1): generator / forward+backward / exchange / optimiser show up as named spans in nsys."""
    if _NVTX:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if _NVTX:
            torch.cuda.nvtx.range_pop()


class TrainingEngine:
    def __init__(self, plan, batchsize=1, nb_features=24, nb_levels=5, conv_size=3, feat_mult=2, nb_conv_per_level=2,
                 nb_labels=None, lr=1e-4, lr_decay=0., metric='l1', work_with_residual_channel=None,
                 loss_cropping=None, conv_impl=None, seed=0, device='cuda', rank=0, world_size=1, seg=None, net_cls=None,
                 net_kwargs=None):
        """seg: optional synthsr_b200.seg_loss.SegRegulariser (segmentation-regularised loss, metrics_model.py:136-215).
        net_cls / net_kwargs: a UNet3D subclass (and extra keyword arguments) to train instead, e.g. the adversarial
        fine-tuner's network (synthsr_b200/adversary.py)."""
        # conv_impl None: SSR_CONV_IMPL or 'tc3' (compensated forward, the parity-gated mode); see synthsr_b200/unet.py
        conv_impl = conv_impl or os.environ.get('SSR_CONV_IMPL', 'tc3')
        self.plan, self.B = plan, int(batchsize)
        self.device = torch.device(device)
        self.rank, self.world = int(rank), int(world_size)
        self.gen = SynthGenerator(plan, batchsize, device)
        nb_labels = plan.n_target_channels if nb_labels is None else nb_labels
        self.seg = seg
        if net_cls is not None:
            self.net = net_cls(plan.image_shape, nb_features, nb_levels, conv_size, nb_labels, feat_mult, nb_conv_per_level,
                               batchsize, device, conv_impl, seed=seed, seg=seg, **(net_kwargs or {}))
        elif seg is None:
            self.net = UNet3D(plan.image_shape, nb_features, nb_levels, conv_size, nb_labels, feat_mult, nb_conv_per_level,
                              batchsize, device, conv_impl, seed=seed)    # same seed on every rank: identical replicas
        else:
            from .seg_loss import SegRegularisedUNet3D
            assert plan.crop_shape == plan.output_shape, 'the segmentation target lives on the crop grid (target_res = atlas_res)'
            self.net = SegRegularisedUNet3D(plan.image_shape, nb_features, nb_levels, conv_size, nb_labels, feat_mult,
                                            nb_conv_per_level, batchsize, device, conv_impl, seed=seed, seg=seg)
        self.lr, self.lr_decay, self.metric = lr, lr_decay, metric
        self.residual, self.loss_cropping = work_with_residual_channel, loss_cropping
        self.rng = np.random.default_rng(seed * 1000003 + 7919 * self.rank)   # per-rank augmentation stream
        self.seed = seed * 65537 + self.rank
        self.steps = 0
        self.exchange = None
        if self.world > 1:
            self.exchange = GradientExchange(self.net, self.world)
        # pipelined mode (train_step_pipelined): a second generator instance and a generator stream
        self._gens, self._gen_stream, self._pending, self._pipe_i = None, None, None, 0

    def train_step(self, labels, means, stds, real_image=None, draws=None):
        """labels: int32 cuda [B, *labels_shape]; means/stds [B, L, C] host arrays.  Returns the loss (1-element cuda
        tensor, float64; averaged over ranks when world_size > 1)."""
        if draws is None:
            draws = sample_draws(self.rng, self.plan, self.B)
        with _range('generator'):
            image, target = self.gen.run(labels, means, stds, draws, real_image=real_image, seed=self.seed)
        if hasattr(self.net, 'seg_labels'):
            self.net.seg_labels = self.gen.labels                         # `segmentation_target` of this batch
        return self._train_on(image, target)

    def _train_on(self, image, target):
        with _range('unet fwd+bwd'):
            loss = self.net.loss_and_grad(image, target, self.metric, self.residual, self.loss_cropping)
        scale = 1.
        if self.world > 1:
            with _range('gradient exchange'):
                loss = self._allreduce(loss)
            scale = 1. / self.world
        with _range('adam'):
            self.net.adam_step(self.lr, self.lr_decay, grad_scale=scale)
        self.steps += 1
        if STEP_EVENTS is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            STEP_EVENTS.append(ev)
        return loss

    # -----------------------------------------------------------------------------------------------------------------
    def train_step_pipelined(self, labels, means, stds, real_image=None, draws=None):
        """Same work per call as train_step -- one generator pass, one U-Net training pass -- but software-pipelined like
        the reference's `fit_generator` queue (SynthSR/training.py:449-453: the Keras generator runs ahead of the
        optimiser): the batch passed in is generated on a second stream while the network trains on the batch of the
        PREVIOUS call, so the latency-bound generator kernels fill the SMs the backward chain leaves idle.  Batches are
        trained exactly once, in order.  Returns the loss of the previous call's batch (None on the first call);
        `flush()` trains the last pending batch.  labels / real_image may be pinned host tensors (copied on the generator
        stream into per-slot device buffers, so nothing the caller allocated is read across streams)."""
        if self._gens is None:
            self._gens = [self.gen, SynthGenerator(self.plan, self.B, self.device)]
            self._gen_stream = torch.cuda.Stream(device=self.device)
            self._gen_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._lab_dev, self._real_dev = [None, None], [None, None]
        i = self._pipe_i
        k = i % 2
        cur = torch.cuda.current_stream()
        if draws is None:
            draws = sample_draws(self.rng, self.plan, self.B)
        # the generator of this call re-uses the buffers the training pass of call i-1 (batch i-2) read, and may read
        # tensors the caller just produced on the current stream
        self._gen_stream.wait_stream(cur)
        with torch.cuda.stream(self._gen_stream):
            if not labels.is_cuda:
                if self._lab_dev[k] is None:
                    self._lab_dev[k] = torch.empty(labels.shape, dtype=torch.int32, device=self.device)
                self._lab_dev[k].copy_(labels, non_blocking=True)
                labels = self._lab_dev[k]
            if real_image is not None and not real_image.is_cuda:
                if self._real_dev[k] is None:
                    self._real_dev[k] = torch.empty(real_image.shape, dtype=torch.float32, device=self.device)
                self._real_dev[k].copy_(real_image, non_blocking=True)
                real_image = self._real_dev[k]
            self._gens[k].philox_step = max(g.philox_step for g in self._gens)    # one noise-counter sequence for both
            with _range('generator'):
                image, target = self._gens[k].run(labels, means, stds, draws, real_image=real_image, seed=self.seed)
            self._gen_done[k].record()
        loss = self._train_pending()
        self._pending = (image, target, k)
        self._pipe_i += 1
        return loss

    def _train_pending(self):
        if self._pending is None:
            return None
        image, target, k = self._pending
        self._pending = None
        torch.cuda.current_stream().wait_event(self._gen_done[k])
        if hasattr(self.net, 'seg_labels'):
            self.net.seg_labels = self._gens[k].labels
        return self._train_on(image, target)

    def flush(self):
        """train on the batch generated by the last train_step_pipelined call (end of an epoch / of training)."""
        return self._train_pending()

    def _allreduce(self, loss):
        return self.exchange.finish(loss)


def exchange_inplace(comm, n_grads, split, world, stage):
    """The data-parallel exchange of one step on the buffer comm = [gradients (n_grads) | BN moving stats | loss], IN PLACE
    (no staging copy), as one logical SUM all-reduce issued in two pieces so that the first overlaps the backward pass:
      stage 0: comm[:split]  -- the gradient prefix that is complete once the deep levels are differentiated (the buffer
               is laid out in backward-completion order, UNet3D.__init__; the deep layers hold > 90 % of the bytes);
      stage 1: comm[split:]  -- the shallow levels' gradients, the moving statistics and the loss; the tail
               [n_grads:] is then averaged (gradients stay SUMMED: Adam applies the 1/world scale).
    Device agnostic (NCCL on the GPUs, gloo in the CPU tests); stream placement is the caller's business."""
    import torch.distributed as dist
    if stage == 0:
        if split > 0:
            dist.all_reduce(comm[:split], op=dist.ReduceOp.SUM)
        return
    dist.all_reduce(comm[split:], op=dist.ReduceOp.SUM)
    comm[n_grads:].mul_(1. / world)


class GradientExchange:
    """Issues exchange_inplace for a UNet3D: stage 0 from the network's grads_ready_hook (on a communication stream, behind
    events of the backward-chain stream and the weight-gradient side stream), stage 1 after the backward pass."""

    def __init__(self, net, world, split_level=None):
        self.net, self.world = net, int(world)
        L = net.L
        # the prefix is sent once encoder level `split_level` is done: levels below it (>= 85 % of the backward time of
        # the 160^3 net is still to come at level 2) hide the transfer of the deep layers' gradients
        if split_level is None and os.environ.get('SSR_EXCHANGE_SPLIT_LEVEL'):      # 0: one all-reduce after the backward pass
            split_level = int(os.environ['SSR_EXCHANGE_SPLIT_LEVEL'])
        self.split_level = min(2, L - 1) if split_level is None else int(split_level)
        self.split = net.level_end.get(self.split_level, 0) if self.split_level > 0 else 0
        self.cuda = net.comm.is_cuda
        self.stream = torch.cuda.Stream(device=net.device) if self.cuda else None
        self._sent = False
        if self.split > 0:
            net.grads_ready_hook = self._on_level

    def _on_level(self, level):
        if level != self.split_level:
            return
        net = self.net
        if self.cuda:
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)                       # bias / BatchNorm gradients + data-gradient chain
            if net._side is not None:
                self.stream.wait_stream(net._side)             # weight gradients enqueued so far
            with torch.cuda.stream(self.stream):
                exchange_inplace(net.comm, net.n_params, self.split, self.world, 0)
        else:
            exchange_inplace(net.comm, net.n_params, self.split, self.world, 0)
        self._sent = True

    def finish(self, loss):
        """after loss_and_grad: exchanges the rest and returns the mean loss (1-element float64 tensor)."""
        net = self.net
        split = self.split if self._sent else 0
        self._sent = False
        net.comm[-1:].copy_(loss.reshape(-1)[:1])
        if self.cuda:
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                exchange_inplace(net.comm, net.n_params, split, self.world, 1)
            cur.wait_stream(self.stream)
        else:
            exchange_inplace(net.comm, net.n_params, split, self.world, 1)
        return net.comm[-1:].double()
