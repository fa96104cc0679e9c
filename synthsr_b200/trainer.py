"""One training step of SynthSR.training.training() (SynthSR/training.py:330-453) on the B200 engine:
on-the-fly generator -> U-Net forward/backward -> (data-parallel gradient all-reduce) -> Adam.

Data parallelism (one process per GPU, torch.distributed / NCCL): every rank generates and trains on its own
mini-batch shard; the ONLY exchange per step is one all-reduce of the flat gradient buffer (the BN moving statistics
ride at its tail so all replicas keep identical state).  BN batch statistics stay rank-local (equals the reference's
batchsize-per-GPU semantics per shard; documented deviation from a single big batch).
"""
import numpy as np
import torch

from .draws import sample_draws
from .generator import SynthGenerator
from .unet import UNet3D


class TrainingEngine:
    def __init__(self, plan, batchsize=1, nb_features=24, nb_levels=5, conv_size=3, feat_mult=2, nb_conv_per_level=2,
                 nb_labels=None, lr=1e-4, lr_decay=0., metric='l1', work_with_residual_channel=None,
                 loss_cropping=None, conv_impl='tc3', seed=0, device='cuda', rank=0, world_size=1, seg=None):
        """seg: optional synthsr_b200.seg_loss.SegRegulariser (segmentation-regularised loss, metrics_model.py:136-215)."""
        self.plan, self.B = plan, int(batchsize)
        self.device = torch.device(device)
        self.rank, self.world = int(rank), int(world_size)
        self.gen = SynthGenerator(plan, batchsize, device)
        nb_labels = plan.n_target_channels if nb_labels is None else nb_labels
        self.seg = seg
        if seg is None:
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
        if self.world > 1:
            n_mv = sum(t.numel() for t in self.net.moving.values())
            self.flat = torch.zeros(self.net.n_params + n_mv + 1, dtype=torch.float32, device=self.device)
        # pipelined mode (train_step_pipelined): a second generator instance and a generator stream
        self._gens, self._gen_stream, self._pending, self._pipe_i = None, None, None, 0

    def train_step(self, labels, means, stds, real_image=None, draws=None):
        """labels: int32 cuda [B, *labels_shape]; means/stds [B, L, C] host arrays.  Returns the loss (1-element cuda
        tensor, float64; averaged over ranks when world_size > 1)."""
        if draws is None:
            draws = sample_draws(self.rng, self.plan, self.B)
        image, target = self.gen.run(labels, means, stds, draws, real_image=real_image, seed=self.seed)
        if self.seg is not None:
            self.net.seg_labels = self.gen.labels                         # `segmentation_target` of this batch
        loss = self.net.loss_and_grad(image, target, self.metric, self.residual, self.loss_cropping)
        scale = 1.
        if self.world > 1:
            loss = self._allreduce(loss)
            scale = 1. / self.world
        self.net.adam_step(self.lr, self.lr_decay, grad_scale=scale)
        self.steps += 1
        return loss

    # -----------------------------------------------------------------------------------------------------------------
    def train_step_pipelined(self, labels, means, stds, real_image=None, draws=None):
        """Same work per call as train_step -- one generator pass, one U-Net training pass -- but software-pipelined like
        the reference's `fit_generator` queue (SynthSR/training.py:449-453: the Keras generator runs ahead of the
        optimiser): the batch passed in is generated on a second stream while the network trains on the batch of the
        PREVIOUS call, so the latency-bound generator kernels fill the SMs the backward chain leaves idle.  Batches are
        trained exactly once, in order.  Returns the loss of the previous call's batch (None on the first call);
        `flush()` trains the last pending batch.  labels may be a pinned host tensor (copied on the generator stream)."""
        if self._gens is None:
            self._gens = [self.gen, SynthGenerator(self.plan, self.B, self.device)]
            self._gen_stream = torch.cuda.Stream(device=self.device)
            self._gen_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._lab_dev = [None, None]
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
            self._gens[k].philox_step = max(g.philox_step for g in self._gens)    # one noise-counter sequence for both
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
        if self.seg is not None:
            self.net.seg_labels = self._gens[k].labels
        loss = self.net.loss_and_grad(image, target, self.metric, self.residual, self.loss_cropping)
        scale = 1.
        if self.world > 1:
            loss = self._allreduce(loss)
            scale = 1. / self.world
        self.net.adam_step(self.lr, self.lr_decay, grad_scale=scale)
        self.steps += 1
        return loss

    def flush(self):
        """train on the batch generated by the last train_step_pipelined call (end of an epoch / of training)."""
        return self._train_pending()

    def _allreduce(self, loss):
        return allreduce_step(self.net.grads, list(self.net.moving.values()), loss, self.flat, self.world)


def allreduce_step(grads, moving, loss, flat, world):
    """THE one collective of the data-parallel step: a single SUM all-reduce of [gradients | BN moving stats | loss].
    Gradients are left SUMMED in `grads` (Adam applies the 1/world scale), moving statistics and the loss are
    averaged in place.  Device agnostic (NCCL on GPUs, gloo in the CPU tests).  Returns the mean loss (float64)."""
    import torch.distributed as dist
    n = grads.numel()
    flat[:n].copy_(grads)
    o = n
    for t in moving:
        flat[o:o + t.numel()].copy_(t.reshape(-1))
        o += t.numel()
    flat[o] = loss.reshape(-1)[0].to(flat.dtype)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    grads.copy_(flat[:n])
    o = n
    for t in moving:
        t.copy_((flat[o:o + t.numel()] / world).reshape(t.shape))
        o += t.numel()
    return (flat[o:o + 1] / world).double()
