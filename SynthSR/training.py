"""`SynthSR.training.training()` with the reference's signature (SynthSR/training.py:38-89) on the B200 engine.

Every step: host sampler (label map + GMM parameters) -> CUDA generator -> U-Net forward/backward (tcgen05 TF32
convolutions) -> [single NCCL all-reduce of the flat gradient buffer when launched with torchrun] -> fused Adam.
One checkpoint per epoch named %03d.h5 (Keras ModelCheckpoint layout, written by synthsr_b200/h5lite.py)."""
import os
import time

import numpy as np

from ext.lab2im import utils
from ext.neuron import models as nrn_models

from .brain_generator import BrainGenerator
from .metrics_model import IdentityLoss, add_seg_loss_to_model, metrics_model  # noqa: F401


def training(labels_dir,
             model_dir,
             prior_means,
             prior_stds,
             path_generation_labels,
             segmentation_label_list=None,
             segmentation_label_equivalency=None,
             segmentation_model_file=None,
             fs_header_segnet=False,
             relative_weight_segmentation=0.25,
             prior_distributions='normal',
             images_dir=None,
             path_generation_classes=None,
             FS_sort=True,
             batchsize=1,
             input_channels=True,
             output_channel=0,
             target_res=None,
             output_shape=None,
             flipping=True,
             padding_margin=None,
             scaling_bounds=0.15,
             rotation_bounds=15,
             shearing_bounds=0.02,
             translation_bounds=5,
             nonlin_std=4.,
             nonlin_shape_factor=0.03125,
             simulate_registration_error=True,
             data_res=None,
             thickness=None,
             randomise_res=None,
             downsample=True,
             blur_range=1.15,
             build_reliability_maps=True,
             bias_field_std=.3,
             bias_shape_factor=0.03125,
             n_levels=5,
             nb_conv_per_level=2,
             conv_size=3,
             unet_feat_count=24,
             feat_multiplier=2,
             dropout=0,
             activation='elu',
             lr=1e-4,
             lr_decay=0,
             epochs=100,
             steps_per_epoch=1000,
             regression_metric='l1',
             work_with_residual_channel=None,
             loss_cropping=None,
             checkpoint=None,
             model_file_has_different_lhood_layer=False):
    n_channels = len(utils.reformat_to_list(input_channels))
    if output_channel is not None:
        output_channel = list(utils.reformat_to_list(output_channel))
        n_output_channels = len(output_channel)
    else:
        n_output_channels = 1
    # same checks and messages as the reference (training.py:252-268)
    if (images_dir is None) & (output_channel is None):
        raise Exception('please provide a value for output_channel or image_dir')
    elif (images_dir is not None) & (output_channel is not None):
        raise Exception('please provide a value either for output_channel or image_dir, but not both at the same time')
    if output_channel is not None:
        if any(x >= n_channels for x in output_channel):
            raise Exception('indices in output_channel cannot be greater than the total number of channels')
    if work_with_residual_channel is not None:
        work_with_residual_channel = utils.reformat_to_list(work_with_residual_channel)
        if output_channel is not None:
            if len(work_with_residual_channel) != len(output_channel):
                raise Exception('The number or residual channels and output channels must be the same')
        if any(x >= n_channels for x in work_with_residual_channel):
            raise Exception('indices in work_with_residual_channel cannot be greater than the total number of channels')
        if build_reliability_maps:
            # Reference behaviour kept on purpose (training.py:270-271, pinned by executing training() itself:
            # tests/golden/make_reference_training_goldens.py): `2 * work_with_residual_channel` REPEATS the list instead
            # of doubling the indices, so metrics_model adds image_out channel c -- of [ch0, rel0, ch1, rel1, ...] -- twice
            # and Keras' Add broadcasts the single predicted channel over the two copies.  Mean and gradient of the L1 / L2
            # loss over two identical copies equal those over one, i.e. the step is that of residual channel c, UN-doubled
            # (for c >= 1 that is a reliability map: the reference's quirk, reproduced).  With several output channels the
            # repeated list no longer broadcasts against the prediction and Keras refuses to build the Add layer.
            if len(work_with_residual_channel) > 1:
                raise ValueError('Operands could not be broadcast together with shapes (..., %d) (..., %d): the reference '
                                 'repeats work_with_residual_channel when build_reliability_maps is set, which only works '
                                 'for a single output channel' % (2 * len(work_with_residual_channel),
                                                                  len(work_with_residual_channel)))
            work_with_residual_channel = [int(c) for c in work_with_residual_channel]
    # options the engine does not implement fail here, before any GPU work, instead of silently training something else
    if activation != 'elu':
        raise NotImplementedError("activation %r: the engine implements the reference's default 'elu' only "
                                  "(ELU and its derivative are fused into the convolution epilogues)" % (activation,))
    if regression_metric not in ('l1', 'l2'):
        metrics_model(None, metrics=regression_metric)      # raises like the reference / NotImplementedError for ssim, laplace

    generation_labels, n_neutral_labels = utils.get_list_labels(label_list=path_generation_labels,
                                                                labels_dir=labels_dir, FS_sort=FS_sort)
    utils.mkdir(model_dir)
    if loss_cropping == 0:
        padding_margin = None
    elif padding_margin is None:
        padding_margin = utils.get_padding_margin(output_shape, loss_cropping)

    brain_generator = BrainGenerator(labels_dir=labels_dir, images_dir=images_dir, generation_labels=generation_labels,
                                     n_neutral_labels=n_neutral_labels, padding_margin=padding_margin,
                                     batchsize=batchsize, input_channels=input_channels, output_channel=output_channel,
                                     target_res=target_res, output_shape=output_shape, output_div_by_n=2 ** n_levels,
                                     generation_classes=path_generation_classes, prior_means=prior_means,
                                     prior_stds=prior_stds, prior_distributions=prior_distributions, flipping=flipping,
                                     scaling_bounds=scaling_bounds, rotation_bounds=rotation_bounds,
                                     shearing_bounds=shearing_bounds, translation_bounds=translation_bounds,
                                     nonlin_std=nonlin_std, nonlin_shape_factor=nonlin_shape_factor,
                                     simulate_registration_error=simulate_registration_error,
                                     randomise_res=randomise_res, data_res=data_res, thickness=thickness,
                                     downsample=downsample, blur_range=blur_range,
                                     build_reliability_maps=build_reliability_maps, bias_field_std=bias_field_std,
                                     bias_shape_factor=bias_shape_factor)
    if dropout:
        raise NotImplementedError('dropout is not part of this build (the reference recommends dropout=0)')
    nb_labels_unet = n_output_channels
    plan = brain_generator.labels_to_image_model.plan

    from synthsr_b200.trainer import TrainingEngine
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        if not dist.is_initialized():
            dist.init_process_group('nccl')
    seg = None
    if segmentation_model_file is not None:          # training.py:371-411: frozen segmentation U-Net + Dice regulariser
        from synthsr_b200 import h5lite
        from synthsr_b200.seg_loss import SegRegulariser
        segmentation_labels = np.load(segmentation_label_list)
        seg_sd, _ = h5lite.load_keras_weights(segmentation_model_file)
        if images_dir is None:
            m = M = None
        else:                                        # clip the synthesised image at the 2nd / 98th percentile of the first scan
            im = utils.load_volume(utils.list_images_in_folder(images_dir)[0], im_only=True).flatten()
            m, M = np.percentile(im, 2), np.percentile(im, 98)
        seg = SegRegulariser(plan.output_shape, batchsize, seg_sd, len(segmentation_labels), generation_labels,
                             utils.load_array_if_path(segmentation_label_equivalency), relative_weight_segmentation,
                             loss_cropping=loss_cropping, m=m, M=M, fs_header=fs_header_segnet, nb_features=unet_feat_count,
                             nb_levels=n_levels, conv_size=conv_size, feat_mult=feat_multiplier,
                             nb_conv_per_level=nb_conv_per_level)
    # augmentation / initialisation seed: fresh OS entropy per run like the reference's unseeded tf.random / glorot draws
    # (SSR_SEED pins it); rank 0's value is shared so that every replica initialises the same weights
    seed = int(os.environ['SSR_SEED']) if os.environ.get('SSR_SEED') else int.from_bytes(os.urandom(4), 'little') >> 2
    if world > 1:
        box = [seed]
        dist.broadcast_object_list(box, src=0)
        seed = int(box[0])
    engine = TrainingEngine(plan, batchsize=batchsize, nb_features=unet_feat_count, nb_levels=n_levels, seed=seed,
                            conv_size=conv_size, feat_mult=feat_multiplier, nb_conv_per_level=nb_conv_per_level,
                            nb_labels=nb_labels_unet, lr=lr, lr_decay=lr_decay, metric=regression_metric,
                            work_with_residual_channel=work_with_residual_channel, loss_cropping=loss_cropping,
                            rank=rank, world_size=world, seg=seg)
    # the Keras-style handles the reference builds (kept so client code can introspect the same objects)
    model = metrics_model(nrn_models.UnetModel(engine.net, brain_generator.labels_to_image_model),
                          loss_cropping=loss_cropping, metrics=regression_metric,
                          work_with_residual_channel=work_with_residual_channel)
    input_generator = utils.build_training_generator(brain_generator.model_inputs_generator, batchsize)

    init_epoch = 0
    if checkpoint is not None:
        print('loading', checkpoint)
        init_epoch = load_checkpoint(engine, checkpoint, model_file_has_different_lhood_layer)
    train_model(engine, input_generator, lr, lr_decay, epochs, steps_per_epoch, model_dir, init_epoch)
    return model


FLAT_LAYOUT_VERSION = 2      # 2: flat parameter buffer in backward-completion order (synthsr_b200/unet.py)


def load_checkpoint(engine, path, different_lhood_layer=False):
    """weights by Keras layer name (+ this engine's Adam state when the file holds it).  Accepts the reference's own
    checkpoints (Keras .h5: ModelCheckpoint full-model files or `save_weights` files, training.py:353-369, 429) and the
    earlier .npz format.  A file whose name ends with a 3-digit epoch resumes at that epoch like the reference
    (training.py:435: int(path_checkpoint[-6:-3]))."""
    if str(path).endswith('.h5'):
        from synthsr_b200 import h5lite
        sd, _ = h5lite.load_keras_weights(path)
        opt = h5lite.load_extra(path)
    else:
        raw = dict(np.load(path))
        sd = {k: v for k, v in raw.items() if not k.startswith('optimizer/')}
        opt = {k.split('/', 1)[1]: v for k, v in raw.items() if k.startswith('optimizer/')}
    if different_lhood_layer:
        sd = {k: v for k, v in sd.items() if not k.startswith('unet_likelihood')}
    engine.net.load_state_dict(sd, strict=False)
    # the flat Adam moments are positional: only a file written with the same buffer layout may restore them
    same_layout = 'layout' in opt and int(np.asarray(opt['layout']).reshape(-1)[0]) == FLAT_LAYOUT_VERSION
    if 'm' in opt and same_layout and not different_lhood_layer and np.size(opt['m']) == engine.net.n_params:
        import torch
        engine.net.adam_m.copy_(torch.as_tensor(np.asarray(opt['m'], dtype=np.float32)))
        engine.net.adam_v.copy_(torch.as_tensor(np.asarray(opt['v'], dtype=np.float32)))
        engine.net.iterations = int(np.asarray(opt['iterations']).reshape(-1)[0])
        if 'aug_state' in opt:       # continue the augmentation counters instead of replaying the first steps' noise
            steps, philox = [int(v) for v in np.asarray(opt['aug_state']).reshape(-1)[:2]]
            engine.steps = steps
            engine.gen.philox_step = philox
    stem = os.path.splitext(os.path.basename(path))[0]
    return int(stem[-3:]) if stem[-3:].isdigit() else 0


def save_checkpoint(engine, path):
    """'%03d.h5' in the layout Keras' ModelCheckpoint writes (weights under /model_weights with the Keras layer names, so
    the reference's `load_weights(by_name=True)` / predict scripts read it); the flat Adam moments ride in
    /optimizer_weights for an exact resume on this engine."""
    from synthsr_b200 import h5lite
    from synthsr_b200.unet import keras_layer_order
    philox = max(g.philox_step for g in (engine._gens or [engine.gen]))
    extra = {'m': engine.net.adam_m.cpu().numpy(), 'v': engine.net.adam_v.cpu().numpy(),
             'iterations': np.array([engine.net.iterations], dtype=np.int64),
             'layout': np.array([FLAT_LAYOUT_VERSION], dtype=np.int64),
             'aug_state': np.array([engine.steps, philox], dtype=np.int64)}
    if str(path).endswith('.h5'):
        h5lite.save_keras_weights(path, engine.net.state_dict(), keras_layer_order(engine.net.L), extra=extra,
                                  full_model=True)
    else:
        sd = engine.net.state_dict()
        sd.update({'optimizer/' + k: v for k, v in extra.items()})
        np.savez(path, **sd)


class _PinnedInputs:
    """Host side of the per-step input feed: label maps (and real images) as PINNED int32 / float32 tensors the engine copies
    to the device asynchronously on its generator stream.  With batchsize 1 the sampler hands out views of its cached
    volumes, so each distinct volume is converted (int64 -> int32) and pinned ONCE and then re-used every time it is drawn;
    anything else (batchsize > 1, evicted volumes) goes through a small ring of pinned buffers guarded by copy events.
    (The reference re-decodes and feeds an int64 array through feed_dict every step, model_inputs.py:91, training.py:449.)"""

    def __init__(self, cuda, budget_bytes=None, ring=4):
        import torch
        self.torch, self.cuda = torch, cuda
        self.budget = int(float(os.environ.get('SSR_PINNED_INPUTS_GB', '4')) * (1 << 30)) if budget_bytes is None else budget_bytes
        self.used, self.cache = 0, {}
        self.ring, self.ring_i, self.nring = {}, 0, ring

    def _new(self, shape, dtype):
        return self.torch.empty(tuple(shape), dtype=dtype, pin_memory=self.cuda)

    def get(self, arr, dtype, gen_stream=None):
        """arr: numpy array [B, X, Y, Z] (any integer / float dtype) -> (pinned tensor, release) ; call release(stream)
        after the engine has enqueued its copy."""
        torch = self.torch
        root = arr
        while isinstance(root.base, np.ndarray):
            root = root.base
        nbytes = arr.size * 4
        if root.size == arr.size and self.used + nbytes <= self.budget or (id(root), dtype) in self.cache:
            hit = self.cache.get((id(root), dtype))
            if hit is not None and hit[0] is root:
                return hit[1], None
            t = self._new(arr.shape, dtype)
            np.copyto(t.numpy(), arr, casting='unsafe')
            if root.size == arr.size:
                self.cache[(id(root), dtype)] = (root, t)         # holding `root` keeps id() unique
                self.used += nbytes
            return t, None
        key = (tuple(arr.shape), dtype)
        slots = self.ring.setdefault(key, [[self._new(arr.shape, dtype), None] for _ in range(self.nring)])
        slot = slots[self.ring_i % self.nring]
        self.ring_i += 1
        if slot[1] is not None:
            slot[1].synchronize()                                 # the copy that last read this buffer has executed
        np.copyto(slot[0].numpy(), arr, casting='unsafe')

        def release(stream):
            if self.cuda:
                slot[1] = slot[1] or torch.cuda.Event()
                slot[1].record(stream)
        return slot[0], release


def train_model(engine, generator, learning_rate, lr_decay, n_epochs, n_steps, model_dir, init_epoch=0):
    """epochs x steps loop of the reference's fit_generator call (training.py:449-453) with one checkpoint per epoch
    ('%03d', :429), TensorBoard scalars (epoch_loss, :425-431) and a plain-text loss log under model_dir/logs.
    Per step: pinned host label map -> async H2D on the generator stream -> generator -> U-Net step; the loss of every
    step is read back to a pinned host buffer asynchronously (Keras reads it for its progress bar) and consumed at the end
    of the epoch, so the host never waits for the GPU inside an epoch."""
    import torch
    log_dir = os.path.join(model_dir, 'logs')
    utils.mkdir(log_dir)
    engine.lr, engine.lr_decay = learning_rate, lr_decay
    is_main = engine.rank == 0
    log = open(os.path.join(log_dir, 'loss.csv'), 'a') if is_main else None
    tb = None
    if is_main:
        from synthsr_b200.tbevents import EventWriter
        tb = EventWriter(log_dir)
    cuda = engine.device.type == 'cuda'
    feed = _PinnedInputs(cuda)
    host_losses = torch.zeros(n_steps, dtype=torch.float64, pin_memory=cuda)
    for epoch in range(init_epoch, n_epochs):
        t0, n_loss = time.time(), 0
        for step in range(n_steps):
            inputs, _ = next(generator)
            lab, rel_lab = feed.get(np.asarray(inputs[0])[..., 0], torch.int32)
            real, rel_real = (None, None)
            if len(inputs) > 3:
                real, rel_real = feed.get(np.asarray(inputs[3])[..., 0], torch.float32)
            # pipelined like fit_generator's queue: this batch is generated while the previous one is trained on; the
            # last batch of the epoch is flushed before the loss is reported and the checkpoint written
            l = engine.train_step_pipelined(lab, inputs[1], inputs[2], real_image=real)
            for rel in (rel_lab, rel_real):
                if rel is not None:
                    rel(engine._gen_stream)
            if l is not None:
                host_losses[n_loss:n_loss + 1].copy_(l, non_blocking=True)
                n_loss += 1
        host_losses[n_loss:n_loss + 1].copy_(engine.flush(), non_blocking=True)
        n_loss += 1
        if cuda:
            torch.cuda.current_stream().synchronize()
        loss = float(host_losses[:n_loss].mean().item())
        if not np.isfinite(loss):
            raise FloatingPointError('Loss not finite')              # tf.debugging.check_numerics (metrics_model.py:228)
        if is_main:
            dt = time.time() - t0
            print('Epoch %d/%d - %ds - loss: %.4f - %.2f volumes/s' % (epoch + 1, n_epochs, dt, loss,
                                                                       n_steps * engine.B * engine.world / dt))
            log.write('%d,%.6f,%.3f\n' % (epoch + 1, loss, dt))
            log.flush()
            tb.scalar('epoch_loss', loss, epoch)                     # what Keras' TensorBoard callback logs per epoch
            tb.scalar('volumes_per_second', n_steps * engine.B * engine.world / dt, epoch)
            save_checkpoint(engine, os.path.join(model_dir, '%03d.h5' % (epoch + 1)))
    if log:
        log.close()
    if tb:
        tb.close()
