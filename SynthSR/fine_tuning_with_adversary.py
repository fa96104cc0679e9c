"""`SynthSR.fine_tuning_with_adversary.training()` with the reference's signature (SynthSR/fine_tuning_with_adversary.py:37-
92) on the B200 engine: WGAN-GP fine-tuning of the super-resolution U-Net against a discriminator that sees the real (or
synthetic target) scans.

Per step, like the reference's loop (:440-459): `training_ratio` discriminator updates (each on a freshly generated batch, the
U-Net frozen), then one U-Net update on  (1 - w_d [- w_s]) * L1 + w_d * mean(-D(prediction)) [+ w_s * Dice].
Generator, U-Net and the optional segmentation regulariser are this package's CUDA kernels; the discriminator runs under torch
autograd because its gradient penalty needs second derivatives (synthsr_b200/adversary.py says why and what is restated).

Files written per epoch: generator_%0Nd.h5 / discriminator_%0Nd.h5 (Keras weight layout, readable with
`load_weights(by_name=True)`; the reference's `model.save` additionally stores its Keras optimizer state, here the engines'
flat Adam moments ride in /optimizer_weights) and logs/{generator,discriminator}_loss.npy (:462-467).
"""
import os

import numpy as np

from ext.lab2im import utils

from .brain_generator import BrainGenerator


def training(labels_dir,
             images_dir,
             model_dir,
             prior_means,
             prior_stds,
             path_generation_labels,
             path_segmentation_equivalency=None,
             segmentation_model_file=None,
             prior_distributions='normal',
             path_generation_classes=None,
             FS_sort=True,
             batchsize=1,
             input_channels=True,
             output_channel=None,
             target_res=None,
             output_shape=None,
             flipping=True,
             padding_margin=None,
             scaling_bounds=0.2,
             rotation_bounds=20,
             shearing_bounds=0.03,
             translation_bounds=5,
             nonlin_std=5.,
             nonlin_shape_factor=0.04,
             simulate_registration_error=False,
             data_res=None,
             thickness=None,
             randomise_res=True,
             downsample=True,
             blur_range=1.03,
             build_reliability_maps=False,
             bias_field_std=.4,
             bias_shape_factor=0.04,
             n_levels=5,
             nb_conv_per_level=2,
             conv_size=3,
             unet_feat_count=24,
             feat_multiplier=2,
             dropout=0,
             activation='elu',
             lr_decay=0,
             epochs=100,
             steps_per_epoch=1000,
             work_with_residual_channel=None,
             loss_cropping=None,
             lr_generator=1e-4,
             lr_discriminator=1e-4,
             relative_weight_segmentation=0.25,
             relative_weight_discriminator=0.01,
             checkpoint_generator=None,
             gradient_penalty_weight=10,
             first_training_ratio=100,
             training_ratio=10,
             labels_to_mask=None):
    """See the reference's docstring (fine_tuning_with_adversary.py:93-236) for the parameters; they mean the same here.
    `work_with_residual_channel` is validated and then unused, as in the reference (its generator loss is built on the raw
    U-Net output, :410-413)."""
    n_channels = len(utils.reformat_to_list(input_channels))
    if output_channel is not None:
        output_channel = list(utils.reformat_to_list(output_channel))
        n_output_channels = len(output_channel)
    else:
        n_output_channels = 1

    # the reference's checks and messages (:248-266)
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
    # what this build does not cover fails here, before any GPU work
    if activation != 'elu':
        raise NotImplementedError("activation %r: the engine implements the reference's default 'elu' only" % (activation,))
    if dropout:
        raise NotImplementedError('dropout is not part of this build (the reference recommends dropout=0)')
    if n_output_channels != 1:
        raise NotImplementedError('the adversarial fine-tuner judges one output channel (the head gradient of the '
                                  'discriminator term is implemented for a single-channel prediction)')

    generation_labels, n_neutral_labels = utils.get_list_labels(label_list=path_generation_labels, labels_dir=labels_dir,
                                                                FS_sort=FS_sort)
    utils.mkdir(model_dir)
    log_dir = os.path.join(model_dir, 'logs')
    utils.mkdir(log_dir)
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
    input_generator = utils.build_training_generator(brain_generator.model_inputs_generator, batchsize)
    plan = brain_generator.labels_to_image_model.plan

    import torch
    from synthsr_b200 import h5lite
    from synthsr_b200.adversary import AdversarialEngine, AdversarialUNet3D
    from synthsr_b200.trainer import TrainingEngine
    from .training import load_checkpoint
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        if not dist.is_initialized():
            dist.init_process_group('nccl')
    seed = int(os.environ['SSR_SEED']) if os.environ.get('SSR_SEED') else int.from_bytes(os.urandom(4), 'little') >> 2
    if world > 1:
        box = [seed]
        dist.broadcast_object_list(box, src=0)
        seed = int(box[0])

    # frozen segmentation network + Dice term (:335-357, :389-399)
    seg = None
    if segmentation_model_file is not None:
        from synthsr_b200.seg_loss import SegRegulariser
        segmentation_label_equivalency = np.load(path_segmentation_equivalency)
        seg_sd, _ = h5lite.load_keras_weights(segmentation_model_file)
        im = utils.load_volume(utils.list_images_in_folder(images_dir)[0], im_only=True)
        m, M = np.percentile(im, 2), np.percentile(im, 98)
        seg = SegRegulariser(plan.output_shape, batchsize, seg_sd, len(segmentation_label_equivalency), generation_labels,
                             segmentation_label_equivalency, relative_weight_segmentation, loss_cropping=loss_cropping,
                             m=m, M=M, nb_features=unet_feat_count, nb_levels=n_levels, conv_size=conv_size,
                             feat_mult=feat_multiplier, nb_conv_per_level=nb_conv_per_level,
                             gt_by_value=True)       # `x[..., -1] == ll`, the label VALUE (:551), unlike metrics_model.py:188

    # discriminator (:331-332) and its optional label mask (:368-372)
    mask_input = labels_to_mask is not None
    unet_input_shape = brain_generator.model_output_shape
    discriminator = make_discriminator([*unet_input_shape[:-1], n_output_channels], mask_input=mask_input,
                                       seed=seed + 1)
    mask_lut = None
    if mask_input:
        labels_to_mask = np.asarray(utils.load_array_if_path(labels_to_mask))
        lut = np.zeros(int(np.max(generation_labels)) + 1, dtype=np.float32)
        lut[np.asarray(generation_labels, dtype=np.int64)] = labels_to_mask          # layers.ConvertLabels
        mask_lut = torch.from_numpy(lut).cuda()

    engine = TrainingEngine(plan, batchsize=batchsize, nb_features=unet_feat_count, nb_levels=n_levels, seed=seed,
                            conv_size=conv_size, feat_mult=feat_multiplier, nb_conv_per_level=nb_conv_per_level,
                            nb_labels=n_output_channels, lr=lr_generator, lr_decay=lr_decay, metric='l1',
                            loss_cropping=loss_cropping, rank=rank, world_size=world, seg=seg, net_cls=AdversarialUNet3D,
                            net_kwargs=dict(disc=discriminator, discr_weight=relative_weight_discriminator,
                                            mask_lut=mask_lut))
    if checkpoint_generator is not None:
        print('loading', checkpoint_generator)
        load_checkpoint(engine, checkpoint_generator)                                # by name, like :327-329
        # `load_weights` brings weights only: the fine-tuning run starts with a fresh optimizer even when the file is one of
        # this engine's own checkpoints (which also carry the Adam moments of the run that wrote them)
        engine.net.adam_m.zero_()
        engine.net.adam_v.zero_()
        engine.net.iterations = 0
    adv = AdversarialEngine(engine, discriminator, lr_discriminator=lr_discriminator, lr_decay=lr_decay,
                            gradient_penalty_weight=gradient_penalty_weight, seed=seed)

    def next_batch():                       # host sampler -> device label map (+ real scan), GMM parameters stay on the host
        inputs, _ = next(input_generator)
        lab = torch.from_numpy(np.ascontiguousarray(np.asarray(inputs[0])[..., 0]).astype(np.int32)).cuda()
        real = None
        if len(inputs) > 3:
            real = torch.from_numpy(np.ascontiguousarray(np.asarray(inputs[3])[..., 0], dtype=np.float32)).cuda()
        return lab, inputs[1], inputs[2], real

    # ------------------------------------------------ training loop (:437-474) ------------------------------------------------
    # per step: `training_ratio` discriminator updates (`first_training_ratio` on the very first step), then one U-Net update;
    # per epoch: the averaged losses appended to logs/*.npy and both models written
    width = len(str(epochs))
    logs = {'discriminator': [], 'generator': []}
    for epoch in range(epochs):
        mean_d = mean_g = 0.
        for step in range(int(steps_per_epoch)):
            ratio = first_training_ratio if epoch == 0 and step == 0 else training_ratio
            for j in range(ratio):
                d_loss = float(adv.discriminator_step(*next_batch()).item())
                mean_d += d_loss / (steps_per_epoch * ratio)
                if rank == 0:
                    print('epoch %d step %d/%d  discriminator update %d/%d  loss %.6g' % (epoch + 1, step + 1, steps_per_epoch,
                                                                                       j + 1, ratio, d_loss))
            g_loss = float(adv.generator_step(*next_batch()).item())
            if not (np.isfinite(g_loss) and np.isfinite(mean_d)):
                raise FloatingPointError('Loss not finite')
            mean_g += g_loss / steps_per_epoch
            if rank == 0:
                print('epoch %d step %d/%d  generator loss %.6g' % (epoch + 1, step + 1, steps_per_epoch, g_loss))
        logs['discriminator'].append(mean_d)
        logs['generator'].append(mean_g)
        if rank == 0:
            print('Epoch %d/%d   average discriminator loss %.6g   average generator loss %.6g   saving models' % (
                epoch + 1, epochs, mean_d, mean_g))
            for k, v in logs.items():
                np.save(os.path.join(log_dir, '%s_loss.npy' % k), np.asarray(v))
            save_models(engine, discriminator, model_dir, '%0*d' % (width, epoch + 1))
    return engine, discriminator


def save_models(engine, discriminator, model_dir, tag):
    """generator_<tag>.h5 (the U-Net, as SynthSR.training writes its checkpoints) and discriminator_<tag>.h5"""
    from synthsr_b200 import h5lite
    from .training import save_checkpoint
    save_checkpoint(engine, os.path.join(model_dir, 'generator_%s.h5' % tag))
    extra = {'m': discriminator.adam_m.cpu().numpy(), 'v': discriminator.adam_v.cpu().numpy(),
             'iterations': np.array([discriminator.iterations], dtype=np.int64)}
    h5lite.save_keras_weights(os.path.join(model_dir, 'discriminator_%s.h5' % tag), discriminator.state_dict(),
                              [name for name, _, _ in discriminator.layers], extra=extra, full_model=True)


def make_discriminator(input_shape, n_filters=32, n_levels=4, mask_input=False, device='cuda', seed=0):
    """fine_tuning_with_adversary.py:482-508 -> synthsr_b200.adversary.Discriminator (callable: D(x[, mask]) -> [B, 1])."""
    from synthsr_b200.adversary import Discriminator
    return Discriminator(input_shape, n_filters=n_filters, n_levels=n_levels, mask_input=mask_input, device=device, seed=seed)


def dummy_loss(y_true, y_predicted):
    """The metric is computed inside the model (fine_tuning_with_adversary.py:599-602)."""
    return y_predicted
