"""Adversarial fine-tuner on a GPU (SynthSR/fine_tuning_with_adversary.py on the engine): the U-Net step on
build_generator_loss against the float64 oracle (oracle/unet.py + oracle/adversary.py), the frozen forward of the
discriminator steps, the discriminator's loss / gradients / Adam on the device against the oracle, and the whole
`training()` call end to end on small label maps."""
import os

import numpy as np
import pytest
import torch

from helpers import gpu_pool_routing

pytestmark = pytest.mark.gpu

PRED_TOL, LOSS_TOL, GRAD_TOL = 1e-3, 1e-3, 1e-2


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _make(disc_kw=None, net_kw=None, dims=(32, 32, 32), seed=0, conv_impl='tc3', feats=8, levels=3):
    from synthsr_b200.adversary import AdversarialUNet3D, Discriminator
    disc = Discriminator([*dims, 1], n_filters=4, n_levels=2, seed=seed + 1, **(disc_kw or {}))
    rng = np.random.default_rng(seed)
    for k in disc.p:                                      # non-trivial biases
        if k.endswith('bias'):
            disc.p[k].copy_(_t(rng.normal(size=tuple(disc.p[k].shape)) * .1))
    net = AdversarialUNet3D([*dims, 1], feats, levels, 3, 1, 2, 2, 1, 'cuda', conv_impl, seed=seed, seg=None, disc=disc,
                            **(net_kw or {}))
    return net, disc


def _oracle_generator_step(net, disc, image, target, discr_weight, loss_cropping=None, mask=None, routing=None):
    from oracle import adversary as OA
    from oracle import unet as OU
    sd = net.state_dict()
    params = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in sd.items()}
    names = OU.trainable_names(params)
    leaves = {k: params[k].clone().requires_grad_(True) for k in names}
    p = {k: leaves.get(k, params[k]) for k in params}
    img, tgt = torch.tensor(image, dtype=torch.float64), torch.tensor(target, dtype=torch.float64)
    pred = OU.forward(p, img, training=True, nb_levels=net.L, pool_routing=routing)
    dparams = {k: torch.tensor(v, dtype=torch.float64) for k, v in disc.state_dict().items()}
    d_out = OA.discriminator_forward(dparams, pred, None if mask is None else torch.tensor(mask, dtype=torch.float64),
                                     n_levels=2)
    loss = OA.generator_loss(tgt, pred, d_out, discr_weight, loss_cropping)
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[k] for k in names])))
    return pred.detach().numpy(), float(loss), {k: v.numpy() for k, v in grads.items()}


@pytest.mark.parametrize('conv_impl,discr_weight,loss_cropping,use_mask',
                         [('ref', .05, None, False), ('ref', .3, 16, True), ('tc3', .01, None, False), ('tc3', .05, 16, True)])
def test_generator_step_matches_float64_oracle(conv_impl, discr_weight, loss_cropping, use_mask):
    """loss_and_grad of the fine-tuned U-Net = l1_weight * L1 + discr_weight * mean(-D(pred)): loss and every gradient tensor
    against the oracle (autograd through the oracle U-Net AND the oracle discriminator).
    'ref' (exact-fp32 convolutions, a small 3-level net): every tensor to 1e-4 -- the algebra of the step (loss weights, the
    discriminator's input gradient, the extra-gradient path through the head) with nothing to hide behind; the large
    discr_weight makes the adversarial term dominate.  'tc3' (the default mode, reference topology of 24 features / 5
    levels): north_star's bars -- 1e-3 on the prediction and the loss, 1e-2 on every gradient tensor (measured 1.8e-3 ..
    2.7e-3; on the small 3-level net the plain-TF32 backward passes the noise-like discriminator gradient on with 1.7e-2,
    which is why the strict case runs in 'ref' mode: scripts/adv_grad_diag.py)."""
    strict = conv_impl == 'ref'
    topo = dict(feats=8, levels=3) if strict else dict(feats=24, levels=5)
    rng = np.random.default_rng(3)
    lut = None
    labels = mask = None
    if use_mask:
        lut = torch.tensor([0., 1., 1., 0., 1.], device='cuda')
        lab_np = rng.integers(0, 5, size=(1, 32, 32, 32)).astype(np.int32)
        labels = torch.from_numpy(lab_np).cuda()
        mask = lut.cpu().numpy()[lab_np][..., None]
    net, disc = _make(disc_kw=dict(mask_input=use_mask), net_kw=dict(discr_weight=discr_weight, mask_lut=lut),
                      conv_impl=conv_impl, **topo)
    image = rng.uniform(0, 1, size=(1, 32, 32, 32, 1)).astype(np.float32)
    target = rng.uniform(0, 1, size=(1, 32, 32, 32, 1)).astype(np.float32)
    net.seg_labels = labels
    loss = net.loss_and_grad(_t(image), _t(target), 'l1', None, loss_cropping)
    torch.cuda.synchronize()
    # the oracle's MaxPooling3D takes the GPU forward's window winners (tests/test_unet_parity_gpu.py explains why: with 8 / 16
    # channels a level has 32k / 8k windows, and ONE near-tied window resolved differently moves a gradient tensor by 1e-2)
    pred_o, loss_o, grads_o = _oracle_generator_step(net, disc, image, target, discr_weight, loss_cropping, mask,
                                                     gpu_pool_routing(net))
    pred = net.pred.view(1, 32, 32, 32, 1).cpu().numpy().astype(np.float64)
    assert np.linalg.norm(pred - pred_o) <= PRED_TOL * np.linalg.norm(pred_o)
    assert abs(loss.item() - loss_o) <= LOSS_TOL * abs(loss_o), (loss.item(), loss_o)
    gtot = np.sqrt(sum(float((g ** 2).sum()) for g in grads_o.values()))
    worst, whole = 0., 0.
    for k, g in grads_o.items():
        diff = net.g[k].cpu().numpy().astype(np.float64) - g
        e = np.linalg.norm(diff) / max(np.linalg.norm(g), 1e-2 * gtot)
        worst, whole = max(worst, e), whole + float((diff ** 2).sum())
        assert e <= (1e-4 if strict else GRAD_TOL), (k, e)
    whole = np.sqrt(whole) / gtot
    assert whole <= (1e-4 if strict else GRAD_TOL), whole
    # the image term and the adversarial term the head reported
    image_term, w = net.last_terms
    l1 = np.abs(pred_o - target).mean() if loss_cropping is None else \
        np.abs(pred_o - target)[:, 8:24, 8:24, 8:24].mean()
    assert abs(image_term.item() - (1 - discr_weight) * l1) <= 1e-3 * l1
    line = 'adversarial generator step %s w_d=%g crop=%s mask=%d: loss %.6f (oracle %.6f), whole gradient %.2e, worst tensor %.2e' % (
        conv_impl, discr_weight, loss_cropping, use_mask, loss.item(), loss_o, whole, worst)
    print(line)
    try:
        os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out'), exist_ok=True)
        open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'adversary_parity.txt'), 'a').write(line + '\n')
    except OSError:
        pass


def test_generator_step_with_discriminator_and_segmentation_regulariser():
    """all three terms of build_generator_loss (:568-573): (1 - w_d - w_s) * L1 + w_d * mean(-D) + w_s * Dice, the Dice ground
    truth being `labels == label VALUE` (:551) -- exact-fp32 mode, every gradient tensor against float64 autograd through the
    oracle U-Net, the oracle discriminator and the oracle segmentation network"""
    from oracle import adversary as OA
    from oracle import unet as OU
    from synthsr_b200.adversary import AdversarialUNet3D, Discriminator
    from synthsr_b200.seg_loss import SegRegulariser, class_tables
    from synthsr_b200.unet import UNet3D
    rng = np.random.default_rng(11)
    dims, L, F, S, w_d, w_s, crop, m, M = [16, 16, 16], 3, 8, 5, .1, .3, 12, .05, .9
    gen_labels = np.array([0, 1, 2, 3, 4, 14, 15])
    equiv = np.array([0, 14, 14, 3, -1])
    assert class_tables(gen_labels, equiv, gt_by_value=True)[1].tolist() == [0, 3, 14]      # label values, not loop indices
    tmp = UNet3D(dims + [1], nb_features=F, nb_levels=L, nb_labels=S, batchsize=1, conv_impl='ref', seed=5)
    seg_sd = tmp.state_dict()
    for k in seg_sd:
        if k.endswith('gamma'):
            seg_sd[k] = rng.uniform(.7, 1.3, size=seg_sd[k].shape).astype(np.float32)
        elif k.endswith(('beta', 'bias')):
            seg_sd[k] = (rng.normal(size=seg_sd[k].shape) * .1).astype(np.float32)
    del tmp
    seg = SegRegulariser(dims, 1, seg_sd, S, gen_labels, equiv, rel_weight=w_s, loss_cropping=crop, m=m, M=M, nb_features=F,
                         nb_levels=L, conv_impl='ref', gt_by_value=True)
    disc = Discriminator([*dims, 1], n_filters=4, n_levels=2, seed=2)
    net = AdversarialUNet3D(dims + [1], F, L, 3, 1, 2, 2, 1, 'cuda', 'ref', seed=1, seg=seg, disc=disc, discr_weight=w_d)
    assert abs(net.l1_weight - (1 - w_d - w_s)) < 1e-12
    image = rng.uniform(0, 1, size=(1, *dims, 1)).astype(np.float32)
    target = rng.uniform(0, 1, size=(1, *dims, 1)).astype(np.float32)
    labels = gen_labels[rng.integers(0, len(gen_labels), size=(1, *dims))].astype(np.int32)
    net.seg_labels = torch.from_numpy(labels).cuda()
    loss = net.loss_and_grad(_t(image), _t(target), 'l1', None, crop)
    torch.cuda.synchronize()
    t64 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731
    params = {k: t64(v) for k, v in net.state_dict().items()}
    names = OU.trainable_names(params)
    for k in names:
        params[k].requires_grad_(True)
    pred = OU.forward(params, t64(image), training=True, nb_levels=L, pool_routing=gpu_pool_routing(net))
    d_out = OA.discriminator_forward({k: t64(v) for k, v in disc.state_dict().items()}, pred, None, n_levels=2)
    x = (torch.clamp(pred, m, M) - m) / (M - m)                                     # input_normalized, :386
    seg_out = torch.softmax(OU.forward({k: t64(v) for k, v in seg_sd.items()}, x, training=True, nb_levels=L), -1)
    total = OA.generator_loss(t64(target), pred, d_out, w_d, crop, target_seg=torch.from_numpy(labels)[..., None],
                              seg_out=seg_out, generation_labels=gen_labels, segmentation_equivalency=equiv, dice_weight=w_s)
    grads = torch.autograd.grad(total, [params[k] for k in names])
    assert abs(loss.item() - float(total.detach())) <= 2e-4 * abs(float(total.detach())), (loss.item(), float(total.detach()))
    gtot = np.sqrt(sum(float((g ** 2).sum()) for g in grads))
    for k, g in zip(names, grads):
        err = np.linalg.norm(net.g[k].cpu().numpy().astype(np.float64) - g.numpy()) / max(np.linalg.norm(g.numpy()), 1e-2 * gtot)
        assert err < 5e-4, (k, err)


def test_frozen_forward_uses_batch_statistics_and_keeps_the_moving_ones():
    """the U-Net under the discriminator steps (generator.trainable = False inside a fitted Keras model): batch-statistics
    BatchNorm, no moving-average update, no parameter change"""
    from oracle import unet as OU
    net, _ = _make()
    rng = np.random.default_rng(4)
    image = rng.uniform(0, 1, size=(1, 32, 32, 32, 1)).astype(np.float32)
    before = net.state_dict()
    pred = net.forward_frozen(_t(image)).cpu().numpy().astype(np.float64)
    after = net.state_dict()
    assert all(np.array_equal(before[k], after[k]) for k in before)
    params = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in before.items()}
    pred_o = OU.forward(params, torch.tensor(image, dtype=torch.float64), training=True, nb_levels=3).numpy()
    assert np.linalg.norm(pred - pred_o) <= PRED_TOL * np.linalg.norm(pred_o)
    # ... whereas a training forward does move them
    net.forward(_t(image), training=True)
    torch.cuda.synchronize()
    moved = net.state_dict()
    assert any(not np.array_equal(before[k], moved[k]) for k in before if k.endswith('moving_mean'))


def test_discriminator_step_on_the_device_matches_oracle():
    """float32 / cuDNN discriminator: loss, parameter gradients (second-order term included) and one fused Keras-Adam update
    against the float64 oracle"""
    from oracle import adversary as OA
    from synthsr_b200 import adversary as PA
    _, disc = _make()
    rng = np.random.default_rng(5)
    real = rng.uniform(0, 1, size=(1, 32, 32, 32, 1)).astype(np.float32)
    fake = rng.uniform(0, 1, size=(1, 32, 32, 32, 1)).astype(np.float32)
    w = np.array([[[[[.37]]]]], dtype=np.float32)
    leaves = disc.leaves()
    with PA.fp32_convs():
        loss, parts = PA.discriminator_loss(disc, _t(real), _t(fake), _t(w), 10., None, leaves)
        grads = torch.autograd.grad(loss, list(leaves.values()))
    t64 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731
    oleaves = {k: t64(v).requires_grad_(True) for k, v in disc.state_dict().items()}
    loss_o = OA.discriminator_loss(oleaves, t64(real), t64(fake), t64(w), 10., None, n_levels=2)
    grads_o = torch.autograd.grad(loss_o, [oleaves[k] for k in leaves])
    assert abs(loss.item() - float(loss_o)) <= 1e-4 * abs(float(loss_o))
    gtot = float(torch.sqrt(sum((g ** 2).sum() for g in grads_o)))
    for k, a, b in zip(leaves, grads, grads_o):
        e = float(torch.linalg.norm(a.double().cpu() - b)) / max(float(torch.linalg.norm(b)), 1e-2 * gtot)
        assert e <= 1e-3, (k, e)
    # Adam: the fused kernel on the flat buffers against the same update in float64
    p0 = disc.params.double().cpu().clone()
    disc.grads.zero_()
    for k, g in zip(leaves, grads):
        disc.g[k].copy_(g)
    g_flat = disc.grads.double().cpu().clone()
    disc.adam_step(1e-3)
    lr_t = 1e-3 * np.sqrt(1 - .999) / (1 - .9)
    expect = p0 - lr_t * (.1 * g_flat) / ((.001 * g_flat * g_flat).sqrt() + 1e-7)
    assert torch.allclose(disc.params.double().cpu(), expect, rtol=1e-5, atol=1e-7)


@pytest.fixture(scope='module')
def dataset(tmp_path_factory):
    from ext.lab2im import utils
    from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors
    root = tmp_path_factory.mktemp('adv')
    labels_dir, images_dir = root / 'labels', root / 'images'
    os.makedirs(labels_dir)
    os.makedirs(images_dir)
    aff = np.array([[1., 0, 0, 40], [0, 1., 0, 16], [0, 0, 1., 63], [0, 0, 0, 1]])
    rng = np.random.default_rng(0)
    for i in range(2):
        lab = phantom_labels([44, 52, 40], seed=i)
        utils.save_volume(lab.astype(np.float32), aff, None, str(labels_dir / ('brain%d_labels.nii.gz' % i)))
        img = (lab > 0) * (50. + 10. * (lab % 7)) + rng.normal(size=lab.shape) * 2.          # a "real scan" per label map
        utils.save_volume(img.astype(np.float32), aff, None, str(images_dir / ('brain%d.nii.gz' % i)))
    pm, ps = synthetic_priors(14, 1, seed=0)
    paths = {k: str(root / (k + '.npy')) for k in ('labels', 'classes', 'means', 'stds', 'mask')}
    np.save(paths['labels'], GEN_LABELS); np.save(paths['classes'], GEN_CLASSES)
    np.save(paths['means'], pm); np.save(paths['stds'], ps)
    np.save(paths['mask'], (np.asarray(GEN_LABELS) > 0).astype(np.int32))
    return str(labels_dir), str(images_dir), paths, root


@pytest.mark.parametrize('real_targets', [False, True])
def test_fine_tuning_runs_end_to_end(dataset, real_targets, monkeypatch):
    """the reference-facing call: 1 epoch x 2 steps, first_training_ratio 2, training_ratio 1 -> both models and the loss logs
    on disk, finite losses; the discriminator moved during its steps, the U-Net only during its own."""
    import SynthSR.fine_tuning_with_adversary as FT
    from synthsr_b200 import adversary as PA
    from synthsr_b200 import h5lite
    labels_dir, images_dir, p, root = dataset
    model_dir = str(root / ('model_real' if real_targets else 'model_syn'))
    seen = {'d': [], 'g': []}
    d_step, g_step = PA.AdversarialEngine.discriminator_step, PA.AdversarialEngine.generator_step

    def spy_d(self, *a, **k):
        net0, d0 = self.engine.net.params.clone(), self.disc.params.clone()
        out = d_step(self, *a, **k)
        seen['d'].append((torch.equal(net0, self.engine.net.params), torch.equal(d0, self.disc.params)))
        return out

    def spy_g(self, *a, **k):
        net0, d0 = self.engine.net.params.clone(), self.disc.params.clone()
        out = g_step(self, *a, **k)
        seen['g'].append((torch.equal(net0, self.engine.net.params), torch.equal(d0, self.disc.params)))
        return out
    ckpt = os.path.join(str(root / 'model_syn'), 'generator_1.h5')
    ckpt = ckpt if real_targets and os.path.isfile(ckpt) else None            # (absent when this case is run on its own)
    monkeypatch.setattr(PA.AdversarialEngine, 'discriminator_step', spy_d)
    monkeypatch.setattr(PA.AdversarialEngine, 'generator_step', spy_g)
    FT.training(labels_dir, images_dir if real_targets else None, model_dir, p['means'], p['stds'], p['labels'],
                path_generation_classes=p['classes'], output_channel=None if real_targets else 0, output_shape=32,
                n_levels=3, unet_feat_count=8, epochs=1, steps_per_epoch=2, first_training_ratio=2, training_ratio=1,
                loss_cropping=16, relative_weight_discriminator=.05, labels_to_mask=p['mask'] if real_targets else None,
                randomise_res=False, data_res=np.array([1., 1., 2.]),
                # the second case starts from the U-Net the first one wrote (weights by name, fresh optimizer)
                checkpoint_generator=ckpt)
    assert seen['d'] == [(True, False)] * 3 and seen['g'] == [(False, True)] * 2, seen
    for name in ('generator_1.h5', 'discriminator_1.h5'):
        assert os.path.isfile(os.path.join(model_dir, name)), name
    d_log = np.load(os.path.join(model_dir, 'logs', 'discriminator_loss.npy'))
    g_log = np.load(os.path.join(model_dir, 'logs', 'generator_loss.npy'))
    assert d_log.shape == g_log.shape == (1,) and np.isfinite(d_log).all() and np.isfinite(g_log).all()
    sd, _ = h5lite.load_keras_weights(os.path.join(model_dir, 'discriminator_1.h5'))
    assert sd['conv3d_1/kernel'].shape == (3, 3, 3, 1, 32) and sd['dense_2/kernel'].shape == (512, 1)
    gsd, _ = h5lite.load_keras_weights(os.path.join(model_dir, 'generator_1.h5'))
    assert 'unet_likelihood/kernel' in gsd
