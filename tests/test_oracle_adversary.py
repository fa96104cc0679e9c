"""Adversarial fine-tuner on the CPU (no GPU): the oracle against the reference's own loss builders executed on the tf shim
(tests/golden/reference_adversary.npz), and the product's Discriminator / losses (synthsr_b200/adversary.py, torch autograd --
device agnostic) against the oracle, in float64."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import adversary as OA

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'reference_adversary.npz'))
t64 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731


@pytest.mark.parametrize('case', ['plain', 'crop', 'seg', 'seg_crop'])
def test_generator_loss_matches_reference(case):
    c = json.loads(str(G['gen_cases']))[case]
    kw = {}
    if c['seg']:
        kw = dict(target_seg=torch.tensor(G['gen_%s_segt' % case]), seg_out=t64(G['gen_%s_sego' % case]),
                  generation_labels=G['generation_labels'], segmentation_equivalency=G['gen_%s_equiv' % case],
                  dice_weight=c['dice_w'])
    loss = OA.generator_loss(t64(G['gen_%s_target' % case]), t64(G['gen_%s_pred' % case]), t64(G['gen_%s_dout' % case]),
                             c['discr_w'], c['crop'], **kw)
    # (the shim subtracts the float32 operands in float32 before averaging in float64)
    assert abs(float(loss) - float(G['gen_%s_loss' % case])) <= 1e-7 * abs(float(G['gen_%s_loss' % case]))


@pytest.mark.parametrize('case', ['a', 'b'])
def test_discriminator_loss_terms_match_reference(case):
    """the reference's norm axes (spatial only -> one norm per batch element AND channel), penalty and sum, gradients given"""
    g = lambda k: G['disc_%s_%s' % (case, k)]  # noqa: E731
    loss = OA.gradient_penalty_terms(t64(g('real')), t64(g('fake')), t64(g('grads')), float(g('gpw')), 3)
    assert abs(float(loss) - float(g('loss'))) <= 1e-7 * abs(float(g('loss')))


def test_random_weighted_average_matches_reference():
    assert G['rwa_w'].shape == (3, 1, 1, 1, 1)                     # ONE weight per batch element
    out = OA.random_weighted_average(t64(G['rwa_a']), t64(G['rwa_b']), t64(G['rwa_w']))
    assert np.allclose(out.numpy(), G['rwa_out'], rtol=0, atol=1e-6)


def _product_disc(shape, **kw):
    from synthsr_b200.adversary import Discriminator
    return Discriminator(shape, device='cpu', dtype=torch.float64, **kw)


def test_discriminator_wiring_matches_reference():
    """layer sequence / filters / strides / units of make_discriminator as the reference builds it"""
    wiring = json.load(open(os.path.join(HERE, 'golden', 'reference_adversary_wiring.json')))['default']
    convs = [(w[1][0], w[2]['strides']) for w in wiring if w[0] == 'Conv3D']
    dense = [w[1][0] for w in wiring if w[0] == 'Dense']
    assert all(w[2] == {'alpha': 0.2} for w in wiring if w[0] == 'LeakyReLU')
    assert [w[2] for w in wiring if w[0] == 'Flatten'] == [{'data_format': 'channels_last'}]
    d = _product_disc([16, 16, 16, 1])
    got = [(d.p[name + '/kernel'].shape[-1], stride) for name, kind, stride in d.layers if kind == 'conv']
    assert got == convs
    assert [d.p['dense_1/kernel'].shape[1], d.p['dense_2/kernel'].shape[1]] == dense
    assert d.p['dense_1/kernel'].shape[0] == 1 * 256                # 16 -> 8 -> 4 -> 2 -> 1 per axis, 256 channels
    # kinds in order: 8 x (conv, leaky), flatten, dense, leaky, dense
    assert [w[0] for w in wiring] == ['Conv3D', 'LeakyReLU'] * 8 + ['Flatten', 'Dense', 'LeakyReLU', 'Dense']


@pytest.mark.parametrize('shape,mask', [([8, 8, 8, 1], False), ([6, 10, 7, 2], True)])
def test_product_discriminator_equals_oracle(shape, mask):
    """forward, discriminator loss (gradient penalty = double backward) and its parameter gradients; odd sizes exercise
    TensorFlow's asymmetric 'same' padding of the strided convolutions"""
    from synthsr_b200 import adversary as PA
    rng = np.random.default_rng(0)
    d = _product_disc(shape, n_filters=3, n_levels=2, mask_input=mask, seed=3)
    for k in d.p:
        if k.endswith('bias'):
            d.p[k].copy_(t64(rng.normal(size=d.p[k].shape) * .1))
    params = {k: v.clone() for k, v in d.p.items()}
    real, fake = t64(rng.uniform(size=(2, *shape))), t64(rng.uniform(size=(2, *shape)))
    m = t64(rng.integers(0, 2, size=(2, *shape[:3], 1))) if mask else None
    assert torch.allclose(d(real, m), OA.discriminator_forward(params, real, m, n_levels=2), rtol=1e-12, atol=1e-12)
    w = t64(rng.uniform(size=(2, 1, 1, 1, 1)))
    leaves = d.leaves()
    loss_p, _ = PA.discriminator_loss(d, real, fake, w, 10., m, leaves)
    oleaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    loss_o = OA.discriminator_loss(oleaves, real, fake, w, 10., m, n_levels=2)
    assert abs(float(loss_p) - float(loss_o)) <= 1e-10 * abs(float(loss_o))
    gp = torch.autograd.grad(loss_p, [leaves[k] for k in leaves])
    go = torch.autograd.grad(loss_o, [oleaves[k] for k in leaves])
    for k, a, b in zip(leaves, gp, go):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10), k


def test_gradient_penalty_against_finite_differences():
    """d(discriminator loss)/d(parameter) including the second-order term of the gradient penalty"""
    rng = np.random.default_rng(1)
    shape = [4, 4, 4, 1]
    d = _product_disc(shape, n_filters=2, n_levels=1, seed=5)
    params = {k: v.clone() for k, v in d.p.items()}
    real, fake = t64(rng.uniform(size=(1, *shape))), t64(rng.uniform(size=(1, *shape)))
    w = t64([[[[[.3]]]]])

    def f(p):
        return OA.discriminator_loss(p, real, fake, w, 10., None, n_levels=1)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    grads = dict(zip(leaves, torch.autograd.grad(f(leaves), list(leaves.values()))))
    for k in ('conv3d_1/kernel', 'dense_1/kernel', 'dense_2/bias', 'conv3d_2/bias'):
        flat = params[k].reshape(-1)
        for i in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            hi = {kk: vv.clone() for kk, vv in params.items()}
            lo = {kk: vv.clone() for kk, vv in params.items()}
            hi[k].reshape(-1)[i] += 1e-6
            lo[k].reshape(-1)[i] -= 1e-6
            fd = (float(f(hi)) - float(f(lo))) / 2e-6
            assert abs(fd - float(grads[k].reshape(-1)[i])) <= 1e-5 * max(1., abs(fd)), (k, int(i), fd)


def test_discriminator_adam_is_keras_adam():
    """the torch fall-back of Discriminator.adam_step (CPU) = keras.optimizers.Adam as oracle/unet.py restates it"""
    d = _product_disc([4, 4, 4, 1], n_filters=2, n_levels=1, seed=2)
    rng = np.random.default_rng(2)
    p0 = d.params.clone()
    m = v = torch.zeros_like(p0)
    p = p0.clone()
    for t in range(1, 4):
        g = t64(rng.normal(size=p0.shape))
        d.grads.copy_(g)
        d.adam_step(1e-3, lr_decay=.1)
        lr = 1e-3 / (1. + .1 * (t - 1))
        lr_t = lr * np.sqrt(1 - .999 ** t) / (1 - .9 ** t)
        m = .9 * m + .1 * g
        v = .999 * v + .001 * g * g
        p = p - lr_t * m / (v.sqrt() + 1e-7)
        assert torch.allclose(d.params, p, rtol=1e-12, atol=1e-14)


def test_fine_tuning_argument_checks_match_reference():
    import SynthSR.fine_tuning_with_adversary as FT
    with pytest.raises(Exception, match='please provide a value for output_channel or image_dir'):
        FT.training('l', None, 'm', None, None, 'g')
    with pytest.raises(Exception, match='but not both at the same time'):
        FT.training('l', 'i', 'm', None, None, 'g', output_channel=0)
    with pytest.raises(Exception, match='indices in output_channel cannot be greater'):
        FT.training('l', None, 'm', None, None, 'g', output_channel=3)
    with pytest.raises(Exception, match='The number or residual channels and output channels must be the same'):
        FT.training('l', None, 'm', None, None, 'g', input_channels=[True, True], output_channel=[0],
                    work_with_residual_channel=[0, 1])
