"""Diagnostic for tests/test_adversary_gpu.py: per-tensor gradient errors of the fine-tuned U-Net's step against the float64
oracle for (C) the plain L1 step of the same small network, (B) the adversarial head with the discriminator term zeroed,
(A) the full generator loss.   python scripts/adv_grad_diag.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_adversary_gpu as T  # noqa: E402
from helpers import gpu_pool_routing  # noqa: E402
from synthsr_b200 import adversary as PA  # noqa: E402


def run(tag, discr_weight, zero_adv=False, conv_impl='tc3'):
    rng = np.random.default_rng(3)
    net, disc = T._make(net_kw=dict(discr_weight=discr_weight))
    if conv_impl != 'tc3':
        from synthsr_b200.adversary import AdversarialUNet3D
        net = AdversarialUNet3D([32, 32, 32, 1], 8, 3, 3, 1, 2, 2, 1, 'cuda', conv_impl, seed=0, seg=None, disc=disc,
                                discr_weight=discr_weight)
    image = rng.uniform(0, 1, size=(1, 32, 32, 32, 1)).astype(np.float32)
    target = rng.uniform(0, 1, size=(1, 32, 32, 32, 1)).astype(np.float32)
    keep = PA.wasserstein_generator_term
    if zero_adv:
        PA.wasserstein_generator_term = lambda d, p, m=None: (torch.zeros((), device='cuda'), torch.zeros_like(p))
    try:
        loss = net.loss_and_grad(T._t(image), T._t(target), 'l1', None, None)
    finally:
        PA.wasserstein_generator_term = keep
    torch.cuda.synchronize()
    if zero_adv:        # oracle: l1_weight * L1 only
        class Z:
            def state_dict(self):
                return {k: np.zeros_like(v) for k, v in disc.state_dict().items()}
        pred_o, loss_o, grads_o = T._oracle_generator_step(net, Z(), image, target, discr_weight, None, None, gpu_pool_routing(net))
    else:
        pred_o, loss_o, grads_o = T._oracle_generator_step(net, disc, image, target, discr_weight, None, None, gpu_pool_routing(net))
    gtot = np.sqrt(sum(float((g ** 2).sum()) for g in grads_o.values()))
    errs = {k: np.linalg.norm(net.g[k].cpu().numpy().astype(np.float64) - g) / max(np.linalg.norm(g), 1e-2 * gtot) for k, g in grads_o.items()}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print('%-40s loss %.6f oracle %.6f | worst tensors: %s' % (tag, loss.item(), loss_o, ', '.join('%s %.2e' % (k.replace('unet_', ''), v) for k, v in worst)))


run('A full generator loss, w_d = 0.05', .05)
run('A full generator loss, w_d = 0.05, ref mode', .05, conv_impl='ref')
run('B discriminator term zeroed, w_d = 0.05', .05, zero_adv=True)
run('B zeroed, ref mode', .05, zero_adv=True, conv_impl='ref')
run('A w_d = 0.3', .3)
run('A w_d = 0.3, ref mode', .3, conv_impl='ref')
