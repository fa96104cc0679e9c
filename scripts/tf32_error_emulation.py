"""CPU emulation of what TF32 operand rounding does to one U-Net training step (no GPU needed).

The tcgen05 path rounds activations (TMA TFLOAT32 load, round to nearest even) and weights (cvt.rna) to TF32 and
accumulates in fp32; the float64 oracle with the same rounding applied to the operands of every 3x3x3 convolution
reproduces its error level (tests/test_unet_gpu.py::_ConvTF32).  This script measures, against the exact float64
oracle, the error of the prediction / loss / gradients for a chosen size and a chosen set of layers and passes left in
plain TF32 (the others exact = what a 3xTF32 compensated convolution delivers), to decide where compensation is needed.

    python scripts/tf32_error_emulation.py --size 64 [--exact-fwd REGEX] [--exact-bwd REGEX] [--weights h5] [--metric l1]
"""
import argparse
import os
import re
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet as OU  # noqa: E402


def rne_tf32(t):
    u = t.float().contiguous().view(torch.int32)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~0x1FFF
    return u.view(torch.float32).to(t.dtype)


def rna_tf32(t):
    u = t.float().contiguous().view(torch.int32)
    u = (u + 0x1000) & ~0x1FFF
    return u.view(torch.float32).to(t.dtype)


LEVEL2 = None     # regex of layers compensated at level 2
ROUND = 'xw'      # which forward operands are rounded: 'x' (activations), 'w' (weights) or both


class ConvTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, pad, fwd_exact, bwd_exact):
        """fwd_exact: True (level 3: nothing rounded), 'w' (level 2: the weights stay rounded), False (plain TF32)"""
        ctx.save_for_backward(x, w)
        ctx.pad, ctx.bwd_exact = pad, bwd_exact
        if fwd_exact is True:
            return torch.nn.functional.conv3d(x, w, b, padding=pad)
        rnd = fwd_exact if isinstance(fwd_exact, str) else ROUND
        xr = rne_tf32(x) if 'x' in rnd else x
        wr = rna_tf32(w) if 'w' in rnd else w
        return torch.nn.functional.conv3d(xr, wr, b, padding=pad)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        if ctx.bwd_exact:
            gx = torch.nn.grad.conv3d_input(x.shape, w, gy, padding=ctx.pad)
            gw = torch.nn.grad.conv3d_weight(x, w.shape, gy, padding=ctx.pad)
        else:
            gx = torch.nn.grad.conv3d_input(x.shape, rna_tf32(w), rne_tf32(gy), padding=ctx.pad)
            gw = torch.nn.grad.conv3d_weight(rne_tf32(x), w.shape, rne_tf32(gy), padding=ctx.pad)
        return gx, gw, gy.sum((0, 2, 3, 4)), None, None, None


def make_conv(exact_fwd, exact_bwd):
    def conv(x, params, name):
        w = params[name + '/kernel'].permute(4, 3, 0, 1, 2)
        k = w.shape[-1]
        if w.shape[1] % 8 != 0 or k == 1:       # first layer / head are exact fp32 on the GPU too
            return torch.nn.functional.conv3d(x, w, params[name + '/bias'], padding=k // 2)
        fe = exact_fwd is not None and re.search(exact_fwd, name) is not None
        if LEVEL2 is not None and re.search(LEVEL2, name) is not None:
            fe = 'w'                               # activation residual only (2 MMAs): the weight rounding remains
        be = exact_bwd is not None and re.search(exact_bwd, name) is not None
        return ConvTF32.apply(x, w, params[name + '/bias'], k // 2, fe, be)
    return conv


def step(params, image, target, conv, metric):
    names = OU.trainable_names(params)
    leaves = {k: params[k].detach().clone().requires_grad_(True) for k in names}
    p = {k: leaves.get(k, params[k]) for k in params}
    keep = OU._conv
    try:
        if conv is not None:
            OU._conv = conv
        pred = OU.forward(p, image, training=True)
    finally:
        OU._conv = keep
    loss = OU.loss_fn(pred, image, target, metric=metric)
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[k] for k in names])))
    return pred.detach(), float(loss), grads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=64)
    ap.add_argument('--exact-fwd', default=None, help='regex of layer names whose FORWARD is exact (compensated)')
    ap.add_argument('--exact-bwd', default=None, help='regex of layer names whose dgrad/wgrad are exact')
    ap.add_argument('--metric', default='l1')
    ap.add_argument('--weights', default=None, help='.h5 file (trained Keras weights) instead of glorot init')
    ap.add_argument('--image', default='uniform', choices=['uniform', 'smooth'])
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--level2', default=None, help='regex of layers whose forward keeps only the weight rounding')
    ap.add_argument('--round', default='xw', help="forward operands rounded to TF32: x, w or xw")
    args = ap.parse_args()
    global ROUND, LEVEL2
    ROUND, LEVEL2 = args.round, args.level2
    n = args.size
    params = OU.init_params(args.seed, 1, dtype=torch.float64)
    if args.weights:
        from synthsr_b200 import h5lite
        sd, _ = h5lite.load_keras_weights(args.weights)
        for k in params:
            params[k] = torch.tensor(np.asarray(sd[k]), dtype=torch.float64)
    rng = np.random.default_rng(args.seed + 1)
    if args.image == 'uniform':
        image = rng.uniform(0, 1, size=(1, n, n, n, 1))
        target = rng.uniform(0, 1, size=(1, n, n, n, 1))
    else:
        from scipy.ndimage import gaussian_filter
        image = gaussian_filter(rng.uniform(0, 1, size=(n, n, n)), 2.)[None, ..., None]
        image = (image - image.min()) / (image.max() - image.min())
        target = gaussian_filter(image[0, ..., 0], 1.)[None, ..., None]
    image, target = torch.tensor(image, dtype=torch.float64), torch.tensor(target, dtype=torch.float64)
    t0 = time.time()
    pred0, loss0, g0 = step(params, image, target, None, args.metric)
    t1 = time.time()
    pred1, loss1, g1 = step(params, image, target, make_conv(args.exact_fwd, args.exact_bwd), args.metric)
    t2 = time.time()
    e_l2 = float((pred1 - pred0).norm() / pred0.norm())
    e_max = float((pred1 - pred0).abs().max() / pred0.abs().max())
    gtot = np.sqrt(sum(float((g0[k] ** 2).sum()) for k in g0))
    gerr = {k: float((g1[k] - g0[k]).norm()) / max(float(g0[k].norm()), 1e-2 * gtot) for k in g0}
    worst = sorted(gerr.items(), key=lambda kv: -kv[1])[:6]
    print('size %d metric %s weights %s image %s exact_fwd=%s exact_bwd=%s (%.0f s + %.0f s)' % (
        n, args.metric, 'h5' if args.weights else 'glorot', args.image, args.exact_fwd, args.exact_bwd, t1 - t0, t2 - t1))
    print('  pred relL2 %.3e max/max %.3e loss rel %.3e' % (e_l2, e_max, abs(loss1 - loss0) / abs(loss0)))
    print('  grads: ' + ', '.join('%s %.2e' % (k.replace('unet_', ''), e) for k, e in worst))
    gl2 = np.sqrt(sum(float(((g1[k] - g0[k]) ** 2).sum()) for k in g0)) / gtot
    print('  whole-gradient relL2 %.3e' % gl2)


if __name__ == '__main__':
    main()
