"""Activation-gradient errors layer by layer: float64 oracle (pre-activation gradients captured with retain_grad) against the
GPU buffers of one step (ga / gb / ga_e / gb_e hold the gradient w.r.t. each convolution's pre-activation).
   python scripts/actgrad_diag.py 96 [noise|gen] [l1|l2]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import unet as OU
from synthsr_b200.unet import UNet3D

size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
kind = sys.argv[2] if len(sys.argv) > 2 else 'noise'
metric = sys.argv[3] if len(sys.argv) > 3 else 'l1'
dims = [size] * 3
if kind == 'gen':
    from test_unet_parity_gpu import _generated_batch
    image, target = _generated_batch(size)
else:
    rng = np.random.default_rng(1)
    image = rng.uniform(0, 1, size=(1, *dims, 1)).astype(np.float32)
    target = rng.uniform(0, 1, size=(1, *dims, 1)).astype(np.float32)

ref = UNet3D(dims + [1], batchsize=1, conv_impl='ref', seed=0)
params = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in ref.state_dict().items()}
names = OU.trainable_names(params)
leaves = {k: params[k].clone().requires_grad_(True) for k in names}
p = {k: leaves.get(k, params[k]) for k in params}
store = {}
orig = OU._conv
def conv(x, pp, name):
    y = orig(x, pp, name)
    y.retain_grad()
    store[name] = y
    return y
OU._conv = conv
img, tgt = torch.tensor(image, dtype=torch.float64), torch.tensor(target, dtype=torch.float64)
pred = OU.forward(p, img, training=True)
OU._conv = orig
loss = OU.loss_fn(pred, img, tgt, metric=metric)
loss.backward()
L = 5
oracle = {k: v.grad.permute(0, 2, 3, 4, 1).reshape(-1, v.shape[1]).numpy() for k, v in store.items() if v.grad is not None}
wg = {k: leaves[k].grad.numpy() for k in names}
gtot = np.sqrt(sum(float((g ** 2).sum()) for g in wg.values()))
img_t, tgt_t = torch.from_numpy(image).cuda(), torch.from_numpy(target).cuda()


def run(tag, net):
    net.loss_and_grad(img_t, tgt_t, metric)
    torch.cuda.synchronize()
    print('== %s  (%d^3, %s inputs, %s)' % (tag, size, kind, metric))
    rows = []
    for l in range(L - 1):
        d = L - 2 - l
        rows += [('unet_conv_uparm_%d_1' % (L + d), net.ga[l]), ('unet_conv_uparm_%d_0' % (L + d), net.gb[l])]
    for l in range(L - 1, -1, -1):
        rows += [('unet_conv_downarm_%d_1' % l, net.ga_e[l]), ('unet_conv_downarm_%d_0' % l, net.gb_e[l])]
    for name, buf in rows:
        g = buf.cpu().numpy().astype(np.float64)
        o = oracle[name]
        e = g - o
        rms = np.sqrt((o ** 2).mean())
        kerr = np.linalg.norm(net.g[name + '/kernel'].cpu().numpy().astype(np.float64) - wg[name + '/kernel']) / max(
            np.linalg.norm(wg[name + '/kernel']), 1e-2 * gtot)
        print('%-26s dL/dpre: rel L2 %.2e   |channel-mean err| / rms %.2e   |true channel mean| / rms %.2e   kernel grad err %.2e' % (
            name.replace('unet_conv_', ''), np.linalg.norm(e) / np.linalg.norm(o), np.abs(e.mean(0)).max() / rms,
            np.abs(o.mean(0)).max() / rms, kerr), flush=True)


run('ref', ref)
del ref
torch.cuda.empty_cache()
run('tc3', UNet3D(dims + [1], batchsize=1, conv_impl='tc3', seed=0))
