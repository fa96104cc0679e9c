"""Where does the gradient error of a full-size step come from?  One float64 oracle step (CPU), then the GPU step in several
configurations; per-tensor relative L2 of the first layers' gradients.   python scripts/grad_error_diag.py 96"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from test_unet_parity_gpu import _oracle64
from synthsr_b200.unet import UNet3D

size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
dims = [size] * 3
rng = np.random.default_rng(1)
image = rng.uniform(0, 1, size=(1, *dims, 1)).astype(np.float32)
target = rng.uniform(0, 1, size=(1, *dims, 1)).astype(np.float32)
ref = UNet3D(dims + [1], batchsize=1, conv_impl='ref', seed=0)
pred_o, loss_o, grads_o = _oracle64(ref.state_dict(), image, target, 5)
gtot = np.sqrt(sum(float((g ** 2).sum()) for g in grads_o.values()))
img_t, tgt_t = torch.from_numpy(image).cuda(), torch.from_numpy(target).cuda()


def report(tag, net):
    net.loss_and_grad(img_t, tgt_t)
    torch.cuda.synchronize()
    e = {k: np.linalg.norm(net.g[k].cpu().numpy().astype(np.float64) - g) / max(np.linalg.norm(g), 1e-2 * gtot)
         for k, g in grads_o.items()}
    top = sorted(e.items(), key=lambda kv: -kv[1])[:5]
    pred = net.pred.view(pred_o.shape).cpu().numpy().astype(np.float64)
    print('%-46s pred %.2e | %s' % (tag, np.linalg.norm(pred - pred_o) / np.linalg.norm(pred_o),
                                    ', '.join('%s %.2e' % (k.replace('unet_conv_', '').replace('unet_', ''), v) for k, v in top)), flush=True)


report('ref (exact fp32 CUDA cores)', ref)
del ref
torch.cuda.empty_cache()
for tag, env, attrs in [('tc3', {}, {}),
                        ('tc3, weight gradients on the CUDA cores', {}, {'wgrad_tc': False}),
                        ('tc3, no fused epilogues', {'SSR_NO_EPI_FUSION': '1', 'SSR_NO_EPI_FUSION_GENERIC': '1'}, {}),
                        ('tc3, no pool+BN fusion, no head BN sums', {'SSR_NO_POOL_BN_FUSION': '1', 'SSR_NO_HEAD_BN_SUMS': '1'}, {}),
                        ('tc3, no parity path', {'SSR_NO_UP_PARITY': '1'}, {}),
                        ('tc3, no wgrad overlap (single stream)', {'SSR_NO_WGRAD_OVERLAP': '1'}, {})]:
    for k, v in env.items():
        os.environ[k] = v
    net = UNet3D(dims + [1], batchsize=1, conv_impl='tc3', seed=0)
    for k, v in attrs.items():
        setattr(net, k, v)
    report(tag, net)
    del net
    torch.cuda.empty_cache()
    for k in env:
        del os.environ[k]
