import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synthsr_b200.unet import UNet3D
dims, rng = [32, 48, 32], np.random.default_rng(11)
image = torch.from_numpy(rng.uniform(0, 1, size=(1, *dims, 1)).astype(np.float32)).cuda()
target = torch.from_numpy(rng.uniform(0, 1, size=(1, *dims, 1)).astype(np.float32)).cuda()
nets = []
for mode in ('all', 'k2n_only', 'none'):
    net = UNet3D(dims + [1], nb_levels=3, batchsize=1, conv_impl='tc', seed=3)
    net.epi_fusion = mode != 'none'
    net.epi_fusion_generic = mode == 'all'
    net.loss_and_grad(image, target)
    torch.cuda.synchronize()
    nets.append(net)
a, b, c = nets
def rel(x, y): return ((x - y).abs().max() / (y.abs().max() + 1e-30)).item()
for nm, A, B_, C in (('fused-all vs none', a, c, None), ('k2n-only vs none', b, c, None)):
    print(nm)
    for l in range(3):
        print('  enc', l, 'h1', rel(A.h1[l], B_.h1[l]), 'stats', rel(A.stats_enc[l], B_.stats_enc[l]),
              [rel(A.stats_enc[l][i * A.feats[l]:(i + 1) * A.feats[l]], B_.stats_enc[l][i * A.feats[l]:(i + 1) * A.feats[l]]) for i in range(4)])
    for l in range(2):
        print('  dec', l, 'g1', rel(A.g1[l], B_.g1[l]), 'stats', rel(A.stats_dec[l], B_.stats_dec[l]))
    print('  pred', rel(A.pred, B_.pred), 'grads', ((A.grads - B_.grads).norm() / B_.grads.norm()).item())
    for k in A.g:
        e = ((A.g[k] - B_.g[k]).norm() / (B_.g[k].norm() + 1e-30)).item()
        if e > 1e-4: print('   grad', k, e)
