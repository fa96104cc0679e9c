"""Shared synthetic inputs for the parity tests (seeded; nothing here reads /root/reference)."""
import numpy as np

from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels  # noqa: F401

SIDED_LABELS = np.array([0, 14, 15, 16, 2, 3, 4, 41, 42, 43])                                  # 4 neutral, 3 L, 3 R


def gmm_params(rng, n_labels, n_channels, batch=1):
    means = rng.uniform(20, 225, size=(batch, n_labels, n_channels)).astype(np.float32)
    stds = rng.uniform(3, 25, size=(batch, n_labels, n_channels)).astype(np.float32)
    means[:, 0] = 0
    stds[:, 0] = 0
    return means, stds


def gpu_pool_routing(net):
    """max-pool winners of the GPU forward that has just run, per encoder level, in F.max_pool3d's index convention:
    BatchNorm of the level's output through the library's own kernel (ssr_bn_apply mode 0), torch's max_pool3d for the
    indices -- checked bit for bit against the pooled tensor the forward itself produced (BN + pool fused, mode 1)."""
    import torch
    from synthsr_b200._lib import lib, stream_ptr
    routing = []
    for l in range(net.L - 1):
        d, c = net.ldims[l], net.feats[l]
        bn = torch.empty_like(net.h1[l])
        lib.ssr_bn_apply(net.h1[l], bn, net.stats_enc[l], net.B, *d, c, 0, 0, 0, stream_ptr())
        pooled, idx = torch.nn.functional.max_pool3d(bn.view(net.B, *d, c).permute(0, 4, 1, 2, 3), 2, return_indices=True)
        assert torch.equal(pooled.permute(0, 2, 3, 4, 1).reshape(net.inp[l + 1].shape), net.inp[l + 1]), l
        routing.append(idx.cpu())
    return routing
