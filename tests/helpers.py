"""Shared synthetic inputs for the parity tests (seeded; nothing here reads /root/reference)."""
import numpy as np

GEN_LABELS = np.array([0, 14, 15, 16, 2, 3, 4, 5, 7, 8, 10, 11, 12, 13, 17, 18, 26, 28, 31])   # fixture label list
GEN_CLASSES = np.array([0, 3, 3, 4, 1, 2, 3, 3, 1, 2, 5, 6, 7, 8, 9, 10, 11, 12, 13])
SIDED_LABELS = np.array([0, 14, 15, 16, 2, 3, 4, 41, 42, 43])                                  # 4 neutral, 3 L, 3 R


def phantom_labels(shape, label_list, seed=0, n_seeds=40):
    """Smooth Voronoi phantom: neighbouring voxels share labels like anatomy."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0, 1, size=(n_seeds, 3)) * np.array(shape)
    labs = rng.choice(label_list, size=n_seeds)
    g = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing='ij'), -1).astype(np.float32)
    d = ((g[..., None, :] - pts[None, None, None]) ** 2).sum(-1)
    lab = labs[np.argmin(d, -1)]
    c = (np.array(shape) - 1) / 2
    r = np.sqrt((((g - c) / (np.array(shape) * 0.45)) ** 2).sum(-1))
    lab[r > 1] = 0
    return lab.astype(np.int32)


def gmm_params(rng, n_labels, n_channels, batch=1):
    means = rng.uniform(20, 225, size=(batch, n_labels, n_channels)).astype(np.float32)
    stds = rng.uniform(3, 25, size=(batch, n_labels, n_channels)).astype(np.float32)
    means[:, 0] = 0
    stds[:, 0] = 0
    return means, stds
