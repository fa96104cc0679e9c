"""Seeded synthetic inputs of the reference's shapes (no dataset / checkpoint is reachable): label phantoms and GMM
priors.  Used by bench.py, smoke() and the tests."""
import numpy as np

# label list / class grouping of the reference's tutorial data (data/labels_classes_priors/*.npy): values only
GEN_LABELS = np.array([0, 14, 15, 16, 2, 3, 4, 5, 7, 8, 10, 11, 12, 13, 17, 18, 26, 28, 31], dtype=np.int32)
GEN_CLASSES = np.array([0, 3, 3, 4, 1, 2, 3, 3, 1, 2, 5, 6, 7, 8, 9, 10, 11, 12, 13], dtype=np.int32)


def phantom_labels(shape, label_list=GEN_LABELS, seed=0, n_seeds=40):
    """Smooth Voronoi phantom inside an ellipsoid: neighbouring voxels share labels like anatomy (a uniform-random
    label map would be unrealistically gather-hostile)."""
    rng = np.random.default_rng(seed)
    shape = [int(s) for s in shape]
    pts = (rng.uniform(0, 1, size=(n_seeds, 3)) * np.array(shape)).astype(np.float32)
    labs = rng.choice(np.asarray(label_list), size=n_seeds)
    out = np.zeros(shape, dtype=np.int32)
    ax = [np.arange(s, dtype=np.float32) for s in shape]
    c = (np.array(shape, dtype=np.float32) - 1) / 2
    for i0 in range(0, shape[0], 16):                      # chunked over the first axis to bound memory
        g = np.stack(np.meshgrid(ax[0][i0:i0 + 16], ax[1], ax[2], indexing='ij'), -1)
        d = ((g[..., None, :] - pts) ** 2).sum(-1)
        lab = labs[np.argmin(d, -1)]
        r = np.sqrt((((g - c) / (np.array(shape, dtype=np.float32) * 0.45)) ** 2).sum(-1))
        lab[r > 1] = 0
        out[i0:i0 + 16] = lab
    return out


def synthetic_priors(n_classes=14, n_channels=1, seed=0):
    """(2*n_channels, K) arrays like prior_means_*.npy / prior_stds_*.npy: row 0 mean, row 1 std of the prior."""
    rng = np.random.default_rng(seed)
    means = np.concatenate([np.stack([rng.uniform(40, 230, n_classes), rng.uniform(3, 15, n_classes)])
                            for _ in range(n_channels)])
    stds = np.concatenate([np.stack([rng.uniform(5, 30, n_classes), rng.uniform(1, 6, n_classes)])
                           for _ in range(n_channels)])
    means[0::2, 0] = 0; means[1::2, 0] = 0; stds[0::2, 0] = 0; stds[1::2, 0] = 0     # background class
    return means, stds
