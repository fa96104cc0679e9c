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
