"""Host-side input sampler with the generator protocol of the reference (SynthSR/model_inputs.py:25-139): picks
random label map(s), draws per-class GMM means/stds from the priors, yields [labels, means, stds(, image)].

Difference from the reference: decoded label maps are cached (the reference re-decodes a gzip NIfTI every step,
:91, which alone would cap throughput below 20 volumes/s).  The cache is least-recently-used with a byte budget
(SSR_LABEL_CACHE_GB, default 8), so a dataset of thousands of label maps degrades to re-decoding instead of exhausting
host memory."""
import os
from collections import OrderedDict

import numpy as np
import numpy.random as npr

from ext.lab2im import utils


class _VolumeCache(OrderedDict):
    def __init__(self, budget_bytes):
        super().__init__()
        self.budget, self.used = int(budget_bytes), 0

    def get_or_load(self, key, loader):
        if key in self:
            self.move_to_end(key)
            return self[key]
        vol = loader()
        if vol.nbytes <= self.budget:
            self[key] = vol
            self.used += vol.nbytes
            while self.used > self.budget:
                _, old = self.popitem(last=False)
                self.used -= old.nbytes
        return vol


def build_model_inputs(path_label_maps, n_labels, prior_means, prior_stds, prior_distributions, path_images=None,
                       batchsize=1, n_channels=1, generation_classes=None, cache=None):
    _, _, n_dims, _, _, _ = utils.get_volume_info(path_label_maps[0])
    if generation_classes is None:
        generation_classes = np.arange(n_labels)
    n_classes = len(np.unique(generation_classes))
    if cache is None:
        cache = _VolumeCache(float(os.environ.get('SSR_LABEL_CACHE_GB', '8')) * (1 << 30))

    def load(path, dtype):
        return cache.get_or_load((path, dtype), lambda: utils.load_volume(path, dtype=dtype, aff_ref=np.eye(4)))

    while True:
        indices = npr.randint(len(path_label_maps), size=batchsize)
        labels, means_l, stds_l, images = [], [], [], []
        for idx in indices:
            labels.append(utils.add_axis(load(path_label_maps[idx], 'int'), axis=[0, -1]))
            if path_images is not None:
                images.append(load(path_images[idx], 'float')[np.newaxis, :, :, :, np.newaxis])
            means, stds = np.empty((1, n_labels, 0)), np.empty((1, n_labels, 0))
            for channel in range(n_channels):
                pm, ps = prior_means, prior_stds
                if isinstance(pm, np.ndarray):
                    if pm.shape[0] / 2 != n_channels:
                        raise ValueError("the number of blocks in prior_means does not match n_channels.")
                    pm = pm[2 * channel:2 * channel + 2, :]
                if isinstance(ps, np.ndarray):
                    if ps.shape[0] / 2 != n_channels:
                        raise ValueError("the number of blocks in prior_stds does not match n_channels.")
                    ps = ps[2 * channel:2 * channel + 2, :]
                cm = utils.draw_value_from_distribution(pm, n_classes, prior_distributions, 125., 100., positive_only=True)
                cs = utils.draw_value_from_distribution(ps, n_classes, prior_distributions, 15., 10., positive_only=True)
                means = np.concatenate([means, utils.add_axis(cm[generation_classes], axis=[0, -1])], axis=-1)
                stds = np.concatenate([stds, utils.add_axis(cs[generation_classes], axis=[0, -1])], axis=-1)
            means_l.append(means)
            stds_l.append(stds)
        inputs = [labels, means_l, stds_l] + ([images] if path_images is not None else [])
        yield [np.concatenate(item, 0) for item in inputs] if batchsize > 1 else [item[0] for item in inputs]
