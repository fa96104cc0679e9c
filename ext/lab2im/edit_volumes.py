"""The two orientation helpers of ext/lab2im/edit_volumes.py that run on the training path (per label-map load and
per generate_brain()): get_ras_axes (reference :591-606) and align_volume_to_ref (:609-654).  The offline batch
tools of that module are out of scope."""
import numpy as np


def get_ras_axes(aff, n_dims=3):
    """index of the voxel axis that maps onto each of the R, A, S world axes."""
    inv = np.linalg.inv(np.asarray(aff, dtype=np.float64))
    axes = np.argmax(np.abs(inv[:n_dims, :n_dims]), axis=0)
    for i in range(n_dims):                      # repair degenerate affines: every axis must appear exactly once
        if i not in axes:
            vals, counts = np.unique(axes, return_counts=True)
            dup = vals[np.argmax(counts)]
            axes[np.where(axes == dup)[0][-1]] = i
    return axes


def align_volume_to_ref(volume, aff, aff_ref=None, return_aff=False, n_dims=None, return_copy=True):
    """permute / flip the voxel axes of `volume` so that its orientation matches `aff_ref` (default identity)."""
    vol = volume.copy() if return_copy else volume
    aff_flo = np.array(aff, dtype=np.float64)
    aff_ref = np.eye(4) if aff_ref is None else np.asarray(aff_ref, dtype=np.float64)
    if n_dims is None:
        n_dims = vol.ndim if vol.shape[-1] > 10 else vol.ndim - 1
    ref_axes, flo_axes = get_ras_axes(aff_ref, n_dims), get_ras_axes(aff_flo, n_dims)
    aff_flo[:, ref_axes] = aff_flo[:, flo_axes]
    for i in range(n_dims):
        if flo_axes[i] != ref_axes[i]:
            vol = np.swapaxes(vol, flo_axes[i], ref_axes[i])
            j = int(np.where(flo_axes == ref_axes[i])[0][0])
            flo_axes[j], flo_axes[i] = flo_axes[i], flo_axes[j]
    dots = np.sum(aff_flo[:3, :3] * aff_ref[:3, :3], axis=0)
    for i in range(n_dims):
        if dots[i] < 0:
            vol = np.flip(vol, axis=i)
            aff_flo[:, i] = -aff_flo[:, i]
            aff_flo[:3, 3] = aff_flo[:3, 3] - aff_flo[:3, i] * (vol.shape[i] - 1)
    return (vol, aff_flo) if return_aff else vol
