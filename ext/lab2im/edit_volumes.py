"""The helpers of ext/lab2im/edit_volumes.py that run on the training path (per label-map load and per
generate_brain()): get_ras_axes (reference :591-606) and align_volume_to_ref (:609-654), and on the inference path
(scripts/predict_command_line.py:113-116): resample_volume (:504-553) and resample_volume_like (:556-588).  Host-side
NumPy / SciPy like the reference (one call per input scan).  The offline batch tools of that module are out of scope
(SURVEY.md 2a #11)."""
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


def resample_volume(volume, aff, new_vox_size, interpolation='linear', blur=True):
    """resize the voxels of `volume` to `new_vox_size` (mm), adjusting the affine so that world coordinates are kept.
    Same arithmetic as the reference (:504-553): voxel size from the affine columns, Gaussian pre-filter with
    sigma = 0.25 / factor on the axes that are downsampled, samples at start + n * step with start = -(f-1)/(2f),
    step = 1/f, clamped to the volume, linear interpolation (scipy RegularGridInterpolator)."""
    from scipy.interpolate import RegularGridInterpolator
    from scipy.ndimage import gaussian_filter
    aff = np.asarray(aff, dtype=np.float64)
    pixdim = np.sqrt(np.sum(aff * aff, axis=0))[:-1]
    new_vox_size = np.array(new_vox_size)
    factor = pixdim / new_vox_size
    sigmas = 0.25 / factor
    sigmas[factor > 1] = 0                                   # no blur when upsampling
    volume_filt = gaussian_filter(volume, sigmas) if blur else volume
    grids = [np.arange(0, n) for n in volume_filt.shape[:3]]
    interp = RegularGridInterpolator(tuple(grids), volume_filt, method=interpolation)
    start = - (factor - 1) / (2 * factor)
    step = 1.0 / factor
    stop = start + step * np.ceil(volume_filt.shape * factor)
    samples = []
    for d in range(3):
        xi = np.arange(start=start[d], stop=stop[d], step=step[d])
        xi[xi < 0] = 0
        xi[xi > (volume_filt.shape[d] - 1)] = volume_filt.shape[d] - 1
        samples.append(xi)
    xig, yig, zig = np.meshgrid(*samples, indexing='ij', sparse=True)
    volume2 = interp((xig, yig, zig))
    aff2 = aff.copy()
    for c in range(3):
        aff2[:-1, c] = aff2[:-1, c] / factor[c]
    aff2[:-1, -1] = aff2[:-1, -1] - np.matmul(aff2[:-1, :-1], 0.5 * (factor - 1))
    return volume2, aff2


def resample_volume_like(vol_ref, aff_ref, vol_flo, aff_flo, interpolation='linear'):
    """reslice the floating image into the voxel grid of the reference image (reference :556-588): coordinates
    T = inv(aff_flo) . aff_ref applied to the reference grid, linear interpolation, zero outside the floating volume."""
    from scipy.interpolate import RegularGridInterpolator
    T = np.matmul(np.linalg.inv(aff_flo), aff_ref)
    grids = tuple(np.arange(0, n) for n in vol_flo.shape[:3])
    interp = RegularGridInterpolator(grids, vol_flo, bounds_error=False, fill_value=0.0, method=interpolation)
    xr, yr, zr = (np.arange(0, n) for n in vol_ref.shape[:3])
    xrg, yrg, zrg = np.meshgrid(xr, yr, zr, indexing='ij', sparse=False)
    n = xrg.size
    coords = np.stack([xrg.reshape([n]), yrg.reshape([n]), zrg.reshape([n]), np.ones(n, dtype=xrg.dtype)])
    coords_new = np.matmul(T, coords)[:-1, :]
    result = interp((coords_new[0, :], coords_new[1, :], coords_new[2, :]))
    return result.reshape(vol_ref.shape[:3])
