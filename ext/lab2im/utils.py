"""Host-side helpers with the names / signatures / error behaviour of the reference's ext/lab2im/utils.py that its
scripts and SynthSR/* call directly.  No TensorFlow, no nibabel: volume I/O goes through
synthsr_b200.nifti.  Reference line numbers are cited per function."""
import glob
import math
import os
import time
from datetime import timedelta

import numpy as np

from synthsr_b200 import nifti

_NUM = (int, float, np.integer, np.floating)


# ----------------------------------------------------------------- I/O (reference :76-206) -------------------------
def load_volume(path_volume, im_only=True, squeeze=True, dtype=None, aff_ref=None):
    assert path_volume.endswith(('.nii', '.nii.gz', '.mgz', '.npz')), 'Unknown data file: %s' % path_volume
    if path_volume.endswith('.npz'):
        volume = np.load(path_volume)['vol_data']
        aff, header = np.eye(4), nifti.blank_header()
    elif path_volume.endswith('.mgz'):
        volume, aff, header = nifti.load_mgz(path_volume)
    else:
        volume, aff, header = nifti.load_nifti(path_volume)
    if squeeze:
        volume = np.squeeze(volume)
    if dtype is not None:
        if 'int' in dtype:
            volume = np.round(volume)
        volume = volume.astype(dtype=dtype)
    if aff_ref is not None:
        from . import edit_volumes
        n_dims, _ = get_dims(list(volume.shape), max_channels=10)
        volume, aff = edit_volumes.align_volume_to_ref(volume, aff, aff_ref=aff_ref, return_aff=True, n_dims=n_dims)
    return volume if im_only else (volume, aff, header)


def save_volume(volume, aff, header, path, res=None, dtype=None, n_dims=3):
    mkdir(os.path.dirname(path))
    if '.npz' in path:
        np.savez_compressed(path, vol_data=volume)
        return
    if isinstance(aff, str) and aff == 'FS':
        aff = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]])
    elif aff is None:
        aff = np.eye(4)
    if dtype is not None and 'int' in dtype:
        volume = np.round(volume)
    header = header.copy() if isinstance(header, nifti.Header) else None
    if res is not None:
        if n_dims is None:
            n_dims, _ = get_dims(volume.shape)
        header = header or nifti.blank_header()
        header.set_zooms(reformat_to_list(res, length=n_dims))
    if path.endswith(('.mgz', '.mgh')):                  # nib.save picks the image class from the extension (utils.py:158-160)
        nifti.save_mgz(path, volume, aff, header, dtype=dtype)
        return
    nifti.save_nifti(path, volume, aff, header, dtype=dtype)


def get_volume_info(path_volume, return_volume=False, aff_ref=None, max_channels=10):
    im, aff, header = load_volume(path_volume, im_only=False)
    im_shape = list(im.shape)
    n_dims, n_channels = get_dims(im_shape, max_channels=max_channels)
    im_shape = im_shape[:n_dims]
    if '.nii' in path_volume:
        data_res = np.array(header['pixdim'][1:n_dims + 1], dtype=np.float64)
    elif '.mgz' in path_volume:
        data_res = np.array(header['delta'], dtype=np.float64)
    else:
        data_res = np.array([1.0] * n_dims)
    if aff_ref is not None:
        from . import edit_volumes
        ras, ras_ref = edit_volumes.get_ras_axes(aff, n_dims=n_dims), edit_volumes.get_ras_axes(aff_ref, n_dims=n_dims)
        im = edit_volumes.align_volume_to_ref(im, aff, aff_ref=aff_ref, n_dims=n_dims)
        shp, res = np.array(im_shape), np.array(data_res)
        shp[ras_ref], res[ras_ref] = np.array(im_shape)[ras], np.array(data_res)[ras]
        im_shape, data_res = shp.tolist(), res
    if return_volume:
        return im, im_shape, aff, n_dims, n_channels, header, data_res
    return im_shape, aff, n_dims, n_channels, header, data_res


# FreeSurfer label taxonomy used by get_list_labels(FS_sort=True) (reference :248-267)
_NEUTRAL_FS = set([0, 14, 15, 16, 21, 22, 23, 24, 72, 77, 80, 85, 165, 258, 259, 260, 530] + list(range(100, 110)) +
                  list(range(200, 211)) + list(range(251, 256)) + list(range(331, 341)) +
                  [502, 506, 507, 508, 509, 511, 512, 514, 515, 516, 517] + list(range(531, 538)))


def _fs_side(la):
    if la in _NEUTRAL_FS:
        return 'n'
    if (0 < la < 14) or (16 < la < 21) or (24 < la < 40) or (135 < la < 139) or (1000 <= la <= 1035) or la == 865 \
            or (20100 < la < 20110):
        return 'l'
    if (39 < la < 72) or (162 < la < 165) or (2000 <= la <= 2035) or (20000 < la < 20010) or la in (139, 866):
        return 'r'
    raise Exception('label {} not in our current FS classification, please update get_list_labels in utils.py'.format(la))


def get_list_labels(label_list=None, labels_dir=None, save_label_list=None, FS_sort=False):
    if label_list is not None:
        label_list = np.array(reformat_to_list(label_list, load_as_numpy=True, dtype='int'))
    elif labels_dir is not None:
        print('Compiling list of unique labels')
        label_list = np.empty(0)
        for path in list_images_in_folder(labels_dir):
            label_list = np.unique(np.concatenate((label_list, np.unique(load_volume(path, dtype='int32'))))).astype('int')
    else:
        raise Exception('either label_list, path_label_list or labels_dir should be provided')
    n_neutral = 0
    if FS_sort:
        groups = {'n': [], 'l': [], 'r': []}
        for la in label_list:
            g = groups[_fs_side(int(la))]
            if la not in g:
                g.append(la)
        label_list = np.concatenate([sorted(groups['n']), sorted(groups['l']), sorted(groups['r'])])
        both, none = bool(groups['l']) and bool(groups['r']), not groups['l'] and not groups['r']
        n_neutral = len(groups['n']) if (both or none) else len(label_list)
    if save_label_list is not None:
        np.save(save_label_list, np.int32(label_list))
    return (np.int32(label_list), n_neutral) if FS_sort else (np.int32(label_list), None)


def load_array_if_path(var, load_as_numpy=True):
    if isinstance(var, str) and load_as_numpy:
        assert os.path.isfile(var), 'No such path: %s' % var
        var = np.load(var)
    return var


# ----------------------------------------------------------------- reformatting (reference :319-397) ----------------
def reformat_to_list(var, length=None, load_as_numpy=False, dtype=None):
    if var is None:
        return None
    var = load_array_if_path(var, load_as_numpy=load_as_numpy)
    if isinstance(var, (bool, np.bool_, str)) or isinstance(var, _NUM):
        var = [var]
    elif isinstance(var, tuple):
        var = list(var)
    elif isinstance(var, np.ndarray):
        var = [var[0]] if var.shape == (1,) else np.squeeze(var).tolist()
        if not isinstance(var, list):
            var = [var]
    if not isinstance(var, list):
        raise TypeError('var should be an int, float, tuple, list, numpy array, or path to numpy array')
    if length is not None:
        if len(var) == 1:
            var = var * length
        elif len(var) != length:
            raise ValueError('if var is a list/tuple/numpy array, it should be of length 1 or {0}, had {1}'.format(length, var))
    if dtype is not None:
        conv = {'int': int, 'float': float, 'bool': bool, 'str': str}
        if dtype not in conv:
            raise ValueError("dtype should be 'str', 'float', 'int', or 'bool'; had {}".format(dtype))
        var = [conv[dtype](v) for v in var]
    return var


def reformat_to_n_channels_array(var, n_dims=3, n_channels=1):
    if var is None:
        return [None] * n_channels
    if isinstance(var, str):
        var = np.load(var)
    if isinstance(var, (int, float, list, tuple)):
        var = np.tile(np.array(reformat_to_list(var, n_dims)), (n_channels, 1))
    elif isinstance(var, np.ndarray):
        if n_channels == 1:
            var = var.reshape((1, n_dims))
        elif np.squeeze(var).shape == (n_dims,):
            var = np.tile(var.reshape((1, n_dims)), (n_channels, 1))
        elif var.shape != (n_channels, n_dims):
            raise ValueError('if array, var should be {0} or {1}'.format((1, n_dims), (n_channels, n_dims)))
    else:
        raise TypeError('var should be int, float, list, tuple or ndarray')
    return np.round(var, 3)


# ----------------------------------------------------------------- paths (reference :403-546) ------------------------
def list_images_in_folder(path_dir, include_single_image=True, check_if_empty=True):
    base = os.path.basename(path_dir)
    if include_single_image and any(e in base for e in ('.nii.gz', '.nii', '.mgz', '.npz')):
        assert os.path.isfile(path_dir), 'file %s does not exist' % path_dir
        return [path_dir]
    if not os.path.isdir(path_dir):
        raise Exception('Folder does not exist: %s' % path_dir)
    out = sorted(glob.glob(os.path.join(path_dir, '*nii.gz')) + glob.glob(os.path.join(path_dir, '*nii')) +
                 glob.glob(os.path.join(path_dir, '*.mgz')) + glob.glob(os.path.join(path_dir, '*.npz')))
    if check_if_empty:
        assert len(out) > 0, 'no .nii, .nii.gz, .mgz or .npz image could be found in %s' % path_dir
    return out


def strip_extension(path):
    for e in ('.nii.gz', '.nii', '.mgz', '.npz'):
        path = path.replace(e, '')
    return path


def mkdir(path_dir):
    if path_dir and path_dir[-1] == '/':
        path_dir = path_dir[:-1]
    if path_dir and not os.path.isdir(path_dir):
        os.makedirs(path_dir, exist_ok=True)


# ----------------------------------------------------------------- shapes (reference :558-614, 928-944) -------------
def get_dims(shape, max_channels=10):
    if shape[-1] <= max_channels:
        return len(shape) - 1, shape[-1]
    return len(shape), 1


def get_resample_shape(patch_shape, factor, n_channels=None):
    factor = reformat_to_list(factor, length=len(patch_shape))
    shape = [math.ceil(patch_shape[i] * factor[i]) for i in range(len(patch_shape))]
    return shape + [n_channels] if n_channels is not None else shape


def add_axis(x, axis=0):
    for ax in reformat_to_list(axis):
        x = np.expand_dims(x, axis=ax)
    return x


def get_padding_margin(cropping, loss_cropping):
    if cropping is None or loss_cropping is None:
        return None
    cropping, loss_cropping = reformat_to_list(cropping), reformat_to_list(loss_cropping)
    n = max(len(cropping), len(loss_cropping))
    cropping, loss_cropping = reformat_to_list(cropping, length=n), reformat_to_list(loss_cropping, length=n)
    margin = [int((cropping[i] - loss_cropping[i]) / 2) for i in range(n)]
    return margin[0] if len(margin) == 1 else margin


def find_closest_number_divisible_by_m(n, m, answer_type='lower'):
    if n % m == 0:
        return n
    lower, higher = int(n / m) * m, (int(n / m) + 1) * m
    if answer_type == 'lower':
        return lower
    if answer_type == 'higher':
        return higher
    if answer_type == 'closer':
        return lower if (n - lower) < (higher - n) else higher
    raise Exception('answer_type should be lower, higher, or closer, had : %s' % answer_type)


# ----------------------------------------------------------------- misc (reference :821-925, 961-1049) --------------
def infer(x):
    try:
        return float(x)
    except ValueError:
        if x == 'False':
            return False
        if x == 'True':
            return True
        if not isinstance(x, str):
            raise TypeError('input should be an int/float/boolean/str, had {}'.format(type(x)))
        return x


class LoopInfo:
    def __init__(self, n_iterations, spacing=10, text='processing', print_time=False):
        self.n, self.spacing, self.text, self.print_time = n_iterations, spacing, text, print_time
        self.t0 = time.time()

    def update(self, idx):
        if idx == 0:
            print(self.text + ' 1/{}'.format(self.n))
        elif idx % self.spacing == self.spacing - 1:
            msg = self.text + ' {}/{}'.format(idx + 1, self.n)
            if self.print_time:
                eta = int((time.time() - self.t0) / (idx + 1) * (self.n - idx - 1))
                msg += '   remaining time: {}'.format(timedelta(seconds=eta))
            print(msg)


def get_mapping_lut(source, dest=None):
    source = np.array(reformat_to_list(source), dtype='int32')
    dest = np.arange(len(source), dtype='int32') if dest is None else np.array(reformat_to_list(dest, dtype='int'))
    assert len(source) == len(dest), 'label_list and new_label_list should have the same length'
    lut = np.zeros(np.max(source) + 1, dtype='int32')
    lut[source] = dest
    return lut


def build_training_generator(gen, batchsize):
    while True:
        yield next(gen), np.zeros((max(batchsize, 1), 1))


def draw_value_from_distribution(hyperparameter, size=1, distribution='uniform', centre=0., default_range=10.0,
                                 positive_only=False, return_as_tensor=False, batchsize=None):
    """NumPy branch of the reference sampler (:996-1016, 1038-1049).  The in-graph branch (return_as_tensor) has no
    meaning here: those draws are made by synthsr_b200.draws."""
    if hyperparameter is False:
        return None
    if return_as_tensor:
        raise NotImplementedError('in-graph sampling is done by synthsr_b200.draws.sample_draws')
    hyperparameter = load_array_if_path(hyperparameter, load_as_numpy=True)
    if not isinstance(hyperparameter, np.ndarray):
        if hyperparameter is None:
            hyperparameter = np.array([[centre - default_range] * size, [centre + default_range] * size])
        elif isinstance(hyperparameter, _NUM):
            hyperparameter = np.array([[centre - hyperparameter] * size, [centre + hyperparameter] * size])
        elif isinstance(hyperparameter, (list, tuple)):
            assert len(hyperparameter) == 2, 'if list, parameter_range should be of length 2.'
            hyperparameter = np.transpose(np.tile(np.array(hyperparameter), (size, 1)))
        else:
            raise ValueError('parameter_range should either be None, a number, a sequence, or a numpy array.')
    else:
        assert hyperparameter.shape[0] % 2 == 0, 'number of rows of parameter_range should be divisible by 2'
        idx = 2 * np.random.randint(int(hyperparameter.shape[0] / 2))
        hyperparameter = hyperparameter[idx: idx + 2, :]
    if distribution == 'uniform':
        value = np.random.uniform(low=hyperparameter[0, :], high=hyperparameter[1, :])
    elif distribution == 'normal':
        value = np.random.normal(loc=hyperparameter[0, :], scale=hyperparameter[1, :])
    else:
        raise ValueError("Distribution not supported, should be 'uniform' or 'normal'.")
    if positive_only:
        value[value < 0] = 0
    return value
