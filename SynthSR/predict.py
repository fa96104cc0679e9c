"""Inference path of the reference's command-line tools on the B200 engine:

    predict()            scripts/predict_command_line.py:63-139           single scan  -> 1 mm MP-RAGE
    predict_hyperfine()  scripts/predict_command_line_hyperfine.py:60-139 T1 + T2 pair -> 1 mm MP-RAGE (residual on T1)

Host-side pre/post-processing (resample to 1 mm, re-orient to RAS, normalise, pad to a multiple of 32, crop back, rescale)
follows the scripts line by line; the U-Net forward (inference-mode BatchNorm, flip test-time augmentation) runs on the
tcgen05 kernels through `ext.neuron.models.unet(input_shape=[None, None, None, C])`.  The reference scripts themselves
import TensorFlow for thread settings and cannot run here; scripts/predict_command_line*.py are thin CLIs over these
functions with the same arguments."""
import os

import numpy as np

from ext.lab2im import edit_volumes, utils
from ext.neuron import models as nrn_models

_HOME = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_unet(n_channels=1, model_file=None, conv_impl='tc3'):
    """the U-Net of the shipped models (predict_command_line.py:66-77) with weights from a Keras .h5 file."""
    unet_model = nrn_models.unet(nb_features=24, input_shape=[None, None, None, n_channels], nb_levels=5, conv_size=3,
                                 nb_labels=1, feat_mult=2, nb_conv_per_level=2, conv_dropout=0,
                                 final_pred_activation='linear', batch_norm=-1, activation='elu', input_model=None,
                                 conv_impl=conv_impl)
    if model_file is not None:
        if not os.path.isfile(model_file):
            raise FileNotFoundError('model weights not found: %s -- pass --model / model=<path to a Keras .h5 file> (the '
                                    "reference ships models/SynthSR_v10_210712*.h5; this repository does not)" % model_file)
        unet_model.load_weights(model_file, by_name=True)
    return unet_model


def _pad_to_32(I):
    """zero-pad [1,X,Y,Z,C] to multiples of 32, centred (predict_command_line.py:121-124) -> padded volume, offsets."""
    W = (np.ceil(np.array(I.shape[1:-1]) / 32.0) * 32).astype('int')
    idx = np.floor((W - I.shape[1:-1]) / 2).astype('int')
    S = np.zeros([1, *W, I.shape[-1]])
    S[0, idx[0]:idx[0] + I.shape[1], idx[1]:idx[1] + I.shape[2], idx[2]:idx[2] + I.shape[3], :] = I
    return S, idx


def _io_lists(path_images, path_predictions):
    """single file or folder (predict_command_line.py:85-106), same naming of the outputs."""
    path_images = os.path.abspath(path_images)
    basename = os.path.basename(path_images)
    path_predictions = os.path.abspath(path_predictions)
    if ('.nii.gz' not in basename) & ('.nii' not in basename) & ('.mgz' not in basename) & ('.npz' not in basename):
        if os.path.isfile(path_images):
            raise Exception('extension not supported for %s, only use: nii.gz, .nii, .mgz, or .npz' % path_images)
        images = utils.list_images_in_folder(path_images)
        utils.mkdir(path_predictions)
        preds = [os.path.join(path_predictions, os.path.basename(image)).replace('.nii', '_SynthSR.nii') for image in images]
        preds = [p.replace('.mgz', '_SynthSR.mgz') for p in preds]
        preds = [p.replace('.npz', '_SynthSR.npz') for p in preds]
    else:
        assert os.path.isfile(path_images), "files does not exist: %s " \
                                            "\nplease make sure the path and the extension are correct" % path_images
        images, preds = [path_images], [path_predictions]
    return images, preds


def preprocess(im, aff, ct=False):
    """-> (network input S [1,X',Y',Z',1], padding offsets, unpadded shape, output affine); :110-124."""
    im = np.array(im, dtype=np.float64)
    if ct:
        im[im < 0] = 0
        im[im > 80] = 80
    im, aff = edit_volumes.resample_volume(im, aff, [1.0, 1.0, 1.0])
    im, aff2 = edit_volumes.align_volume_to_ref(im, aff, aff_ref=np.eye(4), return_aff=True, n_dims=3)
    im = im - np.min(im)
    im = im / np.max(im)
    I = im[np.newaxis, ..., np.newaxis]
    S, idx = _pad_to_32(I)
    return S, idx, I.shape, aff2


def postprocess(output, idx, shape):
    """:131-135."""
    pred = np.squeeze(output)
    pred = 255 * pred
    pred[pred < 0] = 0
    pred[pred > 128] = 128
    return pred[idx[0]:idx[0] + shape[1], idx[1]:idx[1] + shape[2], idx[2]:idx[2] + shape[3]]


def predict_volume(unet_model, im, aff, ct=False, disable_flipping=False):
    S, idx, shape, aff2 = preprocess(im, aff, ct)
    if disable_flipping:
        output = unet_model.predict(S)
    else:                                                 # left-right flip test-time augmentation (:126-129)
        output = 0.5 * unet_model.predict(S) + 0.5 * np.flip(unet_model.predict(np.flip(S, axis=1)), axis=1)
    return postprocess(output, idx, shape), aff2


def predict(path_images, path_predictions, model=None, ct=False, disable_flipping=False, conv_impl='tc3'):
    """super-resolve / synthesise 1 mm MP-RAGEs from the scans in `path_images` (file or folder)."""
    unet_model = build_unet(1, model if model is not None else os.path.join(_HOME, 'models/SynthSR_v10_210712.h5'), conv_impl)
    images, preds = _io_lists(path_images, path_predictions)
    print('Found %d images' % len(images))
    for n, (path_image, path_prediction) in enumerate(zip(images, preds)):
        print('  Working on image %d ' % (n + 1))
        print('  ' + path_image)
        im, aff, hdr = utils.load_volume(path_image, im_only=False, dtype='float')
        pred, aff2 = predict_volume(unet_model, im, aff, ct, disable_flipping)
        utils.save_volume(pred, aff2, None, path_prediction)
    return preds


def predict_hyperfine(path_t1_images, path_t2_images, path_predictions, model=None, conv_impl='tc3'):
    """T1 + T2 Hyperfine pairs (1.5 x 1.5 x 5 mm) -> 1 mm MP-RAGE; the network predicts a residual on the T1 channel
    (predict_command_line_hyperfine.py:108-133, including its intensity scalings)."""
    unet_model = build_unet(2, model if model is not None else os.path.join(_HOME, 'models/SynthSR_v10_210712_hyperfine.h5'),
                            conv_impl)
    t1s, preds = _io_lists(path_t1_images, path_predictions)
    t2s, _ = _io_lists(path_t2_images, path_predictions)
    print('Found %d images' % len(t1s))
    for n, (p1, p2, path_prediction) in enumerate(zip(t1s, t2s, preds)):
        print('  Working on image %d ' % (n + 1))
        print('  ' + p1 + ', ' + p2)
        im1, aff1, _ = utils.load_volume(p1, im_only=False, dtype='float')
        im1, aff1 = edit_volumes.resample_volume(im1, aff1, [1.0, 1.0, 1.0])
        im1, aff1_mod = edit_volumes.align_volume_to_ref(im1, aff1, aff_ref=np.eye(4), return_aff=True, n_dims=3)
        im2, aff2, _ = utils.load_volume(p2, im_only=False, dtype='float')
        im2 = edit_volumes.resample_volume_like(im1, aff1_mod, im2, aff2)
        minimum = np.min(im1)
        im1 = im1 - minimum
        spread = np.max(im1) / 3.0
        im1 = im1 / spread
        im2 = im2 - np.min(im2)
        im2 = im2 / np.max(im2) * 2.0
        I = np.stack([im1, im2], axis=-1)[np.newaxis, ...]
        S, idx = _pad_to_32(I)
        output = unet_model.predict(S)
        res = np.squeeze(output)[idx[0]:idx[0] + I.shape[1], idx[1]:idx[1] + I.shape[2], idx[2]:idx[2] + I.shape[3]]
        pred = minimum + spread * (res + im1)
        pred[pred < 0] = 0
        utils.save_volume(pred, aff1_mod, None, path_prediction)
    return preds
