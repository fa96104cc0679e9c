"""`labels_to_image_model` of the reference (SynthSR/labels_to_image_model.py:32-266) as a B200 kernel pipeline.

Returns a `LabelsToImageModel` with the Keras-`Model` surface the reference's callers use: `.predict(list_inputs) ->
[image, target]` (NumPy in / NumPy out, batch first, channels last), `.inputs`, `.output[i].get_shape().as_list()`.
All in-graph randomness is drawn per call by synthsr_b200.draws (per-voxel GMM noise on the device)."""
import numpy as np

from synthsr_b200.generator import GeneratorPlan, get_shapes  # noqa: F401  (get_shapes re-exported like the reference)


class _Shape:
    def __init__(self, shape):
        self._s = list(shape)

    def as_list(self):
        return list(self._s)


class _Out:
    def __init__(self, shape):
        self.shape = _Shape(shape)

    def get_shape(self):
        return self.shape


class LabelsToImageModel:
    def __init__(self, plan, batchsize=1, seed=None):
        self.plan, self.batchsize = plan, batchsize
        self.inputs = ['labels_input', 'means_input', 'std_devs_input'] + (['real_image_input'] if plan.use_real_image else [])
        self.output = [_Out([None] + plan.image_shape), _Out([None] + plan.target_shape)]
        self.outputs = self.output
        self._gen = None
        self._rng = np.random.default_rng(seed)
        self._seed = int(self._rng.integers(1 << 31))

    @property
    def engine(self):
        if self._gen is None:
            from synthsr_b200.generator import SynthGenerator
            self._gen = SynthGenerator(self.plan, self.batchsize)
        return self._gen

    def predict(self, inputs):
        import torch
        from synthsr_b200.draws import sample_draws
        labels = np.asarray(inputs[0])
        B = labels.shape[0]
        if B != self.batchsize:
            self.batchsize, self._gen = B, None
        dev = self.engine.device                            # 'cuda': SynthGenerator refuses anything else
        lab_t = torch.as_tensor(np.ascontiguousarray(labels[..., 0], dtype=np.int32)).to(dev)
        real_t = None
        if self.plan.use_real_image:
            real_t = torch.as_tensor(np.ascontiguousarray(np.asarray(inputs[3])[..., 0], dtype=np.float32)).to(dev)
        draws = sample_draws(self._rng, self.plan, B)
        image, target = self.engine.run(lab_t, inputs[1], inputs[2], draws, real_image=real_t, seed=self._seed)
        return [image.cpu().numpy(), target.cpu().numpy()]


def labels_to_image_model(labels_shape, input_channels, output_channel, generation_labels, n_neutral_labels, atlas_res,
                          target_res, output_shape=None, output_div_by_n=None, padding_margin=None, flipping=True,
                          aff=None, scaling_bounds=0.15, rotation_bounds=15, shearing_bounds=0.012,
                          translation_bounds=False, nonlin_std=3., nonlin_shape_factor=.0625,
                          simulate_registration_error=True, randomise_res=False, data_res=None, thickness=None,
                          downsample=False, build_reliability_maps=False, blur_range=1.15, bias_field_std=.3,
                          bias_shape_factor=.025, batchsize=1):
    if flipping:
        assert aff is not None, 'aff should not be None if flipping is True'
    plan = GeneratorPlan(labels_shape, input_channels, output_channel, generation_labels, n_neutral_labels, atlas_res,
                         target_res, output_shape=output_shape, output_div_by_n=output_div_by_n,
                         padding_margin=padding_margin, flipping=flipping, aff=aff, scaling_bounds=scaling_bounds,
                         rotation_bounds=rotation_bounds, shearing_bounds=shearing_bounds,
                         translation_bounds=translation_bounds, nonlin_std=nonlin_std,
                         nonlin_shape_factor=nonlin_shape_factor,
                         simulate_registration_error=simulate_registration_error, randomise_res=randomise_res,
                         data_res=data_res, thickness=thickness, downsample=downsample,
                         build_reliability_maps=build_reliability_maps, blur_range=blur_range,
                         bias_field_std=bias_field_std, bias_shape_factor=bias_shape_factor)
    return LabelsToImageModel(plan, batchsize)
