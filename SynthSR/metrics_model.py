"""Loss wiring of the reference (SynthSR/metrics_model.py:29-229).  In the reference the loss is a set of Keras layers
appended to the model; here `metrics_model` records the same options (metric, residual channels, centre cropping)
on the model object and the fused head+loss kernel (ssr_head_loss) evaluates them."""


class IdentityLoss(object):
    """the reference trains on `loss = model output` (metrics_model.py:218-229); kept for API compatibility."""

    def __init__(self, keepdims=True):
        self.keepdims = keepdims

    def loss(self, y_true, y_predicted):
        return y_predicted


class LossModel:
    def __init__(self, input_model, loss_cropping, metrics, work_with_residual_channel):
        self.unet = input_model
        self.loss_cropping = loss_cropping
        self.metrics = metrics
        self.work_with_residual_channel = work_with_residual_channel
        self.inputs = getattr(input_model, 'inputs', None)


def metrics_model(input_model, loss_cropping=16, metrics='l1', work_with_residual_channel=None):
    if metrics not in ('l1', 'l2'):
        if metrics in ('ssim', 'laplace'):
            raise NotImplementedError("regression_metric '%s' is an optional loss outside this build's scope "
                                      "(SURVEY.md 2a #5); use 'l1' or 'l2'" % metrics)
        raise Exception('metrics should either be "l1" or "l2" or "ssim" oro "laplace", got {}'.format(metrics))
    return LossModel(input_model, loss_cropping, metrics, work_with_residual_channel)


def add_seg_loss_to_model(input_model, *args, **kwargs):
    """segmentation-regularised loss (frozen second U-Net + soft Dice, metrics_model.py:136-215).  The reference appends Keras
    layers to `input_model`; on this engine the regulariser is part of the training step itself --
    synthsr_b200.seg_loss.SegRegulariser, attached by SynthSR.training.training(segmentation_model_file=...) (validated
    against the float64 oracle on a B200: tests/test_seg_loss_gpu.py).  Calling this function directly on a loss model has
    nothing to append to."""
    raise NotImplementedError('add_seg_loss_to_model: pass segmentation_model_file / segmentation_label_list / '
                              'segmentation_label_equivalency to SynthSR.training.training(); the regulariser runs inside '
                              'the training step (synthsr_b200/seg_loss.py)')
