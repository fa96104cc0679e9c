"""Generator only (labels_to_image_model on the GPU) at a given size: CUDA-event time per call and per kernel class, for
ncu captures of the generator kernels and the c1 configuration.   python scripts/gen_only.py --size 160 --iters 20
   --defaults training  : SynthSR/training.py:57-73 hyper-parameters (the 160^3 headline generator)
   --defaults brain     : SynthSR/brain_generator.py:30-61 defaults (BASELINE configs[0], 64^3)"""
import argparse
import os
import sys

os.environ.setdefault('OMP_NUM_THREADS', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from synthsr_b200.draws import sample_draws  # noqa: E402
from synthsr_b200.generator import GeneratorPlan, SynthGenerator  # noqa: E402
from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors  # noqa: E402

TRAINING = dict(scaling_bounds=0.15, rotation_bounds=15, shearing_bounds=0.02, translation_bounds=5, nonlin_std=4.,
                nonlin_shape_factor=0.03125, bias_field_std=.3, bias_shape_factor=0.03125, blur_range=1.15,
                build_reliability_maps=False, output_div_by_n=32)
BRAIN = dict()      # GeneratorPlan's own defaults are BrainGenerator's


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=160)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--defaults', default='training', choices=['training', 'brain'])
    args = ap.parse_args()
    n = args.size
    cfg = TRAINING if args.defaults == 'training' else BRAIN
    plan = GeneratorPlan([n] * 3, True, 0, GEN_LABELS, None, 1., None, **cfg)
    gen = SynthGenerator(plan, 1)
    lab = torch.from_numpy(phantom_labels([n] * 3, GEN_LABELS, seed=0)[None].astype(np.int32)).cuda()
    pm, ps = synthetic_priors(int(GEN_CLASSES.max()) + 1, 1, 0)
    rng = np.random.default_rng(0)

    def once():
        m = np.clip(rng.normal(pm[0], pm[1]), 0, None)[GEN_CLASSES][None, :, None].astype(np.float32)
        s = np.clip(rng.normal(ps[0], ps[1]), 0, None)[GEN_CLASSES][None, :, None].astype(np.float32)
        return gen.run(lab, m, s, sample_draws(rng, plan, 1))

    for _ in range(args.warmup):
        once()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
    for a, b in evs:
        a.record()
        once()
        b.record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in evs])
    vox = float(n) ** 3
    print('generator %d^3 (%s defaults): median %.3f ms (p10 %.3f p90 %.3f) per volume; %.1f MB compulsory (20 B/voxel) -> '
          '%.1f GB/s at the median' % (n, args.defaults, np.median(t), np.percentile(t, 10), np.percentile(t, 90),
                                       20 * vox / 1e6, 20 * vox / (np.median(t) * 1e-3) / 1e9))


if __name__ == '__main__':
    main()
