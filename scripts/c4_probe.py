"""BASELINE configs[3]: Hyperfine dual-input (T1+T2) 192x192x64, batch 1 per GPU: generator (registration error, 1.5x1.5x5 mm
acquisition, downsample) + 2-channel U-Net step with the residual on channel 0.  Prints ms/step.  GPU box only."""
import os, sys
os.environ.setdefault('OMP_NUM_THREADS', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synthsr_b200.generator import GeneratorPlan
from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors
from synthsr_b200.trainer import TrainingEngine

shape = [192, 192, 64]
res = np.array([[1.5, 1.5, 5.], [1.5, 1.5, 5.]])
plan = GeneratorPlan(shape, [False, True, True], 0, GEN_LABELS, None, 1., None, output_div_by_n=32, scaling_bounds=.15,
                     rotation_bounds=15, shearing_bounds=.02, translation_bounds=5, nonlin_std=4., nonlin_shape_factor=.03125,
                     bias_field_std=.3, bias_shape_factor=.03125, blur_range=1.15, build_reliability_maps=False,
                     data_res=res, thickness=res, downsample=True, simulate_registration_error=True)
eng = TrainingEngine(plan, batchsize=1, conv_impl='tc', seed=0, work_with_residual_channel=[0])
pm, ps = synthetic_priors(int(GEN_CLASSES.max()) + 1, 3, 0)
lab = torch.from_numpy(phantom_labels(shape, GEN_LABELS, seed=0)[None].astype(np.int32)).cuda()
rng = np.random.default_rng(0)
def gmm():
    m = np.stack([np.clip(rng.normal(pm[2 * c], pm[2 * c + 1]), 0, None)[GEN_CLASSES] for c in range(3)], -1)[None]
    s = np.stack([np.clip(rng.normal(ps[2 * c], ps[2 * c + 1]), 0, None)[GEN_CLASSES] for c in range(3)], -1)[None]
    return m.astype(np.float32), s.astype(np.float32)
for i in range(4):
    l = eng.train_step_pipelined(lab, *gmm())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for i in range(n):
    l = eng.train_step_pipelined(lab, *gmm())
e1.record()
torch.cuda.synchronize()
print('c4 192x192x64 Cin=%d: %.2f ms/step = %.1f volumes/s, loss %.4f, peak mem %.2f GB' % (
    plan.n_image_channels, e0.elapsed_time(e1) / n, n / (e0.elapsed_time(e1) / 1e3), l.item(), torch.cuda.max_memory_allocated() / 2**30))
