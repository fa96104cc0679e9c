"""Time of the two steps of the adversarial fine-tuner at the benchmark size (160^3, reference topology, the reference's
discriminator: 4 levels, 32 filters, Dense(512) on 10^3 x 256 features = 131 M parameters).
    python scripts/adv_step_time.py [size] [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from synthsr_b200.adversary import AdversarialEngine, AdversarialUNet3D, Discriminator  # noqa: E402
from synthsr_b200.generator import GeneratorPlan  # noqa: E402
from synthsr_b200.trainer import TrainingEngine  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 160
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
maps, pm, ps, gl, gc = bench.make_inputs(size, 1, seed=0)
plan = GeneratorPlan([size] * 3, True, 0, gl, None, 1., None, **bench.TRAINING_DEFAULTS)
disc = Discriminator([size, size, size, 1], seed=1)
engine = TrainingEngine(plan, batchsize=1, seed=0, net_cls=AdversarialUNet3D,
                        net_kwargs=dict(disc=disc, discr_weight=.01))
adv = AdversarialEngine(engine, disc)
rng = np.random.default_rng(0)
labels = torch.from_numpy(maps[0][None]).cuda()
print('discriminator parameters: %.1f M; U-Net parameters: %.1f M' % (disc.n_params / 1e6, engine.net.n_params / 1e6))


def timed(fn, n):
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n):
        m, sd = bench.draw_gmm(rng, pm, ps, gc)
        out = fn(labels, m, sd)
    torch.cuda.synchronize()
    return (time.time() - t0) / n * 1e3, float(out.item())


timed(adv.discriminator_step, 2)
timed(adv.generator_step, 2)
d_ms, d_loss = timed(adv.discriminator_step, steps)
g_ms, g_loss = timed(adv.generator_step, steps)
print('%d^3: discriminator step %.1f ms (loss %.4f), generator step %.1f ms (loss %.4f); peak memory %.1f GB' % (
    size, d_ms, d_loss, g_ms, g_loss, torch.cuda.max_memory_allocated() / 2 ** 30))
print('one step of the reference loop (training_ratio 10): %.1f ms' % (10 * d_ms + g_ms))
