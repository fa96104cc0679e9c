"""per-convolution timing of one training step (CUDA events around every conv launch).  GPU box only."""
import os, sys
os.environ.setdefault('OMP_NUM_THREADS', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from synthsr_b200.generator import GeneratorPlan
from synthsr_b200.trainer import TrainingEngine

size = int(sys.argv[1]) if len(sys.argv) > 1 else 160
maps, pm, ps, gl, gc = bench.make_inputs(size, 2, seed=0)
plan = GeneratorPlan([size] * 3, True, 0, gl, None, 1., None, **bench.TRAINING_DEFAULTS)
eng = TrainingEngine(plan, batchsize=1, conv_impl=os.environ.get('SSR_CONV_IMPL', 'tc3'), seed=0)
dev = [torch.from_numpy(m[None]).cuda() for m in maps]
rng = np.random.default_rng(0)
for i in range(3):
    eng.train_step(dev[i % 2], *bench.draw_gmm(rng, pm, ps, gc))
eng.net.prof = []
for i in range(3):
    eng.train_step(dev[i % 2], *bench.draw_gmm(rng, pm, ps, gc))
torch.cuda.synchronize()
n = len(eng.net.prof) // 3
rows = {}
for rep in range(3):
    for j in range(n):
        kind, fl, a, b = eng.net.prof[rep * n + j][:4]
        rows.setdefault(j, [kind, fl, 0.])[2] += a.elapsed_time(b) / 3
tot = {}
for j in range(n):
    kind, fl, ms = rows[j]
    nv = fl / 54.
    print('%3d %-10s cin*cout*nvox=%.3e  %7.3f ms  %6.1f TFLOP/s' % (j, kind, nv, ms, fl / ms / 1e9))
    t = tot.setdefault(kind, [0., 0.]); t[0] += ms; t[1] += fl
for k, (ms, fl) in sorted(tot.items()):
    print('%-10s %7.3f ms  %6.1f TFLOP/s' % (k, ms, fl / ms / 1e9))
