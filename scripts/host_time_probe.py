"""host enqueue time per training step vs GPU time per step (is the step launch-bound?).  GPU box only."""
import os, sys, time
os.environ.setdefault('OMP_NUM_THREADS', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from synthsr_b200.generator import GeneratorPlan
from synthsr_b200.trainer import TrainingEngine
from synthsr_b200._lib import lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 160
maps, pm, ps, gl, gc = bench.make_inputs(size, 2, seed=0)
plan = GeneratorPlan([size] * 3, True, 0, gl, None, 1., None, **bench.TRAINING_DEFAULTS)
eng = TrainingEngine(plan, batchsize=1, conv_impl='tc', seed=0)
dev = [torch.from_numpy(m[None]).cuda() for m in maps]
rng = np.random.default_rng(0)
for i in range(5):
    eng.train_step(dev[i % 2], *bench.draw_gmm(rng, pm, ps, gc))
torch.cuda.synchronize()
n = 20
l0 = lib.ssr_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(n):
    eng.train_step(dev[i % 2], *bench.draw_gmm(rng, pm, ps, gc))
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('size %d: host enqueue %.2f ms/step, gpu %.2f ms/step, wall %.2f ms/step, %d launches/step' % (
    size, (t1 - t0) / n * 1e3, e0.elapsed_time(e1) / n, (t2 - t0) / n * 1e3, (lib.ssr_launch_count() - l0) // n))
# single-step latency with an idle queue: sync before every step (exposes host overhead if the GPU would otherwise wait)
ts = []
for i in range(10):
    torch.cuda.synchronize()
    a = time.perf_counter()
    eng.train_step(dev[i % 2], *bench.draw_gmm(rng, pm, ps, gc))
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - a)
print('size %d: synchronous step %.2f ms (median)' % (size, np.median(ts) * 1e3))
