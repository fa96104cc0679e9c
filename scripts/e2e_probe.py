"""Where does the end-to-end (host buffers) step lose time against the device-resident step?"""
import os, sys, time
os.environ.setdefault('OMP_NUM_THREADS', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from synthsr_b200.generator import GeneratorPlan
from synthsr_b200.trainer import TrainingEngine

size = 160
maps, pm, ps, gl, gc = bench.make_inputs(size, 2, seed=0)
plan = GeneratorPlan([size] * 3, True, 0, gl, None, 1., None, **bench.TRAINING_DEFAULTS)
eng = TrainingEngine(plan, batchsize=1, conv_impl='tc', seed=0)
dev = [torch.from_numpy(m[None]).cuda() for m in maps]
pin = [torch.from_numpy(m[None]).pin_memory() for m in maps]
rng = np.random.default_rng(0)
hl = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]
ev = [torch.cuda.Event() for _ in range(2)]

def loop(n, h2d, d2h):
    thost = 0.
    for i in range(n):
        t0 = time.perf_counter()
        m, s = bench.draw_gmm(rng, pm, ps, gc)
        lab = pin[i % 2].cuda(non_blocking=True) if h2d else dev[i % 2]
        loss = eng.train_step(lab, m, s)
        if d2h:
            hl[i % 2].copy_(loss, non_blocking=True); ev[i % 2].record()
        thost += time.perf_counter() - t0
        if d2h and i > 0:
            ev[(i - 1) % 2].synchronize(); float(hl[(i - 1) % 2][0])
    if d2h:
        ev[(n - 1) % 2].synchronize()
    return thost / n * 1e3

def timed(n, h2d, d2h):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); th = loop(n, h2d, d2h); b.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    return a.elapsed_time(b) / n, wall, th

loop(3, False, False)
for name, h, d in (('device', False, False), ('h2d only', True, False), ('d2h only', False, True), ('h2d+d2h', True, True),
                   ('device', False, False)):
    ms, wall, th = timed(10, h, d)
    print(f'{name:10s} gpu-event {ms:7.2f} ms/step   wall {wall:7.2f}   host enqueue {th:7.2f}')
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for i in range(10): x = pin[i % 2].cuda(non_blocking=True)
b.record(); torch.cuda.synchronize(); print('H2D 16.4 MB: %.3f ms' % (a.elapsed_time(b) / 10))
# host-only cost of the pieces
t0 = time.perf_counter()
for i in range(10): eng.train_step(dev[0], *bench.draw_gmm(rng, pm, ps, gc))
t1 = time.perf_counter(); torch.cuda.synchronize()
print('enqueue-only %.2f ms/step' % ((t1 - t0) / 10 * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(5): eng.train_step(dev[0], *bench.draw_gmm(rng, pm, ps, gc))
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
