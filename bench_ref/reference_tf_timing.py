"""Timing of the unmodified reference on CPU (run by bench_ref/run_reference_tf.sh where TensorFlow 2.x + Keras 2.3.1 exist).
Only the reference's public API is used: BrainGenerator(...).generate_brain() and training(); label maps are the same smooth
phantoms bench.py uses (synthsr_b200.synthetic), written as .npz so that no NIfTI writer is involved.  Prints one JSON line."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ref, threads = sys.argv[1], int(sys.argv[2])
repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, repo)
from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors  # noqa: E402  (no TF, no SynthSR)
sys.path.remove(repo)
for m in [m for m in sys.modules if m == 'SynthSR' or m.startswith(('SynthSR.', 'ext.')) or m == 'ext']:
    del sys.modules[m]
sys.path.insert(0, ref)                                   # from here on `SynthSR` / `ext` are the reference's
import tensorflow as tf  # noqa: E402

tf.config.threading.set_intra_op_parallelism_threads(threads)
from SynthSR.brain_generator import BrainGenerator  # noqa: E402
from SynthSR.training import training  # noqa: E402

out = {'impl': 'reference (TF %s, CPU)' % tf.__version__, 'cores': threads}
tmp = tempfile.mkdtemp()
np.save(os.path.join(tmp, 'labels.npy'), GEN_LABELS)
np.save(os.path.join(tmp, 'classes.npy'), GEN_CLASSES)
pm, ps = synthetic_priors(14, 1, seed=0)
np.save(os.path.join(tmp, 'pm.npy'), pm)
np.save(os.path.join(tmp, 'ps.npy'), ps)
for size in (64, 160):
    d = os.path.join(tmp, 'labels%d' % size)
    os.makedirs(d)
    for i in range(2):
        np.savez(os.path.join(d, 'map%d.npz' % i), vol_data=phantom_labels([size] * 3, seed=i).astype(np.int32))
    gen = BrainGenerator(d, os.path.join(tmp, 'pm.npy'), os.path.join(tmp, 'ps.npy'), 'normal', os.path.join(tmp, 'labels.npy'),
                         generation_classes=os.path.join(tmp, 'classes.npy'))
    gen.generate_brain()                                   # graph build + first call
    t = time.time()
    n = 10 if size == 64 else 3
    for _ in range(n):
        gen.generate_brain()
    out['generate_brain_%d_s_per_volume' % size] = (time.time() - t) / n
t = time.time()
training(os.path.join(tmp, 'labels160'), os.path.join(tmp, 'models'), os.path.join(tmp, 'pm.npy'), os.path.join(tmp, 'ps.npy'),
         os.path.join(tmp, 'labels.npy'), path_generation_classes=os.path.join(tmp, 'classes.npy'), batchsize=1,
         randomise_res=False, build_reliability_maps=False, epochs=1, steps_per_epoch=5)
out['training_160_5_steps_s_incl_graph_build'] = time.time() - t
print(json.dumps(out))
