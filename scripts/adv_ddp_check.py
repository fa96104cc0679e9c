"""Two-rank run of SynthSR.fine_tuning_with_adversary.training() on tiny label maps: both replicas must end with identical U-Net
and discriminator parameters (gradients of both networks are averaged over the ranks every step).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 scripts/adv_ddp_check.py"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ext.lab2im import utils  # noqa: E402
from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors  # noqa: E402
import SynthSR.fine_tuning_with_adversary as FT  # noqa: E402

rank = int(os.environ.get('RANK', 0))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
dist.init_process_group('nccl')
root = tempfile.mkdtemp(prefix='advddp%d_' % rank)
labels_dir = os.path.join(root, 'labels')
os.makedirs(labels_dir)
aff = np.eye(4)
for i in range(2):
    utils.save_volume(phantom_labels([44, 52, 40], seed=i).astype(np.float32), aff, None,
                      os.path.join(labels_dir, 'brain%d_labels.nii.gz' % i))
pm, ps = synthetic_priors(14, 1, seed=0)
p = {k: os.path.join(root, k + '.npy') for k in ('labels', 'classes', 'means', 'stds')}
np.save(p['labels'], GEN_LABELS); np.save(p['classes'], GEN_CLASSES); np.save(p['means'], pm); np.save(p['stds'], ps)
engine, disc = FT.training(labels_dir, None, os.path.join(root, 'model'), p['means'], p['stds'], p['labels'],
                           path_generation_classes=p['classes'], output_channel=0, output_shape=32, n_levels=3,
                           unet_feat_count=8, epochs=1, steps_per_epoch=2, first_training_ratio=2, training_ratio=1,
                           relative_weight_discriminator=.05, randomise_res=False, data_res=np.array([1., 1., 2.]))
torch.cuda.synchronize()
for name, t in (('U-Net', engine.net.params), ('discriminator', disc.params)):
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    if rank == 0:
        print('%s replicas identical: %s (|params| %.4f)' % (name, same, float(t.double().norm())))
    assert same, name
dist.barrier()
dist.destroy_process_group()
