"""world_size-2 gloo tests (CPU) of the data-parallel host logic: the single flat all-reduce of
[gradients | BN moving stats | loss] and the per-rank sharding of random streams."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from synthsr_b200.trainer import allreduce_step
    torch.manual_seed(rank)
    grads = torch.full((1000,), float(rank + 1))
    moving = [torch.full((8,), float(10 * (rank + 1))), torch.full((4,), float(rank))]
    loss = torch.tensor([float(rank + 1)], dtype=torch.float64)
    flat = torch.zeros(1000 + 12 + 1)
    mean_loss = allreduce_step(grads, moving, loss, flat, world)
    out.put((rank, grads[:3].tolist(), moving[0][0].item(), moving[1][0].item(), mean_loss.item()))
    dist.destroy_process_group()


def test_single_flat_allreduce_two_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, m0, m1, l in res:
        assert g == [3.0, 3.0, 3.0]            # gradients SUMMED (Adam applies 1/world)
        assert m0 == 15.0 and m1 == 0.5        # moving stats averaged -> identical replicas
        assert l == 1.5                        # mean loss
    assert res[0][1:] == res[1][1:]


def test_rank_streams_differ_weights_identical():
    """every rank draws its own augmentation stream but initialises identical weights."""
    from synthsr_b200.draws import sample_draws
    from synthsr_b200.generator import GeneratorPlan
    from synthsr_b200.synthetic import GEN_LABELS
    plan = GeneratorPlan([32, 32, 32], True, 0, GEN_LABELS, None, 1., None)
    seeds = [0 * 1000003 + 7919 * r for r in range(2)]          # TrainingEngine's per-rank rng seeding
    d = [sample_draws(np.random.default_rng(s), plan, 1) for s in seeds]
    assert not np.array_equal(d[0]['svf_normal'], d[1]['svf_normal'])
    assert not np.array_equal(d[0]['aff_rotation'], d[1]['aff_rotation'])
