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
    from synthsr_b200.trainer import exchange_inplace
    torch.manual_seed(rank)
    n, split = 1000, 640
    comm = torch.zeros(n + 12 + 1)                       # [gradients | BN moving stats | loss]
    comm[:n] = float(rank + 1)
    comm[n:n + 8] = float(10 * (rank + 1))
    comm[n + 8:n + 12] = float(rank)
    comm[-1] = float(rank + 1)
    exchange_inplace(comm, n, split, world, 0)           # prefix first (overlaps the backward pass on the GPUs) ...
    mid = (comm[0].item(), comm[split].item())
    exchange_inplace(comm, n, split, world, 1)           # ... then the rest
    out.put((rank, comm[:3].tolist() + comm[n - 3:n].tolist(), comm[n].item(), comm[n + 8].item(), comm[-1].item(), mid))
    dist.destroy_process_group()


def _engine_worker(rank, world, port, out):
    """GradientExchange driven the way TrainingEngine drives it, on a stand-in network (CPU tensors, gloo)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from synthsr_b200.trainer import GradientExchange

    class Net:
        L, n_params, level_end, grads_ready_hook, _side, device = 5, 100, {1: 90, 2: 70, 3: 50, 4: 30}, None, None, 'cpu'
        comm = torch.zeros(100 + 4 + 1)
    net = Net()
    ex = GradientExchange(net, world)
    assert ex.split == 70 and net.grads_ready_hook is not None
    res = []
    for step in range(2):
        net.comm[:100] = float(rank + 1 + step)
        net.comm[100:104] = float(4 * rank)
        for level in (4, 3, 2, 1):                       # the backward pass reports the encoder levels deep to shallow
            net.grads_ready_hook(level)
        loss = ex.finish(torch.tensor([float(rank)], dtype=torch.float64))
        res.append((net.comm[0].item(), net.comm[99].item(), net.comm[100].item(), loss.item()))
    out.put((rank, res))
    dist.destroy_process_group()


def _run(worker, world=2):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + id(worker)) % 2000
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_inplace_bucketed_allreduce_two_ranks():
    res = _run(_worker)
    for rank, g, m0, m1, l, mid in res:
        assert g == [3.0] * 6                  # gradients SUMMED (Adam applies 1/world), both buckets
        assert m0 == 15.0 and m1 == 0.5        # moving stats averaged -> identical replicas
        assert l == 1.5                        # mean loss
        assert mid == (3.0, float(rank + 1))   # after stage 0 only the prefix has been exchanged
    assert res[0][1:5] == res[1][1:5]


def test_gradient_exchange_hook_protocol_two_ranks():
    res = _run(_engine_worker)
    assert res[0][1] == res[1][1]
    assert res[0][1][0] == (3.0, 3.0, 2.0, 0.5) and res[0][1][1] == (5.0, 5.0, 2.0, 0.5)


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
