#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): training volumes/sec on 160^3 single-channel volumes, 5-level 24-feature U-Net,
generator + U-Net + Adam every step, at N GPUs of one node (data parallel, one process per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # CPU arm: the oracle restatement of the reference on the host cores

Prints ONE JSON line (rank 0).  A "step" is one pass of the hot path (generate a synthetic scan pair from a label map,
U-Net forward/backward, L1 loss, Adam) over one mini-batch of 1 volume per GPU.
"""
import argparse
import json
import os
import sys
import threading
import time

if '--impl' not in sys.argv or 'reference' not in sys.argv:
    # GPU arm: the host only samples a few hundred random numbers and enqueues kernels; BLAS/OpenMP thread pools of a
    # 100+-core host make those tiny ops slower, not faster (torchrun sets the same for N > 1)
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('MKL_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 160
TRAINING_DEFAULTS = dict(scaling_bounds=0.15, rotation_bounds=15, shearing_bounds=0.02, translation_bounds=5,
                         nonlin_std=4., nonlin_shape_factor=0.03125, bias_field_std=.3, bias_shape_factor=0.03125,
                         blur_range=1.15, build_reliability_maps=False, output_div_by_n=32)   # SynthSR/training.py:57-73


def conv_flops_per_step(size, cin=1):
    """algorithmic conv FLOPs of one training step: fwd + dgrad + wgrad."""
    from synthsr_b200.unet import layer_specs
    v = float(size) ** 3
    fwd, first = 0., None
    lvl = {}
    for name, kind, ci, co in layer_specs(cin):
        if kind == 'bn':
            continue
        if 'downarm' in name:
            l = int(name.split('_')[3])
        elif 'uparm' in name:
            l = 8 - int(name.split('_')[3])
        else:
            l = 0
        k3 = 27 if kind == 'conv' else 1
        f = 2. * k3 * ci * co * v / 8 ** l
        fwd += f
        if first is None:
            first = f
    return fwd, 3 * fwd - first


def make_inputs(size, n_maps=2, seed=0):
    from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors
    maps = []
    for i in range(n_maps):
        lo = phantom_labels([size // 2] * 3, GEN_LABELS, seed=seed + i)
        maps.append(np.ascontiguousarray(np.repeat(np.repeat(np.repeat(lo, 2, 0), 2, 1), 2, 2)[:size, :size, :size]))
    pm, ps = synthetic_priors(int(GEN_CLASSES.max()) + 1, 1, seed)
    return maps, pm, ps, GEN_LABELS, GEN_CLASSES


def draw_gmm(rng, pm, ps, classes):
    """SynthSR/model_inputs.py:118-123 ('normal' priors, negatives clipped)."""
    m = np.clip(rng.normal(pm[0], pm[1]), 0, None)[classes]
    s = np.clip(rng.normal(ps[0], ps[1]), 0, None)[classes]
    return m[None, :, None].astype(np.float32), s[None, :, None].astype(np.float32)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
                 nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake_slowdown'}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': []}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def cpu_reference_steps(size, steps, warmup, threads):
    """the reference's algorithm on the host cores (oracle restatement; TensorFlow itself is not installable here):
    generator (NumPy) + U-Net fwd/bwd + Adam (torch CPU fp32).  Returns seconds per step for a size^3 volume."""
    import torch
    from oracle import generator as OG
    from oracle import unet as OU
    from synthsr_b200.draws import sample_draws
    from synthsr_b200.generator import GeneratorPlan
    torch.set_num_threads(threads)
    maps, pm, ps, gl, gc = make_inputs(size, 1)
    cfg = dict(TRAINING_DEFAULTS)
    plan = GeneratorPlan([size] * 3, True, 0, gl, None, 1., None, **cfg)
    rng = np.random.default_rng(0)
    params = OU.init_params(0, 1)
    opt = OU.adam_init(params)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        draws = sample_draws(rng, plan, 1, gmm_noise=True)
        m, s = draw_gmm(rng, pm, ps, gc)
        image, target = OG.labels_to_image(dict(cfg, generation_labels=gl), [maps[0][None, ..., None], m, s], draws)
        OU.train_step(params, opt, torch.from_numpy(image), torch.from_numpy(target), lr=1e-4)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=SIZE)
    ap.add_argument('--conv-impl', default='tc3', choices=['tc3', 'tc', 'ref'],
                    help="tc3 (default): forward compensated to fp32-class accuracy (the parity-gated mode); tc: plain TF32 "
                         "(fast, outside the 1e-3 bar); ref: exact fp32 CUDA-core convolutions")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-pipeline', action='store_true',
                    help='generate and train on the same batch inside one call (no generator / training overlap)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    fwd_f, step_f = conv_flops_per_step(args.size)
    metric = 'training volumes/sec (160^3, 24-ch 5-level U-Net, generator+U-Net+Adam step)'
    config = {'workload': '%d^3 single-channel label map -> generator (training() defaults) -> 5-level 24-feature '
                          'U-Net fwd/bwd, L1, Adam; batch 1 per GPU (BASELINE configs[1]/[2])' % args.size,
              'global_batch': args.gpus, 'volume': [args.size] * 3, 'parallelism': 'dp%d' % args.gpus,
              'l2_policy': 'per-step working set (>4 GB of activations) exceeds the 126 MB L2; no explicit flush',
              'pipeline': 'generator of batch i+1 overlaps the U-Net step of batch i (one generator pass + one training '
                          'pass per step, as the reference\'s fit_generator queue)' if '--no-pipeline' not in sys.argv
                          else 'none (generate, then train, inside each step)'}

    if args.impl == 'reference':
        if rank != 0:
            return
        threads = min(len(os.sched_getaffinity(0)), 32)     # torch-CPU convs stop scaling (and oversubscribe) beyond ~32
        sample = 64                      # bounded sample: a 64^3 step is 1/15.6 of the 160^3 workload (linear in voxels)
        warm = min(args.warmup, 1)
        sec = cpu_reference_steps(sample, args.steps, warm, threads)
        vps = 1.0 / (sec * (args.size / sample) ** 3)
        out = {'metric': metric, 'value': vps, 'unit': 'volumes/s', 'n_gpus': args.gpus, 'steps': args.steps,
               'warmup': warm, 'ms_per_step': 1e3 / vps, 'higher_is_better': True, 'scaling': 'weak',
               'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config, 'impl': 'reference',
               'cpu_baseline': {'value': vps, 'unit': 'volumes/s', 'cores': threads, 'kind': 'port',
                                'sample': 'oracle restatement (NumPy generator + torch-CPU fp32 U-Net step) on %d^3 '
                                          'sub-volumes, scaled by voxel count to %d^3; reference TF-CPU: not run '
                                          '(TensorFlow 2.0 not installable)' % (sample, args.size)},
               'e2e': {'value': vps, 'unit': 'volumes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    from synthsr_b200._lib import lib
    from synthsr_b200.generator import GeneratorPlan
    from synthsr_b200.trainer import TrainingEngine

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    maps, pm, ps, gl, gc = make_inputs(args.size, 2, seed=rank)
    plan = GeneratorPlan([args.size] * 3, True, 0, gl, None, 1., None, **TRAINING_DEFAULTS)
    eng = TrainingEngine(plan, batchsize=1, conv_impl=args.conv_impl, seed=0, rank=rank, world_size=world)
    dev_maps = [torch.from_numpy(m[None]).cuda() for m in maps]
    pinned = [torch.from_numpy(m[None]).pin_memory() for m in maps]
    rng = np.random.default_rng(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_loss = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]
    lab_dev = [torch.empty_like(dev_maps[0]) for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]

    pipelined = not args.no_pipeline

    def step(lab, m, s):
        """one step = one generator pass + one U-Net training pass.  Pipelined (default): the batch of this call is
        generated on the generator stream while the network trains on the batch of the previous call, like the
        reference's fit_generator queue; the returned loss is the previous batch's (None on the very first call)."""
        return eng.train_step_pipelined(lab, m, s) if pipelined else eng.train_step(lab, m, s)

    def run_steps(n, host_inputs):
        """host_inputs (e2e): every step copies its label map from pinned host memory and reads a loss back to the host;
        the read-back of step i is consumed while step i+1 is being enqueued (one step of lag, like a logging callback)."""
        loss = None
        for i in range(n):
            m, s = draw_gmm(rng, pm, ps, gc)
            if host_inputs:
                if pipelined:
                    l = step(pinned[i % len(pinned)], m, s)                        # H2D of this step's input on the generator stream
                else:
                    lab = lab_dev[i % 2]
                    lab.copy_(pinned[i % len(pinned)], non_blocking=True)          # H2D of this step's input, in-stream
                    l = step(lab, m, s)
                if l is not None:
                    host_loss[i % 2].copy_(l, non_blocking=True)
                loss_ev[i % 2].record()
                if i > 0:
                    loss_ev[(i - 1) % 2].synchronize()
                    loss = float(host_loss[(i - 1) % 2][0])
            else:
                loss = step(dev_maps[i % len(dev_maps)], m, s)
        if host_inputs and n > 0:
            loss_ev[(n - 1) % 2].synchronize()
            loss = float(host_loss[(n - 1) % 2][0])
            assert np.isfinite(loss), 'loss is not finite'
        return loss

    def timed(n, host_inputs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = lib.ssr_launch_count()
        e0.record()
        run_steps(n, host_inputs)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), lib.ssr_launch_count() - l0

    run_steps(max(args.warmup, 3), False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = timed(args.steps, False)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = world * args.steps / (ms / 1e3)
    run_steps(max(args.warmup, 3), True)              # the host-buffer path gets the same warm-up as the device path
    def staged():
        return sum(g.stage.bytes_moved for g in (eng._gens or [eng.gen]))
    sg_before = staged()
    ms_e2e, _ = timed(args.steps, True)
    e2e = world * args.steps / (ms_e2e / 1e3)
    h2d = maps[0].nbytes + (staged() - sg_before) // max(args.steps, 1)

    # ---- roofline of the dominant kernel class: live CUDA-event timing of every convolution launch ------------
    eng.flush()                                  # train the pending batch of the pipelined loop, then profile unpipelined
    torch.cuda.synchronize()
    eng.net.prof = []
    for i in range(2):
        eng.train_step(dev_maps[i % len(dev_maps)], *draw_gmm(rng, pm, ps, gc))
    torch.cuda.synchronize()
    agg = {}
    for kind, fl, a, b in eng.net.prof:
        t, f, n = agg.get(kind, (0., 0., 0))
        agg[kind] = (t + a.elapsed_time(b), f + fl, n + 1)
    eng.net.prof = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('bf16_tflops_sustained', 1590.0)
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained; TF32 runs at half the bf16 rate)' if peaks else \
        'fallback 1590 TF/s bf16 (B200_PROFILING.md)'
    # dominant kernel = the tensor-core convolution kind with the largest share of the step (each kind is one kernel
    # family: wgrad_tc -> wgrad_tc_persistent_kernel, fwd_tc / dgrad_tc -> conv3d_tc_kernel + conv3d_tc_k2n_kernel +
    # conv3d_tc_up_kernel).  FLOPs are ALGORITHMIC (2*27*Cin*Cout*voxels of the reference's layer): the
    # parity path of the decoder convolutions executes 8 instead of 27 taps on the upsampled channels.
    tc = {k: v for k, v in agg.items() if k.endswith('_tc')}
    dom = max(tc, key=lambda k: tc[k][0]) if tc else None
    names = {'wgrad_tc': 'wgrad_tc_persistent_kernel (<0> plain, <1> parity classes of the decoder convolutions)',
             'fwd_tc': 'conv3d_tc_kernel / conv3d_tc_k2n_kernel / conv3d_tc_up_kernel<1> (forward)',
             'dgrad_tc': 'conv3d_tc_kernel / conv3d_tc_k2n_kernel / conv3d_tc_up_kernel<2> (data gradient)'}
    achieved = tc[dom][1] / (tc[dom][0] * 1e-3) / 1e12 if dom else 0.       # algorithmic 2*27*Cin*Cout*voxels per launch
    tc_ms = sum(v[0] for v in tc.values())
    tc_fl = sum(v[1] for v in tc.values())
    all_tc = tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.
    traffic, traffic_note = None, None
    try:      # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture (profiles/)
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
        if dom in tj:
            traffic, traffic_note = tj[dom]['dram_bytes_per_launch'], tj[dom]['launch']
    except Exception:
        pass
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_launch': traffic_note, 'peak_source': peak_src,
                'kernel': '%s (tcgen05 kind::tf32), %.2f ms of the step in %d launches' % (
                    names.get(dom, dom), tc[dom][0] / 2, tc[dom][2] // 2) if dom else None,
                'frac_of_tf32_peak': achieved / (peak / 2),
                'all_tc_convolutions': {'tflops': all_tc, 'frac_of_tf32_peak': all_tc / (peak / 2), 'ms_per_step': tc_ms / 2},
                'per_kind': {k: {'ms_per_step': agg[k][0] / 2, 'tflops': agg[k][1] / (agg[k][0] * 1e-3) / 1e12 if agg[k][0] else 0.,
                                 'launches_per_step': agg[k][2] // 2} for k in sorted(agg)},
                'conv_share_of_step': (sum(v[0] for v in agg.values()) / 2) / (ms / args.steps),
                'note': 'per-kind times are measured with the weight-gradient overlap disabled (serialised launches)'}
    out = {'metric': metric, 'value': value, 'unit': 'volumes/s', 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'tf32' if args.conv_impl in ('tc', 'tc3') else 'f32', 'data': 'synthetic',
           'config': config, 'clocks': sampler.summary(), 'gpu_launches': int(launches),
           'e2e': {'value': e2e, 'unit': 'volumes/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 8},
           'roofline': roofline, 'conv_gflop_per_step': step_f / 1e9}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = min(len(os.sched_getaffinity(0)), 32)
        sample = 64
        sec = cpu_reference_steps(sample, 2, 1, threads)
        vps = 1.0 / (sec * (args.size / sample) ** 3)
        out['cpu_baseline'] = {'value': vps, 'unit': 'volumes/s', 'cores': threads, 'kind': 'port',
                               'sample': '2 oracle steps (NumPy generator + torch-CPU fp32 U-Net fwd/bwd/Adam) on a 64^3 '
                                         'volume, scaled by voxel count to %d^3' % args.size}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
