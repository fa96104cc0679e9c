#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): training volumes/sec on 160^3 single-channel volumes, 5-level 24-feature U-Net,
generator + U-Net + Adam every step, at N GPUs of one node (data parallel, one process per GPU).

    python bench.py --gpus 1 --steps 50 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # CPU arm: the oracle restatement of the reference on the host cores
    python bench.py --config c1|c4|c5           # the other BASELINE.json configs (c2 = the headline, default)

Prints ONE JSON line (rank 0).  A "step" is one pass of the hot path (generate a synthetic scan pair from a label map,
U-Net forward/backward, L1 loss, Adam) over one mini-batch of 1 volume per GPU.

  value   K steps timed on the device (CUDA events, barrier + synchronize on both sides, max over ranks), label maps
          already resident in HBM; per-step percentiles from one event per step
  e2e     the same metric through the drop-in call a user makes -- SynthSR.training.training() on a directory of .npz
          label maps (epochs=2, steps_per_epoch=K; epoch 1 warms up) -- timed from the CUDA events the engine records after
          every step of epoch 2: host sampler, pinned label map -> H2D, generator, U-Net step, loss read-back, every step
  parity  measured live on one batch: prediction / loss of the benchmarked mode against the exact-fp32 mode
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
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
HYPERFINE_RES = [[1.5, 1.5, 5.], [1.5, 1.5, 5.]]


def conv_flops_per_step(shape, cin=1):
    """algorithmic conv FLOPs of one training step: fwd + dgrad + wgrad."""
    from synthsr_b200.unet import layer_specs
    shape = [shape] * 3 if isinstance(shape, (int, np.integer)) else shape
    v = float(np.prod(shape))
    fwd, first = 0., None
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


def make_inputs(shape, n_maps=2, seed=0, n_channels=1):
    from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors
    shape = [int(s) for s in (shape if isinstance(shape, (list, tuple)) else [shape] * 3)]
    maps = []
    for i in range(n_maps):
        lo = phantom_labels([(s + 1) // 2 for s in shape], GEN_LABELS, seed=seed + i)
        maps.append(np.ascontiguousarray(np.repeat(np.repeat(np.repeat(lo, 2, 0), 2, 1), 2, 2)[:shape[0], :shape[1], :shape[2]]))
    pm, ps = synthetic_priors(int(GEN_CLASSES.max()) + 1, n_channels, seed)
    return maps, pm, ps, GEN_LABELS, GEN_CLASSES


def draw_gmm(rng, pm, ps, classes, n_channels=1):
    """SynthSR/model_inputs.py:118-123 ('normal' priors, negatives clipped)."""
    m = np.stack([np.clip(rng.normal(pm[2 * c], pm[2 * c + 1]), 0, None)[classes] for c in range(n_channels)], -1)[None]
    s = np.stack([np.clip(rng.normal(ps[2 * c], ps[2 * c + 1]), 0, None)[classes] for c in range(n_channels)], -1)[None]
    return m.astype(np.float32), s.astype(np.float32)


def workload(config, size):
    """-> (label shape, (input_channels, output_channel), GeneratorPlan kwargs, engine kwargs, description)"""
    if config == 'c4':       # BASELINE configs[3]: Hyperfine T1+T2 (scripts/predict_command_line_hyperfine.py:60-73)
        gen = dict(TRAINING_DEFAULTS, data_res=np.array(HYPERFINE_RES), thickness=np.array(HYPERFINE_RES), downsample=True,
                   simulate_registration_error=True)
        return [192, 192, 64], ([False, True, True], 0), gen, dict(work_with_residual_channel=[0]), \
            'Hyperfine 192x192x64: synthetic 1 mm target + T1/T2 inputs at 1.5x1.5x5 mm with registration error -> 2-channel ' \
            '5-level 24-feature U-Net (residual on channel 0) fwd/bwd, L1, Adam; batch 1 per GPU (BASELINE configs[3])'
    n = 256 if config == 'c5' else size
    return [n] * 3, (True, 0), dict(TRAINING_DEFAULTS), {}, \
        '%d^3 single-channel label map -> generator (training() defaults) -> 5-level 24-feature U-Net fwd/bwd, L1, Adam; ' \
        'batch 1 per GPU (BASELINE configs[%s])' % (n, '4' if config == 'c5' else '1]/[2')


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
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': []}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle restatement; TensorFlow is not installable here)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(shape, steps, warmup, threads, budget_s=None, gen_kw=None, channels=(True, 0), cin=1, loss_kw=None):
    """generator (NumPy) + U-Net fwd/bwd + Adam (torch CPU fp32) on volumes of `shape`: REAL steps at the real size.
    Stops early once `budget_s` seconds of timed steps are spent (at least one).  -> (seconds per step, steps timed)."""
    import torch
    from oracle import generator as OG
    from oracle import unet as OU
    from synthsr_b200.draws import sample_draws
    from synthsr_b200.generator import GeneratorPlan
    torch.set_num_threads(threads)
    shape = [shape] * 3 if isinstance(shape, (int, np.integer)) else list(shape)
    n_ch = len(channels[0]) if isinstance(channels[0], (list, tuple)) else 1
    maps, pm, ps, gl, gc = make_inputs(shape, 1, n_channels=n_ch)
    cfg = dict(gen_kw or TRAINING_DEFAULTS)
    plan = GeneratorPlan(shape, channels[0], channels[1], gl, None, 1., None, **cfg)
    rng = np.random.default_rng(0)
    params = OU.init_params(0, cin)
    opt = OU.adam_init(params)
    ocfg = dict(cfg, generation_labels=gl)
    if n_ch > 1:
        ocfg.update(input_channels=channels[0], output_channel=channels[1])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        draws = sample_draws(rng, plan, 1, gmm_noise=True)
        m, s = draw_gmm(rng, pm, ps, gc, n_ch)
        image, target = OG.labels_to_image(ocfg, [maps[0][None, ..., None], m, s], draws)
        OU.train_step(params, opt, torch.from_numpy(image), torch.from_numpy(target), lr=1e-4, **(loss_kw or {}))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
            if budget_s is not None and sum(times) > budget_s:
                break
    return float(np.mean(times)), len(times)


def cpu_generator_calls(shape, calls):
    """c1 on the host: oracle generator (NumPy float32) with BrainGenerator's defaults -> seconds per volume"""
    from oracle import generator as OG
    from synthsr_b200.draws import sample_draws
    from synthsr_b200.generator import GeneratorPlan
    maps, pm, ps, gl, gc = make_inputs(shape, 1)
    plan = GeneratorPlan(list(shape), True, 0, gl, None, 1., None)
    rng = np.random.default_rng(0)
    times = []
    for it in range(calls + 1):
        t0 = time.perf_counter()
        draws = sample_draws(rng, plan, 1, gmm_noise=True)
        m, s = draw_gmm(rng, pm, ps, gc)
        OG.labels_to_image(dict(generation_labels=gl), [maps[0][None, ..., None], m, s], draws)
        if it:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


# ---------------------------------------------------------------------------------------------------------------------
def write_dataset(root, maps, pm, ps, gl, gc):
    """the on-disk inputs SynthSR.training.training() / BrainGenerator take: a folder of label maps + .npy hyper-parameters"""
    lab_dir = os.path.join(root, 'labels')
    os.makedirs(lab_dir, exist_ok=True)
    for i, m in enumerate(maps):
        np.savez(os.path.join(lab_dir, 'map_%02d.npz' % i), vol_data=m.astype(np.int32))
    paths = {}
    for name, arr in (('prior_means', pm), ('prior_stds', ps), ('generation_labels', gl), ('generation_classes', gc)):
        paths[name] = os.path.join(root, name + '.npy')
        np.save(paths[name], arr)
    return lab_dir, paths


def percentiles(ms):
    ms = np.asarray(ms, dtype=np.float64)
    return {'median': float(np.median(ms)), 'p10': float(np.percentile(ms, 10)), 'p90': float(np.percentile(ms, 90)),
            'n': int(ms.size)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c1', 'c2', 'c4', 'c5'],
                    help='BASELINE.json configs: c1 generator only 64^3 via BrainGenerator.generate_brain(); c2 160^3 training '
                         'step (headline, also configs[2] at N GPUs); c4 Hyperfine 192x192x64 two input channels; c5 256^3')
    ap.add_argument('--size', type=int, default=SIZE)
    ap.add_argument('--conv-impl', default='tc3', choices=['tc3', 'tc', 'ref'],
                    help="tc3 (default): forward compensated to fp32-class accuracy (the parity-gated mode); tc: plain TF32 "
                         "(fast, outside the 1e-3 bar); ref: exact fp32 CUDA-core convolutions")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the SynthSR.training.training() leg')
    ap.add_argument('--no-extras', action='store_true', help='skip the live parity check and the fast-mode secondary number')
    ap.add_argument('--no-pipeline', action='store_true',
                    help='generate and train on the same batch inside one call (no generator / training overlap)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.config == 'c1':
        return bench_generator(args, rank, local_rank, world)
    shape, channels, gen_kw, eng_kw, desc = workload(args.config, args.size)
    cin = sum(bool(c) for c in channels[0]) if isinstance(channels[0], (list, tuple)) else 1
    n_ch = len(channels[0]) if isinstance(channels[0], (list, tuple)) else 1
    fwd_f, step_f = conv_flops_per_step(shape, cin)
    metric = 'training volumes/sec (160^3, 24-ch 5-level U-Net, generator+U-Net+Adam step)'
    if args.config != 'c2' or args.size != SIZE:
        metric = 'training volumes/sec (%s, 24-ch 5-level U-Net, generator+U-Net+Adam step)' % 'x'.join(str(s) for s in shape)
    config = {'workload': desc, 'global_batch': args.gpus, 'volume': shape, 'parallelism': 'dp%d' % args.gpus,
              'l2_policy': 'per-step working set (>4 GB of activations) exceeds the 126 MB L2; no explicit flush',
              'conv_impl': args.conv_impl,
              'precision': {'tc3': 'forward: fp32 operands split into 8-bit pieces, three bf16 MMA terms per convolution (bf16x3; '
                                   'TF32 + one bf16 correction chain on the 24-channel layers), fp32 accumulation in TMEM -- '
                                   'fp32-class results (1e-5 per convolution); backward: TF32 operands, fp32 accumulation; first '
                                   'layer, BatchNorm, loss, Adam: fp32',
                            'tc': 'forward and backward: TF32 operands, fp32 accumulation (outside the 1e-3 parity bar)',
                            'ref': 'fp32 on the CUDA cores'}.get(args.conv_impl),
              'pipeline': 'generator of batch i+1 overlaps the U-Net step of batch i (one generator pass + one training '
                          'pass per step, as the reference\'s fit_generator queue)' if not args.no_pipeline
                          else 'none (generate, then train, inside each step)'}
    loss_kw = dict(work_with_residual_channel=eng_kw['work_with_residual_channel']) if eng_kw else {}

    if args.impl == 'reference':
        if rank != 0:
            return
        threads = min(len(os.sched_getaffinity(0)), 32)     # torch-CPU convs stop scaling (and oversubscribe) beyond ~32
        warm = min(args.warmup, 1)
        sec, done = cpu_reference_steps(shape, args.steps, warm, threads, budget_s=150., gen_kw=gen_kw, channels=channels,
                                        cin=cin, loss_kw=loss_kw)
        vps = 1.0 / sec
        out = {'metric': metric, 'value': vps, 'unit': 'volumes/s', 'n_gpus': args.gpus, 'steps': done,
               'warmup': warm, 'ms_per_step': 1e3 * sec, 'higher_is_better': True, 'scaling': 'weak',
               'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config, 'impl': 'reference',
               'cpu_baseline': {'value': vps, 'unit': 'volumes/s', 'cores': threads, 'kind': 'port',
                                'sample': '%d real steps at the full %s size (requested %d; the loop stops once 150 s of timed '
                                          'steps are spent): oracle restatement of the reference graph -- NumPy generator + '
                                          'torch-CPU fp32 U-Net fwd/bwd/Adam; reference TF-CPU itself: not run (TensorFlow '
                                          '2.0 not installable here)' % (done, 'x'.join(str(s) for s in shape), args.steps)},
               'e2e': {'value': vps, 'unit': 'volumes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    from synthsr_b200 import trainer as T
    from synthsr_b200._lib import lib
    from synthsr_b200.generator import GeneratorPlan
    from synthsr_b200.trainer import TrainingEngine

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    maps, pm, ps, gl, gc = make_inputs(shape, 2, seed=rank, n_channels=n_ch)
    plan = GeneratorPlan(shape, channels[0], channels[1], gl, None, 1., None, **gen_kw)
    rng = np.random.default_rng(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def build(conv_impl):
        return TrainingEngine(plan, batchsize=1, conv_impl=conv_impl, seed=0, rank=rank, world_size=world, **eng_kw)

    def run_device_steps(eng, n, dev_maps):
        loss = None
        for i in range(n):
            m, s = draw_gmm(rng, pm, ps, gc, n_ch)
            if args.no_pipeline:
                loss = eng.train_step(dev_maps[i % len(dev_maps)], m, s)
            else:
                loss = eng.train_step_pipelined(dev_maps[i % len(dev_maps)], m, s)
        return loss

    def timed_device(eng, n, dev_maps):
        """EXACTLY n steps between two events, barrier + synchronize on both sides, max over ranks; one event per step"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        T.STEP_EVENTS = []
        barrier()
        l0 = lib.ssr_launch_count()
        e0.record()
        run_device_steps(eng, n, dev_maps)
        e1.record()
        barrier()
        evs, T.STEP_EVENTS = T.STEP_EVENTS, None
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]
        return ms.item(), lib.ssr_launch_count() - l0, per_step

    eng = build(args.conv_impl)
    dev_maps = [torch.from_numpy(m[None]).cuda() for m in maps]
    run_device_steps(eng, max(args.warmup, 3), dev_maps)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches, per_step = timed_device(eng, args.steps, dev_maps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = world * args.steps / (ms / 1e3)
    eng.flush()
    torch.cuda.synchronize()

    replicas = None
    if world > 1:      # NCCL path correctness: every replica must hold bit-identical parameters after the timed steps
        chk = torch.stack([eng.net.params.double().sum(), eng.net.params.double().abs().sum(),
                           eng.net.comm[eng.net.n_params:].double().sum()])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        replicas = bool(all(torch.equal(allc[0], c) for c in allc))
        assert replicas, 'replicas diverged: parameter checksums differ across ranks'

    # ---- roofline of the dominant kernel class: live CUDA-event timing of every convolution launch ------------
    eng.net.prof = []
    for i in range(2):
        eng.train_step(dev_maps[i % len(dev_maps)], *draw_gmm(rng, pm, ps, gc, n_ch))
    torch.cuda.synchronize()
    agg, executed = {}, {}
    if os.environ.get('SSR_BENCH_DUMP_PROF'):
        for j, (kind, fl, a, b, mult) in enumerate(eng.net.prof):
            print('prof %3d %-9s %8.3f ms  x%d' % (j, kind, a.elapsed_time(b), mult), file=sys.stderr)
    for kind, fl, a, b, mult in eng.net.prof:
        t, f, n = agg.get(kind, (0., 0., 0))
        agg[kind] = (t + a.elapsed_time(b), f + fl, n + 1)
        executed[kind] = executed.get(kind, 0.) + fl * mult
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
    # parity path of the decoder convolutions executes 8 instead of 27 taps on the upsampled channels, the compensated
    # forward executes 1.5 (bf16x3) or 2 (hybrid, the 24-channel layers) MMA chains per algorithmic one.
    tc = {k: v for k, v in agg.items() if k.endswith('_tc')}
    dom = max(tc, key=lambda k: tc[k][0]) if tc else None
    names = {'wgrad_tc': 'wgrad_tc_persistent_kernel (<0> plain, <1> parity classes of the decoder convolutions)',
             'fwd_tc': 'conv3d_tc_kernel / conv3d_tc_k2n_kernel / conv3d_tc_up_kernel<1> (forward%s)' % (
                 '; compensated on every layer but uparm_8_0: bf16x3 = three bf16 K-chunks x1 w1 + x2 w1 + x1 w2 per 64 input '
                 'channels (kind::f16), the 24-channel layers TF32 chain + one bf16 correction chain, '
                 '+ tf32_split_bf16_kernel' if args.conv_impl == 'tc3' else ''),
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
    step_tf = step_f / (ms / args.steps * 1e-3) / 1e12
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_launch': traffic_note, 'peak_source': peak_src,
                'kernel': '%s (tcgen05 kind::tf32 / kind::f16), %.2f ms of the step in %d launches' % (
                    names.get(dom, dom), tc[dom][0] / 2, tc[dom][2] // 2) if dom else None,
                'frac_of_tf32_peak': achieved / (peak / 2),
                'all_tc_convolutions': {'tflops': all_tc, 'frac_of_tf32_peak': all_tc / (peak / 2), 'ms_per_step': tc_ms / 2},
                'whole_step': {'tflops': step_tf, 'frac_of_tf32_peak': step_tf / (peak / 2)},
                'executed': {'note': 'tensor-core work actually issued, in TF32-equivalent FLOPs: a compensated forward '
                                     'convolution runs 1.5 MMA chains per algorithmic one in the bf16x3 scheme (three bf16 '
                                     'K-chunks, each covering twice the K of a TF32 instruction) and 2 on the 24-channel layers '
                                     '(hybrid), so the hardware utilisation of the forward is this figure, not `achieved`',
                             'dominant_kind_tflops': executed.get(dom, 0.) / (tc[dom][0] * 1e-3) / 1e12 if dom else 0.,
                             'dominant_kind_frac_of_tf32_peak': (executed.get(dom, 0.) / (tc[dom][0] * 1e-3) / 1e12) / (peak / 2) if dom else 0.,
                             'all_tc_tflops': sum(executed.get(k, 0.) for k in tc) / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.},
                'per_kind': {k: {'ms_per_step': agg[k][0] / 2, 'tflops': agg[k][1] / (agg[k][0] * 1e-3) / 1e12 if agg[k][0] else 0.,
                                 'launches_per_step': agg[k][2] // 2} for k in sorted(agg)},
                'conv_share_of_step': (sum(v[0] for v in agg.values()) / 2) / (ms / args.steps),
                'note': 'FLOPs are algorithmic (one per reference multiply-add x 2); per-kind times are measured with the '
                        'weight-gradient overlap disabled (serialised launches)'}
    out = {'metric': metric, 'value': value, 'unit': 'volumes/s', 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'tf32' if args.conv_impl in ('tc', 'tc3') else 'f32', 'data': 'synthetic',
           'config': config, 'clocks': sampler.summary(), 'gpu_launches': int(launches),
           'step_ms': percentiles(per_step) if per_step else None,
           'roofline': roofline, 'conv_gflop_per_step': step_f / 1e9}
    if replicas is not None:
        out['replicas_identical'] = replicas

    # ---- live parity of the benchmarked mode on one generated batch, against the exact-fp32 mode -----------------------
    image = eng.gen.image.clone()
    target = eng.gen.target.clone()
    del eng
    torch.cuda.empty_cache()
    if not args.no_extras and args.conv_impl != 'ref' and rank == 0:
        from synthsr_b200.unet import UNet3D
        res = {}
        for impl in ('ref', args.conv_impl):
            net = UNet3D(plan.image_shape, batchsize=1, conv_impl=impl, seed=0, nb_labels=plan.n_target_channels)
            loss = net.loss_and_grad(image, target, 'l1', eng_kw.get('work_with_residual_channel'), None)
            torch.cuda.synchronize()
            res[impl] = (net.pred.double().clone(), loss.item(), net.grads.double().clone())
            del net
            torch.cuda.empty_cache()
        (p0, l0_, g0), (p1, l1_, g1) = res['ref'], res[args.conv_impl]
        out['parity'] = {'mode': args.conv_impl, 'against': "exact-fp32 CUDA-core mode (conv_impl='ref') on the same device, "
                         'same generated batch, random-init weights (seed 0); the float64 oracle comparisons are in '
                         'tests/test_unet_parity_gpu.py',
                         'pred_rel_l2': float((p1 - p0).norm() / p0.norm()),
                         'pred_max_over_max': float((p1 - p0).abs().max() / p0.abs().max()),
                         'loss_rel': abs(l1_ - l0_) / abs(l0_),
                         'grad_rel_l2': float((g1 - g0).norm() / g0.norm()), 'bar': '1e-3 prediction / loss, 1e-2 gradients'}
        del res, p0, p1, g0, g1
    if world > 1:
        dist.barrier()

    # ---- secondary number: the plain-TF32 fast mode (outside the parity bar), same protocol ---------------------------
    if not args.no_extras and args.conv_impl == 'tc3':
        eng2 = build('tc')
        run_device_steps(eng2, max(args.warmup, 3), dev_maps)
        ms2, _, _ = timed_device(eng2, args.steps, dev_maps)
        eng2.flush()
        out['fast_mode'] = {'conv_impl': 'tc', 'value': world * args.steps / (ms2 / 1e3), 'ms_per_step': ms2 / args.steps,
                            'note': 'plain TF32 forward: 2-3e-3 on the prediction of a randomly initialised net, i.e. OUTSIDE '
                                    'the 1e-3 parity bar; not the headline'}
        del eng2
        torch.cuda.empty_cache()
    del dev_maps

    # ---- e2e: the drop-in call, SynthSR.training.training(), on a directory of label maps --------------------------------
    if not args.no_e2e:
        from SynthSR.training import training
        tmp = tempfile.mkdtemp(prefix='ssr_bench_%d_' % rank)
        try:
            lab_dir, paths = write_dataset(tmp, maps, pm, ps, gl, gc)
            kw = dict(gen_kw)
            kw.pop('output_div_by_n')
            if args.config == 'c4':
                kw.update(input_channels=channels[0], output_channel=channels[1], **eng_kw)
            os.environ['SSR_CONV_IMPL'] = args.conv_impl
            os.environ.setdefault('SSR_SEED', '0')
            T.STEP_EVENTS = []
            barrier()
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):         # training() prints its per-epoch line; stdout is ONE JSON line
                training(lab_dir, os.path.join(tmp, 'models'), paths['prior_means'], paths['prior_stds'],
                         paths['generation_labels'], path_generation_classes=paths['generation_classes'], batchsize=1,
                         epochs=2, steps_per_epoch=args.steps, **kw)
            barrier()
            evs, T.STEP_EVENTS = T.STEP_EVENTS, None
            assert len(evs) == 2 * args.steps, len(evs)
            # steady-state steps of epoch 2, measured from its FIRST step's event: K - 1 step intervals (the checkpoint of
            # epoch 1 is written between the last event of epoch 1 and the first of epoch 2, outside the interval)
            ms_e = torch.tensor([evs[args.steps].elapsed_time(evs[-1])], device='cuda')
            if world > 1:
                dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
            e2e = world * (args.steps - 1) / (ms_e.item() / 1e3)
            small = int(4 * (16 + 3 * int(np.prod(plan.svf_small_shape or [1])) + 64 * n_ch + 2 * plan.lut_len * n_ch))
            out['e2e'] = {'value': e2e, 'unit': 'volumes/s', 'h2d_bytes_per_step': int(maps[0].nbytes + small),
                          'd2h_bytes_per_step': 8,
                          'how': 'SynthSR.training.training(labels_dir of .npz maps, epochs=2, steps_per_epoch=%d): CUDA events '
                                 'after every step; %d steady-state step intervals of epoch 2 (epoch 1 = warm-up; the '
                                 'per-epoch checkpoint write lies between the epochs, outside the interval)' % (
                                     args.steps, args.steps - 1)}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    else:
        out['e2e'] = None

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = min(len(os.sched_getaffinity(0)), 32)
        sec, done = cpu_reference_steps(shape, 2, 0, threads, budget_s=20., gen_kw=gen_kw, channels=channels, cin=cin,
                                        loss_kw=loss_kw)
        out['cpu_baseline'] = {'value': 1.0 / sec, 'unit': 'volumes/s', 'cores': threads, 'kind': 'port',
                               'sample': '%d real oracle step(s) at the full %s size (NumPy generator + torch-CPU fp32 U-Net '
                                         'fwd/bwd/Adam); not TensorFlow' % (done, 'x'.join(str(s) for s in shape))}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def bench_generator(args, rank, local_rank, world):
    """BASELINE configs[0]: single 64^3 label map -> BrainGenerator.generate_brain() (SynthSR/brain_generator.py:317-330,
    the class' own defaults), 1 volume per call.  HBM-bound kernels: roofline on 20 algorithmic bytes per voxel."""
    shape = [64, 64, 64] if args.size == SIZE else [args.size] * 3
    vox = float(np.prod(shape))
    metric = 'generated volumes/sec (%d^3 label map -> BrainGenerator.generate_brain())' % shape[0]
    config = {'workload': 'single %d^3 label map -> BrainGenerator defaults (brain_generator.py:30-61) -> generate_brain(): '
                          '1 synthetic image + target per call, returned as NumPy arrays (BASELINE configs[0])' % shape[0],
              'global_batch': args.gpus, 'volume': shape, 'parallelism': 'replicas%d' % args.gpus,
              'l2_policy': 'inputs fit the L2 at this size (5.2 MB compulsory traffic per volume): launch- and host-bound'}
    if args.impl == 'reference':
        if rank != 0:
            return
        n = min(args.steps, 10)
        sec = cpu_generator_calls(shape, n)
        out = {'metric': metric, 'value': 1. / sec, 'unit': 'volumes/s', 'n_gpus': args.gpus, 'steps': n,
               'warmup': 1, 'ms_per_step': 1e3 * sec, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'f32', 'data': 'synthetic', 'config': config, 'impl': 'reference',
               'cpu_baseline': {'value': 1. / sec, 'unit': 'volumes/s', 'cores': 1, 'kind': 'port',
                                'sample': 'oracle generator (NumPy float32, single thread) on the same 64^3 label map; not TF'},
               'e2e': {'value': 1. / sec, 'unit': 'volumes/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(out))
        return
    import torch
    import torch.distributed as dist
    from SynthSR.brain_generator import BrainGenerator
    from synthsr_b200._lib import lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    maps, pm, ps, gl, gc = make_inputs(shape, 1, seed=rank)
    tmp = tempfile.mkdtemp(prefix='ssr_bench_c1_%d_' % rank)
    try:
        lab_dir, paths = write_dataset(tmp, maps, pm, ps, gl, gc)
        bg = BrainGenerator(lab_dir, paths['prior_means'], paths['prior_stds'], 'normal', paths['generation_labels'],
                            generation_classes=paths['generation_classes'])

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(max(args.warmup, 3)):
            bg.generate_brain()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        l0 = lib.ssr_launch_count()
        e0.record()
        for _ in range(args.steps):
            im, tgt = bg.generate_brain()
        e1.record()
        barrier()
        sampler.stop_flag = True
        sampler.join(timeout=2)
        launches = lib.ssr_launch_count() - l0
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = ms.item()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('hbm_gbs', 6650.)
    achieved = 20. * vox * args.steps / (ms * 1e-3) / 1e9
    value = world * args.steps / (ms / 1e3)
    out = {'metric': metric, 'value': value, 'unit': 'volumes/s', 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
           'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config, 'clocks': sampler.summary(),
           'gpu_launches': int(launches),
           'e2e': {'value': value, 'unit': 'volumes/s', 'h2d_bytes_per_step': int(maps[0].nbytes),
                   'd2h_bytes_per_step': int(im.nbytes + tgt.nbytes),
                   'how': 'generate_brain() IS the end-to-end call: host label map in, NumPy image + target out, every call'},
           'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                        'traffic': None, 'kernel': 'whole generator call (deform / gmm_bias / blur3d kernels; at 64^3 the call '
                                                   'is bound by its launches + the blocking NumPy round trip, not by HBM)',
                        'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if peaks else 'fallback 6650 GB/s'}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec = cpu_generator_calls(shape, 5)
        out['cpu_baseline'] = {'value': 1. / sec, 'unit': 'volumes/s', 'cores': 1, 'kind': 'port',
                               'sample': '5 oracle generator calls (NumPy float32, single thread) on the same 64^3 label map'}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
