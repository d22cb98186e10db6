#!/usr/bin/env python
"""Benchmark of the CaSPR reconstruction hot path (BASELINE.json metric: reconstructed points/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one ``CaSPR.reconstruct`` call (TPointNet++ encode -> latent ODE -> CNF decode) over one
batch of synthetic sequences.  Workload = BASELINE.json configs[1] ("rigid cars"): 10 frames x 1024
input points, 2048 reconstructed points per frame, batch 8 per GPU; with N GPUs every rank processes
its own batch of 8 sequences (batch-sharded, no data-path collective; weak scaling).

Prints ONE JSON line on rank 0 (see the driver contract): `value` is device-resident throughput,
`e2e` goes through the public API with host buffers (pinned H2D of the input, CPU-generator base
samples as the reference draws them, D2H of the reconstruction), `roofline` is the dominant kernel
timed with CUDA events inside the timed region, `cpu_baseline` is the CPU oracle on a bounded sample.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch                      # noqa: E402

WORKLOADS = {
    # name: (B per GPU, T, N input pts, P sampled pts, interpolated query steps or None, warping)
    'cars_rigid_T10_N1024_P2048_B8': (8, 10, 1024, 2048, None, False),          # BASELINE configs[1]
    'chairs_rigid_T10_N2048_P2048_B4': (4, 10, 2048, 2048, None, False),        # configs[2]: 32 over 8 GPUs
    'warping_cars_T10_N2048_P2048_S20_B4': (4, 10, 2048, 2048, 20, True),       # configs[3]
}
DEFAULT_WORKLOAD = 'cars_rigid_T10_N1024_P2048_B8'
METRIC = 'reconstructed_points_per_sec'
UNIT = 'points/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--cnf-init', default='vigorous', choices=['vigorous', 'default'])
    ap.add_argument('--engine', default='auto', choices=['auto', 'simt', 'tc'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-clock-sampler', action='store_true',
                    help='do not spawn the nvidia-smi sampler (it hangs under ncu): lets bench.py itself be profiled')
    ap.add_argument('--no-cuda-graph', action='store_true',
                    help='launch the encoder kernels eagerly instead of replaying its CUDA graph (per-kernel ncu lists)')
    ap.add_argument('--no-extra-configs', action='store_true',
                    help='skip the "configs" block (BASELINE configs[2..4] per-GPU shards)')
    ap.add_argument('--ref-budget-s', type=float, default=180.0,
                    help='--impl reference: wall-clock budget for the timed steps (each step is the full batch)')
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'tf_burst': p['bf16_tflops'], 'tf_sustained': p['bf16_tflops_sustained'],
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = str(gpu_index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', self.gpu], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# --------------------------------------------------------------------------------- inputs
def make_inputs(workload, seed, cnf_init):
    from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
    B, T, N, P, S, warping = WORKLOADS[workload]
    sd = synthetic_state_dict(0, cnf_init=cnf_init)
    x, _ = synthetic_sequences(B, T, N, seed=100 + seed, warping=warping, max_timestamp=1.0 if warping else 5.0)
    Tq = S if S is not None else T
    g = torch.Generator().manual_seed(1000 + seed)
    nb = B if S is not None else B * T                     # constant_in_time shares one base cloud per sequence
    y = torch.randn(nb, P, 3, generator=g)
    e = torch.randn(B * Tq, P, 3, generator=g)
    kwargs = {'num_points': P}
    if S is not None:
        kwargs.update(constant_in_time=True, timestamps=torch.linspace(0, 1, S), max_timestamp=1.0)
    return sd, x, y, e, kwargs, (B, T, N, P, Tq)


def cpu_reference_step(oracle, x1, y1, e1, kwargs):
    t0 = time.perf_counter()
    oracle.reconstruct(x1, y=y1, e=e1, **kwargs)
    return time.perf_counter() - t0


def workload_config(workload, world, cnf_init, nfe):
    """The `config` object both arms print (identical keys and values for the same workload)."""
    B, T, N, P, S, _ = WORKLOADS[workload]
    return {'workload': workload, 'batch_per_gpu': B, 'global_batch': B * world, 'frames': T, 'input_points': N,
            'sampled_points': P, 'query_steps': S if S is not None else T, 'cnf_init': cnf_init,
            'nfe_latent_cnf': [int(v) for v in nfe], 'parallelism': 'batch-shard x%d' % world}


class ReferenceModules(object):
    """The reference's own `models.caspr.CaSPR` (unmodified, imported from /root/reference over the oracle shims for its
    absent dependencies) behind the oracle's reconstruct signature; only where /root/reference exists."""

    def __init__(self, sd):
        from oracle.reference_loader import build_reference_caspr
        self.model = build_reference_caspr()
        self.model.load_state_dict(sd)
        self.model.eval()

    def reconstruct(self, x, y=None, e=None, **kwargs):
        # the unmodified modules draw y and e themselves (CPU generator): same cost, different values
        with torch.no_grad():
            return self.model.reconstruct(x, **kwargs)

    def get_nfe(self):
        return [int(v) for v in self.model.get_nfe()]


def run_reference(args, rank, world):
    """The reference's own CPU path (PyTorch + torchdiffeq-0.0.1 semantics) on all host cores: the unmodified reference
    modules over the oracle shims when /root/reference is present (kind "reference"), else the oracle port of the same
    math (kind "port").  Every step is ONE `reconstruct` call over the FULL per-GPU batch of the workload — the same
    batch-global step controllers as the GPU arm.  A full step costs about a minute of CPU, so the number of timed steps
    is capped by --ref-budget-s and warm-up by one step; the line reports the steps actually run."""
    if rank != 0:
        return
    from oracle.reference_loader import reference_available
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, x, y, e, kwargs, (B, T, N, P, Tq) = make_inputs(args.workload, 0, args.cnf_init)
    if reference_available():
        impl, kind = ReferenceModules(sd), 'reference'
    else:
        from oracle.caspr_oracle import CasprOracle
        impl, kind = CasprOracle(sd), 'port'
    t_first = cpu_reference_step(impl, x, y, e, kwargs)               # warm-up (page-in, thread pools)
    n_warm = 1
    while n_warm < args.warmup and t_first * (n_warm + 2) < 0.25 * args.ref_budget_s:
        cpu_reference_step(impl, x, y, e, kwargs)
        n_warm += 1
    times = []
    while len(times) < args.steps and (not times or sum(times) + times[-1] < args.ref_budget_s):
        times.append(cpu_reference_step(impl, x, y, e, kwargs))
    total = sum(times)
    value = B * Tq * P * len(times) / total
    sample = ('full per-GPU batch per step (B=%d,T=%d,N=%d,P=%d), %s, %d threads; %d of %d requested steps timed '
              '(%.0f s budget), %d warm-up' % (B, T, N, P, 'unmodified reference modules over shims' if kind ==
                                              'reference' else 'oracle port', cores, len(times), args.steps,
                                              args.ref_budget_s, n_warm))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': len(times), 'steps_requested': args.steps, 'warmup': n_warm,
            'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.workload, max(args.gpus, 1), args.cnf_init, impl.get_nfe()),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def cpu_training_step(T, N):
    """CPU-baseline leg of the training benchmark (tools/bench_train.py): the reference's training step on the host cores
    — the differentiable oracle port (oracle/train_oracle.py: reference math, torchdiffeq-0.0.1 odeint_adjoint
    restatement) on a bounded sample, ONE sequence of the workload."""
    from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
    from oracle.train_oracle import TrainOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic_state_dict(0, cnf_init='vigorous')
    x, nocs = synthetic_sequences(1, T, N, seed=200)
    e = torch.randn(T, N, 3, generator=torch.Generator().manual_seed(0))
    orc = TrainOracle(sd)
    t0 = time.perf_counter()
    loss = TrainOracle.loss(*orc.forward_train(x, nocs, e))
    loss.backward()
    dt = time.perf_counter() - t0
    return {'metric': 'trained_points_per_sec', 'value': T * N / dt, 'unit': 'points/s', 'cores': cores, 'kind': 'port',
            'sample': '1 sequence (T=%d, N=%d), forward + backward, 1 run, %.1f s, nfe %s' % (T, N, dt, orc.get_nfe())}




def lookup_traffic(engine, n_pts):
    """profiles/traffic.json: {"<engine>:<points per launch>": {"bytes_per_launch": ..., "source": "<ncu summary>"}}."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(path) as f:
            entry = json.load(f).get('%s:%d' % (engine, n_pts))
    except (OSError, ValueError):
        entry = None
    if not entry:
        return None, 'no ncu capture for %s at %d points per launch under profiles/' % (engine, n_pts)
    return float(entry['bytes_per_launch']), entry.get('source')


def guarded(fn):
    """The extra configs must never take the headline line down with them."""
    try:
        return fn()
    except Exception as exc:                                   # noqa: BLE001
        return {'error': '%s: %s' % (type(exc).__name__, str(exc)[:300])}


def _max_over_ranks(values, dev, world):
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def time_reconstruct_config(model, workload, rank, world, dev, cnf_init, steps=3, warmup=3):
    """Device-resident `reconstruct` over this rank's shard of another BASELINE workload (3 warm-up + 3 timed steps,
    barrier + CUDA events, max over ranks)."""
    _, x, y, e, kwargs, (B, T, N, P, Tq) = make_inputs(workload, rank, cnf_init)
    x, y, e = x.to(dev), y.to(dev), e.to(dev)
    kw = dict(kwargs)
    if 'timestamps' in kw:
        kw['timestamps'] = kw['timestamps'].to(dev)
    for _ in range(warmup):
        model.reconstruct(x, y=y, e=e, **kw)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        model.reconstruct(x, y=y, e=e, **kw)
    ev1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks([ev0.elapsed_time(ev1)], dev, world)[0] / steps
    return {'ms_per_step': round(ms, 3), 'value': world * B * Tq * P / (ms * 1e-3), 'unit': UNIT, 'steps': steps,
            'warmup': warmup, 'config': workload_config(workload, world, cnf_init, model.get_nfe())}


def time_training_config(rank, world, dev, steps=3, warmup=3, B=8, T=5, N=1024):
    """BASELINE configs[4] per-GPU shape: one training step = CaSPR.forward (train mode) -> loss (train_utils.py:148-166)
    -> loss.backward() (hand-written encoder backward, CUDA adjoint solves) -> ONE flat NCCL all-reduce of the 16.26 M
    gradient elements (no pack / unpack: `.grad`s are views of the flat buffer) -> Adam (train.py:135) -> broadcast of
    rank 0's MovingBatchNorm statistics.  Weights are re-loaded before every step so the solvers see the same dynamics."""
    from caspr_b200 import _lib
    from caspr_b200.models import CaSPR
    from caspr_b200.sharding import FlatGradients, broadcast_moving_batchnorm
    from caspr_b200.synth import synthetic_state_dict, synthetic_sequences
    sd = synthetic_state_dict(0, cnf_init='vigorous')
    x, nocs = synthetic_sequences(B, T, N, seed=200 + rank)
    x, nocs = x.to(dev), nocs.to(dev)
    e = torch.randn(B * T, N, 3, generator=torch.Generator().manual_seed(rank)).to(dev)
    model = CaSPR().to(dev).train()
    flat = FlatGradients(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    phase = [0.0] * 5

    def loss_fn(nll, tl1):
        return 0.01 * nll.sum(2).mean() + 100.0 * tl1[:, :, :, :4].mean()

    def step(timed):
        model.load_state_dict(sd)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        ev[0].record()
        flat.zero()
        loss = loss_fn(*model(x, nocs, e=e))
        ev[1].record()
        loss.backward()
        ev[2].record()
        if world > 1:
            flat.allreduce(average=True)
        ev[3].record()
        opt.step()
        ev[4].record()
        if world > 1:
            broadcast_moving_batchnorm(model)
        ev[5].record()
        torch.cuda.synchronize()
        if timed:
            for i in range(5):
                phase[i] += ev[i].elapsed_time(ev[i + 1])
        return float(loss.detach())

    for _ in range(warmup):
        step(False)
    n0 = _lib.lib.caspr_launch_count()
    losses = [step(True) for _ in range(steps)]
    launches = _lib.lib.caspr_launch_count() - n0
    # the all-reduce event pair of a rank that finished its backward early also times its wait for the slowest rank:
    # the MIN over ranks is the collective itself (last arrival -> done), the MAX includes the skew
    tot, ar_max, neg_ar_min = _max_over_ranks([sum(phase), phase[2], -phase[2]], dev, world)
    ms, ar_ms, ar_wait_ms = tot / steps, -neg_ar_min / steps, ar_max / steps
    nbytes = flat.flat.numel() * 4
    out = {'ms_per_step': round(ms, 3), 'value': world * B * T * N / (ms * 1e-3), 'unit': 'trained points/s',
           'steps': steps, 'warmup': warmup,
           'config': {'workload': 'airplanes_train_T%d_N%d_B%d' % (T, N, B), 'batch_per_gpu': B, 'global_batch': B * world,
                      'frames': T, 'input_points': N, 'nfe_latent_cnf_forward': [int(v) for v in model.get_nfe()],
                      'optimizer': 'Adam', 'parallelism': 'batch-shard x%d + gradient all-reduce' % world},
           'phase_ms_forward_backward_allreduce_adam_bufsync': [round(v / steps, 3) for v in phase],
           'loss': losses[-1], 'gradient_bytes': nbytes, 'gpu_launches_per_step': int(launches // steps),
           'allreduce_ms': round(ar_ms, 3) if world > 1 else None,
           'allreduce_ms_incl_skew_max_rank': round(ar_wait_ms, 3) if world > 1 else None,
           # NCCL bus bandwidth of a ring/tree all-reduce: 2 (n-1)/n x bytes / time; NVLink 5 offers 900 GB/s per direction
           'allreduce_busbw_gbs': round(2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9, 1) if world > 1 else None}
    return out


# ----------------------------------------------------------------------------------- ours
def run_ours(args, rank, world, local_rank):
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (there is no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from caspr_b200.build import build_library
    build_library()
    from caspr_b200 import _lib
    from caspr_b200 import ops
    from caspr_b200.models import CaSPR
    from caspr_b200.models.cnf import SequentialFlow
    lib = _lib.lib

    if args.engine == 'tc':
        SequentialFlow.engine = ops.CNF_TC_FP16X3
    elif args.engine == 'simt':
        SequentialFlow.engine = ops.CNF_SIMT_FP32
    engine_name = 'tc_fp16x3' if SequentialFlow.engine == ops.CNF_TC_FP16X3 else 'simt_fp32'

    sd, x, y, e, kwargs, (B, T, N, P, Tq) = make_inputs(args.workload, rank, args.cnf_init)
    model = CaSPR().to(dev).eval()
    model.load_state_dict(sd)
    if args.no_cuda_graph:
        model.encoder.use_cuda_graph = False
    x_pin = x.pin_memory()
    x_dev, y_dev, e_dev = x.to(dev), y.to(dev), e.to(dev)
    kw_dev = dict(kwargs)
    if 'timestamps' in kw_dev:
        kw_dev['timestamps'] = kw_dev['timestamps'].to(dev)
    pts_per_step = B * Tq * P

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def step_resident():
        return model.reconstruct(x_dev, y=y_dev, e=e_dev, **kw_dev)

    def step_e2e():
        xd = x_pin.to(dev, non_blocking=True)
        _, _, xr, _ = model.reconstruct(xd, **kw_dev)           # base samples drawn on the CPU like the reference
        return xr.cpu()                                        # device -> host read of the result

    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()

    # ---- timed region 1: device-resident throughput
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
    lib.caspr_profile_enable(1)
    launches0 = lib.caspr_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    launches = lib.caspr_launch_count() - launches0
    lib.caspr_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    nfe = [int(v) for v in model.get_nfe()]
    prof = {}
    for name, kid in (('cnf_mid_layer_kernel', 0), ('cnf_tc_layer_kernel', 1), ('cnf_fused_eval_kernel', 4)):
        tot, cnt = ctypes.c_double(), ctypes.c_longlong()
        lib.caspr_profile_read(kid, ctypes.byref(tot), ctypes.byref(cnt))
        prof[name] = (tot.value, cnt.value)

    # ---- timed region 2: end to end through the public API with host buffers
    step_e2e()
    barrier()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    f0.record()
    for _ in range(args.steps):
        step_e2e()
    f1.record()
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), wall_ms)        # host-side RNG / copies count too

    # ---- phase breakdown (informational, outside the timed regions): encode / latent ODE / CNF decode
    def phases():
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        evs[0].record()
        z0, _ = model.encode(x_dev)
        evs[1].record()
        times = x_dev[:, :, 0, 3] / kw_dev.get('max_timestamp', 5.0) if 'timestamps' not in kw_dev else \
            kw_dev['timestamps'].view(1, -1).repeat(B, 1)
        z = model.aggregate_and_solve_latent(z0, times.to(z0))
        evs[2].record()
        model.decode(z, P, kw_dev.get('constant_in_time', False), y=y_dev, e=e_dev)
        evs[3].record()
        torch.cuda.synchronize()
        return [evs[i].elapsed_time(evs[i + 1]) for i in range(3)]
    phases()
    phase_ms = [round(v, 3) for v in phases()]

    # ---- BASELINE configs[2..4]: per-GPU shards of the other reconstruction workloads and one training step
    extra = None
    if not args.no_extra_configs and args.workload == DEFAULT_WORKLOAD:
        del x_dev, y_dev, e_dev
        extra = {}
        for name in sorted(WORKLOADS):
            if name != args.workload:
                extra[name] = guarded(lambda: time_reconstruct_config(model, name, rank, world, dev, args.cnf_init))
        model = None
        torch.cuda.empty_cache()
        extra['airplanes_train_T5_N1024_B8'] = guarded(lambda: time_training_config(rank, world, dev))

    t = torch.tensor([ms, ms_e2e, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    if rank != 0:
        return

    peaks = load_peaks()
    value = world * pts_per_step * args.steps / (ms * 1e-3)
    e2e_value = world * pts_per_step * args.steps / (ms_e2e * 1e-3)
    # dominant kernel: the dynamics evaluation of the CNF.  Fused engine (default): ONE launch of cnf_fused_eval_kernel
    # evaluates the whole ODEnet with its forward-mode divergence for all points: ALGORITHMIC work per point-evaluation
    # = 2 x 2 x (3 H + H H + H H + H 3) = 2 109 440 FLOP.  Four-kernel path (CASPR_CNF_FUSED=0): one H x H
    # ConcatSquash layer over {activation, tangent} rows per launch.
    H = 512
    n_pts = B * Tq * P
    eval_flop = 2.0 * 2 * n_pts * (3 * H + H * H + H * H + H * 3)
    layer_flop = 2.0 * 2 * n_pts * H * H
    chunks = 2 if (os.environ.get('CASPR_CNF_PIPELINE_HALVES') == '1' and (n_pts + 63) // 64 >= 296) else 1
    if prof['cnf_fused_eval_kernel'][1] > 0:
        kname, engine_key = 'cnf_fused_eval_kernel (tcgen05 cta_group::2, fp16x3: 3 MMAs per algorithmic MAC)', 'fused'
        flop_per_launch, mma_flop = eval_flop, 3 * 2 * layer_flop
        prof[kname] = prof['cnf_fused_eval_kernel']
    elif prof['cnf_tc_layer_kernel'][1] > 0:
        kname, engine_key = 'gemm_fp16x3_pair_kernel<CnfEpilogue> (tcgen05, 3 fp16 MMAs per algorithmic MAC)', 'tc'
        flop_per_launch, mma_flop = layer_flop / chunks, 3 * layer_flop / chunks
        prof[kname] = prof['cnf_tc_layer_kernel']
    else:
        kname, engine_key = 'cnf_mid_layer_kernel', 'simt'
        flop_per_launch, mma_flop = layer_flop, None

    tot_ms, cnt = prof[kname]
    roofline = None
    if cnt > 0:
        avg_ms = tot_ms / cnt
        achieved = flop_per_launch / (avg_ms * 1e-3) / 1e12
        # DRAM bytes per launch: read from the committed summary of the ncu --set full capture of this kernel on this
        # shape (profiles/traffic.json, written from the ncu report); None when no capture matches
        traffic, traffic_src = lookup_traffic(engine_key, n_pts // chunks)
        roofline = {'bound': 'tensor', 'kernel': kname, 'achieved': achieved, 'peak': peaks['tf_sustained'],
                    'unit': 'TFLOP/s', 'frac': achieved / peaks['tf_sustained'], 'traffic': traffic,
                    'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)',
                    'traffic_source': traffic_src,
                    'peak_source': peaks['source'] + ' bf16 dense sustained', 'launches': cnt,
                    'avg_launch_ms': avg_ms, 'share_of_step': tot_ms / (ms / args.steps) / args.steps,
                    'flop_per_launch': flop_per_launch, 'mma_flop_per_launch': mma_flop,
                    'note': 'achieved = algorithmic fp32-grade FLOP / CUDA-event time; the fp16x3 split issues 3x '
                            'that many tensor-core FLOP, so frac <= 1/3 by construction'}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle.caspr_oracle import CasprOracle
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        per_seq_y = y.shape[0] // B
        oracle = CasprOracle(sd)
        dt = cpu_reference_step(oracle, x[:1], y[:per_seq_y], e[:Tq], kwargs)
        cpu_baseline = {'value': Tq * P / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                        'sample': '1 of %d sequences (B=1,T=%d,N=%d,P=%d), 1 run, %.1f s, nfe %s'
                                  % (B, T, N, P, dt, oracle.get_nfe())}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.workload, world, args.cnf_init, nfe),
            'detail': {'engine': engine_name, 'phase_ms_encode_latent_decode': phase_ms,
                       'l2': 'working set per dynamics evaluation (>1 GB of activations) exceeds the 126 MB L2; no flush'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(x.numel() * 4 + y.numel() * 4),
                    'd2h_bytes_per_step': int(B * Tq * P * 3 * 4), 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
            'configs': extra}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit('launch multi-GPU runs with: python -m torch.distributed.run --nnodes=1 '
                             '--nproc-per-node %d --master-addr 127.0.0.1 bench.py --gpus %d ...' % (args.gpus, args.gpus))
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
