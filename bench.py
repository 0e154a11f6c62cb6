"""Benchmark of the loss hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one synthetic batch, exactly as the
training loop runs it (reference train.py:117-128): scale_pyramid ->
reconstruct_pyramid -> TukraUncertaintyLoss forward -> backward, through the
drop-in classes (i.e. through the C ABI of libusl.so).  The backward hands the
gradients w.r.t. the predictions on, as it does in the training loop where the
predictions are the decoder's outputs: `torch.autograd.grad(disp_loss +
error_loss, predictions)`.  (`(disp_loss + error_loss).backward()` with the
predictions as LEAF tensors makes autograd clone every gradient into `.grad`
inside a captured graph -- 89 MB of copies, ~19 us a step, that belong to the
benchmark's leaves and not to the path; that variant is measured too and
reported as `leaf_backward_value`.)

Main line: BASELINE.json configs[1] -- bayesian uncertainty loss, batch 16
synthetic 256x512 stereo pairs per GPU, 4 scales (the configuration the metric
is quoted on).  At N>1 every rank runs that per-GPU batch on its own shard
(weak scaling, global batch 16*N) and the raw term sums are all-reduced over
NCCL inside the forward.

`value`  : whole-job Mpix/s, inputs resident in HBM, each step replayed from a
           CUDA graph (the step is ~0.3 ms of GPU work in ~15 launches; eager
           Python launch overhead would otherwise be what is measured -- the
           eager number is reported as `eager_value`).  Between steps the
           inputs rotate over several sets so nothing is served from L2.
`e2e`    : the same metric through the public API with HOST (pinned) inputs:
           every step copies its stereo pair and predictions to the device
           (ONE copy of one contiguous pinned buffer, on a copy stream, double
           buffered), replays the captured step and reads the two losses back.
`roofline`: the fused loss forward + backward = the column-marching kernels
           (sums and gradients in one pass, four per-scale launches) followed
           by the transposed-warp kernels of the consistency terms, timed
           together, alone, with CUDA events on the launching stream:
           algorithmic bytes of loss forward + backward (SURVEY.md section 8d:
           127.5 B/pixel) over that time, against the measured HBM peak of
           MEASURED_PEAKS.json.  `fused_only_frac` is the same over the column
           kernels alone (round 1's definition).
`cpu_baseline`: the oracle port (same ATen op sequence as the reference's CPU
           path) on the host cores, bounded sample.
`extra`  : the other BASELINE configs, each a measured line of its own:
           c3_strong  config 3 as written: FIXED global batch 64 of 192x384
                      split over the N ranks (64/N per GPU), global Mpix/s;
           c4_adversarial (N=1) config 4's step: 512x1024, batch 8,
                      reconstructions materialised, gradient arriving at them
                      from outside (223.125 B/pixel contract);
           c5_sparsification (N=1) config 5: AUSE evaluation at 1024x1280.
"""
import argparse
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

import torch  # noqa: E402

METRIC = 'loss fwd+bwd Mpix/s (4-scale)'
UNIT = 'Mpix/s'
WORKLOADS = {
    # name: (per-GPU batch, H, W, loss_type, disparity scale)
    'c2': (16, 256, 512, 'bayesian', 0.3),
    'c1': (2, 256, 512, 'l1', 0.3),
    'c3': (8, 192, 384, 'l1', 0.3),
    'c3full': (64, 192, 384, 'l1', 0.3),
    'c4': (8, 512, 1024, 'l1', 0.3),
}
C3_GLOBAL_BATCH = 64
PYRAMID_FACTOR = 1.0 + 0.25 + 0.0625 + 0.015625          # 1.328125
BYTES_FWD = 40.0 * PYRAMID_FACTOR        # images 24 + prediction 16
BYTES_BWD = 56.0 * PYRAMID_FACTOR        # + grad prediction 16
BYTES_PYR = 24.0 + 24.0 * (PYRAMID_FACTOR - 1.0)
BYTES_STEP = BYTES_FWD + BYTES_BWD + BYTES_PYR            # 159.375 B/pixel
BYTES_ADV = BYTES_STEP + 48.0 * PYRAMID_FACTOR            # + recon write, grad_recon read


def loss_config(loss_type):
    return dict(wssim_weight=1.0, consistency_weight=1.0,
                smoothness_weight=1.0, adversarial_weight=0.85,
                perceptual_weight=0.05, predictive_error_weight=1.0,
                wssim_alpha=0.85, perceptual_start=5,
                adversarial_loss_type='mse',
                error_loss_config=dict(loss_type=loss_type,
                                       smoothness_weight=0,
                                       consistency_weight=0.5, pooling=False))


def synth(b, h, w, scale, seed):
    g = torch.Generator().manual_seed(seed)
    stereo = torch.rand(b, 6, h, w, generator=g)
    preds = [scale * torch.sigmoid(torch.randn(b, 4, h >> i, w >> i,
                                               generator=g))
             for i in range(4)]
    return stereo, preds


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------
def cpu_step_seconds(b, h, w, loss_type, scale, repeats, warm=1):
    """Oracle port of the reference's CPU path, all host threads."""
    from oracle import loss_port as P
    stereo, preds = synth(b, h, w, scale, 0)
    cfg = loss_config(loss_type)
    times = []
    for i in range(warm + repeats):
        t0 = time.perf_counter()
        P.step(stereo, preds, cfg)
        if i >= warm:
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path
    (its ATen op sequence, restated in oracle/loss_port.py -- the Python
    reference itself cannot travel to the GPU box) on the host cores."""
    if rank != 0:
        return
    b, h, w, lt, scale = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample: shrink the per-step batch until the run fits ~150 s
    probe = cpu_step_seconds(2, h, w, lt, scale, 1, warm=1)[0] / 2.0
    bs = b
    while bs > 1 and probe * bs * (args.steps + args.warmup) > 150.0:
        bs //= 2
    times = cpu_step_seconds(bs, h, w, lt, scale, args.steps,
                             warm=args.warmup)
    ms = 1e3 * sum(times) / len(times)
    value = bs * h * w / (ms * 1e-3) / 1e6
    sample = f'{args.steps} steps of batch {bs} x {h}x{w} ({lt}), 4 scales'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': config_name(args.workload, b, h, w, lt),
                   'step_batch': bs},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }))


def config_name(key, b, h, w, lt):
    return (f'{key}: {lt} uncertainty loss, batch {b}/GPU synthetic '
            f'{h}x{w} stereo, 4 scales')


# --------------------------------------------------------------------------
class Harness:
    """One workload on one rank: device inputs (rotating sets), the step as
    the training loop runs it, its captured CUDA graphs, timers."""

    def __init__(self, dev, world, rank, b, h, w, lt, scale, nsets,
                 adversarial=False):
        from uncertainty_model_b200.distributed import shard_loss
        from uncertainty_model_b200.train import loss as L
        from uncertainty_model_b200.train import utils as U
        self.U, self.dev, self.world = U, dev, world
        self.shape = (b, h, w)
        self.adversarial = adversarial
        self.fn = L.TukraUncertaintyLoss(**loss_config(lt)).to(dev)
        if world > 1:
            shard_loss(self.fn)
        self.nsets = nsets
        # one contiguous host buffer per set: [stereo | pred0 | .. | pred3]
        self.host = []
        self.sets = []
        for s in range(nsets):
            st, pr = synth(b, h, w, scale, 1000 * rank + s)
            flat = torch.cat([st.reshape(-1)] + [p.reshape(-1) for p in pr])
            self.host.append(flat)
            self.sets.append(self._views(flat.to(dev)))
        self.disc = None
        if adversarial:
            self.disc = FixedGradientDiscriminator(b, h, w).to(dev)
        self.graphs = None
        self.graph_outs = None
        self.grads = [None] * nsets

    def _views(self, flat):
        b, h, w = self.shape
        n = b * 6 * h * w
        stereo = flat[:n].view(b, 6, h, w)
        preds, off = [], n
        for i in range(4):
            m = b * 4 * (h >> i) * (w >> i)
            preds.append(flat[off:off + m].view(b, 4, h >> i, w >> i)
                         .requires_grad_(True))
            off += m
        return flat, stereo, preds

    @property
    def h2d_bytes(self):
        return self.host[0].numel() * 4

    def step(self, k, leaf_backward=False):
        _, stereo, preds = self.sets[k % self.nsets]
        U = self.U
        pyr = U.scale_pyramid(stereo, 4)
        rec = U.reconstruct_pyramid(preds, pyr)
        # (adversarial = BASELINE config 4: the loss hands the reconstructions
        #  to the discriminator and a gradient arrives at them from it)
        dl, el = self.fn(pyr, preds, rec, 0, self.disc)
        if leaf_backward:
            for p in preds:
                p.grad = None
            (dl + el).backward()
        else:
            # the gradients w.r.t. the predictions, handed on (module docstring)
            self.grads[k % self.nsets] = torch.autograd.grad(dl + el, preds)
        return dl, el

    def capture(self):
        try:
            graphs, outs = [], []
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                # Fresh leaves for the predictions, made on the capture stream:
                # autograd binds a leaf's gradient accumulation to the stream
                # that was current when the leaf was first used, and the eager
                # warm-up steps ran on the default stream -- the graph would
                # carry a cross-stream synchronisation after every backward that
                # a training loop (where the predictions are not leaves) does
                # not have.
                self.sets = [(flat, stereo,
                              [p.detach().requires_grad_(True) for p in preds])
                             for flat, stereo, preds in self.sets]
                for s in range(self.nsets):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        outs.append(self.step(s))
                    graphs.append(g)
                self.leaf_graphs = []
                for s in range(self.nsets):
                    self.step(s, leaf_backward=True)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        self.step(s, leaf_backward=True)
                    self.leaf_graphs.append(g)
            torch.cuda.current_stream().wait_stream(side)
            for g in graphs:
                g.replay()
            torch.cuda.synchronize()
            self.graphs, self.graph_outs = graphs, outs
        except Exception as e:         # fall back to eager timing
            print(f'graph capture failed ({e}); timing eagerly',
                  file=sys.stderr)
            self.graphs = None

    def run(self, i):
        if self.graphs is not None:
            self.graphs[i % self.nsets].replay()
        else:
            self.step(i)


class FixedGradientDiscriminator(torch.nn.Module):
    """Stand-in for the reference's discriminator (a conv net outside the
    path, model/discriminator.py): its verdict is a fixed linear functional of
    the reconstruction pyramid, so the gradient arriving at the
    reconstructions is a fixed random tensor times a scalar."""

    def __init__(self, b, h, w):
        super().__init__()
        g = torch.Generator().manual_seed(99)
        for i in range(4):
            self.register_buffer(
                f'w{i}', (torch.rand(1, 6, h >> i, w >> i, generator=g) - 0.5)
                * (4.0 ** i / (h * w)))

    def forward(self, pyramid):
        return sum((p * getattr(self, f'w{i}')).sum(dim=(1, 2, 3))
                   for i, p in enumerate(pyramid))[:, None]


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def timed(run_one, k, world, dev):
    """K calls between two events on the launching stream, barrier +
    synchronize on both sides, max over ranks."""
    barrier(world)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k):
        run_one(i)
    e1.record()
    barrier(world)
    ms = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def launches_per_step(hz):
    """Kernel launches of one step, COUNTED: the library's own counter around
    one eager step (its launches are what a captured graph replays) plus the
    kernels torch itself launches in the step (profiler-free: the autograd
    `add` of the two losses' gradients and the `ones` seed = 2; an adversarial
    step adds one accumulation per scale)."""
    from uncertainty_model_b200 import _lib
    L = _lib.lib()
    torch.cuda.synchronize()
    n0 = L.usl_launch_count()
    hz.step(0)
    torch.cuda.synchronize()
    ours = L.usl_launch_count() - n0
    # (an adversarial step adds the stand-in discriminator's own kernels: not
    #  counted, they are not the path's)
    return int(ours), 2


def measure(hz, steps, warmup, world, dev, e2e=True):
    """value / eager / e2e timings of a harness."""
    for i in range(max(warmup, 3)):
        hz.step(i)
    torch.cuda.synchronize()
    ours, theirs = launches_per_step(hz)
    hz.capture()
    for i in range(warmup):
        hz.run(i)
    ms = timed(hz.run, steps, world, dev) / steps
    out = {'ms_per_step': ms, 'launches_per_step': ours + theirs,
           'library_launches_per_step': ours,
           'launch': 'cuda-graph replay' if hz.graphs is not None else 'eager'}
    out['ms_eager'] = timed(hz.step, steps, world, dev) / steps
    if hz.graphs is not None and getattr(hz, 'leaf_graphs', None):
        out['ms_leaf_backward'] = timed(
            lambda i: hz.leaf_graphs[i % hz.nsets].replay(), steps, world,
            dev) / steps
    if not e2e:
        return out
    # ---- e2e: host inputs, H2D + step + D2H of the losses every step ------
    # The step is the captured graph of the public-API calls; its inputs are
    # the graph's static tensors, filled from pinned host memory on a copy
    # stream (one contiguous buffer -> ONE copy per step), so the copy of step
    # i+1 overlaps the kernels of step i.  Every step pays its own H2D copy and
    # its own D2H read inside the timed region.
    nsets = hz.nsets
    pinned = [f.pin_memory() for f in hz.host]
    out_host = torch.empty(2, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(nsets)]
    consumed = [torch.cuda.Event() for _ in range(nsets)]

    def e2e_step(i):
        k = i % nsets
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])          # buffer free again
            with torch.no_grad():
                hz.sets[k][0].copy_(pinned[k], non_blocking=True)
            copied[k].record(copy_stream)
        main.wait_event(copied[k])
        if hz.graphs is not None:
            hz.graphs[k].replay()
            dl, el = hz.graph_outs[k]
        else:
            dl, el = hz.step(k)
        consumed[k].record(main)
        out_host[0:1].copy_(dl.detach().reshape(1), non_blocking=True)
        out_host[1:2].copy_(el.detach().reshape(1), non_blocking=True)

    for ev in consumed:
        ev.record(torch.cuda.current_stream(dev))
    for i in range(3):
        e2e_step(i)
    out['ms_e2e'] = timed(e2e_step, steps, world, dev) / steps
    return out


# --------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    from uncertainty_model_b200 import functional as K

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    b, h, w, lt, scale = WORKLOADS[args.workload]
    pixels = b * h * w
    hz = Harness(dev, world, rank, b, h, w, lt, scale, args.sets)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m = measure(hz, args.steps, args.warmup, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = m['ms_per_step']
    value = world * pixels / (ms_step * 1e-3) / 1e6

    # ---- per-kernel timing (the roofline kernel) ----------------------------
    kern = kernel_times(K, hz, dev, max(10, min(args.steps, 50)))

    # ---- the other BASELINE configs ------------------------------------------
    extra = {}
    if not args.no_extra:
        extra['c3_strong'] = c3_strong(args, dev, world, rank)
        if world == 1:
            extra['c4_adversarial'] = c4_adversarial(args, dev)
            extra['c5_sparsification'] = c5_sparsification(dev)

    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    dom = 'loss_fused_plus_scatter'
    achieved = (BYTES_FWD + BYTES_BWD) * pixels / (kern[dom] * 1e-3) / 1e9
    fused_only = (BYTES_FWD + BYTES_BWD) * pixels / \
        (kern['loss_fused_main'] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get('loss_fused_plus_scatter_bytes')
        except Exception:
            traffic = None
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': config_name(args.workload, b, h, w, lt),
                   'global_batch': b * world,
                   'parallelism': f'batch-sharded dp{world}',
                   'launch': m['launch'],
                   'l2': f'inputs rotate over {args.sets} sets '
                         f'({args.sets * hz.h2d_bytes / 1e6:.0f} MB) > 126 MB L2'},
        'eager_value': world * pixels / (m['ms_eager'] * 1e-3) / 1e6,
        'leaf_backward_value': (world * pixels / (m['ms_leaf_backward'] * 1e-3) / 1e6
                                if 'ms_leaf_backward' in m else None),
        'e2e': {'value': world * pixels / (m['ms_e2e'] * 1e-3) / 1e6,
                'unit': UNIT, 'h2d_bytes_per_step': hz.h2d_bytes,
                'd2h_bytes_per_step': 8, 'ms_per_step': m['ms_e2e'],
                'h2d_copies_per_step': 1,
                'host_binding': getattr(args, 'numa', None)},
        # counted (bench.launches_per_step), not assumed
        'gpu_launches': m['launches_per_step'] * args.steps,
        'launches_per_step': m['launches_per_step'],
        'library_launches_per_step': m['library_launches_per_step'],
        'kernels_ms': kern,
        'roofline': {'bound': 'hbm',
                     'kernel': 'fused loss fwd+bwd: col_kernel x4 scales '
                               '(sums + gradients in one pass) + '
                               'cons_rows_kernel (transposed warp)',
                     'achieved': achieved,
                     'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'peak_kind': peak_kind, 'traffic': traffic,
                     'ms': kern[dom],
                     'fused_only_frac': fused_only / peak,
                     'fused_only_ms': kern['loss_fused_main'],
                     'note': 'issue/latency bound, not HBM bound: see '
                             'DESIGN.md section 5',
                     'step_frac': BYTES_STEP * pixels / (ms_step * 1e-3) / 1e9
                     / peak},
        'clocks': clocks,
        'extra': extra,
    }
    if args.no_cpu:
        line['cpu_baseline'] = None
    else:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_times = cpu_step_seconds(b, h, w, lt, scale, 2, warm=1)
        cpu_ms = 1e3 * sum(cpu_times) / len(cpu_times)
        line['cpu_baseline'] = {
            'value': pixels / (cpu_ms * 1e-3) / 1e6, 'unit': UNIT,
            'cores': cores, 'kind': 'port',
            'sample': f'2 steps of the full batch {b} x {h}x{w} after 1 '
                      f'warm-up ({cpu_ms:.0f} ms/step)'}
    print(json.dumps(line))


def c3_strong(args, dev, world, rank):
    """BASELINE config 3 as written: global batch 64 of 192x384 (l1), FIXED,
    split over the ranks.  Global Mpix/s; the driver's 1/2/4/8 runs give the
    strong-scaling curve (the N=1 run carries the whole batch of 64)."""
    _, h, w, lt, scale = WORKLOADS['c3']
    per = C3_GLOBAL_BATCH // world
    hz = Harness(dev, world, rank, per, h, w, lt, scale, max(2, args.sets // 2))
    m = measure(hz, max(10, args.steps // 2), args.warmup, world, dev,
                e2e=False)
    peak, _ = measured_peak()
    gpix = C3_GLOBAL_BATCH * h * w
    return {'workload': f'c3: l1, global batch {C3_GLOBAL_BATCH} of {h}x{w} '
                        f'split over {world} GPU(s) ({per}/GPU)',
            'scaling': 'strong', 'n_gpus': world,
            'value': gpix / (m['ms_per_step'] * 1e-3) / 1e6, 'unit': UNIT,
            'ms_per_step': m['ms_per_step'],
            'step_frac_per_gpu': BYTES_STEP * (gpix / world) /
            (m['ms_per_step'] * 1e-3) / 1e9 / peak,
            'launches_per_step': m['launches_per_step']}


def c4_adversarial(args, dev):
    """BASELINE config 4: the adversarial step at 512x1024, batch 8 --
    reconstructions materialised, a gradient arriving at them from outside
    (223.125 B/pixel contract).  The discriminator itself is a conv net
    outside the path; its gradient is a fixed random tensor here."""
    b, h, w, lt, scale = WORKLOADS['c4']
    out = {}
    peak, _ = measured_peak()
    for name, adv in (('adversarial', True), ('plain', False)):
        hz = Harness(dev, 1, 0, b, h, w, lt, scale, 2, adversarial=adv)
        m = measure(hz, max(10, args.steps // 4), args.warmup, 1, dev,
                    e2e=False)
        bytes_px = BYTES_ADV if adv else BYTES_STEP
        out[name] = {'value': b * h * w / (m['ms_per_step'] * 1e-3) / 1e6,
                     'unit': UNIT, 'ms_per_step': m['ms_per_step'],
                     'bytes_per_pixel': bytes_px,
                     'step_frac': bytes_px * b * h * w /
                     (m['ms_per_step'] * 1e-3) / 1e9 / peak,
                     'launches_per_step': m['launches_per_step']}
        del hz
        torch.cuda.empty_cache()
    out['workload'] = f'c4: l1, batch {b} of {h}x{w}, 4 scales'
    return out


def c5_sparsification(dev):
    """BASELINE config 5: AUSE evaluation (oracle curve, predicted curve,
    AUSE) on 1024x1280 maps.  64 frames timed in chunks of 16 (events), and
    one real pass over all 512 frames, generated on the device chunk by
    chunk.  Algorithmic bytes: 16 B per pooled element and curve pair."""
    from uncertainty_model_b200.train import sparsification as S
    hh, ww, k = 1024, 1280, 11
    n = (hh - k + 1) * (ww - k + 1)
    chunk = 16
    g = torch.Generator(device=dev).manual_seed(5)

    def maps(frames):
        err = torch.rand(frames, 2, hh, ww, generator=g, device=dev)
        unc = (err + 0.2 * torch.rand(frames, 2, hh, ww, generator=g,
                                      device=dev)).clamp_(0, 1)
        return err, unc

    def evaluate(err, unc):
        oc = S.curve(err, err, device=dev)
        pc = S.curve(err, unc, device=dev)
        return S.ause(oc, pc)

    sets = [maps(chunk) for _ in range(2)]
    for e, u in sets:
        evaluate(e, u)
    torch.cuda.synchronize()
    reps = 4                                   # 4 x 16 = 64 frames
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        evaluate(*sets[i % 2])
    e1.record()
    e1.synchronize()
    ms64 = e0.elapsed_time(e1)
    # one real pass over 512 frames (generation excluded from the time)
    total = 0.0
    acc = 0.0
    for c in range(512 // chunk):
        e, u = maps(chunk)
        torch.cuda.synchronize()
        e0.record()
        a = evaluate(e, u)
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
        acc += float(a)
    peak, _ = measured_peak()
    per_frame = ms64 / (reps * chunk)
    alg = 2 * 2 * n * 16.0                     # two curves x two views
    return {'workload': 'c5: AUSE (oracle + predicted curve) on 1024x1280 '
                        'error/uncertainty maps, kernel 11, 100 steps',
            'frames_per_s': 1e3 / per_frame, 'ms_per_frame': per_frame,
            'frames_timed': reps * chunk,
            'ms_512_frames_extrapolated': per_frame * 512,
            'ms_512_frames_measured': total,
            'mean_ause_512': acc / (512 // chunk),
            'algorithmic_bytes_per_frame': alg,
            'frac_of_hbm_peak': alg / (per_frame * 1e-3) / 1e9 / peak}


def kernel_times(K, hz, dev, reps):
    """Average duration of each kernel group of the step, launched alone,
    measured with CUDA events on the launching stream, inputs rotating over
    the sets (> L2 in total)."""
    U, fn = hz.U, hz.fn
    st = fn._settings()
    nsets = hz.nsets
    prepared = []
    for _, stereo, preds in hz.sets:
        pyr = U.scale_pyramid(stereo, 4)
        cfgs, fsc, bsc = [], [], []
        keep = []
        for i in range(4):
            bb, _, hh, ww = preds[i].shape
            coefs = st.coefs(i, bb * fn.grad_world_size * hh * ww)
            cfgs.append(K.make_config(st.terms(), st, coefs))
            pd = preds[i].detach()
            g = torch.empty_like(pd)
            ws = torch.empty(bb * 2 * hh * ww * 4, dtype=torch.float32,
                             device=dev)
            keep += [g, ws]
            fsc.append(K.make_scale(pyr[i], pd[:, 0:2], pd[:, 2:4],
                                    shape=(bb, hh, ww)))
            bsc.append(K.make_scale(pyr[i], pd[:, 0:2], pd[:, 2:4],
                                    shape=(bb, hh, ww), grad_disp=g[:, 0:2],
                                    grad_unc=g[:, 2:4], scatter_ws=ws))
        arr = (K._array(K.UslLossConfig, cfgs), K._array(K.UslLossScale, bsc))
        prepared.append((stereo, pyr, cfgs, fsc, bsc, keep, arr))
    one = torch.ones((), device=dev)

    def t(fnc):
        """Mean device time of fnc(i): one CUDA graph per input set, each
        replay between two events on the launching stream (eager calls put the
        host's launch gaps between the concurrent per-scale launches into the
        number); eager if the call cannot be captured."""
        for i in range(max(3, nsets)):
            fnc(i)
        torch.cuda.synchronize()
        graphs = []
        try:
            for i in range(nsets):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fnc(i)
                graphs.append(g)
        except Exception:
            graphs = None
            torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        if graphs:
            for g in graphs:
                g.replay()
            torch.cuda.synchronize()
        total = 0.0
        for i in range(reps):       # one launch at a time: "launched alone"
            e0.record()
            if graphs:
                graphs[i % nsets].replay()
            else:
                fnc(i)
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total / reps

    def grad_call(i, flag):
        arr = prepared[i % nsets][6]
        K.check(K.lib().usl_loss_grad(arr[0], arr[1], 4, None, None, None, flag,
                                      torch.cuda.current_stream().cuda_stream),
                'usl_loss_grad')

    out = {}
    out['pyramid'] = t(lambda i: U.scale_pyramid(prepared[i % nsets][0], 4))
    out['loss_fwd'] = t(lambda i: K.loss_forward(
        prepared[i % nsets][2], prepared[i % nsets][3], dev))
    # the two halves of the one-pass launch sequence, each alone: the fused
    # column kernels (they also leave the scatter inputs behind), then the
    # transposed warp of the consistency terms (all scales in one launch here;
    # in the step every scale's scatter runs right behind its own fused kernel)
    out['loss_fused_main'] = t(lambda i: grad_call(i, K.GRAD_NO_SCATTER))
    out['loss_scatter'] = t(lambda i: grad_call(i, K.GRAD_ONLY_SCATTER))
    # the largest scale's alone: the one that is on the step's critical path
    out['loss_scatter_scale0'] = t(
        lambda i: grad_call(i, K.GRAD_ONLY_SCATTER0))
    # ... and as the step issues them: fused + scatter, per scale, concurrent
    out['loss_fused_plus_scatter'] = t(lambda i: grad_call(i, 0))
    # the training step's loss: that plus reduce + combine
    out['loss_onepass'] = t(lambda i: K.loss_forward(
        prepared[i % nsets][2], prepared[i % nsets][4], dev, with_grad=True))
    # round 1's warp-per-row scatter (recomputes the warp), for reference
    out['loss_bwd_scatter_warp_per_row'] = t(lambda i: K.loss_backward(
        prepared[i % nsets][2], prepared[i % nsets][4], one, one, dev, 1))
    return out


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run on the CPUs of the NUMA node the GPU hangs off,
    so that the pinned staging buffers of the end-to-end leg (first touch) lie
    in that node's memory and the H2D copies do not cross the socket
    interconnect.  What a launcher would do with `numactl`; torchrun does not.
    Returns a short description for the JSON line (None if nothing was done)."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = '%04x:%02x:%02x.0' % (props.pci_domain_id, props.pci_bus_id,
                                    props.pci_device_id)
        base = f'/sys/bus/pci/devices/{bdf}'
        with open(f'{base}/numa_node') as f:
            node = int(f.read().strip())
        with open(f'{base}/local_cpulist') as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(','):
            if '-' in part:
                lo, hi = part.split('-')
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if node < 0 or not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return {'numa_node': node, 'cpus': len(cpus)}
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--sets', type=int, default=4)
    ap.add_argument('--no-cpu', action='store_true',
                    help='(tuning runs) skip the CPU baseline')
    ap.add_argument('--no-extra', action='store_true',
                    help='(tuning runs) skip the c3 / c4 / c5 extras')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        args.numa = bind_to_gpu_numa_node(local_rank)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda',
                                                               local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
