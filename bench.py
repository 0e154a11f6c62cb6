"""Benchmark of the loss hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one synthetic batch, exactly as the
training loop runs it (reference train.py:117-128): scale_pyramid ->
reconstruct_pyramid -> TukraUncertaintyLoss forward -> backward, through the
drop-in classes (i.e. through the C ABI of libusl.so).

Workload at N=1: BASELINE.json configs[1] -- bayesian uncertainty loss, batch
16 synthetic 256x512 stereo pairs, 4 scales.  At N>1 every rank runs the same
per-GPU batch on its own shard (weak scaling, global batch 16*N) and the raw
term sums are all-reduced over NCCL inside the forward.

`value`  : whole-job Mpix/s, inputs resident in HBM, each step replayed from a
           CUDA graph (the step is ~100 us of GPU work; eager Python launch
           overhead would otherwise be what is measured -- the eager number
           is reported as `eager_value`).  Between steps the inputs rotate over
           several sets so nothing is served from L2.
`e2e`    : the same metric through the public API with HOST (pinned) inputs:
           every step copies its stereo pair and predictions to the device
           (copy stream, double buffered), replays the captured step and reads
           the two losses back.
`roofline`: the dominant kernel -- the column-marching fused loss kernel, which
           produces the loss sums AND the gradients in one pass (its four
           per-scale launches run concurrently and are timed together, alone,
           with CUDA events on the launching stream) -- algorithmic bytes of
           loss forward + backward (127.5 B/pixel) over
           its duration, against the measured HBM peak of MEASURED_PEAKS.json.
           `achieved_onepass` is the same with the bytes a one-pass kernel must
           really move (read 40 + write 16 B per scale-pixel = 74.4 B/pixel).
`cpu_baseline`: the oracle port (same ATen op sequence as the reference's CPU
           path) on the host cores, bounded sample.
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
    'c4': (8, 512, 1024, 'l1', 0.3),
}
PYRAMID_FACTOR = 1.0 + 0.25 + 0.0625 + 0.015625          # 1.328125
BYTES_FWD = 40.0 * PYRAMID_FACTOR        # images 24 + prediction 16
BYTES_BWD = 56.0 * PYRAMID_FACTOR        # + grad prediction 16
BYTES_PYR = 24.0 + 24.0 * (PYRAMID_FACTOR - 1.0)
BYTES_STEP = BYTES_FWD + BYTES_BWD + BYTES_PYR            # 159.375 B/pixel


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
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from uncertainty_model_b200 import functional as K
    from uncertainty_model_b200.distributed import shard_loss
    from uncertainty_model_b200.train import loss as L
    from uncertainty_model_b200.train import utils as U

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    b, h, w, lt, scale = WORKLOADS[args.workload]
    pixels = b * h * w
    fn = L.TukraUncertaintyLoss(**loss_config(lt)).to(dev)
    if world > 1:
        shard_loss(fn)

    nsets = args.sets
    host = [synth(b, h, w, scale, 1000 * rank + s) for s in range(nsets)]
    sets = [(st.to(dev), [p.to(dev).requires_grad_(True) for p in pr])
            for st, pr in host]

    def step(stereo, preds):
        for p in preds:
            p.grad = None
        pyr = U.scale_pyramid(stereo, 4)
        rec = U.reconstruct_pyramid(preds, pyr)
        dl, el = fn(pyr, preds, rec, 0, None)
        (dl + el).backward()
        return dl, el

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager) --------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(*sets[i % nsets])
    torch.cuda.synchronize()

    # ---- graph capture, one graph per input set ---------------------------
    graphs = None
    if not args.no_graph:
        try:
            graphs, graph_outs = [], []
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for s in range(nsets):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        graph_outs.append(step(*sets[s]))
                    graphs.append(g)
            torch.cuda.current_stream().wait_stream(side)
            for g in graphs:
                g.replay()
            torch.cuda.synchronize()
        except Exception as e:         # fall back to eager timing
            if rank == 0:
                print(f'graph capture failed ({e}); timing eagerly',
                      file=sys.stderr)
            graphs = None

    def timed(run_one, k):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            run_one(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if graphs is not None:
        for i in range(args.warmup):
            graphs[i % nsets].replay()
        ms_total = timed(lambda i: graphs[i % nsets].replay(), args.steps)
    else:
        ms_total = timed(lambda i: step(*sets[i % nsets]), args.steps)
    ms_step = ms_total / args.steps
    value = world * pixels / (ms_step * 1e-3) / 1e6

    # ---- eager (no graph) number, for the record --------------------------
    ms_eager = timed(lambda i: step(*sets[i % nsets]), args.steps) / args.steps

    # ---- e2e: host inputs, H2D + step + D2H of the losses every step ------
    # The step is the captured graph of the public-API calls (the way a user
    # removes Python launch overhead); its inputs are the graph's static
    # tensors, filled from pinned host memory on a copy stream, so the copy of
    # step i+1 overlaps the kernels of step i.  Every step pays its own H2D
    # copy and its own D2H read inside the timed region.
    pinned = [(st.pin_memory(), [p.pin_memory() for p in pr])
              for st, pr in host]
    out_host = torch.empty(2, dtype=torch.float32).pin_memory()
    h2d = (host[0][0].numel() + sum(p.numel() for p in host[0][1])) * 4
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(nsets)]
    consumed = [torch.cuda.Event() for _ in range(nsets)]

    def e2e_step(i):
        k = i % nsets
        st, pr = pinned[k]
        dst, dpr = sets[k]
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])          # buffer free again
            dst.copy_(st, non_blocking=True)
            with torch.no_grad():
                for a, c in zip(dpr, pr):
                    a.copy_(c, non_blocking=True)
            copied[k].record(copy_stream)
        main.wait_event(copied[k])
        if graphs is not None:
            graphs[k].replay()
            dl, el = graph_outs[k]
        else:
            dl, el = step(dst, dpr)
        consumed[k].record(main)
        out_host[0:1].copy_(dl.detach().reshape(1), non_blocking=True)
        out_host[1:2].copy_(el.detach().reshape(1), non_blocking=True)

    for ev in consumed:
        ev.record(torch.cuda.current_stream(dev))
    for i in range(3):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = world * pixels / (ms_e2e * 1e-3) / 1e6
    # (sampled over all three timed regions: graph replay, eager, end to end)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing (dominant kernel for the roofline) -------------
    kern = kernel_times(K, U, fn, sets, dev, max(10, min(args.steps, 50)))

    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    # the fused column kernel does forward and backward of the loss in one pass
    dom = 'loss_fused_main'
    achieved = (BYTES_FWD + BYTES_BWD) * pixels / (kern[dom] * 1e-3) / 1e9
    achieved_onepass = 56.0 * PYRAMID_FACTOR * pixels / (kern[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get('col_kernel_bytes_per_launch')
        except Exception:
            traffic = None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.no_cpu:
        print(json.dumps({'ms_per_step': ms_step, 'kernels_ms': kern}))
        return
    cpu_times = cpu_step_seconds(b, h, w, lt, scale, 2, warm=1)
    cpu_ms = 1e3 * sum(cpu_times) / len(cpu_times)
    print(json.dumps({
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': config_name(args.workload, b, h, w, lt),
                   'global_batch': b * world,
                   'parallelism': f'batch-sharded dp{world}',
                   'launch': 'cuda-graph replay' if graphs is not None
                   else 'eager',
                   'l2': f'inputs rotate over {nsets} sets '
                         f'({nsets * h2d / 1e6:.0f} MB) > 126 MB L2'},
        'eager_value': world * pixels / (ms_eager * 1e-3) / 1e6,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 8, 'ms_per_step': ms_e2e},
        # pyramid, 4 column kernels + scatter + reduce + combine in forward;
        # 4 + 1 launches in backward (they return at once: unit upstream grads)
        'gpu_launches': 13 * args.steps,
        'kernels_ms': kern,
        'roofline': {'bound': 'hbm',
                     'kernel': 'col_kernel (fused loss fwd+bwd, 4 concurrent '
                               'per-scale launches)',
                     'achieved': achieved,
                     'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'peak_kind': peak_kind, 'traffic': traffic,
                     'achieved_onepass': achieved_onepass,
                     'ms': kern[dom],
                     'note': 'issue/shared-memory bound, not HBM bound: see '
                             'DESIGN.md section 5',
                     'step_frac': BYTES_STEP * pixels / (ms_step * 1e-3) / 1e9
                     / peak},
        'cpu_baseline': {'value': pixels / (cpu_ms * 1e-3) / 1e6,
                         'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'2 steps of the full batch {b} x {h}x{w}'
                                   f' after 1 warm-up ({cpu_ms:.0f} ms/step)'},
        'clocks': clocks,
    }))


def kernel_times(K, U, fn, sets, dev, reps):
    """Average duration of each kernel of the step, launched alone, measured
    with CUDA events on the launching stream, inputs rotating over the sets
    (> L2 in total)."""
    st = fn._settings()
    nsets = len(sets)
    prepared = []
    for stereo, preds in sets:
        pyr = U.scale_pyramid(stereo, 4)
        cfgs, fsc, bsc = [], [], []
        grads = []
        for i in range(4):
            bb, _, hh, ww = preds[i].shape
            coefs = st.coefs(i, bb * fn.grad_world_size * hh * ww)
            cfgs.append(K.make_config(st.terms(), st, coefs))
            pd = preds[i].detach()
            g = torch.empty_like(pd)
            grads.append(g)
            fsc.append(K.make_scale(pyr[i], pd[:, 0:2], pd[:, 2:4],
                                    shape=(bb, hh, ww)))
            ws = torch.empty(bb * 2 * hh * ww * 4, dtype=torch.float32,
                             device=dev)
            grads.append(ws)
            bsc.append(K.make_scale(pyr[i], pd[:, 0:2], pd[:, 2:4],
                                    shape=(bb, hh, ww), grad_disp=g[:, 0:2],
                                    grad_unc=g[:, 2:4], scatter_ws=ws))
        prepared.append((stereo, pyr, cfgs, fsc, bsc, grads))
    one = torch.ones((), device=dev)

    def t(fnc):
        """Mean device time of fnc(i): one CUDA graph per input set, each
        replay between two events on the launching stream (eager calls put the
        host's launch gaps between the four concurrent per-scale launches into
        the number); eager if the call cannot be captured."""
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

    out = {}
    out['pyramid'] = t(lambda i: U.scale_pyramid(prepared[i % nsets][0], 4))
    out['loss_fwd'] = t(lambda i: K.loss_forward(
        prepared[i % nsets][2], prepared[i % nsets][3], dev))
    out['loss_bwd_scatter'] = t(lambda i: K.loss_backward(
        prepared[i % nsets][2], prepared[i % nsets][4], one, one, dev, 1))
    out['loss_fused_main'] = t(lambda i: K.loss_backward(
        prepared[i % nsets][2], prepared[i % nsets][4], one, one, dev, 2))
    # the two halves of the one-pass launch sequence, each alone: the fused
    # kernels (they also leave the scatter inputs behind), then the scatter
    def half(i, flag):
        cfgs, _, bsc = prepared[i % nsets][2], None, prepared[i % nsets][4]
        arr = (K._array(K.UslLossConfig, cfgs), K._array(K.UslLossScale, bsc))
        K.check(K.lib().usl_loss_grad(arr[0], arr[1], 4, None, None, None, flag,
                                      torch.cuda.current_stream().cuda_stream),
                'usl_loss_grad')
    try:
        out['onepass_fused'] = t(lambda i: half(i, K.GRAD_NO_SCATTER))
        out['onepass_scatter'] = t(lambda i: half(i, K.GRAD_ONLY_SCATTER))
    except Exception:
        out['onepass_fused'] = out['onepass_scatter'] = None
    # the training step's path: sums + gradients in one pass (scatter kernel,
    # marching kernel in GRAD mode, reduce, combine)
    try:
        out['loss_onepass'] = t(lambda i: K.loss_forward(
            prepared[i % nsets][2], prepared[i % nsets][4], dev,
            with_grad=True))
    except Exception as e:      # not eligible for this workload
        out['loss_onepass'] = None
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--sets', type=int, default=4)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu', action='store_true',
                    help='(tuning runs) skip the CPU baseline, print timings only')
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
