#!/usr/bin/env python
"""Headline benchmark: point-clouds/sec of one full TRAINING STEP (H2D -> forward -> 4-term loss -> backward ->
gradient all-reduce -> Adam) of the attention model on synthetic clouds (BASELINE.json configs[1] = C2:
models/att architecture, N = 2048 points, batch 32 per GPU, k = 5), plus the roofline of the dominant kernel and the
oracle's CPU timing on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2|C3|C4|C5] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` = clouds/s with the inputs already resident in HBM, device-timed with CUDA events,
max over ranks; `e2e` = the same step through the public module API with inputs in pinned HOST memory (H2D every step, loss
read back every step).  The headline line is C2 under weak scaling (every rank keeps 32 clouds).

The other BASELINE.json configurations ride along in the same line under `extras` (skip with --no-extras) or become the headline
with --config:
  C3  training step at a FIXED global batch of 64 clouds split over the ranks (strong scaling; 8 clouds per GPU at 8 ranks)
  C4  EdgeConv encoder forward + backward, B=16, N=10 000, k=16 (kNN stress; one GPU)
  C5  full-model inference (eval mode, shipped checkpoint when tests/golden/_ckpt/att_state.pt exists), B=128 split over the
      ranks, N in {1024, 2048, 4096, 8192}
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(points=2048, batch_per_gpu=32, k=5)           # BASELINE.json configs[1] (C2)
CPU_SAMPLE_CLOUDS = 8                                          # cpu_baseline sample inside the b200 arm (B reduced to 8)
C3_GLOBAL_BATCH = 64
SEED_INIT = 916143406                                          # models/att/att.yaml:147
FP32_LANES_PER_SM, SMS = 128, 148


def att_configs(k):
    """(data_config, nn_config, loss_config) of the attention model: the values of models/att/att.yaml."""
    from garment_pattern_estimation_b200.configs import ATT_DATA_CONFIG, ATT_NN_CONFIG
    nc = dict(ATT_NN_CONFIG)
    nc['k_neighbors'] = k
    lc = {'loss_components': ['shape', 'loop', 'rotation', 'translation'], 'quality_components': [],
          'loop_loss_weight': 1., 'panel_origin_invariant_loss': False, 'panel_order_inariant_loss': False}
    return dict(ATT_DATA_CONFIG), nc, lc


def synthetic_ground_truth(B, seed, n_panels=23, panel_len=14):
    """Synthetic GT of SURVEY.md section 8d / A.4: N(0,1) targets, num_edges in {0, 3..14}, zero rows past num_edges."""
    g = torch.Generator().manual_seed(seed)
    ne = torch.randint(2, panel_len + 1, (B, n_panels), generator=g)
    ne = torch.where(ne < 3, torch.zeros_like(ne), ne)
    outl = torch.randn(B, n_panels, panel_len, 4, generator=g)
    live = torch.arange(panel_len)[None, None, :] < ne[..., None]
    outl = outl * live[..., None]
    return {'outlines': outl, 'rotations': torch.randn(B, n_panels, 4, generator=g),
            'translations': torch.randn(B, n_panels, 3, generator=g), 'num_edges': ne, 'num_panels': (ne > 0).sum(-1)}


def synthetic_batch(B, N, seed):
    """Positions ~ N(0,1) (the reference standardises its inputs, nn/data/transforms.py:35-49) + GT of SURVEY 8d.  Generated
    here (not by the oracle): the measured arm of this benchmark never imports oracle/."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, N, 3, generator=g), synthetic_ground_truth(B, seed=seed + 1)


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (pure-PyTorch restatement of the reference path; the reference itself cannot be imported on the
# GPU box -- torch_geometric / torch_cluster / sparsemax are not installable and /root/reference is absent there)
# ------------------------------------------------------------------------------------------------------------
def cpu_train_steps(steps, warmup, k, clouds, threads=None, budget_s=None):
    """The oracle (restatement of nn/nets.py + nn/net_blocks.py + the 4 active loss terms, stock Python loops) training on the host
    cores: `clouds` per step.  With `budget_s` the number of timed steps is cut so that the run stays inside the budget (the cut is
    reported).  Returns (clouds/s, seconds per step, threads, timed steps)."""
    from oracle import model as om
    from oracle import thirdparty as tp
    cores = threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    tp.KNN_THREADS = cores
    dc, nc, lc = att_configs(k)
    torch.manual_seed(SEED_INIT)
    model = om.OracleSegmentPattern3D(dc, nc, lc).train()          # the CPU arm is the one place that executes oracle/
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    x, gt = synthetic_batch(clouds, WORKLOAD['points'], seed=1234)
    times, t_start = [], time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = model(x)
        loss, _ = om.main_losses(out, gt)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and i >= warmup + 2 and (time.perf_counter() - t_start) + dt > budget_s:
            break
    sec = sum(times) / len(times)
    return clouds / sec, sec, cores, len(times)


def cpu_best_threads(k):
    """All host threads is not always the fastest setting for these small per-cloud GEMMs: probe the full core count and two
    smaller pools with one quick 2-cloud step each."""
    total = os.cpu_count() or 1
    cands = sorted({total, max(1, total // 2), min(total, 32)}, reverse=True)
    if len(cands) == 1:
        return cands[0]
    probe = {t: cpu_train_steps(1, 1, k, clouds=2, threads=t)[0] for t in cands}
    return max(probe, key=probe.get)


def run_reference_arm(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path (the oracle port -- the reference is Python and its
    third-party operators cannot be installed) at the SAME batch as the b200 arm (32 clouds x 2048 points per
    step), the requested number of steps unless that would exceed ~2.5 minutes of host time (then fewer, stated in `sample`)."""
    if rank != 0:
        return
    k, clouds = WORKLOAD['k'], WORKLOAD['batch_per_gpu']
    threads = cpu_best_threads(k)
    cps, sec, cores, timed = cpu_train_steps(args.steps, max(1, min(args.warmup, 3)), k, clouds=clouds, threads=threads, budget_s=150.0)
    sample = '{} clouds x {} pts per step (the full C2 batch), {} timed steps of {} requested after warm-up, {:.2f} s/step; oracle ' \
             'port of nn/nets.py + nn/net_blocks.py + the 4 loss terms with the reference\'s own Python loops, torch CPU fp32, {} ' \
             'threads'.format(clouds, WORKLOAD['points'], timed, args.steps, sec, cores)
    line = {
        'impl': 'reference', 'metric': 'point-clouds/sec (fwd+bwd+optimizer, training step)', 'value': cps,
        'unit': 'clouds/s', 'n_gpus': args.gpus, 'steps': timed, 'warmup': max(1, min(args.warmup, 3)), 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(1),
        'cpu_baseline': {'value': cps, 'unit': 'clouds/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': cps, 'unit': 'clouds/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def encoder_roofline(groups, ms_per_step, B, N, hbm_peak):
    """BASELINE.json's second metric, 'kNN+EdgeConv HBM GB/s vs peak', in SURVEY.md section 8d's accounting: the COMPULSORY bytes
    of a fully fused encoder (read the layer input once, write its output once: (12 + 600) N bytes for EdgeConv-1 and 1200 N for
    EdgeConv-2 per cloud forward; the backward is counted as twice that) over the measured time of every kNN / EdgeConv / per-point
    MLP kernel of the step.  SURVEY F9 predicts ~1 % by construction: this path is not HBM-bound under compulsory-byte accounting;
    the per-kernel `roofline` object (algorithmic bytes of the kernels as built) is the operative figure."""
    names = [n for n in groups if n.startswith(('nt_knn', 'nt_gemm', 'nt_edge', 'nt_bn', 'nt_maxmin', 'nt_linear_bn'))]
    ms = sum(groups[n] for n in names)
    compulsory = 3.0 * B * N * (12 + 600 + 1200)
    gbs = compulsory / (ms * 1e-3) / 1e9 if ms > 0 else None
    return {'kernels': 'all nt_knn / nt_gemm_* / nt_edge_* / nt_bn_* / nt_maxmin_finish / nt_linear_bn_bwd launches of the step',
            'compulsory_bytes_per_step': compulsory, 'ms_per_step': ms, 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': (gbs / hbm_peak) if gbs else None, 'share_of_step': (ms / ms_per_step) if ms_per_step else None,
            'note': 'compulsory bytes of a perfectly fused encoder (SURVEY 8d) / measured encoder kernel time; ~1 % by '
                    'construction (SURVEY F9: the fused path is compute-bound), reported because BASELINE.json asks for it'}


def workload_config(n_gpus, strong=False):
    if strong:
        return {'workload': 'C3: attention model (models/att NN config) training step, N=2048 pts/cloud, k=5, FIXED global batch {} '
                            '({} clouds per GPU), Adam lr 2e-3, random init seed 916143406'.format(C3_GLOBAL_BATCH, C3_GLOBAL_BATCH // n_gpus),
                'global_batch': C3_GLOBAL_BATCH, 'points': WORKLOAD['points'], 'parallelism': 'dp{}'.format(n_gpus),
                'l2': 'per-step activations exceed the 126 MB L2 and 4 distinct input batches rotate; no flush'}
    return {'workload': 'C2: attention model (models/att NN config) training step, N=2048 pts/cloud, k=5, '
                        'batch 32 clouds per GPU, Adam lr 2e-3, random init seed 916143406',
            'global_batch': WORKLOAD['batch_per_gpu'] * n_gpus, 'points': WORKLOAD['points'],
            'parallelism': 'dp{}'.format(n_gpus),
            'l2': 'per-step activations (~1.9 GB) exceed the 126 MB L2 and 4 distinct input batches rotate; no flush'}


# ------------------------------------------------------------------------------------------------------------
# C4 / C5
# ------------------------------------------------------------------------------------------------------------
def _peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def run_c4(dev, timed, args, as_extra=False):
    """BASELINE.json configs[3]: EdgeConv encoder (models/att NN config with k = 16) forward + backward on 16 clouds of 10 000
    points, one GPU.  Reports clouds/s, the per-kernel-group device times and the two rooflines that matter at this shape: the
    150-d kNN (direct-form-equivalent FP32 work) and the dominant row-GEMM group (algorithmic HBM bytes)."""
    from garment_pattern_estimation_b200 import net_blocks as nb, ops
    from garment_pattern_estimation_b200.configs import ATT_NN_CONFIG
    B, N, k = 16, 10000, 16
    cfg = dict(ATT_NN_CONFIG)
    cfg['k_neighbors'] = k
    torch.manual_seed(SEED_INIT)
    enc = nb.EdgeConvFeatures(250, cfg).to(dev).train()
    pos = [torch.randn(B, N, 3, generator=torch.Generator().manual_seed(77 + i)).to(dev) for i in range(2)]
    gout = torch.randn(B * N, 153, generator=torch.Generator().manual_seed(5)).to(dev)

    def step(i):
        for p in enc.parameters():
            p.grad = None
        _, feats, _ = enc(pos[i % 2], False)
        feats.backward(gout)

    steps = max(3, min(args.steps, 10))
    for i in range(3):
        step(i)
    torch.cuda.reset_peak_memory_stats()
    ms_total, launches = timed(step, steps)
    ms = ms_total / steps
    ops.EVENT_SINK, ops.FLOP_SINK, ops.BYTES_SINK = {}, {}, {}
    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    groups = {n: sum(s.elapsed_time(e) for s, e in evs) / 2.0 for n, evs in ops.EVENT_SINK.items()}
    counts = {n: len(evs) / 2.0 for n, evs in ops.EVENT_SINK.items()}
    gbytes = {n: v / 2.0 for n, v in ops.BYTES_SINK.items()}
    ops.EVENT_SINK = ops.FLOP_SINK = ops.BYTES_SINK = None
    peaks = _peaks()
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    dom = max(gbytes, key=lambda n: groups.get(n, 0.0)) if gbytes else None
    roofline = None
    if dom:
        gbs = gbytes[dom] / (groups[dom] * 1e-3) / 1e9
        roofline = {'group': dom, 'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                    'traffic': None, 'launches_per_step': counts[dom], 'ms_per_step': groups[dom],
                    'peak_source': 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'}
    knn_ms = groups.get('nt_knn[D=150]', 0.0)
    fp32_peak = SMS * FP32_LANES_PER_SM * 2 * peaks.get('sm_max_mhz', 1965.0) * 1e6 / 1e12
    knn = {'launch_ms': knn_ms, 'direct_form_tflops': B * N * N * (3 * 150 - 1) / (knn_ms * 1e-3) / 1e12 if knn_ms else None,
           'fp32_alu_peak_tflops': fp32_peak, 'pair_dims_per_s': B * N * N * 150 / (knn_ms * 1e-3) if knn_ms else None}
    res = {'workload': 'C4: EdgeConv encoder (2 x DynamicEdgeConv 200-200-150, k=16, skip connection) forward + backward, '
                       'B=16 clouds x N=10000 points, 1 GPU; 2.56 M edge rows per GEMM',
           'value': B / (ms * 1e-3), 'unit': 'clouds/s', 'ms_per_step': ms, 'steps': steps,
           'peak_memory_GB': torch.cuda.max_memory_allocated() / 1e9, 'roofline': roofline, 'roofline_knn': knn,
           'kernel_ms_per_step': {n: round(v, 4) for n, v in sorted(groups.items(), key=lambda kv: -kv[1])[:12]}}
    if as_extra:
        return res
    return {'metric': 'point-clouds/sec (EdgeConv encoder fwd+bwd)', 'value': res['value'], 'unit': 'clouds/s', 'n_gpus': 1,
            'steps': steps, 'warmup': 3, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': res['workload']}, 'gpu_launches': launches,
            'roofline': roofline, 'roofline_knn': knn, 'kernel_ms_per_step': res['kernel_ms_per_step'],
            'peak_memory_GB': res['peak_memory_GB']}


def run_c5(dev, rank, world, timed, args, as_extra=False):
    """BASELINE.json configs[4]: full-model inference (eval mode: BatchNorm running statistics folded into the GEMMs), B = 128
    clouds split over the ranks, N in {1024, 2048, 4096, 8192}; shipped checkpoint when its extracted copy is present."""
    import garment_pattern_estimation_b200 as g
    dc, nc, lc = att_configs(WORKLOAD['k'])
    torch.manual_seed(SEED_INIT)
    model = g.GarmentSegmentPattern3D(dc, nc, lc)
    ck = os.path.join(ROOT, 'tests', 'golden', '_ckpt', 'att_state.pt')
    weights = 'random init (seed 916143406)'
    if os.path.exists(ck):
        model.load_state_dict(torch.load(ck))
        weights = 'shipped models/att/neural_tailor_panels.pth'
    model.to(dev).eval()
    B_global = 128
    b = B_global // world
    sweep = {}
    steps = max(3, min(args.steps, 10))
    for N in (1024, 2048, 4096, 8192):
        pos = [torch.randn(b, N, 3, generator=torch.Generator().manual_seed(900 + 10 * rank + i)).to(dev) for i in range(2)]

        def step(i):
            with torch.no_grad():
                model(pos[i % 2])

        for i in range(3):
            step(i)
        ms_total, _ = timed(step, steps)
        ms = ms_total / steps
        sweep[str(N)] = {'ms_per_batch': ms, 'clouds_per_s': b * world / (ms * 1e-3)}
        del pos
    res = {'workload': 'C5: attention model inference (eval mode), B=128 clouds split over {} GPU(s), {}'.format(world, weights),
           'unit': 'clouds/s', 'steps': steps, 'sweep': sweep}
    if as_extra:
        return res
    return {'metric': 'point-clouds/sec (inference forward)', 'value': sweep['2048']['clouds_per_s'], 'unit': 'clouds/s',
            'n_gpus': world, 'steps': steps, 'warmup': 3, 'ms_per_step': sweep['2048']['ms_per_batch'], 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': res['workload'], 'points': 2048, 'global_batch': B_global}, 'sweep': sweep}


# ------------------------------------------------------------------------------------------------------------
def finish(world):
    """End of a rank.  With more than one rank the process leaves through os._exit after a device synchronize: tearing the NCCL
    communicator down while CUDA graphs that captured its all-reduce are still alive hung the interpreter at exit (measured: the
    JSON line was printed, then the ranks sat in destroy_process_group until the launcher's timeout)."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of parallel.GraphedTrainStep')
    ap.add_argument('--config', default='C2', choices=['C2', 'C3', 'C4', 'C5'], help='BASELINE.json configuration of the headline line')
    ap.add_argument('--no-extras', action='store_true', help='skip the C3 / C4 / C5 measurements that ride along with the C2 line')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the B200 hot path has no CPU fallback '
                           '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    import garment_pattern_estimation_b200 as g
    from garment_pattern_estimation_b200 import _lib, ops
    from garment_pattern_estimation_b200.parallel import FlatAdam, FlatDataParallel, GraphedTrainStep

    N, k = WORKLOAD['points'], WORKLOAD['k']
    strong = args.config == 'C3'
    if strong and C3_GLOBAL_BATCH % world != 0:
        raise RuntimeError('C3: the global batch of {} clouds is not divisible by {} ranks'.format(C3_GLOBAL_BATCH, world))
    B = C3_GLOBAL_BATCH // world if strong else WORKLOAD['batch_per_gpu']
    dc, nc, lc = att_configs(k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        start.record()
        for i in range(steps):
            fn(i)
        end.record()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - l0
        ms = torch.tensor([start.elapsed_time(end)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item()), launches

    if args.config in ('C4', 'C5'):
        res = run_c4(dev, timed, args) if args.config == 'C4' else run_c5(dev, rank, world, timed, args)
        if rank == 0:
            print(json.dumps(res))
        finish(world)
        return

    class Training:
        """The training step of the attention model at `b` clouds per rank: model, flat data-parallel wrapper, FlatAdam, four
        synthetic batches (pinned host + resident copies) and the CUDA-graph step."""

        def __init__(self, b):
            self.b = b
            torch.manual_seed(SEED_INIT)
            self.model = g.GarmentSegmentPattern3D(dict(dc), dict(nc), dict(lc)).to(dev).train()
            self.wrapper = FlatDataParallel(self.model, device_ids=[dev], auto_reduce=False)
            self.opt = FlatAdam(self.wrapper, lr=2e-3)      # torch.optim.Adam semantics, one kernel on the flat buffers (csrc/train_step.cu)
            self.host, self.resident = [], []
            for i in range(4):      # 4 distinct synthetic batches per rank, in pinned host memory (e2e) and resident copies (value)
                x, gt = synthetic_batch(b, N, seed=1234 + 100 * rank + i)
                hx = x.pin_memory()
                hgt = {kk: v.pin_memory() for kk, v in gt.items()}
                self.host.append((hx, hgt))
                self.resident.append((hx.to(dev), {kk: v.to(dev) for kk, v in hgt.items()}))
            self.h2d_bytes = self.host[0][0].numel() * 4 + sum(v.numel() * v.element_size() for v in self.host[0][1].values())
            # The product's training-step API: the whole step (forward, loss, backward, NCCL all-reduce of the flat gradient
            # buffer, Adam) captured once into a CUDA graph (parallel.GraphedTrainStep) and replayed per batch.
            self.graphed, self.graph_note, self.launches_per_eager_step = None, 'eager launches (--no-graph)', None
            if not args.no_graph:
                l0 = _lib.launch_count()
                self.eager_step(*self.resident[0])
                self.launches_per_eager_step = _lib.launch_count() - l0
                try:
                    self.graphed = GraphedTrainStep(self.wrapper, self.opt, self.resident[0][0], self.resident[0][1], warmup=3)
                    self.graph_note = ('parallel.GraphedTrainStep: forward + loss + backward{} replayed as one CUDA graph'.format(
                        (' + all-reduce + Adam' if world > 1 else ' + Adam') if self.graphed.capture_update else ' (all-reduce + Adam eager)'))
                except Exception as e:  # noqa: BLE001 -- capture problems must not hide the eager number
                    self.graphed, self.graph_note = None, 'eager launches (graph capture failed: {})'.format(str(e)[:120])
                    torch.cuda.synchronize()

        def eager_step(self, x, gt):
            out = self.wrapper(x)
            loss, _, _ = self.model.loss(out, gt)
            loss.backward()
            self.wrapper.sum_gradients()             # one NCCL all-reduce of the flat gradient buffer (no-op on one GPU)
            self.opt.step(zero_grad=True)            # the 1 / world_size average and zero_grad are folded into the Adam kernel
            return loss

        def step(self, x, gt):
            return self.graphed(x, gt) if self.graphed is not None else self.eager_step(x, gt)

        def e2e_step(self, i):
            """host buffers: H2D of the inputs and D2H of the loss every step"""
            hx, hgt = self.host[i % 4]
            if self.graphed is not None:                 # H2D straight into the graph's static input buffers
                loss = self.graphed(hx, hgt)
            else:
                x = hx.to(dev, non_blocking=True)
                gt = {kk: v.to(dev, non_blocking=True) for kk, v in hgt.items()}
                loss = self.eager_step(x, gt)
            return float(loss.item())                    # device -> host read of the step's result

        def measure(self, steps, warmup, e2e=True):
            for i in range(warmup):
                self.step(*self.resident[i % 4])
            ms_total, launches = timed(lambda i: self.step(*self.resident[i % 4]), steps)
            if self.graphed is not None:   # replayed kernel nodes do not pass through the library's host entry points: count =
                launches = self.launches_per_eager_step * steps     # the libnt_b200 launches of one eager step (captured 1:1) x steps
            out = {'ms_per_step': ms_total / steps, 'launches': launches}
            if e2e:
                for i in range(2):
                    self.e2e_step(i)
                e2e_ms, _ = timed(self.e2e_step, steps)
                out['e2e_ms_per_step'] = e2e_ms / steps
            return out

    tr = Training(B)
    model, wrapper, opt, graphed, graph_note = tr.model, tr.wrapper, tr.opt, tr.graphed, tr.graph_note
    resident, eager_step, h2d_bytes = tr.resident, tr.eager_step, tr.h2d_bytes

    # ---- value (inputs resident in HBM) and e2e (host buffers), clocks sampled across both timed regions
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m = tr.measure(args.steps, args.warmup)
    ms_per_step, launches, e2e_ms = m['ms_per_step'], m['launches'], m['e2e_ms_per_step'] * args.steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B / (m['e2e_ms_per_step'] * 1e-3)
    clocks = sampler.stop() if rank == 0 else None          # sampled (100 ms period) across both timed regions

    # ---- per-kernel-group device times (CUDA events on the launching stream) for the roofline of the dominant kernel
    ops.EVENT_SINK = {}
    ops.FLOP_SINK = {}
    ops.BYTES_SINK = {}
    for i in range(3):
        eager_step(*resident[i % 4])          # eager: CUDA events cannot bracket the nodes of a replayed graph
    torch.cuda.synchronize()
    gemm_flops = {name: v / 3.0 for name, v in ops.FLOP_SINK.items()}
    gemm_bytes = {name: v / 3.0 for name, v in ops.BYTES_SINK.items()}
    ops.FLOP_SINK = None
    ops.BYTES_SINK = None
    groups = {name: sum(s.elapsed_time(e) for s, e in evs) / 3.0 for name, evs in ops.EVENT_SINK.items()}
    counts = {name: len(evs) / 3.0 for name, evs in ops.EVENT_SINK.items()}
    ops.EVENT_SINK = None

    # ---- the other BASELINE.json configurations, measured by every rank (collective timing) before rank 0 assembles the line
    extras = None
    if args.config == 'C2' and not args.no_extras:
        extras = {}
        del tr, graphed
        torch.cuda.empty_cache()
        x_steps, x_warm = max(5, min(args.steps, 20)), 3
        if C3_GLOBAL_BATCH % world == 0:
            try:
                t3 = Training(C3_GLOBAL_BATCH // world)
                m3 = t3.measure(x_steps, x_warm)
                extras['C3_strong'] = {
                    'workload': 'C3: attention model training step, N=2048, FIXED global batch {} = {} clouds per GPU x {} GPUs '
                                '(strong scaling)'.format(C3_GLOBAL_BATCH, C3_GLOBAL_BATCH // world, world),
                    'value': C3_GLOBAL_BATCH / (m3['ms_per_step'] * 1e-3), 'unit': 'clouds/s', 'ms_per_step': m3['ms_per_step'],
                    'e2e': C3_GLOBAL_BATCH / (m3['e2e_ms_per_step'] * 1e-3), 'scaling': 'strong', 'steps': x_steps, 'launch_mode': t3.graph_note}
                del t3
            except Exception as e:  # noqa: BLE001
                extras['C3_strong'] = {'error': str(e)[:200]}
            torch.cuda.empty_cache()
        if world == 1:
            try:
                extras['C4'] = run_c4(dev, timed, args, as_extra=True)
            except Exception as e:  # noqa: BLE001
                extras['C4'] = {'error': str(e)[:200]}
            torch.cuda.empty_cache()
        try:
            extras['C5'] = run_c5(dev, rank, world, timed, args, as_extra=True)
        except Exception as e:  # noqa: BLE001
            extras['C5'] = {'error': str(e)[:200]}

    if rank != 0:
        finish(world)
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'

    # ---- roofline of the DOMINANT kernel group of the step (largest share of device time among this library's kernels).
    # The fused row GEMMs stream edge-sized fp32 operands once and write the result once: HBM is the binding roofline
    # (K <= 200, so arithmetic intensity is ~100 flop/B even with the 3-product split).  achieved = algorithmic bytes of
    # the group's launches (operand rows + result rows + weights, counted in ops.gemm_nt / ops.gemm_tn) / their summed
    # CUDA-event time = launch-weighted average of bytes-per-launch / launch-duration.
    sm_clock = (clocks or {}).get('sm_mhz') or peaks.get('sm_max_mhz', 1965.0)
    dom = max(gemm_bytes, key=lambda n: groups.get(n, 0.0)) if gemm_bytes else None
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'dominant_traffic.json')) as f:
            tj = json.load(f)
            traffic = tj.get(dom, {}).get('dram_bytes_per_launch') if isinstance(tj.get(dom), dict) else None
    except (OSError, ValueError):
        pass
    kernel_names = {'nt_gemm_nt[bnrelu_bwd,plain]': 'nt::gemm_nt_tc4_kernel<BNRELU_BWD> (persistent streaming tcgen05 TF32x3 data-gradient GEMM '
                                                    'fused with the BatchNorm/ReLU backward: raw A k-blocks and aux boxes by TMA tensor-map '
                                                    'loads, A operand split into tensor memory, result boxes by TMA stores; the two '
                                                    'per-point launches of the group run nt::gemm_nt_tc_kernel)',
                    'nt_gemm_nt[relu_stats,plain]': 'nt::gemm_nt_tc4_kernel<RELU_STATS> (same engine, forward Linear + ReLU + BatchNorm statistics)',
                    'nt_gemm_nt[relu_maxmin,plain]': 'nt::gemm_nt_tc4_kernel<RELU_MAXMIN> (same engine, forward Linear + ReLU + max/min over the k edges)',
                    'nt_gemm_tn_centered': 'nt::gemm_tn_mn_kernel + mn_reduce_kernel (weight-gradient GEMM with centred operand: rows consumed as '
                                           'MN-major BF16x3 tcgen05 operands, whole-row bulk copies)',
                    'nt_gemm_tn': 'nt::gemm_tn_mn_kernel + mn_reduce_kernel (weight-gradient GEMM, MN-major BF16x3 tcgen05 operands)'}
    roofline = None
    if dom:
        dom_ms, dom_n = groups[dom], max(counts.get(dom, 1.0), 1.0)
        gbs = gemm_bytes[dom] / (dom_ms * 1e-3) / 1e9
        roofline = {
            'kernel': kernel_names.get(dom, dom) + '; {:.0f} launches per step, C2 shape'.format(dom_n),
            'group': dom, 'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
            'traffic': traffic, 'peak_source': peak_src, 'launch_ms': dom_ms / dom_n,
            'algorithmic_bytes_per_launch': gemm_bytes[dom] / dom_n, 'share_of_step': dom_ms / ms_per_step,
            'traffic_source': 'profiles/dominant_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the '
                              'launches of the group in one eager step)',
            'note': 'algorithmic bytes = fp32 operand rows read once + result rows written once + weights; the same group '
                    'reaches {:.1f} TFLOP/s of useful fp32-equivalent math (see roofline_tensor).  What bounds the group is in DESIGN.md '
                    'section 4 (cycle traces profiles/r02_tc4_cycle_trace.txt, ncu summaries under profiles/)'.format(
                        gemm_flops.get(dom, 0.0) / (dom_ms * 1e-3) / 1e12),
        }

    # ---- kNN on the 150-d EdgeConv features: bf16 tcgen05 filter + exact fp32 re-rank (csrc/knn_tc.cu).  Reported as
    # direct-form-equivalent work: the bit-exact chain costs (3C-1) N^2 flop per cloud on the FP32 pipe.
    knn_name = 'nt_knn[D=150]'
    knn_ms = groups.get(knn_name, 0.0) / max(counts.get(knn_name, 1.0), 1.0)      # average launch duration
    feat = 150
    pair_dims = B * N * N * feat
    flops = B * N * N * (3 * feat - 1)
    fp32_peak = SMS * FP32_LANES_PER_SM * 2 * sm_clock * 1e6 / 1e12
    roofline_knn = {
        'kernel': 'nt_knn D=150: knn_tc_prepare + knn_tc_filter (tcgen05 bf16, M128 N128 K16) + knn_tc_rerank (exact fp32 chain) '
                  '+ knn_tc_fallback; {} clouds x {} pts per call'.format(B, N),
        'launch_ms': knn_ms, 'share_of_step': groups.get(knn_name, 0.0) / ms_per_step if ms_per_step else None,
        'direct_form_equivalent': {'tflops': flops / (knn_ms * 1e-3) / 1e12 if knn_ms else None, 'fp32_alu_peak': fp32_peak,
                                   'ratio_to_fp32_alu_peak': (flops / (knn_ms * 1e-3) / 1e12 / fp32_peak) if knn_ms else None,
                                   'pair_dims_per_s': pair_dims / (knn_ms * 1e-3) if knn_ms else None},
        'note': 'results are bit-identical to the sequential fp32 fma chain; the filter does 2*160*N^2 bf16 tensor flops per '
                'cloud instead of 449 N^2 FP32-ALU flops, so the ratio to the FP32-ALU peak may exceed 1',
    }

    # tensor-core GEMM group (tcgen05, TF32x3): useful flops (2*rows*K*n_out, one product -- the 3x split is overhead)
    tc_names = [n for n in groups if n.startswith('nt_gemm_nt') or n.startswith('nt_gemm_tn')]
    tc_ms = sum(groups[n] for n in tc_names)
    tc_flops = sum(gemm_flops.get(n, 0.0) for n in tc_names)
    tensor_peak = peaks.get('bf16_tflops_sustained', 1414.4)
    roofline_tensor = {
        'kernels': 'nt::gemm_nt_tc4_kernel<*> + nt::gemm_nt_tc_kernel<*> + nt::gemm_tn_mn_kernel (all fused row / weight-gradient GEMMs of the step)',
        'bound': 'tensor', 'achieved': tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else None, 'peak': tensor_peak,
        'unit': 'TFLOP/s', 'frac': (tc_flops / (tc_ms * 1e-3) / 1e12 / tensor_peak) if tc_ms else None,
        'ms_per_step': tc_ms, 'useful_gflop_per_step': tc_flops / 1e9,
        'note': 'peak = measured sustained bf16 cuBLAS GEMM (MEASURED_PEAKS.json); these kernels run kind::tf32 (half the '
                'bf16 rate) with a 3-product error-compensated split, so 1/6 of that peak is the precision-imposed ceiling; '
                'they are streaming kernels (K, n_out <= 200) bound by HBM and shared-memory bandwidth, not by the tensor pipe (profiles/)',
    }

    try:
        roofline_encoder = encoder_roofline(groups, ms_per_step, B, N, hbm_peak)
    except Exception as e:  # noqa: BLE001 -- an auxiliary figure must never cost the bench line
        roofline_encoder = {'error': str(e)[:100]}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cps, sec, cores, timed_steps = cpu_train_steps(3, 1, k, clouds=CPU_SAMPLE_CLOUDS, threads=cpu_best_threads(k), budget_s=60.0)
        cpu_baseline = {'value': cps, 'unit': 'clouds/s', 'cores': cores, 'kind': 'port',
                        'sample': '{} clouds x {} pts per step (BASELINE.md section 3: C2 with B reduced to 8), {} timed steps after 1 '
                                  'warm-up ({:.2f} s/step); oracle port of the reference path with its stock Python loops, torch CPU '
                                  'fp32; `bench.py --impl reference` times the full 32-cloud batch'.format(CPU_SAMPLE_CLOUDS, N,
                                                                                                          timed_steps, sec)}

    line = {
        'metric': 'point-clouds/sec (fwd+bwd+optimizer, training step)', 'value': value, 'unit': 'clouds/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(world, strong),
        'e2e': {'value': e2e_value, 'unit': 'clouds/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': launches, 'gpu_launches_per_step': launches / args.steps, 'launch_mode': graph_note,
        'clocks': clocks, 'roofline': roofline, 'roofline_tensor': roofline_tensor, 'roofline_knn': roofline_knn,
        'roofline_encoder': roofline_encoder,
        'cpu_baseline': cpu_baseline,
        'kernel_ms_per_step': {kk: round(v, 4) for kk, v in sorted(groups.items(), key=lambda kv: -kv[1])},
        'extras': extras,
    }
    print(json.dumps(line))
    finish(world)


if __name__ == '__main__':
    main()
