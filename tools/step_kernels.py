"""Kernel census of one eager C2 training step (development tool): name, launches, total device time -- from torch.profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import garment_pattern_estimation_b200 as g
from garment_pattern_estimation_b200.parallel import FlatAdam, FlatDataParallel

dev = torch.device('cuda:0')
B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 2048)
dc, nc, lc = bench.att_configs(5)
torch.manual_seed(1)
model = g.GarmentSegmentPattern3D(dc, nc, lc).to(dev).train()
wrapper = FlatDataParallel(model, device_ids=[dev], auto_reduce=False)
opt = FlatAdam(wrapper, lr=2e-3)
x, gt = bench.synthetic_batch(B, N, seed=1)
x = x.to(dev); gt = {k: v.to(dev) for k, v in gt.items()}


def step():
    out = wrapper(x)
    loss, _, _ = model.loss(out, gt)
    loss.backward()
    wrapper.sum_gradients()
    opt.step(zero_grad=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, 'device_time_total', None)
    if t is None:
        t = getattr(e, 'cuda_time_total', 0)
    if e.device_type.name == 'CUDA' or (t and e.count and 'void' in e.key or 'nt::' in e.key or 'Memset' in e.key or 'Memcpy' in e.key):
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
ours = sum(r[0] for r in rows if 'nt::' in r[2])
print('total device time %.3f ms over %d launches; nt:: kernels %.3f ms (%.1f %%), other %.3f ms over %d launches' % (
    tot / 1e3, sum(r[1] for r in rows), ours / 1e3, 100 * ours / tot, (tot - ours) / 1e3, sum(r[1] for r in rows if 'nt::' not in r[2])))
for t, c, k in rows[:60]:
    print('%9.1f us %5d  %s' % (t, c, k[:150]))

each = os.environ.get('NT_CENSUS_EACH')
if each:
    print('---- every launch matching %r, in launch order' % each)
    evs = [e for e in prof.events() if e.device_type.name == 'CUDA' and each in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    for e in evs:
        print('%9.1f us  %s' % (e.time_range.elapsed_us(), e.name[:90]))
