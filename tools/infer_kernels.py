"""Kernel census of one C5 inference batch (development tool): name, launches, total device time -- from torch.profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import garment_pattern_estimation_b200 as g

dev = torch.device('cuda:0')
B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 2048)
dc, nc, lc = bench.att_configs(5)
torch.manual_seed(1)
model = g.GarmentSegmentPattern3D(dc, nc, lc)
ck = os.path.join(bench.ROOT, 'tests', 'golden', '_ckpt', 'att_state.pt')
if os.path.exists(ck):
    model.load_state_dict(torch.load(ck))
model.to(dev).eval()
x = torch.randn(B, N, 3, device=dev)
for _ in range(3):
    with torch.no_grad():
        model(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    with torch.no_grad():
        model(x)
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
print('B=%d N=%d: total device time %.3f ms over %d launches' % (B, N, tot / 1e3, sum(r[1] for r in rows)))
for t, c, k in rows[:30]:
    print('%9.1f us %5d  %s' % (t, c, k[:140]))
