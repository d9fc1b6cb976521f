"""In-pipeline kernel timeline of one bench step (torch.profiler / CUPTI): total kernel time vs wall span, i.e. how
much of the step the GPU idles between launches, and the largest gaps.   python tools/gpu_timeline.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gotennet_b200 as g  # noqa: E402
from gotennet_b200.synthetic import synth_batch  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = g.GotenNetWrapper(cutoff_fn=g.CosineCutoff(bench.CUTOFF), max_num_neighbors=bench.MAX_NBR, activation="swish",
                          **bench.MODEL).to(dev)
z, pos, batch = synth_batch("qm9", 1024, seed=1000)
zd, pd, bd = z.to(dev), pos.to(dev), batch.to(dev)
params = list(model.parameters())


class D:
    pass


def step():
    d = D()
    d.z, d.pos, d.batch = zd, pd, bd
    for p in params:
        p.grad = None
    h, X = model(d)
    (h.sum() + X.pow(2).sum()).backward()


for _ in range(4):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
evs = sorted(evs, key=lambda e: e.time_range.start)
# second step only: kernels after the midpoint gap
ks = [(e.time_range.start, e.time_range.end, e.name) for e in evs]
t0, t1 = ks[0][0], ks[-1][1]
busy = sum(b - a for a, b, _ in ks)
print(f"2 steps: span {(t1 - t0) / 1e3:.2f} ms, kernel busy {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms, kernels {len(ks)}")
gaps = sorted(((ks[i + 1][0] - ks[i][1], ks[i][2][:50], ks[i + 1][2][:50]) for i in range(len(ks) - 1)), reverse=True)
print("largest gaps (us):")
for gp, a, b in gaps[:12]:
    print(f"  {gp:8.1f}  after {a:50s} before {b}")
import collections
hist = collections.Counter()
for gp, _, _ in gaps:
    hist[min(int(gp // 2) * 2, 40)] += 1
print("gap histogram (us bucket: count):", sorted(hist.items()))
small = sum(gp for gp, _, _ in gaps if gp < 50)
print(f"sum of gaps < 50 us: {small / 1e3:.2f} ms")
