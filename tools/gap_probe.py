"""Is the ~190 us hole that the CUPTI timeline shows inside forward_embeddings (between a memset and channel_sums) real?
Captures forward_embeddings alone in a CUDA graph and times replays with events; prints the CUPTI kernel list of one replay beside it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch
import vrcoc
from vrcoc import ops
from torch.profiler import ProfilerActivity, profile

torch.manual_seed(0)
dev = "cuda"
model = vrcoc.EfficientVRNet(num_classes=4, num_seg_classes=9, phi="l").eval().to(dev, torch.bfloat16)
net = next(m for m in model.modules() if type(m).__name__ == "VRCoC")
x = torch.randn(8, 3, 512, 512, device=dev).bfloat16()
r = torch.rand(8, 4, 512, 512, device=dev).bfloat16()


def fwd():
    with ops.sums_arena(8, x.device):
        return net.forward_embeddings(x, r)


with torch.no_grad():
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fwd()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(100):
    g.replay()
e.record()
torch.cuda.synchronize()
print(f"forward_embeddings graph replay: {s.elapsed_time(e) * 10:.1f} us per replay (events, 100 replays)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
ev = sorted(((k.time_range.start, k.time_range.end, k.name) for k in prof.events() if k.device_type == torch.autograd.DeviceType.CUDA),
            key=lambda t: t[0])
t0 = ev[0][0]
print(f"CUPTI: span {ev[-1][1] - t0:.1f} us, sum of durations {sum(b - a for a, b, _ in ev):.1f} us")
for a, b, n in ev:
    print(f"  {a - t0:8.1f} {b - a:7.1f}  {n[:70]}")
