"""torch.profiler pass over a few training steps of bench.py's configs[3] setup: GPU-busy time vs wall time, top kernels."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

args = types.SimpleNamespace(phi="l", img=512, no_train_graph=True)
dev = torch.device("cuda", 0)
step, net, model, kind, nparam = bench.train_setup(args, dev, 1, int(os.environ.get("B", "16")))
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
ka = prof.key_averages()
tot_cuda = sum(k.self_device_time_total for k in ka) / 2e3
print(f"GPU busy per step: {tot_cuda:.1f} ms")
rows = sorted(ka, key=lambda k: -k.self_device_time_total)[:45]
for k in rows:
    print(f"{k.self_device_time_total / 2e3:9.2f} ms  n={k.count // 2:5d}  {k.key[:110]}")
print("--- CPU side")
rows = sorted(ka, key=lambda k: -k.self_cpu_time_total)[:25]
for k in rows:
    print(f"{k.self_cpu_time_total / 2e3:9.2f} ms  n={k.count // 2:5d}  {k.key[:110]}")
