"""Which ATen ops (library glue, not vrcoc kernels) does one eager forward launch, and from where?  python tools/torch_glue.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200")); sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
m = bench.build_model("l", torch.bfloat16, torch.device("cuda"))
x, r = bench.synth_batch(8, 1, torch.bfloat16)
x, r = x.cuda(), r.cuda()
with torch.no_grad():
    for _ in range(3):
        m(x, r)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True,
                 experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
        m(x, r)
        torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type.name == "CPU" and e.name.startswith("aten::") and e.cpu_children == [] or False:
        pass
for e in prof.events():
    if e.name.startswith("aten::") and e.device_time_total > 0 and not any(c.name.startswith("aten::") and c.device_time_total > 0 for c in e.cpu_children):
        st = [s for s in (e.stack or []) if "vrcoc/" in s]
        key = (e.name, st[0].split("vrcoc/")[-1][:70] if st else (e.stack[0][-60:] if e.stack else "?"))
        a = rows.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += e.device_time_total
for (n, s), (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{n:28s} n={c:3d} {t:8.1f} us  {s}")
