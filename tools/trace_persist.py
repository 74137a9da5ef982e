"""Per-CTA phase timeline of the persistent tcgen05 projection kernels: python tools/trace_persist.py C O H [g|res]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch
from vrcoc import ops
from vrcoc._lib import lib, ACT_GELU, ACT_NONE

C, O, H = (int(v) for v in sys.argv[1:4])
act = ACT_GELU if len(sys.argv) > 4 and sys.argv[4] == "g" else ACT_NONE
res_mode = len(sys.argv) > 4 and sys.argv[4] == "res"
B, dev = 8, "cuda"
x = torch.randn(B, C, H, H, device=dev).bfloat16()
w = (torch.randn(O, C, device=dev) / C ** 0.5).bfloat16()
out = torch.empty(B, O, H, H, device=dev, dtype=torch.bfloat16)
_, sums = ops.channel_sums(x, want_chan=False, want_sample=True)
if res_mode:
    res = torch.randn(B, O, H, H, device=dev).bfloat16()
    d = ops.conv_desc(x, w, out, e_shift=torch.zeros(O, device=dev), post_scale=torch.ones(O, device=dev), res=res, out_sample_sums=ops.new_sample_sums(B, dev))
else:
    d = ops.conv_desc(x, w, out, gn=(sums, torch.ones(C, device=dev), torch.zeros(C, device=dev), 1e-5), e_shift=torch.zeros(O, device=dev), act=act)
for _ in range(3):
    ops.conv_fwd(d)
torch.cuda.synchronize()
tr = torch.zeros(4096 * 8 * 8, device=dev, dtype=torch.int64)
lib.vrcoc_debug_set_trace(tr.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_()
ops.conv_fwd(d)
torch.cuda.synchronize()
lib.vrcoc_debug_set_trace(None)
t = tr.view(-1, 8).cpu()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
rel = (t[:, :6] - t0).double() / 1e3
print(f"C={C} O={O} H={H}: CTAs {len(t)}  kernel span {rel[:, 5].max():.1f} us")
names = ["start", "setup", "A_built", "acc0_full", "epi_done", "exit"]
d_ = rel[:, 1:] - rel[:, :-1]
for i in range(5):
    print(f"  {names[i]:>9s} -> {names[i + 1]:<9s} mean {d_[:, i].mean():7.2f} us   p90 {d_[:, i].quantile(0.9):7.2f}   max {d_[:, i].max():7.2f}")
print(f"  warp 0: waiting for accumulators {t[:, 6].double().mean() / 1e3:.2f} us, draining them {t[:, 7].double().mean() / 1e3:.2f} us (sum over the CTA's tiles)")
print(f"  CTA lifetime mean {(rel[:, 5] - rel[:, 0]).mean():.2f} us; start times: p50 {rel[:, 0].median():.1f} us, max {rel[:, 0].max():.1f} us")
