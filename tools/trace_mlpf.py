"""Per-CTA phase timeline of the fused channel-MLP kernel: python tools/trace_mlpf.py C HIDDEN H"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch
from vrcoc import ops
from vrcoc._lib import lib

C, O, H = (int(v) for v in sys.argv[1:4])
B, dev = 8, "cuda"
x = torch.randn(B, C, H, H, device=dev).bfloat16()
w1 = (torch.randn(O, C, device=dev) / C ** 0.5).bfloat16()
w2 = (torch.randn(C, O, device=dev) / O ** 0.5).bfloat16()
b1, b2, ls = torch.zeros(O, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
sums = ops.sample_sums_of(x)
osum = ops.new_sample_sums(B, dev)
fn = lambda: ops.mlp_fused_fwd(x, sums, gamma, beta, 1e-5, w1, b1, w2, b2, ls, osum)
for _ in range(3):
    fn()
torch.cuda.synchronize()
tr = torch.zeros(4096 * 8 * 8, device=dev, dtype=torch.int64)
lib.vrcoc_debug_set_trace(tr.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_()
fn()
torch.cuda.synchronize()
lib.vrcoc_debug_set_trace(None)
t = tr.view(-1, 8)[:4096].cpu()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
rel = (t[:, :6] - t0).double() / 1e3
print(f"C={C} hidden={O} H={H}: CTAs {len(t)}  kernel span {rel[:, 5].max():.1f} us")
names = ["start", "setup", "X_normed", "acc1(0)", "epi_done", "exit"]
d_ = rel[:, 1:] - rel[:, :-1]
for i in range(5):
    print(f"  {names[i]:>9s} -> {names[i + 1]:<9s} mean {d_[:, i].mean():7.2f} us   p90 {d_[:, i].quantile(0.9):7.2f}   max {d_[:, i].max():7.2f}")
print(f"  final epilogue, warp 0: waiting for acc2 {t[:, 6].double().mean() / 1e3:.2f} us, draining {t[:, 7].double().mean() / 1e3:.2f} us")
print(f"  CTA lifetime mean {(rel[:, 5] - rel[:, 0]).mean():.2f} us; start times: p50 {rel[:, 0].median():.1f} us, max {rel[:, 0].max():.1f} us")
# VRCOC_MF_TRACE=1 build (VRCOC_LIB=.../libvrcoc_trace.so): barrier wait cycles of the MMA thread and of epilogue warp 0
w = tr.view(-1, 8)[4096:4096 + 4096].cpu().double()
w = w[w[:, 3] > 0]
if len(w):
    mhz = 1965.0
    m = w.mean(0) / mhz
    nh = O // 128
    print(f"  MMA thread ({len(w)} CTAs, us per CTA): weights {m[0]:.2f}  acc1_empty {m[1]:.2f}  h_full {m[2]:.2f}  loop {m[3]:.2f}"
          f"  -> issuing/other {m[3] - m[0] - m[1] - m[2]:.2f};  per chunk {m[3] / nh:.2f} us")
    print(f"  epilogue warp 0: waiting acc1_full {m[4]:.2f}  h_empty {m[5]:.2f}  hidden loop {m[6]:.2f}  -> GELU work {m[6] - m[4] - m[5]:.2f}"
          f" ({(m[6] - m[4] - m[5]) / nh:.2f} per chunk)")
# per-chunk stamps of the C > 128 variant (VRCOC_MF_TRACE build): A = first-GEMM issuer, B = second-GEMM issuer, E = epilogue warp 0
st = tr.view(-1)[4096 * 16:4096 * 16 + 4096 * 128].view(-1, 16, 8).cpu().double()
st = st[st[:, 0, 4] > 0]
if len(st):
    t00 = st[:, 0, 4].view(-1, 1, 1)                      # chunk 0's accumulator seen by the epilogue
    rel = ((st - t00) / 1965.0).mean(0)
    print("  per-chunk timeline (us after the epilogue first sees acc1; mean over CTAs):")
    print("   chunk | A: issue start  A: issued | B: H seen  B: issued | E: acc1 seen  E: GELU done")
    for j in range(O // 128):
        print(f"   {j:5d} | {rel[j, 0]:14.2f} {rel[j, 1]:10.2f} | {rel[j, 2]:9.2f} {rel[j, 3]:10.2f} | {rel[j, 4]:12.2f} {rel[j, 5]:13.2f}")
