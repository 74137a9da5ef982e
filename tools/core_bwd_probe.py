"""Cluster-core backward alone at a live geometry (for ncu / timing): python tools/core_bwd_probe.py [S1|S2|S3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch
from vrcoc import ops

CFG = {"S1": (128, 8, 4), "S2": (64, 4, 4), "S3": (32, 2, 8), "S4": (16, 1, 8)}
name = sys.argv[1] if len(sys.argv) > 1 else "S1"
H, fold, heads = CFG[name]
B, D, dev = 8, 32, "cuda"
ED = heads * D
feat = torch.randn(B, ED, H, H, device=dev)
value = torch.randn(B, ED, H, H, device=dev).bfloat16()
a, b_ = torch.tensor([1.3], device=dev), torch.tensor([-0.2], device=dev)
feat.requires_grad_(True); value.requires_grad_(True); a.requires_grad_(True); b_.requires_grad_(True)
out = ops.ClusterCoreFn.apply(feat, value, a, b_, heads, fold, fold, 2, 2)
g = torch.randn_like(out)
for _ in range(3):
    out.backward(g, retain_graph=True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    out.backward(g, retain_graph=True)
e.record(); torch.cuda.synchronize()
print(name, "core backward (incl. autograd glue):", s.elapsed_time(e) / 10 * 1e3, "us")
