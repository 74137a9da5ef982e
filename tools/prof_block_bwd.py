"""Kernel-level breakdown of one ClusterBlock forward+backward (torch profiler, CUDA time): python tools/prof_block_bwd.py [S1|S3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch
import vrcoc
from torch.profiler import profile, ProfilerActivity

CFG = {"S1": (64, 128, 8, 4, 32, 8), "S2": (128, 64, 4, 4, 32, 8), "S3": (320, 32, 2, 8, 32, 4), "S4": (512, 16, 1, 8, 32, 4)}
name = sys.argv[1] if len(sys.argv) > 1 else "S1"
C, H, fold, heads, hd, ratio = CFG[name]
dt = torch.bfloat16
blk = vrcoc.ClusterBlock(dim=C, mlp_ratio=float(ratio), fold_w=fold, fold_h=fold, heads=heads, head_dim=hd).to("cuda", dt)
x = torch.randn(8, C, H, H, device="cuda").to(dt).requires_grad_(True)
g = torch.randn(8, C, H, H, device="cuda").to(dt)
for _ in range(3):
    blk(x).backward(g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        blk(x).backward(g)
    torch.cuda.synchronize()
print(f"== {name}: 3 x (fwd + bwd)")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
