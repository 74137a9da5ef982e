"""Probe: does one training-mode forward + backward of the product model run (fp32 params, bf16 params, bf16 autocast)?"""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
import torch
import vrcoc

def run(tag, phi, B, param_dtype, autocast):
    torch.manual_seed(0)
    m = vrcoc.EfficientVRNet(4, 9, phi).cuda().train()
    if param_dtype is not None:
        m = m.to(param_dtype)
    dt = param_dtype or torch.float32
    x = torch.randn(B, 3, 512, 512, device="cuda", dtype=dt)
    r = torch.rand(B, 4, 512, 512, device="cuda", dtype=dt)
    try:
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx = torch.autocast("cuda", dtype=autocast) if autocast else torch.autocast("cuda", enabled=False)
            with ctx:
                det, seg = m(x, r)
                loss = sum(d.float().square().mean() for d in det) + seg.float().square().mean()
            loss.backward()
            torch.cuda.synchronize()
            dt_ms = (time.perf_counter() - t0) * 1e3
        ng = sum(1 for p in m.parameters() if p.grad is not None)
        nt = sum(1 for p in m.parameters())
        fin = all(torch.isfinite(p.grad.float()).all().item() for p in m.parameters() if p.grad is not None)
        print(f"[{tag}] ok: loss {loss.item():.4f}  {dt_ms:.1f} ms/step  grads {ng}/{nt} finite={fin}  out dtype {seg.dtype}", flush=True)
    except Exception as e:
        print(f"[{tag}] FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
        traceback.print_exc(limit=6)

run("fp32 nano B2", "nano", 2, None, None)
run("bf16 params nano B2", "nano", 2, torch.bfloat16, None)
run("bf16 params l B4", "l", 4, torch.bfloat16, None)
run("fp32 params + bf16 autocast nano B2", "nano", 2, None, torch.bfloat16)
run("fp32 params + fp16 autocast nano B2", "nano", 2, None, torch.float16)
