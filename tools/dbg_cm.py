"""channel-major tcgen05 kernel vs the CUDA-core engine on projection shapes (debug aid): python tools/dbg_cm.py"""
import sys, torch
sys.path.insert(0, 'asy-vrnet_b200')
from vrcoc import ops
from vrcoc._lib import ACT_GELU, ACT_NONE, ACT_RELU
def rel(a, b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
g = torch.Generator().manual_seed(3)
# (B, C, O, H, W, split, act, mode) mode: gn | plain | res
cases = [(2,64,256,32,32,128,ACT_NONE,'gn'), (2,128,1024,32,32,0,ACT_GELU,'gn'), (2,320,512,32,32,256,ACT_NONE,'gn'), (2,320,1280,16,16,0,ACT_GELU,'gn'),
         (2,512,64,32,32,0,ACT_NONE,'res'), (2,128,64,32,32,0,ACT_NONE,'res'), (2,1280,320,32,32,0,ACT_NONE,'res'), (2,2048,512,16,16,0,ACT_NONE,'res'),
         (2,256,192,32,32,0,ACT_RELU,'plain'), (1,4608,512,16,16,0,ACT_RELU,'plain'), (2,64,72,16,24,0,ACT_NONE,'plain'), (3,136,200,8,24,0,ACT_GELU,'res'),
         (2,64,48,5,8,0,ACT_NONE,'gn'), (1,1152,320,32,32,0,ACT_RELU,'plain')]
for (B, C, O, H, W, split, act, mode) in cases:
    x = (torch.randn(B, C, H, W, generator=g)*1.3+0.2).bfloat16().cuda()
    w = (torch.randn(O, C, generator=g)/C**0.5).bfloat16().cuda()
    bias = torch.randn(O, generator=g).cuda()*0.1
    escale = (torch.rand(O, generator=g)+0.5).cuda()
    kw = dict(e_shift=bias, e_scale=escale, act=act)
    if mode == 'gn':
        gamma = (torch.rand(C, generator=g)+0.5).cuda(); beta = (torch.randn(C, generator=g)*0.1).cuda()
        _, sums = ops.channel_sums(x, want_chan=False, want_sample=True)
        kw['gn'] = (sums, gamma, beta, 1e-5)
    res = None
    if mode == 'res':
        res = torch.randn(B, O, H, W, generator=g).bfloat16().cuda()
        kw.update(res=res, post_scale=(torch.rand(O, generator=g)+0.5).cuda())
    outs, stats = [], []
    for eng in (1, 0):
        if split:
            o1 = torch.empty(B, split, H, W, device='cuda'); o2 = torch.empty(B, O-split, H, W, device='cuda', dtype=torch.bfloat16)
        else:
            o1 = torch.empty(B, O, H, W, device='cuda', dtype=torch.bfloat16); o2 = None
        k2 = dict(kw)
        if mode == 'res':
            k2['out_sample_sums'] = ops.new_sample_sums(B, 'cuda')
        ops.conv_fwd(ops.conv_desc(x, w, o1, out2=o2, engine=eng, **k2))
        torch.cuda.synchronize()
        outs.append(torch.cat([o1.float(), o2.float()], 1) if split else o1.float())
        if mode == 'res':
            stats.append(k2['out_sample_sums'].sum(1))
    e = (outs[1]-outs[0]).abs()
    per_ch = e.amax(dim=(0,2,3))
    bad = (per_ch > 0.1*outs[0].abs().max()).nonzero().flatten()
    s = f"  stats rel {rel(stats[1], stats[0]):.2e}" if stats else ""
    print(f"C={C} O={O} {H}x{W} split={split} {mode}: rel={rel(outs[1], outs[0]):.3e} bad channels n={len(bad)} {bad[:6].tolist()}{s}", flush=True)
