import sys, torch
sys.path.insert(0, 'asy-vrnet_b200'); sys.path.insert(0, 'tests')
from vrcoc import ops
from vrcoc._lib import ACT_GELU, ACT_NONE
def rel(a, b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
g = torch.Generator().manual_seed(3)
for (B, C, O, H, split, act) in [(2,64,128,32,0,ACT_NONE),(2,128,128,32,0,ACT_NONE),(2,256,128,32,0,ACT_NONE),(2,320,128,32,0,ACT_NONE),(2,192,128,32,0,ACT_NONE)]:
    x = (torch.randn(B, C, H, H, generator=g)*1.3+0.2).bfloat16().cuda()
    w = (torch.randn(O, C, generator=g)/C**0.5).bfloat16().cuda()
    bias = torch.randn(O, generator=g).cuda()*0.1
    gamma = (torch.rand(C, generator=g)+0.5).cuda(); beta = (torch.randn(C, generator=g)*0.1).cuda()
    _, sums = ops.channel_sums(x, want_chan=False, want_sample=True)
    outs = []
    for eng in (1, 0):
        if split:
            o1 = torch.empty(B, split, H, H, device='cuda'); o2 = torch.empty(B, O-split, H, H, device='cuda', dtype=torch.bfloat16)
        else:
            o1 = torch.empty(B, O, H, H, device='cuda', dtype=torch.bfloat16); o2 = None
        ops.conv_fwd(ops.conv_desc(x, w, o1, gn=(sums, gamma, beta, 1e-5), e_shift=bias, act=act, out2=o2, engine=eng))
        torch.cuda.synchronize()
        outs.append(torch.cat([o1.float(), o2.float()], 1) if split else o1.float())
    e = (outs[1]-outs[0]).abs()
    per_ch = e.amax(dim=(0,2,3))
    bad = (per_ch > 0.1*outs[0].abs().max()).nonzero().flatten()
    pp = e.amax(dim=(0,1)).flatten(); badp = (pp > 0.1*outs[0].abs().max()).nonzero().flatten()
    print('   bad points', len(badp), badp[:6].tolist(), 'of', pp.numel())
    print(f"C={C} O={O} H={H} split={split}: rel={rel(outs[1], outs[0]):.3e}  bad channels: {bad[:8].tolist()}..{bad[-3:].tolist() if len(bad) else ''} n={len(bad)}")
