"""Differentiable restatements (torch CUDA ops) of the fusion-side modules, used ONLY inside `backward` of the
autograd Functions in fusion.py: the forward of those modules always runs on the native kernels; their backward
recomputes the module with the functions below under autograd and returns torch's gradients.  The CoC block
(ClusterBlock / Cluster / Mlp / GroupNorm) does NOT use this file: its backward is native (ops.py).
Replacing this file by native backward kernels is tracked in DESIGN.md ("not yet native").
"""
import torch
import torch.nn.functional as F

from ._lib import ACT_GELU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SILU


def grads(fn, tensors, dy, needs, slots, total):
    """Run fn(*tensors) under autograd on detached copies and place d(out . dy)/d(tensor_i) at position slots[i] of a
    `total`-long tuple (None elsewhere / where not needed)."""
    ins = []
    for t, need in zip(tensors, needs):
        if t is None:
            ins.append(None)
        else:
            ins.append(t.detach().requires_grad_(bool(need) and t.is_floating_point() and t.numel() > 0))
    with torch.enable_grad():
        out = fn(*ins)
    want = [t for t in ins if t is not None and t.requires_grad]
    got = torch.autograd.grad(out, want, dy.to(out.dtype), allow_unused=True) if want else ()
    res = [None] * total
    it = iter(got)
    for t, slot in zip(ins, slots):
        if t is not None and t.requires_grad:
            g = next(it)
            res[slot] = g if g is not None else torch.zeros_like(t)
        elif t is not None and t.numel() == 0 and t.is_floating_point():
            res[slot] = None
    return tuple(res)


def _act(y, act):
    if act in (ACT_RELU, "relu"):
        return F.relu(y)
    if act in (ACT_SILU, "silu"):
        return y * torch.sigmoid(y)
    if act in (ACT_LRELU, "lrelu"):
        return F.leaky_relu(y, 0.1)
    if act == ACT_GELU:
        return F.gelu(y)
    return y


def _bn(x, w, b, rm, rv, eps, training):
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
    else:
        mean, var = rm.to(x.dtype), rv.to(x.dtype)
    inv = torch.rsqrt(var + eps)
    return (x - mean.view(1, -1, 1, 1)) * (inv * w.to(x.dtype)).view(1, -1, 1, 1) + b.to(x.dtype).view(1, -1, 1, 1)


def conv_act(x, w, b, stride, pad, extra, act, e_scale):
    if extra is not None:
        e = extra if extra.dim() == 4 else extra.unsqueeze(0).expand(x.shape[0], -1, -1, -1)
        x = torch.cat([x, e.to(x.dtype)], dim=1)
    y = F.conv2d(x, w.to(x.dtype), None, stride=stride, padding=pad)
    if e_scale is not None:
        y = y * e_scale.to(y.dtype).view(1, -1, 1, 1)
    if b is not None:
        y = y + b.to(y.dtype).view(1, -1, 1, 1)
    return _act(y, act)


def base_conv(x, w, b, g, h, rm, rv, stride, pad, act, eps, training):
    y = F.conv2d(x, w.to(x.dtype), None if b is None else b.to(x.dtype), stride=stride, padding=pad)
    return _act(_bn(y, g, h, rm, rv, eps, training), act)


def _shuffle(x, groups=2):
    b, c, hh, ww = x.shape
    if c % groups:
        return x
    return x.reshape(b, groups, c // groups, hh, ww).transpose(1, 2).reshape(b, c, hh, ww)


def shuffle_attention(x, cw, cb, sw, sb, gw, gb, G):
    b, c, hh, ww = x.shape
    q = c // (2 * G)
    xg = x.reshape(b, G, 2, q, hh, ww)
    x0, x1 = xg[:, :, 0], xg[:, :, 1]
    v = lambda p: p.to(x.dtype).reshape(1, 1, q, 1, 1)
    xc = x0 * torch.sigmoid(v(cw) * x0.mean(dim=(3, 4), keepdim=True) + v(cb))
    mu = x1.mean(dim=(3, 4), keepdim=True)
    var = x1.var(dim=(3, 4), unbiased=False, keepdim=True)
    xs = x1 * torch.sigmoid(v(sw) * ((x1 - mu) * torch.rsqrt(var + 1e-5) * v(gw) + v(gb)) + v(sb))
    return _shuffle(torch.stack([xc, xs], dim=2).reshape(b, c, hh, ww), 2)


def eca(x, w):
    k = w.shape[-1]
    m = x.mean(dim=(2, 3))
    a = F.conv1d(m.unsqueeze(1), w.to(x.dtype).view(1, 1, k), padding=(k - 1) // 2).squeeze(1)
    return x * torch.sigmoid(a).unsqueeze(-1).unsqueeze(-1)


def image_enhance(image, radar, w, g1, b1, g2, b2, rm1, rv1, rm2, rv2, eps1, eps2, t1, t2):
    k = base_conv(radar.to(image.dtype), w, None, g1, b1, rm1, rv1, 1, 1, "relu", eps1, t1)
    lo, hi = k.min(), k.max()
    y = (1 + (k - lo) / (hi - lo)) * image
    return _bn(y, g2, b2, rm2, rv2, eps2, t2)


def radar_enhance(image, radar, w, g1, b1, g2, b2, eca_w, cw, cb, sw, sb, gw, gb, rm1, rv1, rm2, rv2, eps1, eps2, t1, t2,
                  initial, G):
    image = image.to(radar.dtype)
    ia = image if initial else shuffle_attention(image, cw, cb, sw, sb, gw, gb, G)
    z = eca(_shuffle(torch.cat([ia, radar], dim=1), 2), eca_w)
    u = base_conv(z, w, None, g1, b1, rm1, rv1, 1, 0, "relu", eps1, t1)
    return _bn(u + radar, g2, b2, rm2, rv2, eps2, t2)
