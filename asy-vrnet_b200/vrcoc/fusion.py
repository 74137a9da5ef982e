"""Asymmetric vision<->radar fusion modules and their leaf helpers, reference signatures, native kernels.

Mirrors reference backbone/fusion/vr_coc.py:59-80,303-359 (data_normal, shuffle_channels, ImageEnhanceByRadar,
RadarEnhanceByImage), backbone/attention_modules/shuffle_attention.py:8-72, backbone/attention_modules/eca.py:6-22 and
backbone/conv_utils/normal_conv.py:5-52.

Forward (eval AND train mode) runs on the hand-written kernels:
  ImageEnhanceByRadar  = [3x3 implicit-GEMM + BN + ReLU + global min/max] -> [(1 + minmax-normalise) * image -> BN]
  RadarEnhanceByImage  = [channel sums x2] -> [ShuffleAttention gate params + attended means] -> [ECA + prologue table]
                         -> [1x1 GEMM whose prologue applies attention/ECA/shuffle and whose epilogue applies
                             BN + ReLU + radar residual + BN]
Backward (training) is native as well: the big-tensor passes are kernels of csrc/bwd.cu (k x k convolution backward through
im2col / col2im and the tensor-core weight gradient, BatchNorm + activation backward, the min/max-normalised gate of
ImageEnhanceByRadar, the gated prologue of RadarEnhanceByImage); only O(B*C)-sized statistics algebra runs as torch ops on
tiny tensors (`_prologue_backward`, `_bn_coefficients`).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SILU, VrcocError, check, lib)
from .ops import _dt, _f32, _ptr, _stream, conv_desc, conv_fwd


# ------------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------------
def data_normal(origin_data):
    """reference vr_coc.py:59-67 (whole-tensor min/max).  Stand-alone helper kept for API parity; inside
    ImageEnhanceByRadar the min/max are a side output of the projection kernel."""
    d_min = origin_data.min()
    if d_min < 0:
        origin_data += torch.abs(d_min)
        d_min = origin_data.min()
    d_max = origin_data.max()
    return (origin_data - d_min).true_divide(d_max - d_min)


def shuffle_perm(channels, groups=2):
    """out[k] = in[perm[k]] for the channel shuffle of reference vr_coc.py:70-80; identity when C % groups != 0."""
    if channels % groups:
        return list(range(channels))
    cpg = channels // groups
    return [(k % groups) * cpg + k // groups for k in range(channels)]


def shuffle_channels(x, groups=2):
    """reference vr_coc.py:70-80 / neck/coc_fpn_dual.py:120-130 (pure permutation; inside the fused modules it is
    folded into the consumer's channel map instead of being materialised)."""
    batch_size, channels, h, w = x.size()
    if channels % groups:
        return x
    x = x.view(batch_size, groups, channels // groups, h, w)
    return torch.transpose(x, 1, 2).contiguous().view(batch_size, -1, h, w)


def _bn_affine(bn, chan_sums=None, count=None):
    """(scale, shift) fp32 such that BN(x) = x*scale + shift.  eval: running statistics.  train: batch statistics from
    the per-(b,c) sums (biased variance), with the running-stat update of nn.BatchNorm2d (unbiased variance)."""
    def wb():
        return (bn.weight.detach().float() if bn.affine else None, bn.bias.detach().float() if bn.affine else None)
    use_batch = bn.training or not bn.track_running_stats
    if use_batch and isinstance(bn, nn.SyncBatchNorm) and torch.distributed.is_available() and torch.distributed.is_initialized() \
            and torch.distributed.get_world_size() > 1:
        # train.py:356-357 (sync_bn=True, off by default) converts BatchNorm2d to SyncBatchNorm: the native path computes
        # per-replica statistics and must not silently pretend otherwise
        raise VrcocError("SyncBatchNorm inside a native BaseConv / fusion module is not supported (per-replica batch statistics only): "
                         "train with sync_bn=False (the reference default, train.py:51)")
    if use_batch and chan_sums.is_cuda and chan_sums.dtype == torch.float32 and chan_sums.is_contiguous() and \
            (not bn.track_running_stats or (bn.running_mean.dtype == torch.float32 and bn.momentum is not None)):
        # one launch (csrc/bwd.cu: bn_stats_kernel) for what follows below as ~22 library launches on [C]-sized tensors
        B_, C_ = chan_sums.shape[0], chan_sums.shape[1]
        g32, b32 = wb()
        sc = torch.empty(C_, device=chan_sums.device, dtype=torch.float32)
        sh = torch.empty_like(sc)
        upd = bn.training and bn.track_running_stats
        with torch.no_grad():
            check(lib.vrcoc_bn_stats(_ptr(chan_sums), B_, C_, float(count), _ptr(g32), _ptr(b32), float(bn.eps),
                                     float(bn.momentum) if upd else -1.0, _ptr(bn.running_mean) if upd else None,
                                     _ptr(bn.running_var) if upd else None, _ptr(bn.num_batches_tracked) if upd else None,
                                     _ptr(sc), _ptr(sh), None, None, _stream()), "bn_stats")
        return sc, sh
    if use_batch:
        s = chan_sums.double().sum(0)                       # [C,2]
        n = float(count)
        mean = s[:, 0] / n
        var = (s[:, 1] / n - mean * mean).clamp_min(0)
        if bn.training and bn.track_running_stats:
            with torch.no_grad():
                bn.num_batches_tracked += 1
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - mom).add_(mean.to(bn.running_mean.dtype), alpha=mom)
                bn.running_var.mul_(1 - mom).add_((var * (n / max(n - 1, 1))).to(bn.running_var.dtype), alpha=mom)
        mean, var = mean.float(), var.float()
    else:
        return ops.cached(bn, "eval_affine", [bn.weight, bn.bias, bn.running_mean, bn.running_var], lambda: _fold_bn(
            bn.running_mean.detach().float(), bn.running_var.detach().float(), *wb(), bn.eps))
    return _fold_bn(mean, var, *wb(), bn.eps)


def _fold_bn(mean, var, w, b, eps):
    scale = torch.rsqrt(var + eps)
    if w is not None:
        scale = scale * w
    shift = -mean * scale
    if b is not None:
        shift = shift + b
    return scale.contiguous(), shift.contiguous()


SPLIT_WIDE_PROLOGUE = True  # RadarEnhanceByImage with > 576 input channels: prologue as its own pass + TMA-only GEMM (A/B switch)
PATCH_EMBED_KERNEL = True   # 4x4 / stride-4 patch embedding on its own gather + mma.sync kernel (False: the general engine; A/B switch)
ROW_TAPS = True             # 3x3 stride-1 convolutions: horizontal-tap copies + row-shifted TMA boxes (False: full im2col; A/B switch)


def conv2d_native(x, weight, bias=None, stride=1, pad=0, extra=None, extra_bstride=None, act=ACT_NONE,
                  e_scale=None, out_minmax=None, out_dtype=None):
    """Dense (groups=1) k x k convolution on the implicit-GEMM engine.  Differentiable through `_ConvFn`."""
    return _apply(_ConvFn, x, weight, bias, stride, pad, extra, extra_bstride, act, e_scale, out_minmax, out_dtype)


def _conv_launch(x, weight, bias32, stride, pad, extra, extra_bstride, act, e_scale, out_minmax, out_dtype, dil=1, out_sample_sums=None):
    ops._need_cuda(x)
    x = x.contiguous()
    B, C0, H, W = x.shape
    O, Cin, kh, kw = weight.shape
    C1 = Cin - C0
    if (C1 != 0) != (extra is not None):
        raise VrcocError(f"conv: weight expects {Cin} input channels, got {C0} (+ extra: {extra is not None})")
    if extra is not None:
        extra = extra.contiguous()
        if extra.dim() == 3:
            extra_bstride = 0
        if extra.shape[-3] != C1:
            raise VrcocError("conv: extra channel count mismatch")
    Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    out = torch.empty(B, O, Ho, Wo, device=x.device, dtype=out_dtype or x.dtype)
    if (PATCH_EMBED_KERNEL and kh == kw == stride == 4 and pad == 0 and dil == 1 and act == ACT_NONE and e_scale is None and out_minmax is None
            and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and bias32 is not None
            and (extra is None or extra.dtype == torch.bfloat16)
            and lib.vrcoc_patch_embed_supported(_dt(x), C0, C1, H, W, O, 4)):
        # the 4x4 / stride-4 patch embedding of a handful of channels: HBM-bound gather with a small contraction (csrc/patch_embed.cu)
        ebs = 0 if (extra is None or extra.dim() == 3) else (extra_bstride if extra_bstride is not None else C1 * H * W)
        check(lib.vrcoc_patch_embed(_ptr(x), _ptr(extra), ebs, _ptr(weight.detach().contiguous()), _ptr(bias32), _ptr(out),
                                    _ptr(out_sample_sums), _dt(x), B, C0, C1, H, W, O, 4, _stream()), "patch_embed")
        return out
    if (kh == kw and kh > 1 and kh % 2 == 1 and stride == 1 and pad == dil * (kh // 2) and dil >= H and dil >= W and extra is None
            and weight.dtype == torch.bfloat16):
        # dilation >= the map: every off-centre tap reads zero padding only, the convolution IS the 1x1 projection by its centre tap
        # (ASPP rate 18 on the 16x16 map at 512x512 input: 1/9 of the GEMM and no im2col)
        d = conv_desc(x, ops.center_tap(weight), out, e_scale=e_scale, e_shift=bias32, act=act, out_minmax=out_minmax,
                      out_sample_sums=out_sample_sums)
        conv_fwd(d)
        return out
    # k x k convs on the tensor-core path use the tap-major K order (cheap gathers); few-channel convs (the 512x512 ingest)
    # run on the streaming kernel, which takes the PyTorch order
    tapm = kh * kw > 1 and weight.dtype == torch.bfloat16 and not (O <= 8 and Cin * kh * kw <= 64)
    w2 = ops.tap_major(weight) if tapm else weight.detach().reshape(O, -1).contiguous()
    if (ROW_TAPS and tapm and extra is None and x.dtype == torch.bfloat16 and kh == 3 and kw == 3 and stride == 1 and pad == dil
            and dil == 1 and Cin % 64 == 0 and W % 64 == 0 and (H * W) % 8 == 0):
        # row-tap mode (include/vrcoc.h, k_order 2): only the horizontal taps are materialised (3x the input instead of the 9x of
        # the full im2col below); the vertical taps are TMA boxes of that tensor moved by whole rows in the flattened point index.
        # (Moving a box by ONE element faults in NCHW - the innermost box start must be 16-byte aligned, tools/tma_probe.cu - which
        # is why the horizontal taps are copies.)  Only where a row is a whole number of 128-byte lines (dil * W % 64 == 0): boxes
        # that start mid-line cost two L2 requests per row, and on the 16x16 / 32x32 maps that ate the saving (measured: GEMM 34 ->
        # 46 us at 16x16, 39 -> 43 us at 32x32 against 12 / 8 us less im2col); dense taps only (the dilated ASPP branches live on the
        # 16x16 map, where most shifted boxes are zero fill and the full im2col + GEMM was faster: 53 vs 70 us at rate 12).
        cols = torch.empty(B, 3 * Cin, H, W, device=x.device, dtype=x.dtype)
        check(lib.vrcoc_im2col_rows(_ptr(x), _ptr(cols), _dt(x), B, Cin, H, W, 3, dil, _stream()), "im2col_rows")
        d = conv_desc(cols, w2, out, kh=3, kw=1, stride=1, pad=dil, dil=dil, k_order=2, e_scale=e_scale, e_shift=bias32, act=act,
                      out_minmax=out_minmax, out_sample_sums=out_sample_sums)
        conv_fwd(d)
        return out
    if tapm and extra is None and Cin >= 32 and x.dtype == torch.bfloat16 and (Ho * Wo) % 8 == 0:
        # many channels: materialise the tap-major im2col matrix once and run the TMA-only 1x1 kernel on it, instead of
        # re-gathering the same rows in every N-tile CTA (9x the input: a few MB on the 16x16..64x64 maps where this is used).
        # (Fetching the rows as shifted TMA boxes of x instead does not work in NCHW: the innermost box coordinate must be
        # 16-byte aligned, so the +-1 column taps fault - measured with tools/tma_probe.cu)
        col = torch.empty(B, kh * kw * Cin, Ho, Wo, device=x.device, dtype=x.dtype)
        check(lib.vrcoc_im2col(_ptr(x), _ptr(col), _dt(x), B, Cin, H, W, kh, kw, stride, pad, dil, _stream()), "im2col")
        d = conv_desc(col, w2, out, e_scale=e_scale, e_shift=bias32, act=act, out_minmax=out_minmax, out_sample_sums=out_sample_sums)
        conv_fwd(d)
        return out
    d = conv_desc(x, w2, out, src1=extra, src1_bstride=extra_bstride, kh=kh, kw=kw, stride=stride, pad=pad,
                  e_scale=e_scale, e_shift=bias32, act=act, out_minmax=out_minmax, dil=dil, k_order=1 if tapm else 0,
                  out_sample_sums=out_sample_sums)
    conv_fwd(d)
    return out


def _im2col(x, kh, kw, stride, pad, dil=1):
    B, Cin, H, W = x.shape
    Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    col = torch.empty(B, kh * kw * Cin, Ho, Wo, device=x.device, dtype=x.dtype)
    check(lib.vrcoc_im2col(_ptr(x), _ptr(col), _dt(x), B, Cin, H, W, kh, kw, stride, pad, dil, _stream()), "im2col")
    return col


def _conv_backward(x, weight, dy, stride, pad, extra=None, need_dx=True, need_dw=True, need_db=False):
    """Native backward of y = conv2d(cat[x, extra], weight) (reference vr_coc.py:99-102, normal_conv.py:42-43) given dy = dL/dy:
         dW = dy . im2col(x)^T     tap-major im2col + the 1x1 weight-gradient kernel (tensor cores for bf16), permuted back
         dx = col2im(W^T . dy)     1x1 GEMM with the transposed tap-major weight + the adjoint of im2col
    Returns (dx [B,C0,H,W] in x's dtype or None, dW [O,Cin,kh,kw] fp32 or None, db [O] fp32 or None)."""
    x = x.contiguous()
    dy = dy.contiguous()
    B, C0, H, W = x.shape
    O, Cin, kh, kw = weight.shape
    if dy.dtype != x.dtype:
        dy = dy.to(x.dtype)
    xin = x
    if extra is not None:
        e = extra if extra.dim() == 4 else extra.unsqueeze(0).expand(B, -1, -1, -1)
        xin = torch.cat([x, e.to(x.dtype)], dim=1)
    one = kh == 1 and kw == 1 and stride == 1 and pad == 0
    w2 = weight.detach().reshape(O, -1).contiguous() if one else ops.tap_major(weight)          # [O][K]
    if w2.dtype != x.dtype:
        w2 = w2.to(x.dtype)
    dW = db = dx = None
    if need_dw or need_db:
        col = xin if one else _im2col(xin, kh, kw, stride, pad)
        dWk, db = ops.conv1x1_wgrad(conv_desc(col, w2, dy), dy, want_db=need_db)
        dW = dWk.view(O, Cin, 1, 1) if one else dWk.view(O, kh, kw, Cin).permute(0, 3, 1, 2)
    if need_dx:
        wt = w2.t().contiguous()                                                                 # [K][O]
        dcol = torch.empty(B, wt.shape[0], dy.shape[2], dy.shape[3], device=x.device, dtype=x.dtype)
        conv_fwd(conv_desc(dy, wt, dcol))
        if one:
            dxin = dcol
        else:
            dxin = torch.empty(B, Cin, H, W, device=x.device, dtype=x.dtype)
            check(lib.vrcoc_col2im(_ptr(dcol), _ptr(dxin), _dt(dxin), B, Cin, H, W, kh, kw, stride, pad, 1, _stream()), "col2im")
        dx = dxin if extra is None else dxin[:, :C0].contiguous()
    return dx, dW, db


def _bn_coefficients(S, mean, rstd, gamma, N, training):
    """(ca, cb, cd, dgamma, dbeta) from S = [C,2] = {sum g, sum g*u} (fp64)"""
    S1, S2 = S[:, 0], S[:, 1]
    sgx = (S2 - mean * S1) * rstd                              # sum g * xhat
    ca = gamma * rstd
    if training:
        cb = -gamma * rstd * rstd * sgx / N
        cd = -gamma * rstd * S1 / N - cb * mean
        return ca.float().contiguous(), cb.float().contiguous(), cd.float().contiguous(), sgx.float(), S1.float()
    return ca.float().contiguous(), None, None, sgx.float(), S1.float()


def _stats_from_sums(cs, n):
    """biased batch mean / variance per channel (fp64 [C]) from per-(b, c) sums [B, C, 2]"""
    B, C = cs.shape[0], cs.shape[1]
    if cs.is_cuda and cs.dtype == torch.float32 and cs.is_contiguous():
        mean = torch.empty(C, device=cs.device, dtype=torch.float64)
        var = torch.empty_like(mean)
        check(lib.vrcoc_bn_stats(_ptr(cs), B, C, float(n), None, None, 0.0, -1.0, None, None, None, None, None, _ptr(mean), _ptr(var), _stream()),
              "bn_stats")
        return mean, var
    s = cs.double().sum(0)
    mean = s[:, 0] / n
    return mean, (s[:, 1] / n - mean * mean).clamp_min(0)


def _batch_stats(u):
    """biased batch mean / variance per channel (fp64 [C]) from the native channel sums"""
    B, C, H, W = u.shape
    cs, _ = ops.channel_sums(u)
    n = float(B * H * W)
    if cs.is_cuda and cs.dtype == torch.float32 and cs.is_contiguous():
        mean = torch.empty(C, device=u.device, dtype=torch.float64)
        var = torch.empty_like(mean)
        check(lib.vrcoc_bn_stats(_ptr(cs), B, C, n, None, None, 0.0, -1.0, None, None, None, None, None, _ptr(mean), _ptr(var), _stream()),
              "bn_stats")
        return mean, var
    s = cs.double().sum(0)
    mean = s[:, 0] / n
    return mean, (s[:, 1] / n - mean * mean).clamp_min(0)


def _norm_act_backward(dy, y_act, u, act, bn_w, bn_b, mean, var, eps, training, extra=None):
    """du, dgamma, dbeta for y = act(BN(u)); `extra` (optional) is added to du (a residual branch's gradient)."""
    B, C, H, W = u.shape
    HW, N = H * W, B * H * W
    dy = dy.contiguous()
    if dy.dtype != u.dtype:
        dy = dy.to(u.dtype)
    need_z = act == ACT_SILU or (act in (ACT_RELU, ACT_LRELU) and y_act is None)
    if u.is_cuda and mean.dtype == torch.float64 and var.dtype == torch.float64 and mean.is_contiguous() and var.is_contiguous():
        # two launches of bn_bwd_coef_kernel around the sums pass instead of ~30 library launches on [C]-sized tensors
        dev = u.device
        g32 = bn_w.detach().float().contiguous() if bn_w is not None else None
        b32 = bn_b.detach().float().contiguous() if bn_b is not None else None
        zs = zt = None
        if need_z:
            zs = torch.empty(C, device=dev, dtype=torch.float32)
            zt = torch.empty_like(zs)
            check(lib.vrcoc_bn_bwd_coef(None, B, C, _ptr(mean), _ptr(var), _ptr(g32), _ptr(b32), float(eps), float(N), int(training), _ptr(zs),
                                        _ptr(zt), None, None, None, None, None, _stream()), "bn_bwd_coef")
        ya = y_act if act in (ACT_RELU, ACT_LRELU) else None
        sums = torch.empty(B, C, 2, device=dev, dtype=torch.float32)
        check(lib.vrcoc_chan_bwd_sums(_ptr(dy), _ptr(ya), _ptr(u), _dt(u), act, _ptr(zs), _ptr(zt), B, C, HW, _ptr(sums), _stream()), "chan_bwd_sums")
        ca = torch.empty(C, device=dev, dtype=torch.float32)
        cb = torch.empty_like(ca) if training else None
        cd = torch.empty_like(ca) if training else None
        dgamma, dbeta = torch.empty_like(ca), torch.empty_like(ca)
        check(lib.vrcoc_bn_bwd_coef(_ptr(sums), B, C, _ptr(mean), _ptr(var), _ptr(g32), _ptr(b32), float(eps), float(N), int(training), None, None,
                                    _ptr(ca), _ptr(cb), _ptr(cd), _ptr(dgamma), _ptr(dbeta), _stream()), "bn_bwd_coef")
        du = torch.empty_like(u)
        check(lib.vrcoc_chan_bwd_apply(_ptr(dy), _ptr(ya), _ptr(u), _ptr(extra), _ptr(du), _dt(u), act, _ptr(ca), _ptr(cb), _ptr(cd), _ptr(zs),
                                       _ptr(zt), B, C, HW, _stream()), "chan_bwd_apply")
        return du, dgamma, dbeta
    rstd = torch.rsqrt(var + eps)
    gamma = bn_w.detach().double() if bn_w is not None else torch.ones_like(mean)
    zs = zt = None
    if need_z:
        # the activation derivative from the recomputed pre-activation z = u * zs + zt
        beta = bn_b.detach().double() if bn_b is not None else torch.zeros_like(mean)
        zs = (gamma * rstd).float().contiguous()
        zt = (beta - mean * gamma * rstd).float().contiguous()
    ya = y_act if act in (ACT_RELU, ACT_LRELU) else None
    sums = torch.empty(B, C, 2, device=u.device, dtype=torch.float32)
    check(lib.vrcoc_chan_bwd_sums(_ptr(dy), _ptr(ya), _ptr(u), _dt(u), act, _ptr(zs), _ptr(zt), B, C, HW, _ptr(sums), _stream()), "chan_bwd_sums")
    ca, cb, cd, dgamma, dbeta = _bn_coefficients(sums.double().sum(0), mean, rstd, gamma, N, training)
    du = torch.empty_like(u)
    check(lib.vrcoc_chan_bwd_apply(_ptr(dy), _ptr(ya), _ptr(u), _ptr(extra), _ptr(du), _dt(u), act, _ptr(ca), _ptr(cb), _ptr(cd), _ptr(zs),
                                   _ptr(zt), B, C, HW, _stream()), "chan_bwd_apply")
    return du, dgamma, dbeta


@ops.amp_function
class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, extra, extra_bstride, act, e_scale, out_minmax, out_dtype):
        out = _conv_launch(x, weight, _f32(bias), stride, pad, extra, extra_bstride, act, _f32(e_scale), out_minmax, out_dtype)
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x, weight, bias, extra, e_scale)
            ctx.meta = (stride, pad, act)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, extra, e_scale = ctx.saved_tensors
        stride, pad, act = ctx.meta
        if act != ACT_NONE or e_scale is not None:
            raise VrcocError("conv2d_native: backward is implemented for the plain convolution (PointRecuder, vr_coc.py:99-102); "
                             "conv + BatchNorm + activation trains through BaseConv")
        dx, dW, db = _conv_backward(x, weight, dy, stride, pad, extra, need_dx=ctx.needs_input_grad[0], need_dw=ctx.needs_input_grad[1],
                                    need_db=bias is not None and ctx.needs_input_grad[2])
        return (dx, dW, db) + (None,) * 8


_ACT_CODE = {"relu": ACT_RELU, "silu": ACT_SILU, "lrelu": ACT_LRELU}


def conv_bn_relu_infer(x, conv, bn):
    """Gradient-free eval-mode nn.Sequential(Conv2d(bias, dilation), BatchNorm2d, ReLU) (ASPP branches, reference
    neck/coc_fpn_dual.py:49-79) as one implicit-GEMM launch: BN and the conv bias folded into the epilogue."""
    sc, sh = _bn_affine(bn)
    if conv.bias is not None:
        sh = ops.cached(bn, "bias_fold", [conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var],
                        lambda: sh + _f32(conv.bias) * sc)
    return _conv_launch(x, conv.weight, sh, conv.stride[0], conv.padding[0], None, None, ACT_RELU, sc, None, None,
                        dil=conv.dilation[0])


class SiLU(nn.Module):
    """reference normal_conv.py:5-8"""
    @staticmethod
    def forward(x):
        return x * torch.sigmoid(x)


def get_activation(name="silu", inplace=True):
    """reference normal_conv.py:11-20"""
    if name == "silu":
        return SiLU()
    if name == "relu":
        return nn.ReLU(inplace=inplace)
    if name == "lrelu":
        return nn.LeakyReLU(0.1, inplace=inplace)
    raise AttributeError("Unsupported act type: {}".format(name))


@ops.amp_function
class _DWConv3Fn(torch.autograd.Function):
    """depthwise 3x3 / stride 1 / pad 1 with autograd on the native kernels: forward and input gradient = vrcoc_dwconv (the latter
    on dy with the flipped kernel), weight / bias gradient = vrcoc_dwconv3_wgrad"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        B, C, H, W = x.shape
        w = weight.detach().to(x.dtype).reshape(C, 9).contiguous()
        y = torch.empty_like(x)
        check(lib.vrcoc_dwconv(_ptr(x), _ptr(w), _ptr(_f32(bias)), _ptr(y), _dt(x), B, C, H, W, 3, 1, 1, _stream()), "dwconv")
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.wshape = weight.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        B, C, H, W = x.shape
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            check(lib.vrcoc_dwconv(_ptr(dy), _ptr(w.flip(1).contiguous()), None, _ptr(dx), _dt(x), B, C, H, W, 3, 1, 1, _stream()), "dwconv")
        dW = torch.empty(C, 9, device=x.device, dtype=torch.float32)
        db = torch.empty(C, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        check(lib.vrcoc_dwconv3_wgrad(_ptr(x), _ptr(dy), _dt(x), B, C, H, W, _ptr(dW), _ptr(db), _stream()), "dwconv3_wgrad")
        return dx, dW.reshape(ctx.wshape), db


class DWConv(nn.Module):
    """reference normal_conv.py:23-33: depthwise k x k + pointwise 1x1.  Only used by the detection head (outside the
    CoC/fusion hot path).  With autograd the 3x3 / stride-1 depthwise part runs on the native kernels (`_DWConv3Fn`); the
    pointwise part is the nn.Conv2d module (BaseConv's gradient-free path folds it with BatchNorm into one native GEMM)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        self.dconv = nn.Conv2d(in_channels, in_channels, kernel_size=kernel_size, stride=stride, groups=in_channels,
                               padding=padding, dilation=dilation, bias=bias)
        self.pconv = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, groups=1, bias=bias)

    def forward(self, x):
        d = self.dconv
        if (x.is_cuda and d.kernel_size == (3, 3) and d.stride == (1, 1) and d.padding == (1, 1) and d.dilation == (1, 1)
                and d.groups == x.shape[1] and x.shape[-1] % 8 == 0 and (x.dtype in (torch.float32, torch.bfloat16) or torch.is_autocast_enabled("cuda"))):
            return self.pconv(_DWConv3Fn.apply(x, d.weight, d.bias))
        return self.pconv(self.dconv(x))


class BaseConv(nn.Module):
    """reference normal_conv.py:36-52: conv(pad=(k-1)//2, no bias) -> BatchNorm2d(eps 1e-3, momentum 0.03) -> act,
    one launch in eval mode (BN folded into the GEMM epilogue), conv + stats + affine pass in train mode."""

    def __init__(self, in_channels, out_channels, ksize, stride, groups=1, bias=False, act="relu", ds_conv=False):
        super().__init__()
        pad = (ksize - 1) // 2
        if ds_conv is False:
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=ksize, stride=stride, padding=pad,
                                  groups=groups, bias=bias)
        else:
            self.conv = DWConv(in_channels, out_channels, kernel_size=ksize, stride=stride, padding=pad, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels, eps=0.001, momentum=0.03)
        self.act = get_activation(act, inplace=True)
        self._act_name = act

    def _native(self):
        return isinstance(self.conv, nn.Conv2d) and self.conv.groups == 1

    def _ds_infer(self, x):
        """eval / no-grad DWConv + BN + act: native depthwise kernel, then the pointwise conv with BN and both biases
        folded into one GEMM epilogue (head/decouplehead.py:24-37)."""
        dconv, pconv, bn = self.conv.dconv, self.conv.pconv, self.bn
        x = x.contiguous()
        B, C, H, W = x.shape
        k, stride, pad = dconv.kernel_size[0], dconv.stride[0], dconv.padding[0]
        Ho, Wo = ops.out_hw(H, W, k, stride, pad)
        y = torch.empty(B, C, Ho, Wo, device=x.device, dtype=x.dtype)
        wd = dconv.weight.detach().reshape(C, k * k)
        check(lib.vrcoc_dwconv(_ptr(x), _ptr(wd), _ptr(_f32(dconv.bias)), _ptr(y), _dt(x), B, C, H, W, k, stride, pad, _stream()), "dwconv")
        sc, sh = _bn_affine(bn)
        if pconv.bias is not None:
            sh = ops.cached(self, "ds_bias_fold", [pconv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var],
                            lambda: sh + _f32(pconv.bias) * sc)
        return _conv_launch(y, pconv.weight, sh, 1, 0, None, None, _ACT_CODE[self._act_name], sc, None, None)

    def forward(self, x, out_minmax=None):
        if not self._native():
            ds = isinstance(self.conv, DWConv) and self.conv.dconv.kernel_size[0] in (3, 5) and self.conv.dconv.dilation[0] == 1
            if ds and x.is_cuda and not torch.is_grad_enabled() and not self.bn.training and x.dtype in (torch.float32, torch.bfloat16) \
                    and self.conv.dconv.weight.dtype == x.dtype:
                return self._ds_infer(x)
            return self.act(self.bn(self.conv(x)))        # training / other configurations: library path (out of scope §8f)
        if not x.is_cuda:
            raise VrcocError("vrcoc BaseConv needs a CUDA tensor (no CPU fallback exists)")
        return _apply(_BaseConvFn, x, self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias, self, out_minmax)

    def fuseforward(self, x):
        return self.act(self.conv(x))


def _base_conv_forward(mod, x, out_minmax, weight=None):
    """`weight`: the (possibly autocast-cast) weight tensor the autograd Function received; defaults to the module's own"""
    conv, bn = mod.conv, mod.bn
    weight = conv.weight if weight is None else weight
    act = _ACT_CODE[mod._act_name]
    stride, pad = conv.stride[0], conv.padding[0]
    bias32 = _f32(conv.bias)
    if x.dtype != weight.dtype and x.is_floating_point():
        x = x.to(weight.dtype)
    if bn.training or not bn.track_running_stats:
        u = _conv_launch(x, weight, bias32, stride, pad, None, None, ACT_NONE, None, None, None)
        B, O, H, W = u.shape
        cs, _ = ops.channel_sums(u)
        sc, sh = _bn_affine(bn, cs, B * H * W)
        check(lib.vrcoc_chan_affine(_ptr(u), _dt(u), None, 0, _ptr(u), _dt(u), _ptr(sc), _ptr(sh), act, None, None,
                                    B, O, H * W, None, _ptr(out_minmax), _stream()), "chan_affine")
        return u
    sc, sh = _bn_affine(bn)
    if bias32 is not None:
        sh = ops.cached(mod, "bias_fold", [conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var], lambda: sh + bias32 * sc)
    return _conv_launch(x, weight, sh, stride, pad, None, None, act, sc, out_minmax, None)


class _NoGradCtx:
    """stand-in for the autograd context when a Function's forward body is run without autograd"""
    needs_input_grad = (False,) * 32

    def save_for_backward(self, *tensors):
        pass


def _apply(fn, *args):
    """fn.apply(*args) — or, without autograd, fn.forward alone: inside Function.apply `ctx.needs_input_grad` still reports
    the parameters' requires_grad, which made every gradient-free call clone the BatchNorm statistics it would need for
    backward (64 copy kernels per forward, profiles/)"""
    if torch.is_grad_enabled():
        return fn.apply(*args)
    return fn.forward(_NoGradCtx(), *args)


@ops.amp_function
class _BaseConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, bn_w, bn_b, mod, out_minmax):
        training = mod.bn.training or not mod.bn.track_running_stats
        need_bwd = any(ctx.needs_input_grad)
        rm = None if training or not need_bwd else mod.bn.running_mean.detach().clone()
        rv = None if training or not need_bwd else mod.bn.running_var.detach().clone()
        out = _base_conv_forward(mod, x, out_minmax, weight)
        if need_bwd:
            ctx.save_for_backward(x, weight, bias, bn_w, bn_b, rm, rv, out)
            ctx.meta = (mod.conv.stride[0], mod.conv.padding[0], mod._act_name, mod.bn.eps, training)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, bn_w, bn_b, rm, rv, y = ctx.saved_tensors
        stride, pad, act, eps, training = ctx.meta
        dx, dW, db, dg, dbt = _base_conv_backward(x, weight, bias, bn_w, bn_b, rm, rv, y, stride, pad, act, eps, training, dy,
                                                  need_dx=ctx.needs_input_grad[0])
        return dx, dW, db, dg, dbt, None, None


def _base_conv_backward(x, weight, bias, bn_w, bn_b, rm, rv, y, stride, pad, act, eps, training, dy, need_dx=True):
    """Native backward of BaseConv (normal_conv.py:36-52): recompute the convolution output u (not saved: it was rewritten in
    place by the BatchNorm + activation pass), BatchNorm + activation backward, convolution backward."""
    if x.dtype != weight.dtype:
        x = x.to(weight.dtype)
    u = _conv_launch(x, weight, _f32(bias), stride, pad, None, None, ACT_NONE, None, None, None)
    mean, var = _batch_stats(u) if training else (rm.double(), rv.double())
    du, dgamma, dbeta = _norm_act_backward(dy, y, u, _ACT_CODE[act], bn_w, bn_b, mean, var, eps, training)
    dx, dW, db = _conv_backward(x, weight, du, stride, pad, None, need_dx=need_dx, need_dw=True, need_db=bias is not None)
    return dx, dW, db, dgamma, dbeta


# ------------------------------------------------------------------------------------------------------------
# attention leaves
# ------------------------------------------------------------------------------------------------------------
class ShuffleAttention(nn.Module):
    """reference shuffle_attention.py:8-72 (same parameters; `channel=3, G=4` yields zero-size parameters and a
    GroupNorm(0,0) that is constructed but never called, as in reference vr_coc.py:325)."""

    def __init__(self, channel=512, reduction=16, G=8):
        super().__init__()
        self.G = G
        self.channel = channel
        q = channel // (2 * G)
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.gn = _group_norm_maybe_empty(q)
        self.cweight = nn.Parameter(torch.zeros(1, q, 1, 1))
        self.cbias = nn.Parameter(torch.ones(1, q, 1, 1))
        self.sweight = nn.Parameter(torch.zeros(1, q, 1, 1))
        self.sbias = nn.Parameter(torch.ones(1, q, 1, 1))
        self.sigmoid = nn.Sigmoid()
        # channel_shuffle(., 2) as a gather map for the stand-alone forward (built here: no H2D copy at call time)
        self.register_buffer("_perm", torch.tensor(shuffle_perm(channel, 2), dtype=torch.int32), persistent=False)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    @staticmethod
    def channel_shuffle(x, groups):
        b, c, h, w = x.shape
        return x.reshape(b, groups, -1, h, w).permute(0, 2, 1, 3, 4).reshape(b, -1, h, w)

    def params32(self):
        src = (self.cweight, self.cbias, self.sweight, self.sbias, self.gn.weight, self.gn.bias)
        return ops.cached(self, "params32", list(src), lambda: tuple(p.detach().float().reshape(-1).contiguous() for p in src))

    def gate_table(self, x):
        """attn [B,C,4] = {scale, gate_a, gate_c, mean of the attended channel} for every input channel"""
        B, Cc, H, W = x.shape
        cs = ops.channel_sums(x)[0] if H * W > 65536 else None          # small planes: each block reduces its own plane first
        attn = torch.empty(B, Cc, 4, device=x.device, dtype=torch.float32)
        cw, cb, sw, sb, gw, gb = self.params32()
        check(lib.vrcoc_sa_gate_sums(_ptr(x), _dt(x), B, Cc, H * W, self.G, _ptr(cs), _ptr(cw), _ptr(cb), _ptr(sw), _ptr(sb),
                                     _ptr(gw), _ptr(gb), _ptr(attn), _stream()), "sa_gate_sums")
        return attn

    def forward(self, x):
        if not x.is_cuda:
            raise VrcocError("vrcoc ShuffleAttention needs a CUDA tensor (no CPU fallback exists)")
        return _apply(_ShuffleAttentionFn, x, self.cweight, self.cbias, self.sweight, self.sbias, self.gn.weight, self.gn.bias, self)


def _group_norm_maybe_empty(q):
    if q > 0:
        return nn.GroupNorm(q, q)
    gn = nn.GroupNorm(1, 1)
    gn.num_groups, gn.num_channels = 0, 0
    gn.weight = nn.Parameter(torch.empty(0))
    gn.bias = nn.Parameter(torch.empty(0))
    return gn


def _prologue_backward(dzf, image, radar, chan_src, sa, eca_w, initial=False, extra_radar=None):
    """Backward of the gated / scaled prologue  z'_k = x * s * sigmoid(ga*x + gc) * e_k  in front of RadarEnhanceByImage's projection
    (vr_coc.py:344-350: ShuffleAttention on the image channels, cat + shuffle, eca_block) — and of ShuffleAttention / eca_block alone.

    dzf  [B,K,H,W]  gradient w.r.t. z' in LOGICAL channel order;  image [B,Ci,H,W] / radar [B,Cr,H,W] (either may be None);
    chan_src int32 [K]: logical k -> channel of the virtual concat [image | radar] (None = identity);
    sa = (cw, cb, sw, sb, gw, gb, G) or None;  eca_w conv1d weight or None.

    Two native passes over the maps per source (six sums per (b, channel); the elementwise apply) around the O(B*K) chain
    statistics -> gates -> ECA, which is evaluated on [B,K]-sized tensors with autograd: the sums enter as first-order models so
    that its gradients are exact.  Returns (dimage, dradar, sa_param_grads or None, deca_w or None)."""
    B, K, H, W = dzf.shape
    HW = H * W
    dev = dzf.device
    dzf = dzf.contiguous()
    Ci = 0 if image is None else image.shape[1]
    Cr = 0 if radar is None else radar.shape[1]
    if chan_src is None:
        inv = torch.arange(K, device=dev, dtype=torch.int32)
        src = inv.long()
    else:
        src = chan_src.long()
        inv = torch.empty(K, device=dev, dtype=torch.int32)
        inv[src] = torch.arange(K, device=dev, dtype=torch.int32)
    kidx_i, kidx_r = inv[:Ci].contiguous(), inv[Ci:].contiguous()
    D = torch.float64

    def leaf(t):
        return t.detach().to(D).requires_grad_(True)

    with torch.enable_grad():
        params = None
        Sx_i = Sxx_i = Sx_r = J1 = None
        gate = None
        if Ci:
            cs_i, _ = ops.channel_sums(image)
            Sx_i, Sxx_i = leaf(cs_i[..., 0]), leaf(cs_i[..., 1])
            if sa is not None and not initial:
                cw, cb, sw, sb, gw, gb, G = sa
                q = Ci // (2 * G)
                params = [leaf(t.reshape(-1)) for t in (cw, cb, sw, sb, gw, gb)]
                c = torch.arange(Ci, device=dev)
                j, half = c % q, (c // q) % 2
                m = Sx_i / HW
                rstd = torch.rsqrt(Sxx_i / HW - m * m + 1e-5)
                pc = [t[j] for t in params]
                s_i = torch.where(half == 0, torch.sigmoid(pc[0] * m + pc[1]), torch.ones_like(m))
                ga = torch.where(half == 1, pc[2] * pc[4] * rstd, torch.zeros_like(m))
                gc = torch.where(half == 1, pc[2] * (pc[5] - pc[4] * m * rstd) + pc[3], torch.full_like(m, 88.0))
                gate = torch.stack([ga.detach(), gc.detach()], dim=-1).float().contiguous()
            else:
                s_i = torch.ones_like(Sx_i)
                ga = gc = None
            sums_i = torch.empty(B, Ci, 6, device=dev, dtype=torch.float32)
            check(lib.vrcoc_table_bwd_sums(_ptr(dzf), _ptr(image), _dt(image), _ptr(kidx_i), _ptr(gate), B, Ci, K, HW, _ptr(sums_i), _stream()),
                  "table_bwd_sums")
            si = sums_i.to(D)
            J1 = leaf(si[..., 3])
            J1v = J1
            if gate is not None:
                J1v = J1 + si[..., 4] * (ga - ga.detach()) + si[..., 5] * (gc - gc.detach())
            mean_att = s_i * J1v / HW
        if Cr:
            cs_r, _ = ops.channel_sums(radar)
            Sx_r = leaf(cs_r[..., 0])
            sums_r = torch.empty(B, Cr, 6, device=dev, dtype=torch.float32)
            check(lib.vrcoc_table_bwd_sums(_ptr(dzf), _ptr(radar), _dt(radar), _ptr(kidx_r), None, B, Cr, K, HW, _ptr(sums_r), _stream()),
                  "table_bwd_sums")
            sr = sums_r.to(D)
        M_src = torch.cat([t for t in ((mean_att if Ci else None), (Sx_r / HW if Cr else None)) if t is not None], dim=1)
        ew = None
        if eca_w is not None:
            ew = leaf(eca_w.reshape(-1))
            k = ew.numel()
            a = F.conv1d(M_src[:, src].unsqueeze(1), ew.view(1, 1, k), padding=(k - 1) // 2).squeeze(1)
            e_src = torch.sigmoid(a)[:, inv.long()]
        else:
            e_src = torch.ones_like(M_src)
        e_i, e_r = e_src[:, :Ci], e_src[:, Ci:]
        L = 0.0
        if Ci:
            es = e_i * s_i
            L = L + (es * si[..., 0]).sum()
            if gate is not None:
                L = L + (es.detach() * si[..., 1] * ga).sum() + (es.detach() * si[..., 2] * gc).sum()
        if Cr:
            L = L + (e_r * sr[..., 0]).sum()
        wrt = [t for t in (Sx_i, Sxx_i, J1, Sx_r, ew) if t is not None] + (params or [])
        got = torch.autograd.grad(L, wrt, allow_unused=True)
    g = {id(t): (v if v is not None else torch.zeros_like(t)) for t, v in zip(wrt, got)}
    dimage = dradar = None
    if Ci:
        coef = torch.stack([es.detach(), g[id(J1)], g[id(Sx_i)], g[id(Sxx_i)]], dim=-1).float().contiguous()
        dimage = torch.empty_like(image)
        check(lib.vrcoc_table_bwd_apply(_ptr(dzf), _ptr(image), None, _ptr(dimage), _dt(image), _ptr(kidx_i), _ptr(gate), _ptr(coef), B, Ci, K, HW,
                                        _stream()), "table_bwd_apply")
    if Cr:
        z = torch.zeros_like(e_r)
        coef = torch.stack([e_r.detach(), z, g[id(Sx_r)], z], dim=-1).float().contiguous()
        dradar = torch.empty_like(radar)
        ex = None if extra_radar is None else extra_radar.contiguous()
        check(lib.vrcoc_table_bwd_apply(_ptr(dzf), _ptr(radar), _ptr(ex), _ptr(dradar), _dt(radar), _ptr(kidx_r), None, _ptr(coef), B, Cr, K, HW,
                                        _stream()), "table_bwd_apply")
    sa_grads = None if params is None else [g[id(t)].float() for t in params]
    return dimage, dradar, sa_grads, (None if ew is None else g[id(ew)].float())


@ops.amp_function
class _ShuffleAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cw, cb, sw, sb, gw, gb, mod):
        x = x.contiguous()
        B, Cc, H, W = x.shape
        attn = mod.gate_table(x)
        table = torch.stack([attn[..., 0], torch.zeros_like(attn[..., 0]), attn[..., 1], attn[..., 2]], dim=-1).contiguous()
        perm = mod._perm
        table = table[:, perm.long()].contiguous()
        out = torch.empty_like(x)
        d = conv_desc(x, x, out, chan_src=perm, table=table, has_gate=True)
        check(lib.vrcoc_table_apply(d, _stream()), "table_apply")
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x, cw, cb, sw, sb, gw, gb)
            ctx.G = mod.G
            ctx.perm = perm
        return out

    @staticmethod
    def backward(ctx, dy):
        x, cw, cb, sw, sb, gw, gb = ctx.saved_tensors
        dx, _, pg, _ = _prologue_backward(dy.to(x.dtype), x, None, ctx.perm, (cw, cb, sw, sb, gw, gb, ctx.G), None)
        shapes = [t.shape for t in (cw, cb, sw, sb, gw, gb)]
        return (dx,) + tuple(p.reshape(sh) for p, sh in zip(pg, shapes)) + (None,)


class eca_block(nn.Module):
    """reference eca.py:6-22."""

    def __init__(self, channel, b=1, gamma=2):
        super().__init__()
        kernel_size = int(abs((math.log(channel, 2) + b) / gamma))
        kernel_size = kernel_size if kernel_size % 2 else kernel_size + 1
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv = nn.Conv1d(1, 1, kernel_size=kernel_size, padding=(kernel_size - 1) // 2, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        if not x.is_cuda:
            raise VrcocError("vrcoc eca_block needs a CUDA tensor (no CPU fallback exists)")
        return _apply(_EcaFn, x, self.conv.weight)


@ops.amp_function
class _EcaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        x = x.contiguous()
        B, Cc, H, W = x.shape
        cs = ops.channel_sums(x)[0] if H * W > 65536 else None
        k = w.shape[-1]
        # attn records with scale 1 / gate off / mean = channel mean, then the same ECA table kernel as the fused path
        attn = torch.empty(B, Cc, 4, device=x.device, dtype=torch.float32)
        check(lib.vrcoc_sa_gate_sums(_ptr(x), _dt(x), B, Cc, H * W, 0, _ptr(cs), None, None, None, None, None, None,
                                     _ptr(attn), _stream()), "sa_gate_sums")
        table = torch.empty(B, Cc + 1, 4, device=x.device, dtype=torch.float32)
        zero_cs = torch.zeros(B, 1, 2, device=x.device, dtype=torch.float32)
        # a dummy zero radar channel at the end does not influence the zero-padded conv1d of the real channels
        check(lib.vrcoc_radar_enh_table(_ptr(attn), _ptr(zero_cs), None, _ptr(_f32(w).reshape(-1)), k, B, Cc, 1, H * W,
                                        _ptr(table), _stream()), "radar_enh_table")
        table = table[:, :Cc].contiguous()
        out = torch.empty_like(x)
        check(lib.vrcoc_table_apply(conv_desc(x, x, out, table=table, has_gate=False), _stream()), "table_apply")
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        _, dx, _, dw = _prologue_backward(dy.to(x.dtype), None, x, None, None, w)
        return dx, dw.reshape(w.shape)


# ------------------------------------------------------------------------------------------------------------
# fusion modules
# ------------------------------------------------------------------------------------------------------------
class ImageEnhanceByRadar(nn.Module):
    """reference vr_coc.py:303-316: out = BN((1 + data_normal(ReLU(BN(conv3x3(radar))))) * image)."""

    def __init__(self, radar_in_channels, image_in_channels):
        super().__init__()
        self.radar_in_channels = radar_in_channels
        self.image_in_channels = image_in_channels
        self.radar_projection = BaseConv(in_channels=radar_in_channels, out_channels=image_in_channels, ksize=3, stride=1)
        self.norm = nn.BatchNorm2d(image_in_channels)

    def forward(self, image_map, radar_map):
        if not image_map.is_cuda:
            raise VrcocError("vrcoc ImageEnhanceByRadar needs CUDA tensors (no CPU fallback exists)")
        rp = self.radar_projection
        return _apply(_ImageEnhanceFn, image_map, radar_map, rp.conv.weight, rp.bn.weight, rp.bn.bias,
                                     self.norm.weight, self.norm.bias, self)


@ops.amp_function
class _ImageEnhanceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, radar, w, g1, b1, g2, b2, mod):
        image = image.contiguous()
        radar = radar.contiguous()
        rp, bn2 = mod.radar_projection, mod.norm
        train1 = rp.bn.training or not rp.bn.track_running_stats
        train2 = bn2.training or not bn2.track_running_stats
        stats = None
        if any(ctx.needs_input_grad):
            stats = tuple(None if t else s.detach().clone() for t, s in
                          ((train1, rp.bn.running_mean), (train1, rp.bn.running_var), (train2, bn2.running_mean), (train2, bn2.running_var)))
        B, Ci, H, W = image.shape
        minmax = torch.zeros(2, device=image.device, dtype=torch.int32)
        k = _base_conv_forward(rp, radar if radar.dtype == image.dtype else radar.to(image.dtype), minmax, w)
        out = torch.empty_like(image)
        if train2:
            cs = torch.empty(B, Ci, 2, device=image.device, dtype=torch.float32)
            check(lib.vrcoc_img_enh_finish(_ptr(k), _dt(k), _ptr(image), _dt(image), _ptr(out), _dt(out), _ptr(minmax),
                                           None, None, B, Ci, H * W, _ptr(cs), _stream()), "img_enh_finish")
            sc, sh = _bn_affine(bn2, cs, B * H * W)
            check(lib.vrcoc_chan_affine(_ptr(out), _dt(out), None, 0, _ptr(out), _dt(out), _ptr(sc), _ptr(sh), ACT_NONE,
                                        None, None, B, Ci, H * W, None, None, _stream()), "chan_affine")
        else:
            sc, sh = _bn_affine(bn2)
            check(lib.vrcoc_img_enh_finish(_ptr(k), _dt(k), _ptr(image), _dt(image), _ptr(out), _dt(out), _ptr(minmax),
                                           _ptr(sc), _ptr(sh), B, Ci, H * W, None, _stream()), "img_enh_finish")
        if stats is not None:
            ctx.save_for_backward(image, radar, w, g1, b1, g2, b2, k, minmax, *stats)
            ctx.meta = (rp.bn.eps, bn2.eps, train1, train2)
        return out

    @staticmethod
    def backward(ctx, dy):
        """native: recompute y = (1 + kn) * image -> BatchNorm backward -> tail backward (+ the min / max paths) -> BaseConv backward"""
        image, radar, w, g1, b1, g2, b2, k, minmax, rm1, rv1, rm2, rv2 = ctx.saved_tensors
        eps1, eps2, t1, t2 = ctx.meta
        B, Ci, H, W = image.shape
        HW = H * W
        dev = image.device
        yv = torch.empty_like(image)
        cs = torch.empty(B, Ci, 2, device=dev, dtype=torch.float32)
        check(lib.vrcoc_img_enh_finish(_ptr(k), _dt(k), _ptr(image), _dt(image), _ptr(yv), _dt(yv), _ptr(minmax), None, None, B, Ci, HW,
                                       _ptr(cs), _stream()), "img_enh_finish")
        if t2:
            mean2, var2 = _stats_from_sums(cs, float(B * HW))
        else:
            mean2, var2 = rm2.double(), rv2.double()
        dyv, dg2, db2 = _norm_act_backward(dy, None, yv, ACT_NONE, g2, b2, mean2, var2, eps2, t2)
        dimage = torch.empty_like(image)
        dk = torch.empty_like(k)
        part = torch.empty(B * Ci, 4, device=dev, dtype=torch.float32)
        check(lib.vrcoc_img_enh_bwd(_ptr(dyv), _ptr(image), _ptr(k), _dt(k), _ptr(minmax), B, Ci, HW, _ptr(dimage), _ptr(dk), _ptr(part),
                                    _stream()), "img_enh_bwd")
        # d mn = r^2 * sum dkn*(k - mx),  d mx = -r^2 * sum dkn*(k - mn): O(1)-sized algebra on device scalars, then spread over the ties
        p = part.double().sum(0)
        mm = minmax.view(torch.int32)
        mx = mm[0:1].view(torch.float32).double()
        mn = (~mm[1:2]).view(torch.float32).double()
        r2 = 1.0 / ((mx - mn) * (mx - mn))
        coef = torch.cat([r2 * (p[1] - mx * p[0]) / p[2].clamp_min(1.0), -r2 * (p[1] - mn * p[0]) / p[3].clamp_min(1.0)]).float().contiguous()
        check(lib.vrcoc_minmax_scatter(_ptr(k), _ptr(dk), _dt(k), _ptr(minmax), _ptr(coef), k.numel(), _stream()), "minmax_scatter")
        rad = radar if radar.dtype == image.dtype else radar.to(image.dtype)
        drad, dW, _, dg1, db1 = _base_conv_backward(rad, w, None, g1, b1, rm1, rv1, k, 1, 1, "relu", eps1, t1, dk,
                                                   need_dx=ctx.needs_input_grad[1])
        return dimage, drad, dW, dg1, db1, dg2, db2, None


class RadarEnhanceByImage(nn.Module):
    """reference vr_coc.py:319-359."""

    def __init__(self, radar_in_channels, image_in_channels, initial=False):
        super().__init__()
        self.initial = initial
        self.radar_in_channels = radar_in_channels
        self.image_in_channels = image_in_channels
        self.image_attn = ShuffleAttention(channel=image_in_channels, G=4)
        self.channel_attn = eca_block(channel=radar_in_channels + image_in_channels)
        self.inverse_projection = BaseConv(in_channels=radar_in_channels + image_in_channels,
                                           out_channels=radar_in_channels, ksize=1, stride=1)
        self.norm = nn.BatchNorm2d(radar_in_channels)
        # logical channel k of shuffle_channels(cat[image_attn(image), radar], 2) -> channel of the virtual concat
        # [image | radar] as stored in memory; ShuffleAttention's own channel_shuffle(.,2) is folded in.
        Ci, Cr = image_in_channels, radar_in_channels
        outer = shuffle_perm(Ci + Cr, 2)
        inner = list(range(Ci)) if initial else shuffle_perm(Ci, 2)
        cmap = [inner[j] if j < Ci else j for j in outer]
        self.register_buffer("_chan_src", torch.tensor(cmap, dtype=torch.int32), persistent=False)

    def forward(self, image_map, radar_map):
        if not image_map.is_cuda:
            raise VrcocError("vrcoc RadarEnhanceByImage needs CUDA tensors (no CPU fallback exists)")
        ip, sa = self.inverse_projection, self.image_attn
        return _apply(_RadarEnhanceFn, image_map, radar_map, ip.conv.weight, ip.bn.weight, ip.bn.bias, self.norm.weight,
                                     self.norm.bias, self.channel_attn.conv.weight, sa.cweight, sa.cbias, sa.sweight,
                                     sa.sbias, sa.gn.weight, sa.gn.bias, self)


def _concat_order_weight(w2, chan_src):
    """W'[:, chan_src[k]] = W[:, k]: the projection weight for inputs in the memory order of the virtual concat"""
    wn = torch.empty_like(w2)
    wn[:, chan_src.long()] = w2
    return wn.contiguous()


@ops.amp_function
class _RadarEnhanceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, radar, w, g1, b1, g2, b2, eca_w, cw, cb, sw, sb, gw, gb, mod):
        image = image.contiguous()
        radar = radar.contiguous()
        if image.dtype != radar.dtype:
            image = image.to(radar.dtype)
        ip, bn2, sa = mod.inverse_projection, mod.norm, mod.image_attn
        train1 = ip.bn.training or not ip.bn.track_running_stats
        train2 = bn2.training or not bn2.track_running_stats
        stats = None
        if any(ctx.needs_input_grad):
            stats = tuple(None if t else s.detach().clone() for t, s in
                          ((train1, ip.bn.running_mean), (train1, ip.bn.running_var), (train2, bn2.running_mean), (train2, bn2.running_var)))
        B, Ci, H, W = image.shape
        Cr = radar.shape[1]
        HW = H * W
        dev = image.device
        # statistics + attention parameters + ECA -> conv prologue table
        attn = torch.empty(B, Ci, 4, device=dev, dtype=torch.float32)
        pc = [None] * 6 if mod.initial else sa.params32()
        G = 0 if mod.initial else sa.G
        if HW <= 65536:
            # plane sums of both maps and the attention gates in one launch (a block per plane reduces its own plane)
            cs_rad = torch.empty(B, Cr, 2, device=dev, dtype=torch.float32)
            check(lib.vrcoc_fusion_stats(_ptr(image), _ptr(radar), _dt(image), B, Ci, Cr, HW, G, *[_ptr(p) for p in pc], _ptr(attn),
                                         _ptr(cs_rad), _stream()), "fusion_stats")
        else:
            cs_img, _ = ops.channel_sums(image)
            cs_rad, _ = ops.channel_sums(radar)
            check(lib.vrcoc_sa_gate_sums(_ptr(image), _dt(image), B, Ci, HW, G, _ptr(cs_img), *[_ptr(p) for p in pc],
                                         _ptr(attn), _stream()), "sa_gate_sums")
        table = torch.empty(B, Ci + Cr, 4, device=dev, dtype=torch.float32)
        ew = _f32(eca_w).reshape(-1)
        w2 = w.detach().reshape(Cr, -1).contiguous()
        out = torch.empty_like(radar)
        gate = not mod.initial
        # eval, bf16, the image/radar boundary on a 64-channel k slab: the contraction does not care about the order of k, so the
        # shuffle moves from the activations to the weight columns (W'[:, chan_src[k]] = W[:, k], memoised) and the GEMM reads
        # [image | radar] in memory order - both sources by TMA, attention / ECA applied in place (channel-major kernel)
        concat_order = (not (train1 or train2) and not any(ctx.needs_input_grad) and image.dtype == torch.bfloat16 and Ci % 64 == 0
                        and w2.dtype == torch.bfloat16 and HW % 8 == 0)
        if concat_order:
            check(lib.vrcoc_radar_enh_table_concat_order(_ptr(attn), _ptr(cs_rad), _ptr(mod._chan_src), _ptr(ew), ew.numel(), B, Ci,
                                                         Cr, HW, _ptr(table), _stream()), "radar_enh_table_concat_order")
            w_nat = ops.cached(mod, "w_concat_order", [w], lambda: _concat_order_weight(w2, mod._chan_src))
            s1, t1 = _bn_affine(ip.bn)
            s2, t2 = _bn_affine(bn2)
            if SPLIT_WIDE_PROLOGUE and Ci + Cr > 576:
                # more than nine k-slabs (stages 3 and 4): the in-place-prologue GEMM cannot keep [image | radar] resident and the
                # transform-on-load kernel re-applies attention / gate per output-tile CTA (50 us at stage 4).  One streaming pass
                # writes the transformed operand (same bf16 values), the projection then takes both operands by TMA.
                z = torch.empty(B, Ci + Cr, *image.shape[-2:], device=dev, dtype=image.dtype)
                check(lib.vrcoc_table_apply(conv_desc(image, image, z, src1=radar, table=table, has_gate=gate), _stream()), "table_apply")
                conv_fwd(conv_desc(z, w_nat, out, e_scale=s1, e_shift=t1, act=ACT_RELU, res=radar, f_scale=s2, f_shift=t2))
                return out
            conv_fwd(conv_desc(image, w_nat, out, src1=radar, table=table, has_gate=gate,
                               e_scale=s1, e_shift=t1, act=ACT_RELU, res=radar, f_scale=s2, f_shift=t2))
            return out
        check(lib.vrcoc_radar_enh_table(_ptr(attn), _ptr(cs_rad), _ptr(mod._chan_src), _ptr(ew), ew.numel(), B, Ci, Cr, HW,
                                        _ptr(table), _stream()), "radar_enh_table")
        if train1 or train2:
            u = torch.empty_like(radar)
            conv_fwd(conv_desc(image, w2, u, src1=radar, chan_src=mod._chan_src, table=table, has_gate=gate))
            cs_u, _ = ops.channel_sums(u)
            s1, t1 = _bn_affine(ip.bn, cs_u, B * HW)
            cs_t = torch.empty(B, Cr, 2, device=dev, dtype=torch.float32)
            check(lib.vrcoc_chan_affine(_ptr(u), _dt(u), _ptr(radar), _dt(radar), _ptr(u), _dt(u), _ptr(s1), _ptr(t1), ACT_RELU,
                                        None, None, B, Cr, HW, _ptr(cs_t), None, _stream()), "chan_affine")
            s2, t2 = _bn_affine(bn2, cs_t, B * HW)
            check(lib.vrcoc_chan_affine(_ptr(u), _dt(u), None, 0, _ptr(out), _dt(out), _ptr(s2), _ptr(t2), ACT_NONE, None, None,
                                        B, Cr, HW, None, None, _stream()), "chan_affine")
        else:
            s1, t1 = _bn_affine(ip.bn)
            s2, t2 = _bn_affine(bn2)
            conv_fwd(conv_desc(image, w2, out, src1=radar, chan_src=mod._chan_src, table=table, has_gate=gate,
                               e_scale=s1, e_shift=t1, act=ACT_RELU, res=radar, f_scale=s2, f_shift=t2))
        if stats is not None:
            ctx.save_for_backward(image, radar, w, g1, b1, g2, b2, eca_w, cw, cb, sw, sb, gw, gb, table, *stats)
            ctx.meta = (ip.bn.eps, bn2.eps, train1, train2, mod.initial, sa.G)
            ctx.chan_src = mod._chan_src
        return out

    @staticmethod
    def backward(ctx, dy):
        """native: recompute u = W . prologue(image | radar) and t = ReLU(BN1(u)) + radar; BN2 backward; ReLU + BN1 backward; weight
        gradient (prologue applied inside the kernel) and dgrad GEMM; prologue backward (gates, shuffle, ECA)."""
        tsv = ctx.saved_tensors
        image, radar, w, g1, b1, g2, b2, eca_w, cw, cb, sw, sb, gw, gb, table = tsv[:15]
        rm1, rv1, rm2, rv2 = tsv[15:]
        eps1, eps2, t1, t2, initial, G = ctx.meta
        chan_src = ctx.chan_src
        B, Ci, H, W = image.shape
        Cr = radar.shape[1]
        HW = H * W
        dev = image.device
        gate = not initial
        w2 = w.detach().reshape(Cr, -1).contiguous()
        if w2.dtype != radar.dtype:
            w2 = w2.to(radar.dtype)
        u = torch.empty_like(radar)
        z = None
        if radar.dtype == torch.bfloat16 and image.dtype == torch.bfloat16 and HW % 8 == 0:
            # bf16: materialise z = prologue([image | radar]) (shuffle, gates, ECA scale) once - the recomputed projection and, above all,
            # the weight gradient then run on the tensor core as plain 1x1 problems (the two-source / gated form of the weight gradient
            # only exists on the CUDA cores: 0.65 ms per fusion stage in the training profile, profiles/r02_prof_train.txt)
            z = torch.empty(B, Ci + Cr, H, W, device=dev, dtype=radar.dtype)
            check(lib.vrcoc_table_apply(conv_desc(image, image, z, src1=radar, chan_src=chan_src, table=table, has_gate=gate), _stream()),
                  "table_apply")
            conv_fwd(conv_desc(z, w2, u))
        else:
            conv_fwd(conv_desc(image, w2, u, src1=radar, chan_src=chan_src, table=table, has_gate=gate))
        mean1, var1 = _batch_stats(u) if t1 else (rm1.double(), rv1.double())
        s1, sh1 = _fold_bn(mean1.float(), var1.float(), g1.detach().float(), b1.detach().float(), eps1)
        t = torch.empty_like(radar)
        cs_t = torch.empty(B, Cr, 2, device=dev, dtype=torch.float32)
        check(lib.vrcoc_chan_affine(_ptr(u), _dt(u), _ptr(radar), _dt(radar), _ptr(t), _dt(t), _ptr(s1), _ptr(sh1), ACT_RELU, None, None,
                                    B, Cr, HW, _ptr(cs_t), None, _stream()), "chan_affine")
        if t2:
            mean2, var2 = _stats_from_sums(cs_t, float(B * HW))
        else:
            mean2, var2 = rm2.double(), rv2.double()
        dt, dg2, db2 = _norm_act_backward(dy, None, t, ACT_NONE, g2, b2, mean2, var2, eps2, t2)
        du, dg1, db1 = _norm_act_backward(dt, None, u, ACT_RELU, g1, b1, mean1, var1, eps1, t1)
        if z is not None:
            dWk, _ = ops.conv1x1_wgrad(conv_desc(z, w2, du), du, want_db=False)
        else:
            dWk, _ = ops.conv1x1_wgrad(conv_desc(image, w2, du, src1=radar, chan_src=chan_src, table=table, has_gate=gate), du, want_db=False)
        dzf = torch.empty(B, Ci + Cr, H, W, device=dev, dtype=radar.dtype)
        conv_fwd(conv_desc(du, w2.t().contiguous(), dzf))
        sa = None if initial else (cw, cb, sw, sb, gw, gb, G)
        dimage, dradar, pg, dew = _prologue_backward(dzf, image, radar, chan_src, sa, eca_w, initial=initial, extra_radar=dt)
        if pg is None:
            pg = [None] * 6
        else:
            pg = [p.reshape(tt.shape) for p, tt in zip(pg, (cw, cb, sw, sb, gw, gb))]
        return (dimage, dradar, dWk.reshape(w.shape), dg1, db1, dg2, db2, dew.reshape(eca_w.shape), *pg, None)
