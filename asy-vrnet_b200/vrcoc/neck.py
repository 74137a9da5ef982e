"""Dual FPN neck: reference neck/coc_fpn_dual.py:15-224 (CoCUpsample, CoC_Conv, ASPP, SpatialPyramidPooling,
shuffle_channels, CoCFpnDual), same constructor arguments, module tree and state-dict keys.

Hot-path content here = the three CoC_Conv ClusterBlocks, the 1x1 BaseConvs and the
ShuffleAttention gates, all on the native kernels; so are the "next" rows of SURVEY §8f in this file: ASPP's dilated 3x3 convs
(im2col + tcgen05 GEMM), the bilinear upsamples (csrc/stats.cu) and, for serving, the fused upsample + arg-max class map.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import check, lib
from .context_cluster import ClusterBlock
from .fusion import BaseConv, DWConv, ShuffleAttention, eca_block, shuffle_channels  # noqa: F401
from .vr_coc import coc_medium, coc_small  # noqa: F401


@ops.amp_function
class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        x = x.contiguous()
        B, C, H, W = x.shape
        Ho, Wo = int(H * scale), int(W * scale)
        out = torch.empty(B, C, Ho, Wo, device=x.device, dtype=x.dtype)
        check(lib.vrcoc_upsample_bilinear(x.data_ptr(), out.data_ptr(), ops._dt(x), B * C, H, W, Ho, Wo, ops._stream()),
              "upsample_bilinear")
        ctx.shape = (H, W, scale)
        return out

    @staticmethod
    def backward(ctx, dy):
        # adjoint of a fixed linear map, native (gather form, the forward kernel's own weights)
        H, W, scale = ctx.shape
        dy = dy.contiguous()
        B, C, Ho, Wo = dy.shape
        dx = torch.empty(B, C, H, W, device=dy.device, dtype=dy.dtype)
        check(lib.vrcoc_upsample_bilinear_bwd(dy.data_ptr(), dx.data_ptr(), ops._dt(dy), B * C, H, W, Ho, Wo, ops._stream()),
              "upsample_bilinear_bwd")
        return dx, None


class BilinearUpsample(nn.Upsample):
    """nn.Upsample(mode='bilinear', align_corners=True) on the native kernel (the x4 upsample of the segmentation logits
    alone cost 3.5 ms per 8-frame batch through the ATen kernel, profiles/r01_launches_first.csv)."""

    def forward(self, x):
        if x.is_cuda and self.mode == "bilinear" and self.align_corners and x.dtype in (torch.float32, torch.bfloat16):
            if not torch.is_grad_enabled():
                return _UpsampleFn.forward(type("_Ctx", (), {})(), x, float(self.scale_factor))
            return _UpsampleFn.apply(x, float(self.scale_factor))
        return super().forward(x)


def upsample_argmax(x, scale):
    """class map [B, H*scale, W*scale] uint8 = argmax over dim 1 of the align_corners=True bilinear upsample of the logits x
    [B, C, H, W] (rounded to x.dtype, lowest index on ties): bit-identical to BilinearUpsample followed by argmax(1), one launch,
    the full-size logits never reach memory (csrc/stats.cu: upsample_argmax_kernel; reference deeplab.py:149-167)."""
    x = x.contiguous()
    B, C, H, W = x.shape
    Ho, Wo = int(H * scale), int(W * scale)
    out = torch.empty(B, Ho, Wo, device=x.device, dtype=torch.uint8)
    check(lib.vrcoc_upsample_argmax(x.data_ptr(), out.data_ptr(), ops._dt(x), B, C, H, W, Ho, Wo, ops._stream()), "upsample_argmax")
    return out


def upsample_argmax_ok(x, scale):
    B, C, H, W = x.shape
    return x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and \
        bool(lib.vrcoc_upsample_argmax_supported(C, H, W, int(H * scale), int(W * scale)))


class CoCUpsample(nn.Module):
    """reference coc_fpn_dual.py:15-26"""

    def __init__(self, in_channels, out_channels, scale=2, ds_conv=False):
        super().__init__()
        self.upsample = nn.Sequential(
            BaseConv(in_channels, out_channels, 1, 1, act='relu', ds_conv=ds_conv),
            BilinearUpsample(scale_factor=scale, mode='bilinear', align_corners=True))

    def forward(self, x):
        return self.upsample(x)


class CoC_Conv(nn.Module):
    """reference coc_fpn_dual.py:29-39: default-argument ClusterBlock followed by a 1x1 BaseConv"""

    def __init__(self, in_channels, out_channels, ksize=1, stride=1, act="relu", ds_conv=False):
        super().__init__()
        self.coc = ClusterBlock(dim=in_channels)
        self.conv_att = BaseConv(in_channels, out_channels, ksize=ksize, stride=stride, act=act, ds_conv=ds_conv)

    def forward(self, x):
        return self.conv_att(self.coc(x))


class ASPP(nn.Module):
    """reference coc_fpn_dual.py:46-104"""

    def __init__(self, dim_in, dim_out, rate=1, bn_mom=0.1):
        super().__init__()

        def branch(k, dil):
            return nn.Sequential(nn.Conv2d(dim_in, dim_out, k, 1, padding=0 if k == 1 else dil, dilation=dil, bias=True),
                                 nn.BatchNorm2d(dim_out, momentum=bn_mom), nn.ReLU(inplace=True))
        self.branch1 = branch(1, rate)
        self.branch2 = branch(3, 6 * rate)
        self.branch3 = branch(3, 12 * rate)
        self.branch4 = branch(3, 18 * rate)
        self.branch5_conv = nn.Conv2d(dim_in, dim_out, 1, 1, 0, bias=True)
        self.branch5_bn = nn.BatchNorm2d(dim_out, momentum=bn_mom)
        self.branch5_relu = nn.ReLU(inplace=True)
        self.conv_cat = nn.Sequential(nn.Conv2d(dim_out * 5, dim_out, 1, 1, padding=0, bias=True),
                                      nn.BatchNorm2d(dim_out, momentum=bn_mom), nn.ReLU(inplace=True))

    def forward(self, x):
        b, c, row, col = x.size()
        native = x.is_cuda and not torch.is_grad_enabled() and not self.training and x.dtype in (torch.float32, torch.bfloat16)
        if native:     # eval, no autograd: every conv+BN+ReLU is one implicit-GEMM launch (dilation in the im2col gather)
            from .fusion import conv_bn_relu_infer
            # the dilated branches are independent and each fills only part of the chip at 16x16: three of them on side streams
            forks = [ops.Fork(lambda br=br: conv_bn_relu_infer(x, br[0], br[1]), lane=8 + i)
                     for i, br in enumerate((self.branch2, self.branch3, self.branch4))]
            feats = [conv_bn_relu_infer(x, self.branch1[0], self.branch1[1])] + [f.join() for f in forks]
        else:
            feats = [self.branch1(x), self.branch2(x), self.branch3(x), self.branch4(x)]
        g = torch.mean(torch.mean(x, 2, True), 3, True)
        g = self.branch5_relu(self.branch5_bn(self.branch5_conv(g)))
        # bilinear interpolation of a 1x1 map with align_corners=True is a broadcast (the ATen kernel took 0.37 ms for it)
        g = g.expand(-1, -1, row, col) if native else F.interpolate(g, (row, col), None, 'bilinear', True)
        cat = torch.cat(feats + [g], dim=1)
        if native:
            from .fusion import conv_bn_relu_infer
            return conv_bn_relu_infer(cat, self.conv_cat[0], self.conv_cat[1])
        return self.conv_cat(cat)


class SpatialPyramidPooling(nn.Module):
    """reference coc_fpn_dual.py:107-117 (unused by the live model)"""

    def __init__(self, pool_sizes=[5, 9, 13]):
        super().__init__()
        self.maxpools = nn.ModuleList([nn.MaxPool2d(p, 1, p // 2) for p in pool_sizes])

    def forward(self, x):
        return torch.cat([m(x) for m in self.maxpools[::-1]] + [x], dim=1)


def cat_shuffle(a, b):
    """shuffle_channels(torch.cat([a, b], 1)) (reference coc_fpn_dual.py:196-197) in one pass: a two-source gather with the
    interleave folded into the channel map (no concatenated intermediate)."""
    if not (a.is_cuda and a.dtype == b.dtype and a.dtype in (torch.float32, torch.bfloat16)) or torch.is_grad_enabled():
        return shuffle_channels(torch.cat([a, b], dim=1))
    from .fusion import shuffle_perm
    a, b = a.contiguous(), b.contiguous()
    B, Ca, H, W = a.shape
    C = Ca + b.shape[1]
    key = (Ca, C, a.device)
    perm = _PERM_CACHE.get(key)
    if perm is None:
        perm = _PERM_CACHE[key] = torch.tensor(shuffle_perm(C, 2), dtype=torch.int32, device=a.device)
    out = torch.empty(B, C, H, W, device=a.device, dtype=a.dtype)
    check(lib.vrcoc_table_apply(ops.conv_desc(a, a, out, src1=b, chan_src=perm), ops._stream()), "table_apply")
    return out


_PERM_CACHE = {}


class CoCFpnDual(nn.Module):
    """reference coc_fpn_dual.py:133-224"""

    def __init__(self, num_seg_class=9, depth=1.0, width=1.0, in_features=("dark2", "dark3", "dark4", "dark5"),
                 in_channels=[64, 128, 320, 512], aspp_channel=1024):
        super().__init__()
        Conv = CoC_Conv
        self.backbone = coc_small(pretrained=False, width=width)
        self.in_features = in_features
        self.num_seg_class = num_seg_class
        in_channels = [int(item * width) for item in in_channels]
        self.aspp = ASPP(dim_in=in_channels[-1], dim_out=in_channels[-1])
        self.upsample5_4 = CoCUpsample(in_channels=in_channels[-1], out_channels=in_channels[-2])
        self.sc_attn_seg4 = ShuffleAttention(channel=in_channels[-2] * 2)
        self.upsample4_3 = CoCUpsample(in_channels=in_channels[-2] * 2, out_channels=in_channels[-3])
        self.sc_attn_seg3 = ShuffleAttention(channel=in_channels[-3] * 2)
        self.upsample3_2 = CoCUpsample(in_channels=in_channels[-3] * 2, out_channels=in_channels[0])
        self.sc_attn_seg2 = ShuffleAttention(channel=in_channels[0] * 2)
        self.upsample2_0 = CoCUpsample(in_channels=in_channels[0] * 2, out_channels=self.num_seg_class, scale=4)
        self.p5_out_det = Conv(in_channels=in_channels[-1], out_channels=in_channels[-1])
        self.p5_4_det = CoCUpsample(in_channels=in_channels[-1], out_channels=in_channels[-2])
        self.p4_out_det = Conv(in_channels=in_channels[-2] * 2, out_channels=in_channels[-2])
        self.p4_3_det = CoCUpsample(in_channels=in_channels[-2], out_channels=in_channels[-3])
        self.p3_out_det = Conv(in_channels=in_channels[-3] * 2, out_channels=in_channels[-3])

    # Optional hook (an ATTRIBUTE, so that forward keeps the reference signature (x, x_radar), coc_fpn_dual.py:184): a callable
    # (k, p_k) applied to each detection map as soon as it exists.  EfficientVRNet sets it to DecoupleHead.forward_level for the
    # duration of its own forward, so that the head of a coarse level runs next to the neck of the finer ones and the whole
    # detection half next to the segmentation half; forward then returns the head outputs in place of (p3, p4, p5).
    det_level_hook = None
    # Serving switch (an attribute for the same reason; vrcoc.InferenceSession sets it): True makes a gradient-free forward return
    # the per-pixel CLASS MAP [B,H,W] uint8 in place of the full-size segmentation logits — the x4 upsample and the arg-max the
    # reference's deeplab.py:149-167 applies to them fused into one kernel (neck.upsample_argmax), bit-identical to
    # forward(...)[1].argmax(1).
    seg_class_map = False

    def forward(self, x, x_radar):
        with ops.sums_arena(x.shape[0], x.device):
            return self._forward(x, x_radar, self.det_level_hook)

    def _forward(self, x, x_radar, det_level):
        x_out, x_radar_out = self.backbone(x, x_radar)
        s2, s3, s4, s5 = x_out
        r2, r3, r4, r5 = x_radar_out

        def seg_branch():          # image features
            t = self.aspp(s5)
            t = self.sc_attn_seg4(cat_shuffle(s4, self.upsample5_4(t)))
            t = self.sc_attn_seg3(cat_shuffle(self.upsample4_3(t), s3))
            t = self.sc_attn_seg2(cat_shuffle(self.upsample3_2(t), s2))
            conv, up = self.upsample2_0.upsample[0], self.upsample2_0.upsample[1]
            if self.seg_class_map and not torch.is_grad_enabled() and getattr(up, "mode", "") == "bilinear" and up.align_corners:
                t = conv(t)
                if upsample_argmax_ok(t, float(up.scale_factor)):
                    return upsample_argmax(t, float(up.scale_factor))
                return up(t).argmax(dim=1).to(torch.uint8)
            return self.upsample2_0(t)

        def det_branch():          # radar features
            p5 = self.p5_out_det(r5)
            if det_level is None:
                p4 = self.p4_out_det(torch.cat([r4, self.p5_4_det(p5)], dim=1))
                p3 = self.p3_out_det(torch.cat([r3, self.p4_3_det(p4)], dim=1))
                return p3, p4, p5
            o5 = ops.Fork(lambda: det_level(2, p5), lane=2)
            p4 = self.p4_out_det(torch.cat([r4, self.p5_4_det(p5)], dim=1))
            o4 = ops.Fork(lambda: det_level(1, p4), lane=3)
            p3 = self.p3_out_det(torch.cat([r3, self.p4_3_det(p4)], dim=1))
            o3 = det_level(0, p3)
            return [o3, o4.join(), o5.join()]

        if x.is_cuda:
            det, seg = ops.run_pair(det_branch, seg_branch)
        else:
            seg = seg_branch()
            det = det_branch()
        return det, seg
