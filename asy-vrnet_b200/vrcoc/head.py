"""YOLOX-style decoupled head: reference head/decouplehead.py:7-88, same constructor, module tree and state-dict keys.
Outside the CoC/fusion hot path (SURVEY §8f "next"), all on the native kernels: the 1x1 stems and the depthwise-separable 3x3
towers through BaseConv (vrcoc_dwconv + the GEMM engine); gradient-free, the three prediction convs of a level and the
torch.cat of their outputs (decouplehead.py:80-87) are ONE two-source projection on the GEMM engine writing [reg | obj | cls] directly.  With
autograd the prediction convs are the nn.Conv2d modules (library calls; their parameters train through them)."""
import torch
import torch.nn as nn

from .fusion import BaseConv


class DecoupleHead(nn.Module):
    def __init__(self, num_classes, width=1.0, in_channels=[128, 320, 512], act="relu", depthwise=False):
        super().__init__()
        Conv = BaseConv
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        self.cls_preds = nn.ModuleList()
        self.reg_preds = nn.ModuleList()
        self.obj_preds = nn.ModuleList()
        self.stems = nn.ModuleList()
        c = int(256 * width)
        for i in range(len(in_channels)):
            self.stems.append(Conv(in_channels=int(in_channels[i] * width), out_channels=c, ksize=1, stride=1, act=act))
            self.cls_convs.append(nn.Sequential(Conv(c, c, ksize=3, stride=1, act=act, ds_conv=True),
                                                Conv(c, c, ksize=3, stride=1, act=act, ds_conv=True)))
            self.cls_preds.append(nn.Conv2d(c, num_classes, kernel_size=1, stride=1, padding=0))
            self.reg_convs.append(nn.Sequential(Conv(c, c, ksize=3, stride=1, act=act, ds_conv=True),
                                                Conv(c, c, ksize=3, stride=1, act=act, ds_conv=True)))
            self.reg_preds.append(nn.Conv2d(c, 4, kernel_size=1, stride=1, padding=0))
            self.obj_preds.append(nn.Conv2d(c, 1, kernel_size=1, stride=1, padding=0))

    def _pred_weights(self, k, dtype):
        """[4 + 1 + nc, 2c] block weight over the virtual concat [reg_feat | cls_feat] (zero where a prediction does not read a
        tower) and the bias, outputs in the order of the reference's torch.cat([reg_output, obj_output, cls_output], 1)"""
        from . import ops
        rp, op, cp = self.reg_preds[k], self.obj_preds[k], self.cls_preds[k]

        def build():
            c = rp.weight.shape[1]
            nr, no, ncls = rp.weight.shape[0], op.weight.shape[0], cp.weight.shape[0]
            w = torch.zeros(nr + no + ncls, 2 * c, device=rp.weight.device, dtype=dtype)
            w[:nr, :c] = rp.weight.detach().reshape(nr, c)
            w[nr:nr + no, :c] = op.weight.detach().reshape(no, c)
            w[nr + no:, c:] = cp.weight.detach().reshape(ncls, c)
            b = torch.cat([m.bias.detach().float() if m.bias is not None else torch.zeros(m.weight.shape[0], device=w.device)
                           for m in (rp, op, cp)])
            return w, b.contiguous()
        return ops.cached(self, f"pred_w{k}_{dtype}", [rp.weight, rp.bias, op.weight, op.bias, cp.weight, cp.bias], build)

    def forward_level(self, k, x):
        """one pyramid level (reference decouplehead.py:70-88 loop body); the classification tower runs on a side stream
        next to the regression tower when there is no autograd"""
        from . import ops
        x = self.stems[k](x)
        if x.is_cuda and not torch.is_grad_enabled() and x.dtype in (torch.bfloat16, torch.float32):
            cls = ops.Fork(lambda: self.cls_convs[k](x), lane=5 + k)
            reg_feat = self.reg_convs[k](x).contiguous()
            cls_feat = cls.join().contiguous()
            # one two-source projection on the GEMM engine (same-box A/B against the three library convs + cat: 2836 vs 2832
            # frames/s; a CUDA-core kernel for the 9 thin outputs was slower than both, 2821)
            w, b = self._pred_weights(k, x.dtype)
            B, _, H, W = reg_feat.shape
            out = torch.empty(B, w.shape[0], H, W, device=x.device, dtype=x.dtype)
            ops.conv_fwd(ops.conv_desc(reg_feat, w, out, src1=cls_feat, e_shift=b))
            return out
        cls = None
        reg_feat = self.reg_convs[k](x)
        reg_output = self.reg_preds[k](reg_feat)
        obj_output = self.obj_preds[k](reg_feat)
        cls_output = cls.join() if cls is not None else self.cls_preds[k](self.cls_convs[k](x))
        return torch.cat([reg_output, obj_output, cls_output], 1)

    def forward(self, inputs):
        return [self.forward_level(k, x) for k, x in enumerate(inputs)]
