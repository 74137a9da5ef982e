"""YOLOX-style decoupled head: reference head/decouplehead.py:7-88, same constructor, module tree and state-dict keys.
Outside the CoC/fusion hot path (SURVEY §8f "next"): 1x1 stems run on the native GEMM engine through BaseConv, the
depthwise-separable 3x3 towers and the tiny prediction convs are cuDNN library calls."""
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

    def forward_level(self, k, x):
        """one pyramid level (reference decouplehead.py:70-88 loop body); the classification tower runs on a side stream
        next to the regression tower when there is no autograd"""
        from . import ops
        x = self.stems[k](x)
        if x.is_cuda and not torch.is_grad_enabled():
            cls = ops.Fork(lambda: self.cls_preds[k](self.cls_convs[k](x)), lane=5 + k)
        else:
            cls = None
        reg_feat = self.reg_convs[k](x)
        reg_output = self.reg_preds[k](reg_feat)
        obj_output = self.obj_preds[k](reg_feat)
        cls_output = cls.join() if cls is not None else self.cls_preds[k](self.cls_convs[k](x))
        return torch.cat([reg_output, obj_output, cls_output], 1)

    def forward(self, inputs):
        return [self.forward_level(k, x) for k, x in enumerate(inputs)]
