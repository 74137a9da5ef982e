"""Whole-model wrapper: reference nets/efficient_vrnet.py:13-27."""
import torch.nn as nn

from .head import DecoupleHead
from .neck import CoCFpnDual


class EfficientVRNet(nn.Module):
    def __init__(self, num_classes, num_seg_classes, phi):
        super().__init__()
        depth_dict = {'nano': 0.33, 'tiny': 0.33, 's': 0.33, 'm': 0.67, 'l': 1.00}
        width_dict = {'nano': 0.25, 'tiny': 0.375, 's': 0.50, 'm': 0.75, 'l': 1.00}
        depth, width = depth_dict[phi], width_dict[phi]
        self.backbone = CoCFpnDual(width=width, num_seg_class=num_seg_classes)
        self.head = DecoupleHead(num_classes, width, depthwise=True)

    def forward(self, x, x_radar):
        # reference: fpn_outs, seg = backbone.forward(x, x_radar); det = head.forward(fpn_outs).  Same calls, except that the head
        # is handed to the neck as a per-level hook so that its levels overlap the rest of the neck (CoCFpnDual.det_level_hook).
        self.backbone.det_level_hook = self.head.forward_level
        try:
            det_outputs, seg_outputs = self.backbone.forward(x, x_radar)
        finally:
            self.backbone.det_level_hook = None
        return det_outputs, seg_outputs
