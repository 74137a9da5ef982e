"""reference backbone/conv_utils/normal_conv.py -> vrcoc"""
from vrcoc.fusion import BaseConv, DWConv, SiLU, get_activation  # noqa: F401
