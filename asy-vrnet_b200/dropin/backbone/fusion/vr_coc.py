"""reference backbone/fusion/vr_coc.py -> vrcoc (same names, constructors, forward signatures, state-dict keys)"""
from vrcoc.context_cluster import (Cluster, ClusterBlock, DropPath, GroupNorm, Mlp, PointRecuder, basic_blocks,  # noqa: F401
                                   pairwise_cos_sim, to_2tuple)
from vrcoc.fusion import (BaseConv, DWConv, ImageEnhanceByRadar, RadarEnhanceByImage, ShuffleAttention,  # noqa: F401
                          data_normal, eca_block, shuffle_channels)
from vrcoc.vr_coc import (VRCoC, coc_medium, coc_small, coc_tiny, coc_tiny2, default_cfgs, register_model)  # noqa: F401
