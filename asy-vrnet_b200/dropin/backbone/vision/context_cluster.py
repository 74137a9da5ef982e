"""reference backbone/vision/context_cluster.py (the copy neck/coc_fpn_dual.py:10 imports ClusterBlock from) -> vrcoc"""
from vrcoc.context_cluster import (Cluster, ClusterBlock, DropPath, GroupNorm, Mlp, PointRecuder, basic_blocks,  # noqa: F401
                                   pairwise_cos_sim, to_2tuple)
