"""reference backbone/radar/context_cluster.py (same API, named in the north star; imported by nothing in the reference) -> vrcoc"""
from vrcoc.context_cluster import (Cluster, ClusterBlock, DropPath, GroupNorm, Mlp, PointRecuder, basic_blocks,  # noqa: F401
                                   pairwise_cos_sim, to_2tuple)
