"""vrcoc: B200-native (sm_100a) implementation of the ASY-VRNet context-cluster block and asymmetric vision<->radar
fusion, behind the reference's nn.Module API.  Importing this package loads libvrcoc.so and fails loudly if it is
missing: there is no CPU or PyTorch fallback for the hot path."""
from . import _lib  # noqa: F401  (loads the shared library)
from ._lib import VrcocError, version  # noqa: F401
from .context_cluster import (Cluster, ClusterBlock, DropPath, GroupNorm, Mlp, PointRecuder, basic_blocks,  # noqa: F401
                              pairwise_cos_sim)
from .fusion import (BaseConv, DWConv, ImageEnhanceByRadar, RadarEnhanceByImage, ShuffleAttention, SiLU,  # noqa: F401
                     data_normal, eca_block, get_activation, shuffle_channels)
from . import losses  # noqa: F401
from .head import DecoupleHead  # noqa: F401
from .neck import ASPP, CoC_Conv, CoCFpnDual, CoCUpsample  # noqa: F401
from .nets import EfficientVRNet  # noqa: F401
from .session import InferenceSession, decode_outputs  # noqa: F401
from .vr_coc import VRCoC, coc_medium, coc_small, coc_tiny, coc_tiny2, replace_pos_buffers  # noqa: F401
