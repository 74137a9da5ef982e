"""reference backbone/attention_modules/shuffle_attention.py -> vrcoc"""
from vrcoc.fusion import ShuffleAttention  # noqa: F401
