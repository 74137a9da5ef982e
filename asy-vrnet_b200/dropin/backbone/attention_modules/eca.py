"""reference backbone/attention_modules/eca.py -> vrcoc"""
from vrcoc.fusion import eca_block  # noqa: F401
