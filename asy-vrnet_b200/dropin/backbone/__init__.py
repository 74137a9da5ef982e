"""Drop-in overlay for the reference's `backbone` package.

Put `asy-vrnet_b200/dropin` and `asy-vrnet_b200` in front of the reference root on sys.path: the reference's
`neck/coc_fpn_dual.py`, `head/decouplehead.py`, `nets/efficient_vrnet.py`, `yolo.py` and `deeplab.py` then import the
B200-native modules below through their usual `from backbone.... import ...` lines, unchanged."""
