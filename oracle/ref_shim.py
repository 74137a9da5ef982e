"""
Import shim for the UNMODIFIED reference under /root/reference.   *** TEST INFRASTRUCTURE ONLY ***

Used by `oracle/make_golden.py` (fixture generation), by `tests/test_oracle_vs_reference.py` (live re-check when
the read-only reference mount is present) and by the BASELINE legs of `bench.py` (`--impl reference`, `cpu_baseline`,
`reference_eager_gpu`), which import the unmodified reference from the git-ignored copy `baseline/_ref/` that
`__graft_entry__.build()` vendors (it travels to the GPU box; /root/reference does not exist there).  Nothing under
`asy-vrnet_b200/` imports it.

What it does: stubs the python packages the reference imports but this image lacks
(timm, thop, torchinfo), tolerates the `nn.GroupNorm(0, 0)` that `RadarEnhanceByImage(image_in_channels=3)`
constructs (backbone/attention_modules/shuffle_attention.py:15; torch 2.x rejects it, torch 1.9 did not),
and puts the reference root on sys.path.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("VRCOC_REFERENCE_ROOT", "/root/reference")


VENDORED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def available(root=None):
    return os.path.isdir(os.path.join(root or REF_ROOT, "backbone", "fusion"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


def install(root=None):
    """Idempotent.  Returns the reference root (default: the read-only mount; bench.py passes the vendored copy)."""
    root = root or REF_ROOT
    if not available(root):
        raise RuntimeError(f"reference not found at {root}")
    if "timm" not in sys.modules:
        def to_2tuple(v):
            return tuple(v) if isinstance(v, (tuple, list)) else (v, v)
        _stub("timm")
        _stub("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
        _stub("timm.models")
        _stub("timm.models.layers", DropPath=_DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
        _stub("timm.models.registry", register_model=lambda f: f)
        _stub("timm.models.layers.helpers", to_2tuple=to_2tuple)
    if "thop" not in sys.modules:
        _stub("thop", profile=lambda *a, **k: (0, 0), clever_format=lambda v, f=None: v)
    if "torchinfo" not in sys.modules:
        _stub("torchinfo", summary=lambda *a, **k: None)
    if not getattr(nn.GroupNorm, "_vrcoc_zero_ok", False):
        orig = nn.GroupNorm.__init__

        def init(self, num_groups, num_channels, *a, **k):
            if num_groups == 0:
                orig(self, 1, 1, *a, **k)
                self.num_groups, self.num_channels = 0, 0
                if self.affine:
                    self.weight = nn.Parameter(torch.empty(0))
                    self.bias = nn.Parameter(torch.empty(0))
                return
            orig(self, num_groups, num_channels, *a, **k)

        nn.GroupNorm.__init__ = init
        nn.GroupNorm._vrcoc_zero_ok = True
    if root not in sys.path:
        sys.path.insert(0, root)
    return root
