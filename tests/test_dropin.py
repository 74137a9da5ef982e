"""CPU, only where the reference mount exists: the UNMODIFIED reference `nets/efficient_vrnet.py` / `neck/coc_fpn_dual.py`
/ `head/decouplehead.py` import and build on top of the drop-in `backbone` overlay (separate process: the overlay and the
reference share top-level package names)."""
import os
import subprocess
import sys

import pytest

from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted")

_SCRIPT = r'''
import sys
root, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
from oracle import ref_shim
ref_shim.install()                      # stubs timm/thop/torchinfo and puts the reference root on sys.path ...
sys.path.insert(0, root + "/asy-vrnet_b200")
sys.path.insert(0, root + "/asy-vrnet_b200/dropin")          # ... the overlay goes in front of it
import backbone.fusion.vr_coc as V
assert V.__file__.startswith(root), V.__file__
from nets.efficient_vrnet import EfficientVRNet                # the reference's own file
import nets.efficient_vrnet as N, neck.coc_fpn_dual as K, head.decouplehead as Hd
assert N.__file__.startswith(ref) and K.__file__.startswith(ref) and Hd.__file__.startswith(ref)
import vrcoc
m = EfficientVRNet(4, 9, "nano")
blocks = [x for x in m.modules() if x.__class__.__name__ == "ClusterBlock"]
assert len(blocks) == 27 and all(isinstance(b, vrcoc.ClusterBlock) for b in blocks)
ours = vrcoc.EfficientVRNet(4, 9, "nano")
assert list(m.state_dict().keys()) == list(ours.state_dict().keys())
m.load_state_dict(ours.state_dict(), strict=True)
import copy; copy.deepcopy(m)
print("dropin-ok", len(m.state_dict()))
'''


def test_reference_files_run_on_the_overlay(tmp_path):
    script = tmp_path / "d.py"
    script.write_text(_SCRIPT)
    r = subprocess.run([sys.executable, str(script), ROOT, ref_shim.REF_ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "dropin-ok 887" in r.stdout
