"""CPU, only where the read-only reference mount exists (the dev container): the oracle against the LIVE reference
modules, whole model included.  Skipped on the GPU box, where /root/reference is absent."""
import pytest
import torch

from oracle import coc_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted")


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_whole_model_nano_matches_reference():
    ref_shim.install()
    from nets.efficient_vrnet import EfficientVRNet
    torch.manual_seed(0)
    net = EfficientVRNet(4, 9, "nano").eval()
    x, r = torch.randn(1, 3, 512, 512), torch.rand(1, 4, 512, 512)
    with torch.no_grad():
        det, seg = net(x, r)
        det2, seg2 = O.efficient_vrnet_forward(x, r, net.state_dict(), "nano")
    assert _rel(seg2, seg) < 1e-6
    for a, b in zip(det2, det):
        assert _rel(a, b) < 1e-6


def test_product_modules_match_reference_state_dict():
    """the product's EfficientVRNet has the reference's exact key set, shapes and order"""
    ref_shim.install()
    from nets.efficient_vrnet import EfficientVRNet
    import vrcoc
    ref = EfficientVRNet(4, 9, "nano").state_dict()
    ours = vrcoc.EfficientVRNet(4, 9, "nano")
    assert list(ours.state_dict().keys()) == list(ref.keys())
    assert all(ours.state_dict()[k].shape == v.shape for k, v in ref.items())
    ours.load_state_dict(ref, strict=True)


def test_decode_outputs_matches_reference(monkeypatch):
    """oracle.decode_outputs against the reference's utils/utils_bbox.py:32-84 (its `.cuda(local_rank)` calls neutralised)"""
    ref_shim.install()
    from utils.utils_bbox import decode_outputs as ref_decode
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    g = torch.Generator().manual_seed(3)
    outs = [torch.randn(2, 9, s, s, generator=g) for s in (64, 32, 16)]
    ref = ref_decode([o.clone() for o in outs], (512, 512), 0)
    got = O.decode_outputs(outs, (512, 512))
    assert _rel(got, ref) < 1e-6
