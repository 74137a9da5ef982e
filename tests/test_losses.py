"""CPU: the batched, sync-free detection loss (vrcoc/losses.py) against the reference's own YOLOLoss
(nets/yolo_training.py:60) on the same predictions and labels — loss value and gradients.  Needs the reference tree (the read-only
mount in the dev container, or the vendored baseline/_ref copy on the GPU box)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "asy-vrnet_b200"))
from oracle import ref_shim  # noqa: E402

_REF = ref_shim.REF_ROOT if ref_shim.available() else (ref_shim.VENDORED if ref_shim.available(ref_shim.VENDORED) else None)
pytestmark = pytest.mark.skipif(_REF is None, reason="reference tree not available")


def _case(seed, B, counts, scale=0.6, S=512, nc=4):
    g = torch.Generator().manual_seed(seed)
    maps = [torch.randn(B, 5 + nc, S // s, S // s, generator=g) * scale for s in (8, 16, 32)]
    labels = []
    for n in counts:
        cxcy = torch.rand(n, 2, generator=g) * (S - 112) + 56
        wh = torch.rand(n, 2, generator=g) * 120 + 16
        cls = torch.randint(0, nc, (n, 1), generator=g).float()
        labels.append(torch.cat([cxcy, wh, cls], 1))
    return maps, labels


@pytest.mark.parametrize("seed,counts", [(0, [3, 3, 3, 3]), (1, [1, 5, 0, 2]), (2, [8, 8]), (3, [0, 0]), (4, [2, 7, 4])])
def test_batched_simota_loss_matches_reference(seed, counts):
    ref_shim.install(_REF)
    from nets.yolo_training import YOLOLoss as RefLoss
    from vrcoc.losses import YOLOLoss
    B = len(counts)
    maps, labels = _case(seed, B, counts)
    ours_in = [m.clone().requires_grad_(True) for m in maps]
    ref_in = [m.clone().requires_grad_(True) for m in maps]
    lo = YOLOLoss(4, fp16=True)(ours_in, labels)
    lr = RefLoss(4, True)([m.clone() for m in ref_in], [l.clone() for l in labels])      # (the reference decodes its inputs in place)
    assert torch.isfinite(lo)
    assert abs(lo.item() - lr.item()) <= 1e-5 * max(1.0, abs(lr.item())), (lo.item(), lr.item())
    lo.backward()
    lr.backward()
    for a, b in zip(ours_in, ref_in):
        assert (a.grad - b.grad).abs().max().item() <= 1e-6 + 1e-4 * b.grad.abs().max().item()


def test_assignment_is_identical_on_crowded_boxes():
    """many overlapping boxes on one image: anchors claimed by several boxes go to the cheapest one, exactly as the reference decides"""
    ref_shim.install(_REF)
    from nets.yolo_training import YOLOLoss as RefLoss
    from vrcoc.losses import YOLOLoss
    g = torch.Generator().manual_seed(11)
    maps = [torch.randn(1, 9, 512 // s, 512 // s, generator=g) * 0.8 for s in (8, 16, 32)]
    base = torch.tensor([256.0, 256.0, 140.0, 120.0])
    labels = [torch.cat([base[None, :].repeat(6, 1) + torch.randn(6, 4, generator=g) * 12, torch.randint(0, 4, (6, 1), generator=g).float()], 1)]
    lo = YOLOLoss(4, fp16=True)([m.clone() for m in maps], labels)
    lr = RefLoss(4, True)([m.clone() for m in maps], [l.clone() for l in labels])
    assert abs(lo.item() - lr.item()) <= 1e-5 * max(1.0, abs(lr.item()))
