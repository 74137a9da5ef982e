"""CPU: host-side logic of the reference-shaped modules — constructor/state-dict contract, channel maps, error
behaviour without a GPU, and the world_size-2 batch sharding used by bench.py (gloo)."""
import json
import os
import subprocess
import sys

import pytest
import torch

import vrcoc
from golden_util import Fixture
from vrcoc import fusion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_contract_matches_reference_fixtures():
    """every golden fixture's reference state_dict loads strictly into the mirrored module"""
    cases = {
        "cluster_c16": lambda c: vrcoc.Cluster(**c), "mlp_c16": lambda c: vrcoc.Mlp(**c),
        "block_c16": lambda c: vrcoc.ClusterBlock(**c), "block_neck_default": lambda c: vrcoc.ClusterBlock(**c),
        "shuffle_attention_c32_g4": lambda c: vrcoc.ShuffleAttention(**c), "eca_c32": lambda c: vrcoc.eca_block(**c),
        "point_reducer_k3s2": lambda c: vrcoc.PointRecuder(**c),
        "image_enhance_c8_eval": lambda c: vrcoc.ImageEnhanceByRadar(c["radar_in_channels"], c["image_in_channels"]),
        "radar_enhance_c16_eval": lambda c: vrcoc.RadarEnhanceByImage(c["radar_in_channels"], c["image_in_channels"]),
        "radar_enhance_initial_eval": lambda c: vrcoc.RadarEnhanceByImage(c["radar_in_channels"], c["image_in_channels"], initial=True),
        "vrcoc_mini_eval": lambda c: vrcoc.VRCoC(norm_layer=vrcoc.GroupNorm, **c),
    }
    for name, ctor in cases.items():
        fx = Fixture(name)
        m = ctor(fx.cfg)
        res = m.load_state_dict(fx.sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys, name
        assert list(m.state_dict().keys()) == list(fx.sd.keys()), f"{name}: key order differs from the reference"


def test_whole_model_manifest():
    """887 state-dict entries / 4.137 M parameters for phi='nano', incl. the zero-size attention tensors and the
    persistent position buffers"""
    m = vrcoc.EfficientVRNet(4, 9, "nano")
    sd = m.state_dict()
    assert len(sd) == 887
    assert sum(p.numel() for p in m.parameters()) == 4137016
    assert sd["backbone.backbone.fea_pos"].shape == (512, 512, 2) and sd["backbone.backbone.fea_pos_r"].shape == (512, 512, 2)
    zero = [k for k, v in sd.items() if v.numel() == 0]
    assert len(zero) == 6 and all("radar_enhance_by_image1.image_attn" in k for k in zero)
    manifest = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_nano.json")))
    assert {k: list(v.shape) for k, v in sd.items()} == manifest
    # weights_init-style traversal (reference nets/yolo_training.py:485) sees real Conv children
    convs = [n for n, mod in m.named_modules() if mod.__class__.__name__.find("Conv") != -1 and hasattr(mod, "weight")]
    assert any(n.endswith("token_mixer.fc1") for n in convs) and any(n.endswith("mlp.fc2") for n in convs)
    import copy
    copy.deepcopy(m)      # ModelEMA deep-copies the model (reference nets/yolo_training.py:457)


def test_position_grid_values():
    m = vrcoc.coc_small(width=0.25)
    p = m.fea_pos
    assert p[0, 0, 0] == -0.5 and p[511, 0, 0] == 0.5 and p[0, 511, 1] == 0.5 and p[0, 0, 1] == -0.5
    assert torch.equal(p, m.fea_pos_r)
    assert abs(p[100, 7, 0].item() - (100 / 511.0 - 0.5)) < 1e-7


def test_channel_maps():
    # shuffle_perm is the gather form of shuffle_channels
    for C in (6, 7, 32):
        x = torch.arange(C, dtype=torch.float32).view(1, C, 1, 1)
        perm = fusion.shuffle_perm(C, 2)
        assert torch.equal(fusion.shuffle_channels(x.clone(), 2).flatten(), x.flatten()[perm])
    # RadarEnhanceByImage map == shuffle_channels(cat[channel_shuffle(image), radar]) composed, on labelled channels
    Ci = Cr = 8
    m = vrcoc.RadarEnhanceByImage(Cr, Ci)
    img = torch.arange(Ci, dtype=torch.float32).view(1, Ci, 1, 1)
    rad = (100 + torch.arange(Cr, dtype=torch.float32)).view(1, Cr, 1, 1)
    ia = vrcoc.ShuffleAttention.channel_shuffle(img, 2)
    z = fusion.shuffle_channels(torch.cat([ia, rad], 1), 2).flatten()
    cat = torch.cat([img, rad], 1).flatten()
    assert torch.equal(z, cat[m._chan_src.long()])
    # initial=True: 3 + 4 = 7 channels -> the shuffle is the identity (vr_coc.py:73)
    m0 = vrcoc.RadarEnhanceByImage(4, 3, initial=True)
    assert m0._chan_src.tolist() == list(range(7))
    assert "_chan_src" not in m.state_dict()


def test_no_cpu_fallback():
    blk = vrcoc.ClusterBlock(dim=16, heads=2, head_dim=8)
    with pytest.raises(vrcoc.VrcocError, match="no CPU fallback"):
        blk(torch.randn(1, 16, 8, 8))
    with pytest.raises(vrcoc.VrcocError, match="no CPU fallback"):
        vrcoc.ImageEnhanceByRadar(4, 4)(torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8))
    with pytest.raises(vrcoc.VrcocError, match="deprecated"):
        vrcoc.Cluster(8, 8, return_center=True)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "asy-vrnet_b200", "vrcoc")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# weak scaling: every rank draws its own batch (distinct seeds), identical replicas, no data-path collective
x, r = bench.synth_batch(2, 100 + rank * 10, torch.float32)
gathered = [torch.zeros_like(x[:, :, :4, :4]) for _ in range(world)]
dist.all_gather(gathered, x[:, :, :4, :4].contiguous())
assert not torch.equal(gathered[0], gathered[1]), "ranks must process different frames"
ms = torch.tensor([10.0 + rank])
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
assert ms.item() == 10.0 + world - 1      # value = all frames / max-over-ranks time
print("ok", rank)
dist.destroy_process_group()
'''


def test_world_size_2_sharding_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29571", str(script), ROOT], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_center_tap_identity_and_memo():
    """the shortcut fusion._conv_launch takes when dilation >= map size (ASPP rate 18 on the 16x16 map, reference
    neck/coc_fpn_dual.py:55-67): every off-centre tap of the dilated k x k convolution reads only zero padding, so it IS the 1x1
    projection by its centre tap; and ops.center_tap memoises that slice on the weight (refreshed in place after an update)"""
    import torch.nn.functional as F
    from vrcoc import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 6, 16, 16, generator=g, dtype=torch.float64)
    w = torch.nn.Parameter(torch.randn(5, 6, 3, 3, generator=g, dtype=torch.float64))
    for dil in (16, 18, 40):
        full = F.conv2d(x, w, padding=dil, dilation=dil)
        assert torch.equal(full, F.conv2d(x, w[:, :, 1:2, 1:2]))
    assert not torch.allclose(F.conv2d(x, w, padding=15, dilation=15), F.conv2d(x, w[:, :, 1:2, 1:2]))     # one row / column short
    c = ops.center_tap(w)
    assert c.shape == (5, 6) and c.is_contiguous() and torch.equal(c, w.detach()[:, :, 1, 1])
    assert ops.center_tap(w) is c                                   # memo hit
    with torch.no_grad():
        w.mul_(2.0)                                                  # bumps the version counter
    c2 = ops.center_tap(w)
    assert c2 is c and torch.equal(c2, w.detach()[:, :, 1, 1])      # same storage (graph-capture contract), new values


def test_row_tap_decomposition_matches_conv():
    """the algebra behind the convolution engine's row-tap mode (include/vrcoc.h, k_order 2): a 3x3 / pad d / dilation d convolution
    = a 3x1 convolution (vertical taps, rows zero-padded) over the horizontal-tap expansion cols[b, kx*C + c, y, x] = x[b, c, y,
    x + (kx-1)*d] with the tap-major weight [O][ky][kx*C + c] - the layout vrcoc_im2col_rows writes and ops.tap_major produces"""
    import torch.nn.functional as F
    from vrcoc import ops
    g = torch.Generator().manual_seed(8)
    B, C, O, H, W = 2, 4, 3, 6, 8
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(O, C, 3, 3, generator=g, dtype=torch.float64)
    for d in (1, 2):
        xp = F.pad(x, (d, d, 0, 0))
        cols = torch.cat([xp[..., kx * d: kx * d + W] for kx in range(3)], 1)                  # [B, 3C, H, W]
        wt = ops.tap_major(w).view(O, 3, 3 * C)                                                 # [O][ky][kx*C + c]
        got = F.conv2d(cols, wt.permute(0, 2, 1).reshape(O, 3 * C, 3, 1), padding=(d, 0), dilation=(d, 1))
        assert torch.allclose(got, F.conv2d(x, w, padding=d, dilation=d), atol=1e-12)


def test_arena_lanes_are_unique():
    """every concurrent pipeline slot of InferenceSession captures its forward under its own statistics-arena lane"""
    from vrcoc import ops
    a, b = ops.sums_arena.new_lane(), ops.sums_arena.new_lane()
    assert a != b and a > 0 and b > 0 and ops.sums_arena.lane == 0
