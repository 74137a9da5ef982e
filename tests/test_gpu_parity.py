"""GPU parity tests proper: the CUDA path (through the C-ABI, via the reference-shaped nn.Modules) against
 (a) the committed golden vectors produced by the unmodified reference, and
 (b) the CPU oracle on seeded inputs at the live configurations,
plus size-independent properties at full BASELINE sizes.

Gates (BASELINE.json north_star): relative L2 error <= 1e-4 in fp32 and <= 2e-2 in bf16 on outputs and gradients;
cluster-assignment indices bit-exact wherever the oracle's top-2 similarity margin exceeds 1e-5."""
import pytest
import torch

from golden_util import Fixture, names, rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2
MARGIN = 1e-5


@pytest.fixture(scope="module")
def V():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import vrcoc
    from vrcoc import _lib
    assert _lib.lib.vrcoc_device_ok() == 1, "not an sm_100 device"
    return vrcoc


def cu(t, dtype=torch.float32):
    return t.to("cuda", dtype) if t.is_floating_point() else t.to("cuda")


# ---------------------------------------------------------------------------------------------------------------
# cluster core vs golden
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", names("core_"))
def test_core_golden_fp32(V, name):
    from vrcoc import ops
    fx = Fixture(name)
    c = fx.cfg
    feat, value = cu(fx.inp["feat"]), cu(fx.inp["value"])
    alpha, beta = cu(fx.inp["alpha"]), cu(fx.inp["beta"])
    args = (c["E"], c["fold_w"], c["fold_h"], c["proposal_w"], c["proposal_h"])
    out, idx, smax = ops.cluster_core_fwd(feat, value, alpha, beta, *args, save_aux=True)
    assert rel_err(out, fx.out["y"]) < FP32_TOL
    safe = fx.out["margin"] > MARGIN
    got_idx = idx.cpu().to(torch.int32)
    B, E = c["B"], c["E"]
    ref_idx = fx.out["idx"].reshape(got_idx.shape)
    assert torch.equal(got_idx[safe.reshape(got_idx.shape)], ref_idx[safe.reshape(got_idx.shape)]), \
        "assignment mismatch outside the 1e-5 margin"
    assert safe.float().mean() > 0.99
    assert rel_err(smax.reshape(-1), fx.out["sim_max"].reshape(-1)) < FP32_TOL
    # backward
    dfeat, dvalue, dab = ops.cluster_core_bwd(feat, value, cu(fx.inp["gout"]), idx, smax, alpha, beta, *args)
    assert rel_err(dvalue, fx.grad_in["value"]) < FP32_TOL
    assert rel_err(dfeat, fx.grad_in["feat"]) < FP32_TOL
    assert abs(dab[0].item() - fx.grad_in["alpha"].item()) <= FP32_TOL * max(1.0, abs(fx.grad_in["alpha"].item())) * 10
    assert abs(dab[1].item() - fx.grad_in["beta"].item()) <= FP32_TOL * max(1.0, abs(fx.grad_in["beta"].item())) * 10


@pytest.mark.parametrize("name", ["core_16x16_f2_p2", "core_live_s1", "core_8x8_d24", "core_16x16_d24"])
def test_core_golden_bf16_storage(V, name):
    """bf16 value/out storage with fp32 similarity operand (the layout the bf16 block uses): same assignments as the
    fp32 reference because feat is untouched; outputs within the bf16 gate."""
    from vrcoc import ops
    fx = Fixture(name)
    c = fx.cfg
    feat = cu(fx.inp["feat"])
    value = cu(fx.inp["value"], torch.bfloat16)
    alpha, beta = cu(fx.inp["alpha"]), cu(fx.inp["beta"])
    out, idx, smax = ops.cluster_core_fwd(feat, value, alpha, beta, c["E"], c["fold_w"], c["fold_h"], c["proposal_w"],
                                          c["proposal_h"], save_aux=True)
    assert out.dtype == torch.bfloat16
    assert rel_err(out.float(), fx.out["y"]) < BF16_TOL
    safe = (fx.out["margin"] > MARGIN).reshape(idx.shape)
    assert torch.equal(idx.cpu().to(torch.int32)[safe], fx.out["idx"].reshape(idx.shape)[safe])


# ---------------------------------------------------------------------------------------------------------------
# modules vs golden (forward + gradients)
# ---------------------------------------------------------------------------------------------------------------
def _run_module(mod, fx, inputs, dtype=torch.float32, train=None):
    mod = mod.to("cuda", dtype)
    missing = mod.load_state_dict(fx.sd_as(dtype, "cuda"), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    if train is not None:
        mod.train(train)
    xs = [cu(fx.inp[k], dtype).requires_grad_("gout" in fx.inp) for k in inputs]
    y = mod(*xs)
    return mod, xs, y


def _check_grads(mod, fx, xs, inputs, y, tol, dtype=torch.float32):
    y.backward(cu(fx.inp["gout"], dtype))
    for x, k in zip(xs, inputs):
        assert rel_err(x.grad.float(), fx.grad_in[k]) < tol, f"d/d{k}"
    params = dict(mod.named_parameters())
    for k, g in fx.grad_sd.items():
        got = params[k].grad
        assert got is not None, k
        ref = g
        denom = ref.double().norm().item()
        if denom < 1e-12:
            assert got.double().norm().item() < 1e-6, k
        else:
            assert rel_err(got.float(), ref) < tol * (3 if ref.numel() <= 2 else 1), k


MODULE_CASES = [
    ("cluster_c16", "Cluster", ["x"]),
    ("cluster_vision_c24", "Cluster", ["x"]),
    ("mlp_c16", "Mlp", ["x"]),
    ("block_c16", "ClusterBlock", ["x"]),
    ("block_live_s1_small", "ClusterBlock", ["x"]),
    ("block_neck_default", "ClusterBlock", ["x"]),
]


@pytest.mark.parametrize("name,cls,inputs", MODULE_CASES)
def test_module_golden_fp32(V, name, cls, inputs):
    fx = Fixture(name)
    mod, xs, y = _run_module(getattr(V, cls)(**fx.cfg), fx, inputs)
    assert rel_err(y, fx.out["y"]) < FP32_TOL
    _check_grads(mod, fx, xs, inputs, y, FP32_TOL)


def test_block_stock_layer_scale(V):
    """stock 1e-5 layer scale: outputs are ~input; the gate is on the tiny increment too"""
    fx = Fixture("block_stock_ls")
    mod, xs, y = _run_module(V.ClusterBlock(**fx.cfg), fx, ["x"])
    assert rel_err(y, fx.out["y"]) < 1e-6
    inc, ref_inc = y.detach().double().cpu() - fx.inp["x"].double(), fx.out["y"] - fx.inp["x"].double()
    assert rel_err(inc, ref_inc) < 1e-1      # fp32 resolution: the increment is ~1e-6 of an O(1) signal (eps/1e-6 ~ 6e-2)
    _check_grads(mod, fx, xs, ["x"], y, FP32_TOL)


@pytest.mark.parametrize("name,cls,inputs", [c for c in MODULE_CASES if c[0] in ("cluster_c16", "mlp_c16", "block_c16", "block_live_s1_small")])
def test_module_golden_bf16(V, name, cls, inputs):
    """bf16 oracle = the fp32 reference on the same bf16-rounded inputs/weights; here the golden
    fp32-input result is used and the tolerance covers the input rounding (2e-2 gate)."""
    fx = Fixture(name)
    mod, xs, y = _run_module(getattr(V, cls)(**fx.cfg), fx, inputs, dtype=torch.bfloat16)
    assert y.dtype == torch.bfloat16
    assert rel_err(y.float(), fx.out["y"]) < BF16_TOL
    y.backward(cu(fx.inp["gout"], torch.bfloat16))
    assert rel_err(xs[0].grad.float(), fx.grad_in["x"]) < 3 * BF16_TOL


def test_leaves_golden(V):
    fx = Fixture("shuffle_attention_c32_g4")
    mod, xs, y = _run_module(V.ShuffleAttention(channel=32, G=4), fx, ["x"])
    assert rel_err(y, fx.out["y"]) < FP32_TOL
    _check_grads(mod, fx, xs, ["x"], y, FP32_TOL)

    fx = Fixture("eca_c32")
    mod, xs, y = _run_module(V.eca_block(channel=32), fx, ["x"])
    assert rel_err(y, fx.out["y"]) < FP32_TOL
    _check_grads(mod, fx, xs, ["x"], y, FP32_TOL)

    fx = Fixture("point_reducer_k3s2")
    mod, xs, y = _run_module(V.PointRecuder(**fx.cfg), fx, ["x"])
    assert rel_err(y, fx.out["y"]) < FP32_TOL
    _check_grads(mod, fx, xs, ["x"], y, FP32_TOL)


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_base_conv_golden(V, mode):
    fx = Fixture(f"base_conv_k3_{mode}")
    c = fx.cfg
    mod, xs, y = _run_module(V.BaseConv(c["in_channels"], c["out_channels"], c["ksize"], c["stride"]), fx, ["x"],
                             train=c["training"])
    assert rel_err(y, fx.out["y"]) < FP32_TOL
    if c["training"]:
        sd = mod.state_dict()
        for k in ("bn.running_mean", "bn.running_var"):
            assert rel_err(sd[k], fx.out["new." + k]) < FP32_TOL, k
        assert int(sd["bn.num_batches_tracked"]) == int(fx.out["new.bn.num_batches_tracked"]) if "new.bn.num_batches_tracked" in fx.out else True
    _check_grads(mod, fx, xs, ["x"], y, FP32_TOL)


@pytest.mark.parametrize("name", names("image_enhance_") + names("radar_enhance_"))
def test_fusion_golden(V, name):
    fx = Fixture(name)
    c = fx.cfg
    if name.startswith("image_enhance"):
        m = V.ImageEnhanceByRadar(radar_in_channels=c["radar_in_channels"], image_in_channels=c["image_in_channels"])
    else:
        m = V.RadarEnhanceByImage(radar_in_channels=c["radar_in_channels"], image_in_channels=c["image_in_channels"],
                                  initial=c["initial"])
    mod, xs, y = _run_module(m, fx, ["image", "radar"], train=c["training"])
    assert rel_err(y, fx.out["y"]) < FP32_TOL
    if c["training"]:
        sd = mod.state_dict()
        for k, v in fx.out.items():
            if k.startswith("new.") and "running" in k:
                assert rel_err(sd[k[4:]], v) < FP32_TOL, k
    if "gout" in fx.inp:
        _check_grads(mod, fx, xs, ["image", "radar"], y, FP32_TOL)


@pytest.mark.parametrize("Ci,Cr,H,W,B", [(64, 64, 32, 32, 2), (128, 128, 16, 24, 3), (320, 320, 8, 8, 2), (64, 128, 16, 16, 1), (512, 512, 16, 16, 2),
                                         (320, 320, 32, 32, 2)])
def test_radar_enhance_concat_order_path_vs_oracle(V, Ci, Cr, H, W, B):
    """gradient-free bf16 RadarEnhanceByImage with the image/radar boundary on a 64-channel slab: the shuffle is moved to the
    weight columns and the projection reads [image | radar] in memory order (channel-major tcgen05 kernel, both sources by
    TMA; 640 / 1024 input channels: the prologue as a streaming pass + the TMA-only GEMM) - against the fp64 oracle on the same
    bf16-rounded weights and inputs"""
    from oracle import coc_oracle as O
    torch.manual_seed(5)
    m = V.RadarEnhanceByImage(radar_in_channels=Cr, image_in_channels=Ci).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.numel() and n.split(".")[-1] in ("cweight", "cbias", "sweight", "sbias"):
                p.normal_()
        for n, b in m.named_buffers():
            if n.endswith("running_var"):
                b.uniform_(0.5, 1.5)
            elif n.endswith("running_mean"):
                b.normal_(0, 0.2)
    m = m.to(torch.bfloat16)
    img = torch.randn(B, Ci, H, W).to(torch.bfloat16)
    rad = torch.rand(B, Cr, H, W).to(torch.bfloat16)
    sd = {k: v.double() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.radar_enhance_by_image(img.double(), rad.double(), sd, "")
        got = m.cuda()(img.cuda(), rad.cuda())
        got2 = m(img.cuda(), rad.cuda())
    assert got.dtype == torch.bfloat16
    assert rel_err(got.float(), ref) < BF16_TOL
    assert torch.equal(got, got2)


def test_vrcoc_mini_golden(V):
    fx = Fixture("vrcoc_mini_eval")
    m = V.VRCoC(norm_layer=V.GroupNorm, **fx.cfg).to("cuda").eval()
    m.load_state_dict(fx.sd_as(torch.float32, "cuda"), strict=True)
    with torch.no_grad():
        outs, outs_r = m(cu(fx.inp["x"]), cu(fx.inp["x_radar"]))
    for i in range(4):
        assert rel_err(outs[i], fx.out[f"img{i}"]) < FP32_TOL, f"img{i}"
        assert rel_err(outs_r[i], fx.out[f"radar{i}"]) < FP32_TOL, f"radar{i}"


# ---------------------------------------------------------------------------------------------------------------
# live configurations vs the CPU oracle (seeded), SURVEY §8 table
# ---------------------------------------------------------------------------------------------------------------
LIVE = {  # id: (C, H, fold, heads, head_dim, mlp_ratio)
    "S1": (64, 128, 8, 4, 32, 8), "S2": (128, 64, 4, 4, 32, 8), "S3": (320, 32, 2, 8, 32, 4), "S4": (512, 16, 1, 8, 32, 4),
    "N5": (512, 16, 2, 4, 24, 4), "N4": (640, 32, 2, 4, 24, 4), "N3": (256, 64, 2, 4, 24, 4),
}


def _seeded_block(V, cid, dtype=torch.float32):
    from oracle import coc_oracle as O  # noqa: F401
    C, H, fold, heads, hd, r = LIVE[cid]
    torch.manual_seed(0)
    m = V.ClusterBlock(dim=C, mlp_ratio=float(r), fold_w=fold, fold_h=fold, heads=heads, head_dim=hd)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "layer_scale" in n:
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif n.endswith("sim_alpha"):
                p.copy_(torch.rand(1, generator=g) * 1.5 + 0.5)
            elif n.endswith("sim_beta"):
                p.copy_(torch.rand(1, generator=g) - 0.5)
            elif "norm" in n and n.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif p.ndim == 4:
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / p.shape[1]) ** 0.5)
    x = torch.randn(2, C, H, H, generator=g)
    return m, x, (heads, fold, fold, 2, 2)


@pytest.mark.parametrize("cid", list(LIVE))
def test_live_block_vs_oracle_fp32(V, cid):
    from oracle import coc_oracle as O
    m, x, (heads, fw, fh, pw, ph) = _seeded_block(V, cid)
    sd = {k: v.double() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.cluster_block(x.double(), sd, "", heads, fw, fh, pw, ph)
        got = m.cuda()(x.cuda())
    assert rel_err(got, ref) < FP32_TOL


@pytest.mark.parametrize("cid", list(LIVE))
def test_live_block_vs_oracle_bf16(V, cid):
    from oracle import coc_oracle as O
    m, x, (heads, fw, fh, pw, ph) = _seeded_block(V, cid)
    m = m.to(torch.bfloat16)
    xb = x.to(torch.bfloat16)
    sd = {k: v.double() for k, v in m.state_dict().items()}          # the bf16-rounded weights, evaluated in fp64
    with torch.no_grad():
        ref = O.cluster_block(xb.double(), sd, "", heads, fw, fh, pw, ph)
        got = m.cuda()(xb.cuda())
    assert rel_err(got.float(), ref) < BF16_TOL


# ---------------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE sizes; no oracle needed)
# ---------------------------------------------------------------------------------------------------------------
def test_core_properties_full_size(V):
    """S1 at B=8: (1) linear in `value` for fixed `feat` (assignments depend on feat only), (2) samples independent,
    (3) dispatch invariant: every point of a region/head assigned to the same centre gets out/sim_max equal."""
    from vrcoc import ops
    torch.manual_seed(3)
    B, E, D, H = 8, 4, 32, 128
    feat = torch.randn(B, E * D, H, H, device="cuda")
    v1 = torch.randn(B, E * D, H, H, device="cuda")
    v2 = torch.randn(B, E * D, H, H, device="cuda")
    a = torch.tensor([1.3], device="cuda")
    b = torch.tensor([-0.2], device="cuda")
    args = (E, 8, 8, 2, 2)
    o1, idx, smax = ops.cluster_core_fwd(feat, v1, a, b, *args, save_aux=True)
    o2, _, _ = ops.cluster_core_fwd(feat, v2, a, b, *args)
    o12, idx2, _ = ops.cluster_core_fwd(feat, 0.5 * v1 - 2.0 * v2, a, b, *args, save_aux=True)
    assert torch.equal(idx, idx2)
    assert rel_err(o12, 0.5 * o1 - 2.0 * o2) < 1e-5
    o_half, _, _ = ops.cluster_core_fwd(feat[4:].contiguous(), v1[4:].contiguous(), a, b, *args)
    assert torch.equal(o_half, o1[4:])
    # dispatch invariant on one region/head
    r = (o1[0, :D, :16, :16] / smax[0, 0, :16, :16]).reshape(D, -1)
    k = idx[0, 0, :16, :16].reshape(-1)
    for m in range(4):
        sel = r[:, k == m]
        if sel.shape[1] > 1:
            assert (sel - sel[:, :1]).abs().max() < 1e-4 * sel.abs().max()
    assert int(idx.max()) <= 3


@pytest.mark.parametrize("E,D,H,fold", [(4, 32, 128, 8), (8, 32, 32, 2), (8, 32, 16, 1), (4, 24, 32, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_core_compile_time_kernel_vs_general_fast_path(V, E, D, H, fold, dtype):
    """The compile-time 16x16 kernel (cluster_core_fast2.cu) against the run-time-geometry TMA kernel it replaces, at the
    live sizes S1 / S3 / S4 / N4 with B=8 (S1: several region-heads per persistent CTA): identical assignments, same
    sim_max and outputs up to summation order / 3-ulp sigmoid."""
    import os
    from vrcoc import ops
    torch.manual_seed(11)
    B = 8
    feat = torch.randn(B, E * D, H, H, device="cuda")
    value = torch.randn(B, E * D, H, H, device="cuda").to(dtype)
    a = torch.tensor([0.9], device="cuda")
    b = torch.tensor([0.15], device="cuda")
    args = (E, fold, fold, 2, 2)
    assert os.environ.get("VRCOC_CORE_FAST2") is None
    o_new, idx_new, s_new = ops.cluster_core_fwd(feat, value, a, b, *args, save_aux=True)
    o_inf, _, _ = ops.cluster_core_fwd(feat, value, a, b, *args)            # without the auxiliary outputs
    os.environ["VRCOC_CORE_FAST2"] = "0"
    try:
        o_old, idx_old, s_old = ops.cluster_core_fwd(feat, value, a, b, *args, save_aux=True)
    finally:
        del os.environ["VRCOC_CORE_FAST2"]
    torch.cuda.synchronize()
    assert torch.equal(o_new, o_inf)
    assert torch.equal(idx_new, idx_old)
    assert rel_err(s_new, s_old) < 1e-6
    assert rel_err(o_new.float(), o_old.float()) < (1e-5 if dtype == torch.float32 else 4e-3)
    assert (o_new.float() - o_old.float()).abs().max() <= (1e-5 if dtype == torch.float32 else 2 ** -7) * o_old.float().abs().max()


def test_reference_assert_is_mirrored(V):
    """Cluster raises like the reference's assert (vr_coc.py:163) when the map is not divisible by the fold"""
    m = V.Cluster(8, 8, fold_w=4, fold_h=4, heads=2, head_dim=4).cuda()
    with pytest.raises(RuntimeError, match="can be divided by fold"):
        m(torch.randn(1, 8, 10, 10, device="cuda"))


# ---------------------------------------------------------------------------------------------------------------
# whole model (backbone + neck + head) vs the CPU oracle, and the neck's native upsample
# ---------------------------------------------------------------------------------------------------------------
def test_upsample_matches_interpolate(V):
    import torch.nn.functional as F
    from vrcoc.neck import BilinearUpsample
    g = torch.Generator().manual_seed(4)
    for shape, scale in (((2, 9, 32, 32), 4), ((1, 5, 7, 12), 2), ((2, 3, 16, 8), 2)):
        x = torch.randn(*shape, generator=g).cuda()
        ref = F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=True)
        got = BilinearUpsample(scale_factor=scale, mode="bilinear", align_corners=True)(x)
        assert rel_err(got, ref) < 1e-6
        gb = BilinearUpsample(scale_factor=scale, mode="bilinear", align_corners=True)(x.bfloat16())
        assert rel_err(gb.float(), ref) < 5e-3
    xr = torch.randn(1, 2, 8, 8, generator=g).cuda().requires_grad_(True)
    BilinearUpsample(scale_factor=2, mode="bilinear", align_corners=True)(xr).square().sum().backward()
    xr2 = xr.detach().clone().requires_grad_(True)
    F.interpolate(xr2, scale_factor=2, mode="bilinear", align_corners=True).square().sum().backward()
    assert rel_err(xr.grad, xr2.grad) < 1e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_point_reducer_emits_groupnorm_sums(V, dtype):
    """gradient-free PointRecuder: the per-sample sums riding on the output (conv epilogue side channel) are the sums of that
    output (up to its own storage rounding), for the patch embed with the position grid and for the 3x3 stride-2 reducers"""
    from vrcoc import ops
    g = torch.Generator().manual_seed(5)
    for cin, cout, k, s, p, H, extra in ((3, 64, 4, 4, 0, 128, 2), (64, 128, 3, 2, 1, 64, 0), (320, 512, 3, 2, 1, 32, 0)):
        m = V.PointRecuder(patch_size=k, stride=s, padding=p, in_chans=cin + extra, embed_dim=cout).cuda().to(dtype)
        x = torch.randn(3, cin, H, H, generator=g).cuda().to(dtype)
        pos = torch.randn(extra, H, H, generator=g).cuda().to(dtype) if extra else None
        with torch.no_grad():
            y = m(x, extra=pos)
        ss = ops.sample_sums_of(y)
        assert ss is y._vrcoc_sums
        got = ss.sum(dim=1)                                               # [B, 2] over the slots
        ref = torch.stack([y.double().sum(dim=(1, 2, 3)), y.double().square().sum(dim=(1, 2, 3))], dim=1)
        tol = 2e-3 if dtype == torch.bfloat16 else 1e-5
        n = y[0].numel()
        assert ((got[:, 0] - ref[:, 0]).abs() / n <= tol * y.double().abs().mean()).all()
        assert rel_err(got[:, 1], ref[:, 1]) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_head_level_fused_predictions(V, dtype, tol):
    """DecoupleHead.forward_level without autograd (towers + ONE prediction launch) against the same modules with autograd
    enabled (reg_preds / obj_preds / cls_preds as nn.Conv2d + torch.cat, decouplehead.py:80-87)"""
    torch.manual_seed(2)
    head = V.DecoupleHead(4, 1.0, in_channels=[128, 320, 512], depthwise=True).cuda().to(dtype).eval()
    g = torch.Generator().manual_seed(9)
    for k, (c, h) in enumerate(((128, 64), (320, 32), (512, 16))):
        x = torch.randn(2, c, h, h, generator=g).cuda().to(dtype)
        with torch.no_grad():
            got = head.forward_level(k, x)
        ref = head.forward_level(k, x).detach()                         # grad mode: the library path of the prediction convs
        assert got.shape == ref.shape == (2, 9, h, h)
        assert rel_err(got.float(), ref.float()) < tol


@pytest.mark.parametrize("shape", [(2, 16, 32, 32), (1, 8, 16, 24), (2, 5, 8, 8)])
def test_col2im_3x3_stride1_bf16_is_fold(V, shape):
    """the 128-bit col2im path (3x3, stride 1, pad 1, bf16): dx = the adjoint of the tap-major im2col = F.fold of the same columns"""
    import torch.nn.functional as F
    from vrcoc import ops
    B, C, H, W = shape
    g = torch.Generator().manual_seed(13)
    dcol = torch.randn(B, 9 * C, H, W, generator=g).cuda().bfloat16()
    dx = torch.empty(B, C, H, W, device="cuda", dtype=torch.bfloat16)
    ops.check(ops.lib.vrcoc_col2im(dcol.data_ptr(), dx.data_ptr(), 1, B, C, H, W, 3, 3, 1, 1, 1, ops._stream()), "col2im")
    cols = dcol.float().view(B, 9, C, H * W).permute(0, 2, 1, 3).reshape(B, C * 9, H * W)       # tap-major -> F.fold's channel-major
    ref = F.fold(cols, output_size=(H, W), kernel_size=3, padding=1, stride=1)
    assert (dx.float() - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()          # one bf16 rounding of a 9-term fp32 sum


@pytest.mark.parametrize("shape", [(2, 16, 32, 32), (1, 8, 16, 24), (2, 5, 8, 8)])
def test_col2im_3x3_stride2_bf16_is_fold(V, shape):
    """the 64-bit-load col2im path of the point reducers (3x3, stride 2, pad 1, bf16) = F.fold of the same columns"""
    import torch.nn.functional as F
    from vrcoc import ops
    B, C, H, W = shape
    Ho, Wo = H // 2, W // 2
    g = torch.Generator().manual_seed(17)
    dcol = torch.randn(B, 9 * C, Ho, Wo, generator=g).cuda().bfloat16()
    dx = torch.empty(B, C, H, W, device="cuda", dtype=torch.bfloat16)
    ops.check(ops.lib.vrcoc_col2im(dcol.data_ptr(), dx.data_ptr(), 1, B, C, H, W, 3, 3, 2, 1, 1, ops._stream()), "col2im")
    cols = dcol.float().view(B, 9, C, Ho * Wo).permute(0, 2, 1, 3).reshape(B, C * 9, Ho * Wo)
    ref = F.fold(cols, output_size=(H, W), kernel_size=3, padding=1, stride=2)
    assert (dx.float() - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 5, 32, 32), (1, 3, 16, 24)])
def test_col2im_4x4_stride4_bf16_is_fold(V, shape):
    """the patch-embedding col2im path (4x4, stride 4, pad 0, bf16: a permutation) = F.fold of the same columns, exactly"""
    import torch.nn.functional as F
    from vrcoc import ops
    B, C, H, W = shape
    Ho, Wo = H // 4, W // 4
    g = torch.Generator().manual_seed(19)
    dcol = torch.randn(B, 16 * C, Ho, Wo, generator=g).cuda().bfloat16()
    dx = torch.empty(B, C, H, W, device="cuda", dtype=torch.bfloat16)
    ops.check(ops.lib.vrcoc_col2im(dcol.data_ptr(), dx.data_ptr(), 1, B, C, H, W, 4, 4, 4, 0, 1, ops._stream()), "col2im")
    cols = dcol.float().view(B, 16, C, Ho * Wo).permute(0, 2, 1, 3).reshape(B, C * 16, Ho * Wo)
    ref = F.fold(cols, output_size=(H, W), kernel_size=4, padding=0, stride=4)
    assert torch.equal(dx.float(), ref)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_dwconv_training_path_vs_conv2d(V, dtype, tol):
    """DWConv with autograd: the native depthwise forward / input gradient / weight gradient against F.conv2d(groups=C) autograd"""
    import torch.nn.functional as F
    torch.manual_seed(3)
    m = V.DWConv(16, 24, kernel_size=3, stride=1, padding=1, bias=True).cuda().to(dtype)
    x = torch.randn(3, 16, 16, 24, device="cuda").to(dtype).requires_grad_(True)
    g = torch.randn(3, 24, 16, 24, device="cuda").to(dtype)
    y = m(x)
    y.backward(g)
    xr = x.detach().double().requires_grad_(True)
    wd, bd = m.dconv.weight.detach().double().requires_grad_(True), m.dconv.bias.detach().double().requires_grad_(True)
    yr = F.conv2d(F.conv2d(xr, wd, bd, padding=1, groups=16), m.pconv.weight.detach().double(), m.pconv.bias.detach().double())
    yr.backward(g.double())
    assert rel_err(y.float(), yr) < tol
    assert rel_err(x.grad.float(), xr.grad) < tol
    assert rel_err(m.dconv.weight.grad.float(), wd.grad) < tol
    assert rel_err(m.dconv.bias.grad.float(), bd.grad) < tol


def test_upsample_rows_kernel_shapes(V):
    """the two-pass (strip) kernel at the live shapes (16->32 ... 128->512, also 256->1024) against F.interpolate in fp32, and in
    bf16 against the fp32 result rounded once (the kernel interpolates in fp32 and rounds the output only)"""
    import torch.nn.functional as F
    from vrcoc.neck import BilinearUpsample
    g = torch.Generator().manual_seed(7)
    for shape, scale in (((2, 5, 16, 16), 2), ((1, 3, 32, 32), 2), ((1, 2, 64, 64), 2), ((2, 9, 128, 128), 4), ((1, 2, 256, 256), 4), ((1, 3, 24, 40), 2)):
        x = torch.randn(*shape, generator=g).cuda()
        up = BilinearUpsample(scale_factor=scale, mode="bilinear", align_corners=True)
        ref = F.interpolate(x.double(), scale_factor=scale, mode="bilinear", align_corners=True)
        got = up(x)
        assert got.shape == ref.shape and rel_err(got, ref) < 2e-5      # fp32 source coordinates (ox * (W-1)/(Wo-1)) vs the fp64 reference
        xb = x.bfloat16()
        gb = up(xb)
        refb = F.interpolate(xb.double(), scale_factor=scale, mode="bilinear", align_corners=True)
        assert (gb.double() - refb).abs().max().item() <= 2.0 ** -8 * refb.abs().max().item() + 1e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_upsample_argmax_is_upsample_then_argmax(V, dtype):
    """serving tail: the fused class map == BilinearUpsample(x4) followed by argmax(1), exactly (ties to the lowest class);
    bf16 logits make ties common, so the tie rule is exercised"""
    from vrcoc.neck import BilinearUpsample, upsample_argmax, upsample_argmax_ok
    g = torch.Generator().manual_seed(11)
    for shape, scale in (((8, 9, 128, 128), 4), ((2, 9, 32, 48), 4), ((1, 21, 64, 64), 2), ((1, 9, 256, 256), 4)):
        x = (torch.randn(*shape, generator=g) * 0.5).cuda().to(dtype)
        x[:, 3] = x[:, 1]                                               # exact ties between two classes wherever they win
        assert upsample_argmax_ok(x, scale)
        got = upsample_argmax(x, scale)
        up = BilinearUpsample(scale_factor=scale, mode="bilinear", align_corners=True)(x)
        ref = up.argmax(dim=1)
        # torch.argmax does not promise an index on ties: compare through the values, and the tie rule explicitly
        picked = torch.gather(up, 1, got.long().unsqueeze(1)).squeeze(1)
        assert got.dtype == torch.uint8 and got.shape == ref.shape
        assert torch.equal(picked, up.max(dim=1).values)
        first = (up == up.max(dim=1, keepdim=True).values).float().argmax(dim=1)
        assert torch.equal(got.long(), first)
        assert (got != 3).all()                                         # class 3 always ties with class 1 and must lose


def test_neck_class_map_switch(V):
    """CoCFpnDual.seg_class_map: the gradient-free forward returns the class map, equal to argmax of the logits it returns otherwise"""
    m = _randomised_model(V, "nano").cuda().bfloat16()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 512, 512, generator=g).cuda().bfloat16()
    r = torch.randn(2, 4, 512, 512, generator=g).cuda().bfloat16()
    with torch.no_grad():
        det0, seg = m(x, r)
        m.backbone.seg_class_map = True
        try:
            det1, cls = m(x, r)
        finally:
            m.backbone.seg_class_map = False
    assert cls.dtype == torch.uint8 and cls.shape == (2, 512, 512)
    # two forwards are not bit-identical (the GroupNorm statistics are accumulated with atomics): the class picked by the second
    # must hold the maximum of the first's logits up to that noise, and nearly every pixel must agree outright
    mx = seg.float().max(dim=1).values
    picked = torch.gather(seg.float(), 1, cls.long().unsqueeze(1)).squeeze(1)
    assert (picked >= mx - 2.0 ** -6 * mx.abs().clamp_min(1e-3)).all()
    assert (cls.long() == seg.argmax(dim=1)).float().mean().item() > 0.995
    for a, b in zip(det0, det1):
        assert rel_err(a.float(), b.float()) < 1e-2


def _randomised_model(V, phi):
    torch.manual_seed(0)
    m = V.EfficientVRNet(4, 9, phi).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.numel() == 0:
                continue
            leaf = n.split(".")[-1]
            if leaf.startswith("layer_scale"):
                p.copy_(torch.rand(p.shape, generator=g) * 0.5 + 0.25)
            elif leaf == "sim_alpha":
                p.copy_(torch.rand(1, generator=g) + 0.5)
            elif leaf == "sim_beta":
                p.copy_(torch.rand(1, generator=g) - 0.5)
        for n, b in m.named_buffers():
            if n.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
            elif n.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
    return m


def test_whole_model_vs_oracle_fp32(V):
    """EfficientVRNet(phi='nano'), B=2, 512x512: the product (no-grad inference path, CUDA) against the CPU oracle on the
    same state_dict; O(1) layer scales so that the CoC blocks actually contribute."""
    from oracle import coc_oracle as O
    m = _randomised_model(V, "nano")
    g = torch.Generator().manual_seed(2)
    x, r = torch.randn(2, 3, 512, 512, generator=g), torch.rand(2, 4, 512, 512, generator=g)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        det_ref, seg_ref = O.efficient_vrnet_forward(x, r, sd, "nano")
        m = m.cuda()
        det, seg = m(x.cuda(), r.cuda())
        det2, seg2 = m(x.cuda(), r.cuda())       # second call exercises the memoised parameter views
    assert rel_err(seg, seg_ref) < 2e-3
    for a, b in zip(det, det_ref):
        assert rel_err(a, b) < 2e-3
    # run-to-run differences come only from the order of the fp64 atomics behind the GroupNorm statistics
    assert rel_err(seg2, seg) < 1e-3 and all(rel_err(a, b) < 1e-3 for a, b in zip(det2, det))


def test_whole_model_bf16_runs_and_tracks_fp32(V):
    m = _randomised_model(V, "nano").cuda()
    g = torch.Generator().manual_seed(2)
    x, r = torch.randn(1, 3, 512, 512, generator=g).cuda(), torch.rand(1, 4, 512, 512, generator=g).cuda()
    with torch.no_grad():
        det32, seg32 = m(x, r)
        mb = m.bfloat16()
        det16, seg16 = mb(x.bfloat16(), r.bfloat16())
    assert seg16.dtype == torch.bfloat16 and torch.isfinite(seg16.float()).all()
    assert rel_err(seg16.float(), seg32) < 0.15      # end-to-end bf16 drift over 27 blocks (information, loose)
