"""GPU: the two variants of the convolution-as-GEMM engine against each other and against a plain fp32 reference of
the same op (torch conv2d on the same, already bf16-rounded, operands; TF32 disabled).  The tcgen05 variant must agree
with the CUDA-core variant to accumulation-order noise: both multiply the same bf16 values with fp32 accumulation."""
import pytest
import torch
import torch.nn.functional as F

from golden_util import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [  # B, C, O, H, W, k, stride, pad
    (2, 64, 128, 16, 16, 1, 1, 0),        # one k-slab, one 128-point tile pair
    (1, 64, 256, 128, 128, 1, 1, 0),      # S1 fc1|fc_v: N tile 256
    (2, 320, 1280, 32, 32, 1, 1, 0),      # S3 mlp.fc1: 5 k-slabs, 5 N tiles
    (2, 1280, 320, 32, 32, 1, 1, 0),      # S3 mlp.fc2: 20 k-slabs, 2 N tiles of 160
    (1, 96, 512, 16, 16, 1, 1, 0),        # neck fc2: K = 96 (1.5 slabs)
    (2, 7, 4, 64, 64, 1, 1, 0),           # ingest inverse projection: K = 7, O = 4
    (1, 24, 40, 8, 8, 1, 1, 0),           # 64 points: half-empty M tile
    (2, 4, 3, 64, 64, 3, 1, 1),           # ingest radar projection 3x3: K = 36 (not a multiple of 8)
    (1, 64, 128, 32, 32, 3, 2, 1),        # point reducer 3x3 / stride 2
    (2, 5, 64, 64, 64, 4, 4, 0),          # patch embed 4x4 / stride 4
    (1, 128, 128, 16, 16, 3, 1, 1),       # fusion radar projection 3x3
]


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from vrcoc import ops
    return ops


def _run(ops, x, w, engine, stride, pad, out_dtype=torch.float32, **kw):
    from vrcoc._lib import check, lib
    B, C, H, W = x.shape
    O, _, kh, kwid = w.shape
    Ho, Wo = ops.out_hw(H, W, kh, stride, pad)
    out = torch.empty(B, O, Ho, Wo, device=x.device, dtype=out_dtype)
    d = ops.conv_desc(x, w.reshape(O, -1).contiguous(), out, kh=kh, kw=kwid, stride=stride, pad=pad, engine=engine, **kw)
    ops.conv_fwd(d)
    return out


@pytest.mark.parametrize("shape", SHAPES, ids=[f"C{s[1]}_O{s[2]}_{s[3]}x{s[4]}_k{s[5]}s{s[6]}" for s in SHAPES])
def test_tcgen05_matches_reference_and_simt(env, shape):
    ops = env
    B, C, O, H, W, k, stride, pad = shape
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(O, C, k, k, generator=g) / (C * k * k) ** 0.5).to(torch.bfloat16).cuda()
    ref = F.conv2d(x.float(), w.float(), None, stride=stride, padding=pad)
    simt = _run(ops, x, w, 1, stride, pad)
    tc = _run(ops, x, w, 2, stride, pad)
    assert rel_err(simt, ref) < 1e-5
    assert rel_err(tc, ref) < 1e-5, "tcgen05 engine disagrees with the fp32 reference"
    assert rel_err(tc, simt) < 1e-5


def test_tcgen05_full_epilogue_and_prologue(env):
    """GroupNorm prologue (transform-on-load rounds the normalised activation to bf16 once), bias + GELU, layer-scale,
    residual, final affine, split outputs and the per-sample side statistics"""
    ops = env
    g = torch.Generator().manual_seed(11)
    B, C, O, H, W = 2, 64, 192, 32, 32
    x = (torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3).to(torch.bfloat16).cuda()
    w = (torch.randn(O, C, 1, 1, generator=g) / C ** 0.5).to(torch.bfloat16).cuda()
    gamma = (torch.rand(C, generator=g) + 0.5).cuda()
    beta = (torch.randn(C, generator=g) * 0.1).cuda()
    bias = (torch.randn(O, generator=g) * 0.1).cuda()
    ls = (torch.rand(O, generator=g) + 0.5).cuda()
    fs, fh = (torch.rand(O, generator=g) + 0.5).cuda(), torch.randn(O, generator=g).cuda()
    res = torch.randn(B, O, H, W, generator=g).to(torch.bfloat16).cuda()
    _, sums = ops.channel_sums(x, want_chan=False, want_sample=True)
    xn = F.group_norm(x.float(), 1, gamma, beta, 1e-5)
    ref = F.gelu(F.conv2d(xn, w.float(), bias)) * ls.view(1, -1, 1, 1) + res.float()
    ref = ref * fs.view(1, -1, 1, 1) + fh.view(1, -1, 1, 1)
    outs = {}
    for engine in (1, 2):
        from vrcoc._lib import ACT_GELU
        o1 = torch.empty(B, 64, H, W, device="cuda", dtype=torch.float32)
        o2 = torch.empty(B, O - 64, H, W, device="cuda", dtype=torch.bfloat16)
        ss = ops.new_sample_sums(B, "cuda")
        d = ops.conv_desc(x, w.reshape(O, C).contiguous(), o1, gn=(sums, gamma, beta, 1e-5), e_shift=bias, act=ACT_GELU,
                          post_scale=ls, res=res, f_scale=fs, f_shift=fh, out2=o2, out_sample_sums=ss, engine=engine)
        ops.conv_fwd(d)
        got = torch.cat([o1, o2.float()], 1)
        outs[engine] = got
        tol = 1e-5 if engine == 1 else 6e-3          # engine 2 rounds GN(x) to bf16 before the tensor core
        assert rel_err(got[:, :64], ref[:, :64]) < tol, engine
        assert rel_err(got, ref) < 5e-3, engine      # bf16 storage of the second output
        full = ref.double()
        assert abs(ss[..., 0].sum().item() - full.sum().item()) < 2e-2 * full.abs().sum().item() ** 0.5 + 1e-3 * abs(full.sum().item())
        assert abs(ss[..., 1].sum().item() / (full ** 2).sum().item() - 1) < 1e-2


@pytest.mark.parametrize("shape", [(2, 64, 96, 16, 16, 3, 1, 1, 1), (1, 128, 64, 32, 32, 3, 2, 1, 1), (2, 64, 64, 16, 16, 3, 1, 6, 6),
                                   (1, 320, 128, 16, 16, 3, 1, 12, 12), (2, 4, 3, 64, 64, 3, 1, 1, 1), (2, 7, 4, 64, 64, 1, 1, 0, 1),
                                   # the 3x3 im2col fast path (stride 1 / 2) + GEMM at the live map sizes, 5x5 and odd maps on the generic one
                                   (1, 64, 64, 128, 128, 3, 1, 1, 1), (2, 128, 128, 64, 64, 3, 1, 1, 1), (2, 320, 320, 32, 32, 3, 1, 1, 1),
                                   (2, 512, 512, 16, 16, 3, 1, 1, 1), (1, 64, 192, 32, 32, 3, 1, 2, 2), (1, 64, 96, 64, 64, 5, 1, 2, 1),
                                   (1, 128, 64, 8, 16, 3, 1, 1, 1), (2, 64, 128, 64, 64, 3, 2, 1, 1), (1, 320, 512, 32, 32, 3, 2, 1, 1),
                                   # dilation >= the map: the centre-tap 1x1 shortcut (ASPP rate 18 at 16x16) and the case just below it
                                   (2, 512, 512, 16, 16, 3, 1, 18, 18), (2, 128, 96, 16, 16, 3, 1, 16, 16), (1, 128, 96, 16, 16, 3, 1, 15, 15)])
def test_tap_major_dilation_and_small_kernel(env, shape):
    """fusion._conv_launch picks: tap-major K order on the tensor-core path for k x k convs, dilation (ASPP), and the
    few-channel streaming kernel for the ingest convs; all against torch conv2d on the same bf16 operands"""
    from vrcoc import fusion
    B, C, O, H, W, k, stride, pad, dil = shape
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(O, C, k, k, generator=g) / (C * k * k) ** 0.5).to(torch.bfloat16).cuda()
    bias = torch.randn(O, generator=g).cuda()
    ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=pad, dilation=dil)
    got = fusion._conv_launch(x, w, bias, stride, pad, None, None, 0, None, None, torch.float32, dil=dil)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-5


@pytest.mark.parametrize("case", [(2, 64, 16, 128, 1, torch.bfloat16), (1, 128, 32, 32, 1, torch.bfloat16), (2, 64, 16, 16, 6, torch.bfloat16),
                                  (1, 64, 8, 16, 12, torch.bfloat16), (1, 3, 5, 12, 2, torch.float32)],
                         ids=lambda c: "x".join(map(str, c[:5])))
def test_im2col_rows(env, case):
    """vrcoc_im2col_rows: cols[b][kx*C + c][y][x] = x[b][c][y][x + (kx-1)*dil], zero outside (fast bf16 kernel and the generic one)"""
    from vrcoc._lib import check, lib
    B, C, H, W, dil, dtype = case
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, H, W, generator=g).to(dtype).cuda()
    cols = torch.full((B, 3 * C, H, W), float("nan"), device="cuda", dtype=dtype)
    check(lib.vrcoc_im2col_rows(x.data_ptr(), cols.data_ptr(), 0 if dtype == torch.float32 else 1, B, C, H, W, 3, dil,
                                torch.cuda.current_stream().cuda_stream), "im2col_rows")
    xp = F.pad(x, (dil, dil, 0, 0))
    ref = torch.cat([xp[..., kx * dil: kx * dil + W] for kx in range(3)], dim=1)
    assert torch.equal(cols, ref)


@pytest.mark.parametrize("shape", [(1, 64, 64, 128, 128, 1), (2, 128, 128, 64, 64, 1), (2, 320, 320, 32, 32, 1), (2, 512, 512, 16, 16, 1),
                                   (2, 512, 512, 16, 16, 6), (1, 512, 512, 16, 16, 12), (1, 512, 512, 16, 16, 18), (2, 64, 96, 16, 16, 1),
                                   (1, 128, 64, 8, 16, 1), (3, 64, 320, 32, 64, 1), (1, 64, 128, 24, 64, 1)],
                         ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32], ids=["bf16out", "f32out"])
def test_row_tap_conv3x3_matches_full_im2col_and_torch(env, shape, out_dtype):
    """3x3 stride-1 convolutions in row-tap mode (vrcoc.h k_order 2: horizontal-tap copies + TMA boxes shifted by rows, both TMA
    kernels, every live map width incl. dilated ASPP branches) against the full-im2col path and torch conv2d
    (reference vr_coc.py:99-102,313; coc_fpn_dual.py:55-67)"""
    from vrcoc import fusion
    B, C, O, H, W, dil = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(O, C, 3, 3, generator=g) / (C * 9) ** 0.5).to(torch.bfloat16).cuda()
    bias = torch.randn(O, generator=g).cuda()
    scale = (torch.rand(O, generator=g) + 0.5).cuda()
    ref = F.relu(F.conv2d(x.float(), w.float(), None, padding=dil, dilation=dil) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    from vrcoc._lib import ACT_RELU, check, lib
    ops = env
    # the row-tap launch itself (fusion._conv_launch dispatches it only where a row is a whole number of 128-byte lines)
    cols = torch.empty(B, 3 * C, H, W, device="cuda", dtype=torch.bfloat16)
    check(lib.vrcoc_im2col_rows(x.data_ptr(), cols.data_ptr(), 1, B, C, H, W, 3, dil, torch.cuda.current_stream().cuda_stream), "im2col_rows")
    got = torch.full((B, O, H, W), float("nan"), device="cuda", dtype=out_dtype)
    ops.conv_fwd(ops.conv_desc(cols, ops.tap_major(w), got, kh=3, kw=1, stride=1, pad=dil, dil=dil, k_order=2, e_scale=scale, e_shift=bias,
                               act=ACT_RELU))
    fusion.ROW_TAPS = False
    try:
        full = fusion._conv_launch(x, w, bias, 1, dil, None, None, ACT_RELU, scale, None, out_dtype, dil=dil)
    finally:
        fusion.ROW_TAPS = True
    if dil == 1 and W % 64 == 0:
        via_dispatch = fusion._conv_launch(x, w, bias, 1, dil, None, None, ACT_RELU, scale, None, out_dtype, dil=dil)
        assert torch.equal(via_dispatch, got)
    torch.cuda.synchronize()
    assert got.shape == ref.shape and torch.isfinite(got.float()).all()
    tol = 1e-5 if out_dtype == torch.float32 else 4e-3
    assert rel_err(got.float(), ref) < tol
    assert rel_err(got.float(), full.float()) < (2e-6 if out_dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("case", [(2, 3, 2, 512, 512, True), (2, 4, 2, 512, 512, True), (1, 4, 2, 128, 256, False), (3, 6, 0, 64, 128, False),
                                  (1, 3, 2, 1024, 1024, True)], ids=lambda c: "x".join(map(str, c)))
def test_patch_embed_kernel(env, case):
    """vrcoc_patch_embed (csrc/patch_embed.cu): the 4x4 / stride-4 patch embedding of cat([x, pos]) (reference vr_coc.py:83-102 from
    :575-587) against torch conv2d on the same bf16 operands and against the general engine, incl. the GroupNorm statistics that
    leave with the output and the batch-broadcast position source"""
    from vrcoc import fusion
    ops = env
    B, C0, C1, H, W, bcast = case
    g = torch.Generator().manual_seed(13)
    x = torch.randn(B, C0, H, W, generator=g).to(torch.bfloat16).cuda()
    extra = None
    if C1:
        extra = (torch.rand(C1, H, W, generator=g) if bcast else torch.rand(B, C1, H, W, generator=g)).to(torch.bfloat16).cuda()
    w = (torch.randn(64, C0 + C1, 4, 4, generator=g) / ((C0 + C1) * 16) ** 0.5).to(torch.bfloat16).cuda()
    bias = torch.randn(64, generator=g).cuda()
    full = x.float() if extra is None else torch.cat([x.float(), (extra.float().expand(B, -1, -1, -1) if bcast else extra.float())], 1)
    ref = F.conv2d(full, w.float(), bias, stride=4)
    assert fusion.PATCH_EMBED_KERNEL and ops.lib.vrcoc_patch_embed_supported(1, C0, C1, H, W, 64, 4)
    s_new = ops.new_sample_sums(B, "cuda")
    got = fusion._conv_launch(x, w, bias, 4, 0, extra, None, 0, None, None, None, out_sample_sums=s_new)
    fusion.PATCH_EMBED_KERNEL = False
    try:
        s_old = ops.new_sample_sums(B, "cuda")
        old = fusion._conv_launch(x, w, bias, 4, 0, extra, None, 0, None, None, None, out_sample_sums=s_old)
    finally:
        fusion.PATCH_EMBED_KERNEL = True
    torch.cuda.synchronize()
    assert got.shape == ref.shape and got.dtype == torch.bfloat16
    assert rel_err(got.float(), ref) < 4e-3                       # bf16 output rounding
    assert rel_err(got.float(), old.float()) < 2e-3
    ss = s_new.sum(1).cpu()                                       # [B, 2]: sum, sum of squares of the fp32 values
    assert torch.allclose(ss[:, 0], ref.double().sum((1, 2, 3)).cpu(), rtol=1e-3, atol=1e-2 * ref[0].numel() ** 0.5)
    assert torch.allclose(ss[:, 1], (ref.double() ** 2).sum((1, 2, 3)).cpu(), rtol=1e-3)
    assert rel_err(s_new.sum(1), s_old.sum(1)) < 1e-4


def test_auto_engine_picks_tcgen05_for_bf16_weights(env):
    """fp32 weights -> exact CUDA-core path; bf16 weights -> tensor cores.  Seen through the numerics: with fp32
    activations and bf16 weights the tcgen05 path rounds the activation operand to bf16, the CUDA-core path does not."""
    ops = env
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 64, 16, 16, generator=g).cuda()
    w = (torch.randn(64, 64, 1, 1, generator=g) / 8).to(torch.bfloat16).cuda()
    ref = F.conv2d(x, w.float())
    auto = _run(ops, x, w, 0, 1, 0)
    simt = _run(ops, x, w, 1, 1, 0)
    assert rel_err(simt, ref) < 1e-6
    e = rel_err(auto, ref)
    assert 1e-4 < e < 6e-3, e


CM_CASES = [  # B, C, O, H, W, split, act, mode
    (2, 64, 256, 32, 32, 128, "none", "gn"),       # S1 fc1|fc_v: fp32 | bf16 split outputs, GroupNorm prologue
    (2, 128, 1024, 32, 32, 0, "gelu", "gn"),       # S2 mlp.fc1: 8 output tiles per CTA range
    (2, 320, 512, 32, 32, 256, "none", "gn"),      # 5 resident k-slabs
    (2, 1024, 128, 32, 32, 0, "none", "res"),      # S2 mlp.fc2: both operands by TMA, residual + layer scale + statistics
    (2, 2048, 512, 16, 16, 0, "none", "res"),      # streamed X, 32 k-slabs
    (1, 4608, 512, 16, 16, 0, "relu", "plain"),    # ASPP branch GEMM
    (2, 64, 72, 16, 24, 0, "none", "gn"),          # O not a multiple of 32: clipped TMA store rows
    (3, 136, 200, 8, 24, 0, "gelu", "gn"),         # K = 2.1 slabs, 192 points: clipped point columns
    (2, 64, 48, 5, 8, 0, "none", "gn"),            # 40 points: one partial tile
]


@pytest.mark.parametrize("case", CM_CASES, ids=[f"C{c[1]}_O{c[2]}_{c[3]}x{c[4]}_{c[7]}" for c in CM_CASES])
def test_channel_major_kernel_matches_point_major(env, case):
    """conv_tc_cm_kernel (weights as the M operand, TMA-staged epilogue) against the point-major tcgen05 kernels (switched
    with vrcoc_debug_set_cm) and the CUDA-core engine: same bf16 operands, fp32 accumulation."""
    ops = env
    from vrcoc._lib import ACT_GELU, ACT_NONE, ACT_RELU, lib
    B, C, O, H, W, split, act_name, mode = case
    act = {"none": ACT_NONE, "gelu": ACT_GELU, "relu": ACT_RELU}[act_name]
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(B, C, H, W, generator=g) * 1.3 + 0.2).bfloat16().cuda()
    w = (torch.randn(O, C, generator=g) / C ** 0.5).bfloat16().cuda()
    kw = dict(e_shift=(torch.randn(O, generator=g) * 0.1).cuda(), e_scale=(torch.rand(O, generator=g) + 0.5).cuda(), act=act)
    if mode == "gn":
        _, sums = ops.channel_sums(x, want_chan=False, want_sample=True)
        kw["gn"] = (sums, (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.1).cuda(), 1e-5)
    if mode == "res":
        kw.update(res=torch.randn(B, O, H, W, generator=g).bfloat16().cuda(), post_scale=(torch.rand(O, generator=g) + 0.5).cuda())
    outs, stats, mm = {}, {}, {}
    try:
        for name, engine, cm in (("simt", 1, 1), ("cm", 2, 1), ("pm", 2, 0)):
            lib.vrcoc_debug_set_cm(cm)
            if split:
                o1 = torch.full((B, split, H, W), float("nan"), device="cuda")
                o2 = torch.full((B, O - split, H, W), float("nan"), device="cuda", dtype=torch.bfloat16)
            else:
                o1, o2 = torch.full((B, O, H, W), float("nan"), device="cuda", dtype=torch.bfloat16), None
            k2 = dict(kw)
            if mode == "res":
                k2["out_sample_sums"] = ops.new_sample_sums(B, "cuda")
            ops.conv_fwd(ops.conv_desc(x, w, o1, out2=o2, engine=engine, **k2))
            torch.cuda.synchronize()
            outs[name] = torch.cat([o1.float(), o2.float()], 1) if split else o1.float()
            if mode == "res":
                stats[name] = k2["out_sample_sums"].sum(1)
    finally:
        lib.vrcoc_debug_set_cm(1)
    assert torch.isfinite(outs["cm"]).all(), "channel-major kernel left part of the output unwritten"
    assert rel_err(outs["cm"], outs["pm"]) < 1e-5          # same operands, same accumulation: rounding-order noise only
    assert rel_err(outs["cm"], outs["simt"]) < (6e-3 if mode == "gn" else 3e-3)   # GN(x) rounded to bf16 before the MMA
    if stats:
        assert rel_err(stats["cm"], stats["simt"]) < 1e-4


@pytest.mark.parametrize("case", [(2, 256, 64, 64, 3, 1, 1), (2, 16, 22, 24, 3, 1, 1), (1, 8, 9, 10, 3, 1, 1), (2, 12, 16, 16, 3, 2, 1),
                                  (1, 6, 12, 16, 5, 1, 2)], ids=lambda c: "x".join(map(str, c)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_depthwise_conv(env, case, dtype):
    """vrcoc_dwconv (DWConv.dconv, normal_conv.py:26-27): 3x3/s1 fast path (8 columns x 4 rows per thread) and the generic one"""
    from vrcoc._lib import check, lib
    B, C, H, W, k, stride, pad = case
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, H, W, generator=g).to(dtype).cuda()
    w = (torch.randn(C, 1, k, k, generator=g) / k).to(dtype).cuda()
    bias = torch.randn(C, generator=g).cuda()
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.full((B, C, Ho, Wo), float("nan"), device="cuda", dtype=dtype)
    check(lib.vrcoc_dwconv(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), 0 if dtype == torch.float32 else 1, B, C, H, W, k,
                           stride, pad, torch.cuda.current_stream().cuda_stream), "dwconv")
    ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=pad, groups=C)
    assert rel_err(y.float(), ref) < (1e-5 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("case", [(2, 64, 512, 32, 32), (1, 128, 1024, 16, 24), (2, 64, 256, 5, 8), (3, 128, 128, 8, 8),
                                  (2, 320, 1280, 32, 32), (1, 192, 384, 8, 24), (2, 384, 256, 16, 16)],
                         ids=lambda c: "x".join(map(str, c)))
def test_fused_mlp_matches_two_launches(env, case):
    """vrcoc_mlp_fused_fwd (hidden layer kept on chip) against the two-GEMM path and an fp32 torch restatement of
    x + ls * fc2(gelu(fc1(GroupNorm(1, C)(x))))  (reference vr_coc.py:208-228, :270-275)"""
    ops = env
    from vrcoc._lib import ACT_GELU
    B, C, hid, H, W = case
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(B, C, H, W, generator=g) * 1.2 + 0.1).bfloat16().cuda()
    w1 = (torch.randn(hid, C, generator=g) / C ** 0.5).bfloat16().cuda()
    w2 = (torch.randn(C, hid, generator=g) / hid ** 0.5).bfloat16().cuda()
    b1, b2 = (torch.randn(hid, generator=g) * 0.1).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    ls = (torch.rand(C, generator=g) + 0.5).cuda()
    sums = ops.sample_sums_of(x)
    assert ops.mlp_fused_ok(x, hid) or C > ops.FUSED_MLP_MAX_C      # wide shapes: kernel tested here, not dispatched by policy
    s_f = ops.new_sample_sums(B, "cuda")
    fused = ops.mlp_fused_fwd(x, sums, gamma, beta, 1e-5, w1, b1, w2, b2, ls, s_f)
    h = torch.empty(B, hid, H, W, device="cuda", dtype=torch.bfloat16)
    ops.conv_fwd(ops.conv_desc(x, w1, h, gn=(sums, gamma, beta, 1e-5), e_shift=b1, act=ACT_GELU))
    two = torch.empty_like(x)
    s_t = ops.new_sample_sums(B, "cuda")
    ops.conv_fwd(ops.conv_desc(h, w2, two, e_shift=b2, post_scale=ls, res=x, out_sample_sums=s_t))
    torch.cuda.synchronize()
    xf = x.float()
    ref = xf + ls.view(1, -1, 1, 1) * F.conv2d(F.gelu(F.conv2d(F.group_norm(xf, 1, gamma, beta, 1e-5), w1.float().view(hid, C, 1, 1), b1)),
                                              w2.float().view(C, hid, 1, 1), b2)
    assert torch.isfinite(fused.float()).all()
    assert rel_err(fused.float(), two.float()) < 2e-3            # same arithmetic; both round the hidden layer to bf16
    assert rel_err(fused.float(), ref) < 8e-3                    # bf16 GN(x), bf16 hidden, bf16 output
    assert rel_err(s_f.sum(1), s_t.sum(1)) < 2e-3


@pytest.mark.parametrize("case", [(8, 64, 512, 32, 32, True), (2, 320, 1280, 32, 32, True), (2, 128, 64, 64, 64, False),
                                  (3, 72, 96, 10, 20, True), (1, 640, 200, 8, 8, False), (4, 512, 64, 128, 128, False)],
                         ids=lambda c: f"B{c[0]}_C{c[1]}_O{c[2]}_{c[3]}x{c[4]}_{'gn' if c[5] else 'plain'}")
def test_wgrad_tensor_core_matches_reference_and_cuda_core(env, case):
    """dW = dY . prologue(x)^T, db = sum dY: the tcgen05 path (raw bf16 operands by TMA, GroupNorm coefficients applied to
    the per-sample partials) against fp64 torch on the same bf16 values and against the CUDA-core path it replaces"""
    import os
    ops = env
    B, C, O, H, W, use_gn = case
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(B, C, H, W, generator=g) * 1.7 + 0.6).to(torch.bfloat16).cuda()
    dy = torch.randn(B, O, H, W, generator=g).to(torch.bfloat16).cuda()
    w = torch.zeros(O, C, dtype=torch.bfloat16, device="cuda")
    gamma = (torch.rand(C, generator=g) + 0.5).cuda()
    beta = torch.randn(C, generator=g).cuda()
    gn = None
    xh = x.double()
    if use_gn:
        _, sums = ops.channel_sums(x, want_chan=False, want_sample=True)
        gn = (sums, gamma, beta, 1e-5)
        mu = xh.mean(dim=(1, 2, 3), keepdim=True)
        var = xh.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
        xh = (xh - mu) / torch.sqrt(var + 1e-5) * gamma.double().view(1, C, 1, 1) + beta.double().view(1, C, 1, 1)
    ref_w = torch.einsum("bop,bcp->oc", dy.double().flatten(2), xh.flatten(2))
    ref_b = dy.double().sum(dim=(0, 2, 3))
    assert os.environ.get("VRCOC_WGRAD_TC") is None
    dW, db = ops.conv1x1_wgrad(ops.conv_desc(x, w, dy, gn=gn), dy)
    os.environ["VRCOC_WGRAD_TC"] = "0"
    try:
        dW0, db0 = ops.conv1x1_wgrad(ops.conv_desc(x, w, dy, gn=gn), dy)
    finally:
        del os.environ["VRCOC_WGRAD_TC"]
    assert rel_err(dW0, ref_w) < 1e-5 and rel_err(db0, ref_b) < 1e-5
    assert rel_err(dW, ref_w) < 2e-5, "tensor-core weight gradient disagrees with the fp64 reference"
    assert rel_err(db, ref_b) < 1e-5
    dW2, _ = ops.conv1x1_wgrad(ops.conv_desc(x, w, dy, gn=gn), dy)
    assert torch.equal(dW, dW2), "the tensor-core weight gradient must be deterministic"


@pytest.mark.parametrize("case", [(3, 4, 3, True, True), (1, 7, 4, True, False), (1, 3, 3, False, False), (1, 4, 4, False, True)],
                         ids=lambda c: f"k{c[0]}_{c[1]}to{c[2]}_{'relu' if c[3] else 'none'}_{'stats' if c[4] else 'plain'}")
def test_ingest_convs_compile_time_kernel(env, case):
    """the four ingest shapes (conv_ingest_kernel: bf16 in / out, compile-time channel counts, BN-style affine, ReLU, residual,
    whole-batch min/max side output) against torch conv2d in fp32 on the same bf16 operands; frames wide enough for several
    8-pixel groups per row and rows that touch both borders"""
    ops = env
    k, C, O, relu, stats = case
    B, H, W = 2, 24, 64
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(O, C, k, k, generator=g) / (C * k * k) ** 0.5).to(torch.bfloat16).cuda()
    es, eh = (torch.rand(O, generator=g) + 0.5).cuda(), torch.randn(O, generator=g).cuda()
    res = torch.randn(B, O, H, W, generator=g).to(torch.bfloat16).cuda() if not stats else None
    ref = F.conv2d(x.float(), w.float(), None, padding=(k - 1) // 2) * es.view(1, O, 1, 1) + eh.view(1, O, 1, 1)
    if relu:
        ref = ref.relu()
    if res is not None:
        ref = ref + res.float()
    out = torch.empty(B, O, H, W, device="cuda", dtype=torch.bfloat16)
    mm = torch.tensor([0, 0], dtype=torch.int32, device="cuda") if stats else None       # {max bits, ~min bits} (atomicMax on both)
    from vrcoc._lib import ACT_RELU, ACT_NONE
    d = ops.conv_desc(x, w.reshape(O, -1).contiguous(), out, kh=k, kw=k, stride=1, pad=(k - 1) // 2, e_scale=es, e_shift=eh,
                      act=ACT_RELU if relu else ACT_NONE, res=res, out_minmax=mm)
    ops.conv_fwd(d)
    assert rel_err(out.float(), ref) < 4e-3                     # one bf16 rounding of the result
    assert (out.float() - ref).abs().max() <= 2 ** -7 * ref.abs().max()
    if stats and relu:
        got_max = torch.tensor([mm[0].item()], dtype=torch.int32).view(torch.float32).item()
        assert abs(got_max - ref.max().item()) <= 2 ** -7 * abs(ref.max().item())
