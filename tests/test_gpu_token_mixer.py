"""The fused token-mixer kernel (csrc/token_mixer_fused.cu: GN-folded fc1|fc_v -> cluster core -> fc2 + layer scale + residual in
one persistent launch) against (a) the three-launch path it replaces (same folded weights, so the two differ only by summation
order) and (b) the CPU oracle in fp64 on the bf16-rounded weights and inputs.  reference: vr_coc.py:155-192, :264-267."""
import pytest
import torch

from golden_util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vrcoc
    return vrcoc


def _params(C, g, ED=128):
    p = dict(
        w1=(torch.randn(ED, C, generator=g) / C ** 0.5).to(torch.bfloat16), b1=torch.randn(ED, generator=g) * 0.1,
        wv=(torch.randn(ED, C, generator=g) / C ** 0.5).to(torch.bfloat16), bv=torch.randn(ED, generator=g) * 0.1,
        w2=(torch.randn(C, ED, generator=g) / ED ** 0.5).to(torch.bfloat16), b2=torch.randn(C, generator=g) * 0.1,
        gamma=torch.rand(C, generator=g) + 0.5, beta=torch.randn(C, generator=g) * 0.1, ls=torch.rand(C, generator=g) + 0.5,
        alpha=torch.tensor([1.3]), sbeta=torch.tensor([-0.2]))
    return p


def _run_both(V, B, C, H, seed=0):
    from vrcoc import ops
    g = torch.Generator().manual_seed(seed)
    p = _params(C, g)
    x = (torch.randn(B, C, H, H, generator=g) * 1.3 + 0.4).to(torch.bfloat16)
    pc = {k: v.cuda() for k, v in p.items()}
    xc = x.cuda()
    fold = H // 16
    assert ops.token_mixer_fused_ok(xc, 4, 32, fold, fold, 2, 2)
    sums = ops.channel_sums(xc, want_chan=False, want_sample=True)[1]
    w_fold, k0, k1 = ops.fold_gn_weights(pc["w1"], pc["b1"], pc["wv"], pc["bv"], pc["gamma"], pc["beta"])
    # fused
    osum_f = ops.new_sample_sums(B, xc.device)
    out_f, idx_f, smax_f = ops.token_mixer_fused_fwd(xc, sums, 1e-5, w_fold, k0, k1, pc["alpha"], pc["sbeta"], pc["w2"], pc["b2"], pc["ls"],
                                                     osum_f, 4, 32, fold, fold, save_aux=True)
    # three launches
    feat = torch.empty(B, 128, H, H, device="cuda", dtype=torch.float32)
    value = torch.empty(B, 128, H, H, device="cuda", dtype=torch.bfloat16)
    ops.conv_fwd(ops.conv_desc(xc, w_fold, feat, gn_fold=(sums, k1, 1e-5), e_shift=k0, out2=value))
    o, idx_u, smax_u = ops.cluster_core_fwd(feat, value, pc["alpha"], pc["sbeta"], 4, fold, fold, 2, 2, out_dtype=torch.bfloat16, save_aux=True)
    osum_u = ops.new_sample_sums(B, xc.device)
    out_u = torch.empty_like(xc)
    ops.conv_fwd(ops.conv_desc(o, pc["w2"], out_u, e_shift=pc["b2"], post_scale=pc["ls"], res=xc, out_sample_sums=osum_u))
    torch.cuda.synchronize()
    return p, x, (out_f, idx_f, smax_f, osum_f), (out_u, idx_u, smax_u, osum_u), (feat, value, o)


@pytest.mark.parametrize("B,C,H", [(1, 64, 16), (2, 64, 32), (2, 128, 32), (8, 64, 128), (8, 128, 64), (3, 64, 48)])
def test_fused_token_mixer_vs_three_launches(V, B, C, H):
    p, x, (out_f, idx_f, smax_f, osum_f), (out_u, idx_u, smax_u, osum_u), _ = _run_both(V, B, C, H)
    assert torch.isfinite(out_f.float()).all()
    flips = (idx_f != idx_u).float().mean().item()
    e_s = rel_err(smax_f, smax_u)
    e_o = rel_err(out_f.float(), out_u.float())
    d_sum = (osum_f.sum(1) - osum_u.sum(1)).abs().max().item() / osum_u.sum(1).abs().max().item()
    print(f"B={B} C={C} H={H}: assignment flips {flips:.2e}  sim_max {e_s:.2e}  out {e_o:.2e}  stats {d_sum:.2e}")
    assert flips < 2e-4                    # both paths see the same fp32 feat up to summation order: only ~1e-6-margin points may differ
    assert e_s < 1e-4 and e_o < 4e-3 and d_sum < 1e-4


@pytest.mark.parametrize("B,C,H", [(2, 64, 32), (2, 128, 32)])
def test_fused_token_mixer_vs_oracle(V, B, C, H):
    from oracle import coc_oracle as O
    p, x, (out_f, idx_f, smax_f, _), _, _ = _run_both(V, B, C, H, seed=3)
    d = {k: v.double() for k, v in p.items()}
    gn = O.group_norm1(x.double(), d["gamma"], d["beta"])
    sd = {"fc1.weight": d["w1"][:, :, None, None], "fc1.bias": d["b1"], "fc_v.weight": d["wv"][:, :, None, None], "fc_v.bias": d["bv"],
          "fc2.weight": d["w2"][:, :, None, None], "fc2.bias": d["b2"], "sim_alpha": d["alpha"], "sim_beta": d["sbeta"]}
    fold = H // 16
    y, idx, gmap, margin = O.cluster(gn, sd, "", 4, fold, fold, 2, 2, aux=True)
    ref = x.double() + d["ls"].view(1, -1, 1, 1) * y
    safe = (margin > 1e-4).reshape(idx_f.shape)
    got_idx = idx_f.cpu().long()
    mism = (got_idx != idx.reshape(idx_f.shape))[safe].float().mean().item()
    e = rel_err(out_f.float(), ref)
    print(f"fused token mixer vs oracle C={C}: out {e:.2e}, assignment mismatches outside a 1e-4 margin {mism:.2e}, safe {safe.float().mean():.4f}")
    assert mism == 0.0
    assert e < 2e-2


@pytest.mark.parametrize("B,heads", [(1, 8), (2, 8), (8, 8), (5, 4), (40, 8)])
def test_fused_projection_core_stage3_vs_two_launches(V, B, heads):
    """the wide-stage kernel (C = 320: GN-folded fc1|fc_v + cluster core per (region, four heads)) against projection + core as two
    launches on the same folded weights; B = 40 makes a CTA loop over several units"""
    from vrcoc import ops
    C, H, ED = 320, 32, heads * 32
    g = torch.Generator().manual_seed(4)
    w1 = (torch.randn(ED, C, generator=g) / C ** 0.5).to(torch.bfloat16).cuda()
    wv = (torch.randn(ED, C, generator=g) / C ** 0.5).to(torch.bfloat16).cuda()
    b1, bv = (torch.randn(ED, generator=g) * 0.1).cuda(), (torch.randn(ED, generator=g) * 0.1).cuda()
    gamma, beta = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    x = (torch.randn(B, C, H, H, generator=g) * 1.3 + 0.4).to(torch.bfloat16).cuda()
    alpha, sbeta = torch.tensor([1.3]).cuda(), torch.tensor([-0.2]).cuda()
    assert ops.token_mixer_core_ok(x, heads, 32, 2, 2, 2, 2)
    sums = ops.channel_sums(x, want_chan=False, want_sample=True)[1]
    w_fold, k0, k1 = ops.fold_gn_weights(w1, b1, wv, bv, gamma, beta)
    o_f, idx_f, smax_f = ops.token_mixer_core_fwd(x, sums, 1e-5, w_fold, k0, k1, alpha, sbeta, heads, 32, 2, 2, save_aux=True)
    feat = torch.empty(B, ED, H, H, device="cuda", dtype=torch.float32)
    value = torch.empty(B, ED, H, H, device="cuda", dtype=torch.bfloat16)
    ops.conv_fwd(ops.conv_desc(x, w_fold, feat, gn_fold=(sums, k1, 1e-5), e_shift=k0, out2=value))
    o_u, idx_u, smax_u = ops.cluster_core_fwd(feat, value, alpha, sbeta, heads, 2, 2, 2, 2, out_dtype=torch.bfloat16, save_aux=True)
    torch.cuda.synchronize()
    flips = (idx_f != idx_u).float().mean().item()
    e_s, e_o = rel_err(smax_f, smax_u), rel_err(o_f.float(), o_u.float())
    print(f"stage-3 fused projection+core B={B} heads={heads}: flips {flips:.2e} sim_max {e_s:.2e} out {e_o:.2e}")
    assert torch.isfinite(o_f.float()).all() and flips < 2e-4 and e_s < 1e-4 and e_o < 4e-3
