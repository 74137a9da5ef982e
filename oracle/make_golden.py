"""
Golden-vector generator.   *** TEST INFRASTRUCTURE ONLY — runs in the dev container, never on the GPU box ***

Imports the UNMODIFIED reference from /root/reference (oracle/ref_shim.py), runs its own nn.Modules in
float64 on seeded float32-representable inputs/weights, and writes small .npz fixtures to tests/golden/.
The reference ships no golden vectors of its own, so these fixtures are what pins the
oracle (oracle/coc_oracle.py) and, through it, the CUDA path.

    python oracle/make_golden.py            # regenerate everything (deterministic)

Fixture layout (.npz): `cfg` (json string), `in.*` inputs, `sd.*` state-dict entries (reference key names),
`out.*` float64 reference outputs, `grad.in.*` / `grad.sd.*` float64 reference gradients for the upstream
gradient `in.gout`.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def f32(t):
    """float64 tensor holding float32-representable values."""
    return t.float().double()


def randomize_hidden(mod, gen):
    """Give O(1) values to the parameters whose stock init hides errors."""
    for name, p in mod.named_parameters():
        leaf = name.split(".")[-1]
        if p.numel() == 0:
            continue
        with torch.no_grad():
            if leaf.startswith("layer_scale"):
                p.copy_(f32(torch.rand(p.shape, generator=gen, dtype=torch.double) + 0.5))
            elif leaf == "sim_alpha":
                p.copy_(f32(torch.rand(p.shape, generator=gen, dtype=torch.double) * 1.5 + 0.5))
            elif leaf == "sim_beta":
                p.copy_(f32(torch.rand(p.shape, generator=gen, dtype=torch.double) - 0.5))
            elif leaf in ("cweight", "cbias", "sweight", "sbias"):
                p.copy_(f32(torch.randn(p.shape, generator=gen, dtype=torch.double)))
            elif leaf == "bias":
                p.copy_(f32(torch.randn(p.shape, generator=gen, dtype=torch.double) * 0.1))
            elif leaf == "weight" and p.ndim == 1:          # norm affine weights
                p.copy_(f32(torch.rand(p.shape, generator=gen, dtype=torch.double) + 0.5))
            else:
                p.copy_(f32(p.detach()))
    for name, b in mod.named_buffers():
        leaf = name.split(".")[-1]
        with torch.no_grad():
            if leaf == "running_mean":
                b.copy_(f32(torch.randn(b.shape, generator=gen, dtype=torch.double) * 0.2))
            elif leaf == "running_var":
                b.copy_(f32(torch.rand(b.shape, generator=gen, dtype=torch.double) + 0.5))


def save(name, cfg, inputs, mod, outs, grads_in=None, sd_before=None):
    rec = {"cfg": np.array(json.dumps(cfg))}
    for k, v in inputs.items():
        rec["in." + k] = v.detach().float().numpy()                  # float32-representable by construction
    sd = sd_before if sd_before is not None else (mod.state_dict() if mod is not None else {})
    for k, v in sd.items():
        rec["sd." + k] = v.detach().float().numpy() if v.is_floating_point() else v.detach().numpy()
    for k, v in outs.items():
        rec["out." + k] = v.detach().numpy()
    if grads_in:
        for k, v in grads_in.items():
            rec["grad.in." + k] = v.detach().numpy()
    if mod is not None and grads_in is not None:
        for k, p in mod.named_parameters():
            if p.grad is not None:
                rec["grad.sd." + k] = p.grad.detach().numpy()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    ref_shim.install()
    from backbone.fusion import vr_coc as R
    from backbone.vision import context_cluster as RV
    from backbone.attention_modules.shuffle_attention import ShuffleAttention
    from backbone.attention_modules.eca import eca_block
    from backbone.conv_utils.normal_conv import BaseConv
    from einops import rearrange

    gen = torch.Generator().manual_seed(20261017)
    stream = {"gen": gen}                                            # the extra cases at the end draw from their own stream

    def randn(*s):
        return f32(torch.randn(*s, generator=stream["gen"], dtype=torch.double))

    def rand(*s):
        return f32(torch.rand(*s, generator=stream["gen"], dtype=torch.double))

    # ---- cluster core (the part of Cluster.forward between fc1/fc_v and fc2), via the reference module with
    #      identity projections so that the module's own code computes it -------------------------------------
    core_cases = [
        dict(name="core_16x16_f2_p2", B=2, E=2, D=8, H=32, W=32, fold_w=2, fold_h=2, proposal_w=2, proposal_h=2),
        dict(name="core_live_s1", B=1, E=2, D=32, H=32, W=32, fold_w=2, fold_h=2, proposal_w=2, proposal_h=2),
        dict(name="core_nofold_p3", B=2, E=3, D=8, H=12, W=14, fold_w=1, fold_h=1, proposal_w=3, proposal_h=3),
        dict(name="core_overlap_bins", B=1, E=2, D=6, H=24, W=14, fold_w=2, fold_h=2, proposal_w=5, proposal_h=2),
        dict(name="core_8x8_d24", B=2, E=4, D=24, H=16, W=16, fold_w=2, fold_h=2, proposal_w=2, proposal_h=2),
        dict(name="core_p4_m16", B=1, E=2, D=16, H=32, W=16, fold_w=2, fold_h=1, proposal_w=4, proposal_h=4),
    ]
    # added after the first fixtures were committed: own random stream (so the files above never change) and, with
    # `--extra-only`, generated alone.  D=24 on 16x16 regions = neck level p4, the second shape of the compile-time kernel.
    extra_core_cases = [
        dict(name="core_16x16_d24", B=1, E=2, D=24, H=32, W=32, fold_w=2, fold_h=2, proposal_w=2, proposal_h=2),
    ]
    extra_only = "--extra-only" in sys.argv
    if extra_only:
        stream["gen"] = torch.Generator().manual_seed(20261018)
    for c in (extra_core_cases if extra_only else core_cases):
        ED = c["E"] * c["D"]
        m = R.Cluster(ED, ED, c["proposal_w"], c["proposal_h"], c["fold_w"], c["fold_h"], c["E"], c["D"]).double()
        alpha, beta = f32(rand(1) * 1.5 + 0.5), f32(rand(1) - 0.5)
        with torch.no_grad():
            m.sim_alpha.copy_(alpha)
            m.sim_beta.copy_(beta)
        feat = randn(c["B"], ED, c["H"], c["W"]).requires_grad_(True)
        value = randn(c["B"], ED, c["H"], c["W"]).requires_grad_(True)
        gout = randn(c["B"], ED, c["H"], c["W"])
        # run exactly the reference's lines vr_coc.py:158-190 by calling forward with fc1/fc_v/fc2 swapped for
        # pass-through modules (the module code is unmodified; only its three conv children are replaced)
        class Pick(torch.nn.Module):
            def __init__(self, t):
                super().__init__()
                self.t = t

            def forward(self, _):
                return self.t
        m.fc1, m.fc_v, m.fc2 = Pick(feat), Pick(value), torch.nn.Identity()
        out = m(feat)
        out.backward(gout)
        # assignment data recomputed with the reference's own helper on the reference's own rearranges
        with torch.no_grad():
            x = rearrange(feat, "b (e c) w h -> (b e) c w h", e=c["E"])
            if c["fold_w"] > 1 and c["fold_h"] > 1:
                x = rearrange(x, "b c (f1 w) (f2 h) -> (b f1 f2) c w h", f1=c["fold_w"], f2=c["fold_h"])
            b, cc, w, h = x.shape
            cen = m.centers_proposal(x)
            sim = torch.sigmoid(m.sim_beta + m.sim_alpha * R.pairwise_cos_sim(
                cen.reshape(b, cc, -1).permute(0, 2, 1), x.reshape(b, cc, -1).permute(0, 2, 1)))
            smax, sidx = sim.max(dim=1, keepdim=True)
            top2 = sim.topk(2, dim=1).values
            margin = (top2[:, 0] - top2[:, 1]).reshape(b, 1, w, h)
            idx = sidx.reshape(b, 1, w, h).double()
            smax = smax.reshape(b, 1, w, h)

            def back(t):
                if c["fold_w"] > 1 and c["fold_h"] > 1:
                    t = rearrange(t, "(b f1 f2) c w h -> b c (f1 w) (f2 h)", f1=c["fold_w"], f2=c["fold_h"])
                return rearrange(t, "(b e) c w h -> b (e c) w h", e=c["E"])
            idx, smax, margin = back(idx), back(smax), back(margin)
        cfg = {k: v for k, v in c.items() if k != "name"}
        rec_in = {"feat": feat, "value": value, "alpha": alpha, "beta": beta, "gout": gout}
        save(c["name"], cfg, rec_in, None,
             {"y": out, "idx": idx.to(torch.int32), "sim_max": smax, "margin": margin},
             {"feat": feat.grad, "value": value.grad, "alpha": m.sim_alpha.grad, "beta": m.sim_beta.grad})
    if extra_only:
        return

    # ---- Cluster / Mlp / ClusterBlock modules -----------------------------------------------------------------
    mod_cases = [
        ("cluster_c16", R.Cluster, dict(dim=16, out_dim=16, proposal_w=2, proposal_h=2, fold_w=2, fold_h=2, heads=2, head_dim=8), (2, 16, 16, 16)),
        ("cluster_vision_c24", RV.Cluster, dict(dim=24, out_dim=24, proposal_w=2, proposal_h=2, fold_w=1, fold_h=1, heads=4, head_dim=24), (1, 24, 8, 8)),
        ("mlp_c16", R.Mlp, dict(in_features=16, hidden_features=64), (2, 16, 12, 12)),
        ("block_c16", R.ClusterBlock, dict(dim=16, mlp_ratio=4.0, proposal_w=2, proposal_h=2, fold_w=2, fold_h=2, heads=2, head_dim=8), (2, 16, 16, 16)),
        ("block_live_s1_small", R.ClusterBlock, dict(dim=64, mlp_ratio=8.0, proposal_w=2, proposal_h=2, fold_w=2, fold_h=2, heads=4, head_dim=32), (1, 64, 32, 32)),
        ("block_neck_default", RV.ClusterBlock, dict(dim=32), (2, 32, 16, 16)),
        ("block_stock_ls", R.ClusterBlock, dict(dim=16, mlp_ratio=4.0, heads=2, head_dim=8), (1, 16, 16, 16)),
    ]
    for name, cls, kw, shape in mod_cases:
        m = cls(**kw).double()
        if name != "block_stock_ls":
            randomize_hidden(m, gen)
        else:
            for p in m.parameters():
                p.data = f32(p.data)
        x = randn(*shape).requires_grad_(True)
        y = m(x)
        gout = randn(*y.shape)
        y.backward(gout)
        save(name, kw, {"x": x, "gout": gout}, m, {"y": y}, {"x": x.grad})

    # ---- leaves ------------------------------------------------------------------------------------------------
    m = ShuffleAttention(channel=32, G=4).double()
    randomize_hidden(m, gen)
    x = randn(2, 32, 10, 12).requires_grad_(True)
    y = m(x)
    gout = randn(*y.shape)
    y.backward(gout)
    save("shuffle_attention_c32_g4", dict(channel=32, G=4), {"x": x, "gout": gout}, m, {"y": y}, {"x": x.grad})

    m = eca_block(channel=32).double()
    randomize_hidden(m, gen)
    x = randn(2, 32, 6, 6).requires_grad_(True)
    y = m(x)
    gout = randn(*y.shape)
    y.backward(gout)
    save("eca_c32", dict(channel=32), {"x": x, "gout": gout}, m, {"y": y}, {"x": x.grad})

    for tr in (False, True):
        m = BaseConv(6, 10, 3, 1).double().train(tr)
        randomize_hidden(m, gen)
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        x = randn(2, 6, 9, 11).requires_grad_(True)
        y = m(x)
        gout = randn(*y.shape)
        y.backward(gout)
        outs = {"y": y}
        if tr:
            outs.update({"new." + k: v for k, v in m.state_dict().items() if "running" in k})
        save(f"base_conv_k3_{'train' if tr else 'eval'}", dict(in_channels=6, out_channels=10, ksize=3, stride=1, training=tr),
             {"x": x, "gout": gout}, m, outs, {"x": x.grad}, sd_before=sd0)

    m = R.PointRecuder(patch_size=3, stride=2, padding=1, in_chans=8, embed_dim=12).double()
    randomize_hidden(m, gen)
    x = randn(2, 8, 16, 16).requires_grad_(True)
    y = m(x)
    gout = randn(*y.shape)
    y.backward(gout)
    save("point_reducer_k3s2", dict(patch_size=3, stride=2, padding=1, in_chans=8, embed_dim=12), {"x": x, "gout": gout}, m, {"y": y}, {"x": x.grad})

    x = randn(2, 3, 5, 5)
    save("data_normal", {}, {"x": x, "xpos": x.abs()}, None,
         {"y": R.data_normal(x.clone()), "ypos": R.data_normal(x.abs().clone()),
          "shuffle": R.shuffle_channels(randn(1, 6, 2, 2).clone(), 2)})

    # ---- fusion ------------------------------------------------------------------------------------------------
    for tr in (False, True):
        tag = "train" if tr else "eval"
        m = R.ImageEnhanceByRadar(radar_in_channels=8, image_in_channels=8).double().train(tr)
        randomize_hidden(m, gen)
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        img = randn(2, 8, 16, 16).requires_grad_(True)
        rad = rand(2, 8, 16, 16).requires_grad_(True)
        y = m(img, rad)
        gout = randn(*y.shape)
        y.backward(gout)
        outs = {"y": y}
        if tr:
            outs.update({"new." + k: v for k, v in m.state_dict().items() if "running" in k})
        save(f"image_enhance_c8_{tag}", dict(radar_in_channels=8, image_in_channels=8, training=tr),
             {"image": img, "radar": rad, "gout": gout}, m, outs, {"image": img.grad, "radar": rad.grad}, sd_before=sd0)

        m = R.RadarEnhanceByImage(radar_in_channels=16, image_in_channels=16).double().train(tr)
        randomize_hidden(m, gen)
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        img = randn(2, 16, 16, 16).requires_grad_(True)
        rad = rand(2, 16, 16, 16).requires_grad_(True)
        y = m(img, rad)
        gout = randn(*y.shape)
        y.backward(gout)
        outs = {"y": y}
        if tr:
            outs.update({"new." + k: v for k, v in m.state_dict().items() if "running" in k})
        save(f"radar_enhance_c16_{tag}", dict(radar_in_channels=16, image_in_channels=16, initial=False, training=tr),
             {"image": img, "radar": rad, "gout": gout}, m, outs, {"image": img.grad, "radar": rad.grad}, sd_before=sd0)

    m = R.ImageEnhanceByRadar(radar_in_channels=4, image_in_channels=3).double().eval()
    randomize_hidden(m, gen)
    img, rad = randn(2, 3, 32, 32), rand(2, 4, 32, 32)
    save("image_enhance_initial_eval", dict(radar_in_channels=4, image_in_channels=3, training=False),
         {"image": img, "radar": rad}, m, {"y": m(img, rad)})
    m = R.RadarEnhanceByImage(radar_in_channels=4, image_in_channels=3, initial=True).double().eval()
    randomize_hidden(m, gen)
    img, rad = randn(2, 3, 32, 32), rand(2, 4, 32, 32)
    save("radar_enhance_initial_eval", dict(radar_in_channels=4, image_in_channels=3, initial=True, training=False),
         {"image": img, "radar": rad}, m, {"y": m(img, rad)})

    # ---- a miniature dual-branch backbone (same class, small dims) --------------------------------------------
    kw = dict(layers=[1, 1, 2, 1], embed_dims=[8, 16, 16, 32], mlp_ratios=[2, 2, 2, 2], downsamples=[True] * 4,
              down_patch_size=3, down_stride=2, down_pad=1, img_w=64, img_h=64,
              proposal_w=[2, 2, 2, 2], proposal_h=[2, 2, 2, 2], fold_w=[2, 2, 1, 1], fold_h=[2, 2, 1, 1],
              heads=[2, 2, 2, 2], head_dim=[4, 4, 8, 8])
    m = R.VRCoC(norm_layer=R.GroupNorm, **kw).double().eval()
    randomize_hidden(m, gen)
    x, r = randn(2, 3, 64, 64), rand(2, 4, 64, 64)
    outs, outs_r = m(x, r)
    rec = {f"img{i}": t for i, t in enumerate(outs)}
    rec.update({f"radar{i}": t for i, t in enumerate(outs_r)})
    save("vrcoc_mini_eval", kw, {"x": x, "x_radar": r}, m, rec)


if __name__ == "__main__":
    main()
