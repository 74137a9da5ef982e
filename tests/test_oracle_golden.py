"""CPU: the oracle restatement (oracle/coc_oracle.py) against every golden vector produced by the UNMODIFIED
reference (oracle/make_golden.py).  This is what pins the oracle (the reference has no tests of its own)."""
import pytest
import torch

from golden_util import Fixture, names, rel_err
from oracle import coc_oracle as O

TOL = 1e-10      # float64 oracle vs float64 reference


def _grads(out, gout, tensors):
    gs = torch.autograd.grad(out, tensors, gout, allow_unused=True)
    return gs


@pytest.mark.parametrize("name", names("core_"))
def test_cluster_core(name):
    fx = Fixture(name)
    c = fx.cfg
    feat = fx.inp["feat"].double().requires_grad_(True)
    value = fx.inp["value"].double().requires_grad_(True)
    alpha = fx.inp["alpha"].double().requires_grad_(True)
    beta = fx.inp["beta"].double().requires_grad_(True)
    out, idx, g, margin = O.cluster_core(feat, value, alpha, beta, c["E"], c["fold_w"], c["fold_h"], c["proposal_w"],
                                         c["proposal_h"], aux=True)
    assert rel_err(out, fx.out["y"]) < TOL
    assert torch.equal(idx.to(torch.int32), fx.out["idx"])
    assert rel_err(g, fx.out["sim_max"]) < TOL
    assert rel_err(margin, fx.out["margin"]) < 1e-8
    gs = _grads(out, fx.inp["gout"].double(), [feat, value, alpha, beta])
    for got, key in zip(gs, ["feat", "value", "alpha", "beta"]):
        assert rel_err(got, fx.grad_in[key]) < TOL, key


def _sd64(fx, grad=True):
    return {k: (v.double().requires_grad_(grad) if v.is_floating_point() and v.numel() else v) for k, v in fx.sd.items()}


def _check_param_grads(fx, sd, out, gout, x):
    keys = [k for k in fx.grad_sd]
    gs = _grads(out, gout, [x] + [sd[k] for k in keys])
    assert rel_err(gs[0], fx.grad_in["x"]) < TOL
    for k, g in zip(keys, gs[1:]):
        assert g is not None, k
        assert rel_err(g, fx.grad_sd[k]) < TOL, k


@pytest.mark.parametrize("name", names("cluster_"))
def test_cluster_module(name):
    fx = Fixture(name)
    c = fx.cfg
    sd = _sd64(fx)
    x = fx.inp["x"].double().requires_grad_(True)
    y = O.cluster(x, sd, "", c["heads"], c["fold_w"], c["fold_h"], c["proposal_w"], c["proposal_h"])
    assert rel_err(y, fx.out["y"]) < TOL
    _check_param_grads(fx, sd, y, fx.inp["gout"].double(), x)


def test_mlp():
    fx = Fixture("mlp_c16")
    sd = _sd64(fx)
    x = fx.inp["x"].double().requires_grad_(True)
    y = O.mlp(x, sd, "")
    assert rel_err(y, fx.out["y"]) < TOL
    _check_param_grads(fx, sd, y, fx.inp["gout"].double(), x)


@pytest.mark.parametrize("name", names("block_"))
def test_cluster_block(name):
    fx = Fixture(name)
    c = fx.cfg
    sd = _sd64(fx)
    x = fx.inp["x"].double().requires_grad_(True)
    y = O.cluster_block(x, sd, "", c.get("heads", 4), c.get("fold_w", 2), c.get("fold_h", 2), c.get("proposal_w", 2),
                        c.get("proposal_h", 2))
    assert rel_err(y, fx.out["y"]) < TOL
    _check_param_grads(fx, sd, y, fx.inp["gout"].double(), x)


def test_leaves():
    fx = Fixture("shuffle_attention_c32_g4")
    sd = _sd64(fx)
    x = fx.inp["x"].double().requires_grad_(True)
    y = O.shuffle_attention(x, sd, "", fx.cfg["G"])
    assert rel_err(y, fx.out["y"]) < TOL
    _check_param_grads(fx, sd, y, fx.inp["gout"].double(), x)

    fx = Fixture("eca_c32")
    sd = _sd64(fx)
    x = fx.inp["x"].double().requires_grad_(True)
    y = O.eca_block(x, sd["conv.weight"])
    assert rel_err(y, fx.out["y"]) < TOL
    _check_param_grads(fx, sd, y, fx.inp["gout"].double(), x)

    fx = Fixture("point_reducer_k3s2")
    sd = _sd64(fx)
    x = fx.inp["x"].double().requires_grad_(True)
    y = O.point_reducer(x, sd, "", fx.cfg["stride"], fx.cfg["padding"])
    assert rel_err(y, fx.out["y"]) < TOL
    _check_param_grads(fx, sd, y, fx.inp["gout"].double(), x)

    fx = Fixture("data_normal")
    assert rel_err(O.data_normal(fx.inp["x"].double()), fx.out["y"]) < TOL
    assert rel_err(O.data_normal(fx.inp["xpos"].double()), fx.out["ypos"]) < TOL
    assert O.eca_kernel_size(7) == 1 and O.eca_kernel_size(32) == 3 and O.eca_kernel_size(128) == 5 and O.eca_kernel_size(1024) == 5


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_base_conv(mode):
    fx = Fixture(f"base_conv_k3_{mode}")
    sd = _sd64(fx)
    x = fx.inp["x"].double().requires_grad_(True)
    upd = {}
    y = O.base_conv(x, sd, "", fx.cfg["ksize"], fx.cfg["stride"], training=fx.cfg["training"], update=upd)
    assert rel_err(y, fx.out["y"]) < TOL
    _check_param_grads(fx, sd, y, fx.inp["gout"].double(), x)
    for k, v in upd.items():
        assert rel_err(v, fx.out["new." + k]) < TOL, k


@pytest.mark.parametrize("name", names("image_enhance_") + names("radar_enhance_"))
def test_fusion(name):
    fx = Fixture(name)
    c = fx.cfg
    sd = _sd64(fx)
    img = fx.inp["image"].double().requires_grad_(True)
    rad = fx.inp["radar"].double().requires_grad_(True)
    upd = {}
    if name.startswith("image_enhance"):
        y = O.image_enhance_by_radar(img, rad, sd, "", c["training"], upd)
    else:
        y = O.radar_enhance_by_image(img, rad, sd, "", c["initial"], c["training"], upd)
    assert rel_err(y, fx.out["y"]) < TOL
    if "gout" in fx.inp:
        keys = list(fx.grad_sd)
        gs = _grads(y, fx.inp["gout"].double(), [img, rad] + [sd[k] for k in keys])
        assert rel_err(gs[0], fx.grad_in["image"]) < TOL
        assert rel_err(gs[1], fx.grad_in["radar"]) < TOL
        for k, g in zip(keys, gs[2:]):
            assert rel_err(g, fx.grad_sd[k]) < TOL, k
    for k, v in upd.items():
        assert rel_err(v, fx.out["new." + k]) < TOL, k


def test_vrcoc_mini():
    fx = Fixture("vrcoc_mini_eval")
    sd = _sd64(fx, grad=False)
    cfg = dict(fx.cfg)
    cfg.update(in_stride=4, in_pad=0)
    with torch.no_grad():
        outs, outs_r = O.vrcoc_forward(fx.inp["x"].double(), fx.inp["x_radar"].double(), sd, cfg)
    for i in range(4):
        assert rel_err(outs[i], fx.out[f"img{i}"]) < TOL, i
        assert rel_err(outs_r[i], fx.out[f"radar{i}"]) < TOL, i
