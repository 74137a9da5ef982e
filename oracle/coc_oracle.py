"""
CPU ORACLE for the ASY-VRNet CoC + fusion hot path.   *** TEST INFRASTRUCTURE ONLY ***

This file is a closed-form, functional (state-dict driven) restatement in plain PyTorch of the
algorithm the reference implements for the hot path.  It is NOT part of the product: only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it, and only as the checker / the timed CPU baseline.  The product path
(`asy-vrnet_b200/vrcoc`) never imports it and fails loudly when the CUDA library is missing.

Parity status: PINNED.  The reference ships no tests/golden vectors, so the oracle is
pinned against outputs of the reference itself: `oracle/make_golden.py` imports the unmodified
reference from /root/reference (through `oracle/ref_shim.py`) and writes the fixtures under
`tests/golden/`; `tests/test_oracle_golden.py` checks every function here against them, and
`tests/test_oracle_vs_reference.py` re-checks live (whole model included) whenever /root/reference
is present.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Conventions follow the reference: tensors are [B, C, dim2, dim3]; the reference calls dim2 "w" and
dim3 "h" (backbone/fusion/vr_coc.py:155), so `fold_w`/`proposal_w` act on dim2 and `fold_h`/
`proposal_h` on dim3.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# leaf helpers
# --------------------------------------------------------------------------------------------
def pairwise_cos_sim(x1, x2):
    """backbone/fusion/vr_coc.py:114-125 (== backbone/vision/context_cluster.py:86-97).
    x1 [...,M,D], x2 [...,N,D] -> [...,M,N]; normalize = x / max(||x||, 1e-12)."""
    n1 = x1 / x1.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    n2 = x2 / x2.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return n1 @ n2.transpose(-2, -1)


def data_normal(t):
    """backbone/fusion/vr_coc.py:59-67: WHOLE-tensor min/max normalisation (couples the batch)."""
    lo = t.min()
    if lo < 0:
        t = t + lo.abs()
        lo = t.min()
    hi = t.max()
    return (t - lo) / (hi - lo)


def shuffle_channels(x, groups=2):
    """backbone/fusion/vr_coc.py:70-80, neck/coc_fpn_dual.py:120-130: out[g-interleave]; identity when
    C % groups != 0."""
    b, c, h, w = x.shape
    if c % groups:
        return x
    return x.reshape(b, groups, c // groups, h, w).transpose(1, 2).reshape(b, c, h, w)


def group_norm1(x, weight, bias, eps=1e-5):
    """backbone/fusion/vr_coc.py:105-111: nn.GroupNorm(1, C): per-sample statistics over C*H*W."""
    b = x.shape[0]
    flat = x.reshape(b, -1)
    mu = flat.mean(dim=1).view(b, 1, 1, 1)
    var = flat.var(dim=1, unbiased=False).view(b, 1, 1, 1)
    xh = (x - mu) / torch.sqrt(var + eps)
    return xh * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def batch_norm(x, sd, prefix, training, eps, momentum=0.1, update=None):
    """nn.BatchNorm2d semantics (backbone/conv_utils/normal_conv.py:45, vr_coc.py:310,329).
    eval: running stats.  train: biased batch variance for normalisation; `update`, if given, is a dict
    that receives the new running_mean / running_var (unbiased variance, `momentum`)."""
    w = sd[prefix + "weight"]
    b = sd[prefix + "bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if update is not None:
            n = x.numel() / x.shape[1]
            with torch.no_grad():
                update[prefix + "running_mean"] = (1 - momentum) * sd[prefix + "running_mean"] + momentum * mean
                update[prefix + "running_var"] = (1 - momentum) * sd[prefix + "running_var"] + \
                    momentum * var * (n / max(n - 1, 1))
    else:
        mean = sd[prefix + "running_mean"]
        var = sd[prefix + "running_var"]
    inv = torch.rsqrt(var + eps)
    return (x - mean.view(1, -1, 1, 1)) * (inv * w).view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def base_conv(x, sd, prefix, ksize, stride=1, act="relu", training=False, ds_conv=False, update=None):
    """backbone/conv_utils/normal_conv.py:36-52: conv(pad=(k-1)//2, no bias) -> BN(eps 1e-3, mom .03) -> act.
    ds_conv=True: depthwise k x k (groups=Cin) followed by pointwise 1x1 (normal_conv.py:23-33)."""
    pad = (ksize - 1) // 2
    if ds_conv:
        wd = sd[prefix + "conv.dconv.weight"]
        y = F.conv2d(x, wd, sd.get(prefix + "conv.dconv.bias"), stride=stride, padding=pad, groups=wd.shape[0])
        y = F.conv2d(y, sd[prefix + "conv.pconv.weight"], sd.get(prefix + "conv.pconv.bias"))
    else:
        y = F.conv2d(x, sd[prefix + "conv.weight"], sd.get(prefix + "conv.bias"), stride=stride, padding=pad)
    y = batch_norm(y, sd, prefix + "bn.", training, eps=1e-3, momentum=0.03, update=update)
    if act == "relu":
        return F.relu(y)
    if act == "silu":
        return y * torch.sigmoid(y)
    if act == "lrelu":
        return F.leaky_relu(y, 0.1)
    raise AttributeError("Unsupported act type: {}".format(act))


def eca_kernel_size(channel, b=1, gamma=2):
    """backbone/attention_modules/eca.py:9-10."""
    k = int(abs((math.log(channel, 2) + b) / gamma))
    return k if k % 2 else k + 1


def eca_block(x, conv_w):
    """backbone/attention_modules/eca.py:16-22: x * sigmoid(conv1d_k(mean_hw(x))) along the channel axis."""
    k = conv_w.shape[-1]
    m = x.mean(dim=(2, 3))                                   # [B, C]
    a = F.conv1d(m.unsqueeze(1), conv_w.view(1, 1, k), padding=(k - 1) // 2).squeeze(1)
    return x * torch.sigmoid(a).unsqueeze(-1).unsqueeze(-1)


def shuffle_attention(x, sd, prefix, G):
    """backbone/attention_modules/shuffle_attention.py:48-72.  Channel c = g*(2q) + half*q + j, q = C/(2G).
    half 0: x * sigmoid(cweight_j * mean_hw(x) + cbias_j)
    half 1: x * sigmoid(sweight_j * GN_per_channel(x) + sbias_j), then channel_shuffle(.,2)."""
    b, c, h, w = x.shape
    q = c // (2 * G)
    xg = x.reshape(b, G, 2, q, h, w)
    x0, x1 = xg[:, :, 0], xg[:, :, 1]                        # [b,G,q,h,w]
    cw = sd[prefix + "cweight"].view(1, 1, q, 1, 1)
    cb = sd[prefix + "cbias"].view(1, 1, q, 1, 1)
    sw = sd[prefix + "sweight"].view(1, 1, q, 1, 1)
    sb = sd[prefix + "sbias"].view(1, 1, q, 1, 1)
    gw = sd[prefix + "gn.weight"].view(1, 1, q, 1, 1)
    gb = sd[prefix + "gn.bias"].view(1, 1, q, 1, 1)
    xc = x0 * torch.sigmoid(cw * x0.mean(dim=(3, 4), keepdim=True) + cb)
    mu = x1.mean(dim=(3, 4), keepdim=True)
    var = x1.var(dim=(3, 4), unbiased=False, keepdim=True)
    xs = x1 * torch.sigmoid(sw * ((x1 - mu) / torch.sqrt(var + 1e-5) * gw + gb) + sb)
    out = torch.stack([xc, xs], dim=2).reshape(b, c, h, w)
    return shuffle_channels(out, 2)                          # ShuffleAttention.channel_shuffle == same permutation


# --------------------------------------------------------------------------------------------
# context cluster
# --------------------------------------------------------------------------------------------
def pool_bins(n, p):
    """AdaptiveAvgPool bin i = [floor(i*n/p), ceil((i+1)*n/p)) (vr_coc.py:151,168)."""
    return [((i * n) // p, -((-(i + 1) * n) // p)) for i in range(p)]


def cluster_core(feat, value, alpha, beta, heads, fold_w, fold_h, proposal_w, proposal_h, aux=False):
    """backbone/fusion/vr_coc.py:158-190 (everything of Cluster.forward between the projections).
    feat, value: [B, E*D, H, W] -> out [B, E*D, H, W].
    Closed form per region-head: c_m = bin-mean(f), vc_m = bin-mean(v),
    s_mn = sigmoid(beta + alpha * cos(c_m, f_n)), k_n = argmax_m (first max), g_n = s_{k_n n},
    a_m = (sum_{k_n=m} g_n v_n + vc_m) / (cnt_m + 1), o_n = g_n * a_{k_n}.
    With aux=True also returns (idx [B,E,H,W] int64, sim_max [B,E,H,W], margin [B,E,H,W]) where margin is
    the top-1 minus top-2 similarity (used for the 'bit-exact where margin > 1e-5' gate)."""
    B, ED, H, W = feat.shape
    E = heads
    D = ED // E
    folded = fold_w > 1 and fold_h > 1
    f1, f2 = (fold_w, fold_h) if folded else (1, 1)
    assert H % f1 == 0 and W % f2 == 0, \
        f"Ensure the feature map size ({H}*{W}) can be divided by fold {fold_w}*{fold_h}"
    w, h = H // f1, W // f2

    def to_regions(t):
        t = t.reshape(B, E, D, f1, w, f2, h).permute(0, 1, 3, 5, 2, 4, 6)   # b e f1 f2 d w h
        return t.reshape(B * E * f1 * f2, D, w, h)

    f = to_regions(feat)
    v = to_regions(value)
    R, N, M = f.shape[0], w * h, proposal_w * proposal_h
    c = F.adaptive_avg_pool2d(f, (proposal_w, proposal_h)).reshape(R, D, M).transpose(1, 2)   # [R,M,D]
    vc = F.adaptive_avg_pool2d(v, (proposal_w, proposal_h)).reshape(R, D, M).transpose(1, 2)
    fp = f.reshape(R, D, N).transpose(1, 2)                                                    # [R,N,D]
    vp = v.reshape(R, D, N).transpose(1, 2)
    sim = torch.sigmoid(beta + alpha * pairwise_cos_sim(c, fp))                                # [R,M,N]
    g, k = sim.max(dim=1)                                                                      # first max wins on CPU
    onehot = F.one_hot(k, M).transpose(1, 2).to(sim.dtype)                                     # [R,M,N]
    sw = sim * onehot
    agg = (sw @ vp + vc) / (onehot.sum(dim=-1, keepdim=True) + 1.0)                            # [R,M,D]
    o = sw.transpose(1, 2) @ agg                                                               # [R,N,D]

    def from_regions(t, d):
        t = t.reshape(B, E, f1, f2, d, w, h).permute(0, 1, 4, 2, 5, 3, 6)                      # b e d f1 w f2 h
        return t.reshape(B, E * d, H, W)

    out = from_regions(o.transpose(1, 2).reshape(R, D, w, h), D)
    if not aux:
        return out
    top2 = sim.topk(min(2, M), dim=1).values
    margin = (top2[:, 0] - top2[:, 1]) if M > 1 else torch.full_like(g, float("inf"))
    idx = from_regions(k.reshape(R, 1, w, h), 1)
    gmap = from_regions(g.reshape(R, 1, w, h), 1)
    mmap = from_regions(margin.reshape(R, 1, w, h), 1)
    return out, idx, gmap, mmap


def cluster(x, sd, prefix, heads, fold_w=2, fold_h=2, proposal_w=2, proposal_h=2, aux=False):
    """Cluster.forward, backbone/fusion/vr_coc.py:155-192: fc_v / fc1 1x1 projections, core, fc2."""
    value = F.conv2d(x, sd[prefix + "fc_v.weight"], sd[prefix + "fc_v.bias"])
    feat = F.conv2d(x, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"])
    core = cluster_core(feat, value, sd[prefix + "sim_alpha"], sd[prefix + "sim_beta"],
                        heads, fold_w, fold_h, proposal_w, proposal_h, aux=aux)
    o = core[0] if aux else core
    y = F.conv2d(o, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])
    return (y,) + tuple(core[1:]) if aux else y


def mlp(x, sd, prefix):
    """Mlp.forward, backbone/fusion/vr_coc.py:217-223 (drop = 0): fc2(GELU_erf(fc1(x)))."""
    hdn = F.gelu(F.conv2d(x, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"]))
    return F.conv2d(hdn, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])


def cluster_block(x, sd, prefix, heads, fold_w=2, fold_h=2, proposal_w=2, proposal_h=2, use_layer_scale=True):
    """ClusterBlock.forward, backbone/fusion/vr_coc.py:264-275 (drop_path = 0)."""
    t = cluster(group_norm1(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"]), sd,
                prefix + "token_mixer.", heads, fold_w, fold_h, proposal_w, proposal_h)
    if use_layer_scale:
        t = sd[prefix + "layer_scale_1"].view(1, -1, 1, 1) * t
    x = x + t
    m = mlp(group_norm1(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"]), sd, prefix + "mlp.")
    if use_layer_scale:
        m = sd[prefix + "layer_scale_2"].view(1, -1, 1, 1) * m
    return x + m


def point_reducer(x, sd, prefix, stride, padding):
    """PointRecuder.forward, backbone/fusion/vr_coc.py:99-102 (norm_layer None)."""
    return F.conv2d(x, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"], stride=stride, padding=padding)


# --------------------------------------------------------------------------------------------
# asymmetric fusion
# --------------------------------------------------------------------------------------------
def image_enhance_by_radar(image, radar, sd, prefix, training=False, update=None):
    """ImageEnhanceByRadar.forward, backbone/fusion/vr_coc.py:312-316."""
    k = base_conv(radar, sd, prefix + "radar_projection.", 3, training=training, update=update)
    y = (1 + data_normal(k)) * image
    return batch_norm(y, sd, prefix + "norm.", training, eps=1e-5, momentum=0.1, update=update)


def radar_enhance_by_image(image, radar, sd, prefix, initial=False, training=False, update=None):
    """RadarEnhanceByImage.forward, backbone/fusion/vr_coc.py:331-359 (ShuffleAttention G=4 skipped when
    initial=True)."""
    ia = image if initial else shuffle_attention(image, sd, prefix + "image_attn.", G=4)
    z = shuffle_channels(torch.cat([ia, radar], dim=1), 2)
    z = eca_block(z, sd[prefix + "channel_attn.conv.weight"])
    u = base_conv(z, sd, prefix + "inverse_projection.", 1, training=training, update=update)
    return batch_norm(u + radar, sd, prefix + "norm.", training, eps=1e-5, momentum=0.1, update=update)


# --------------------------------------------------------------------------------------------
# dual-branch backbone
# --------------------------------------------------------------------------------------------
COC_SMALL = dict(layers=[2, 2, 6, 2], mlp_ratios=[8, 8, 4, 4], proposal_w=[2, 2, 2, 2], proposal_h=[2, 2, 2, 2],
                 fold_w=[8, 4, 2, 1], fold_h=[8, 4, 2, 1], heads=[4, 4, 8, 8], head_dim=[32, 32, 32, 32],
                 down_patch_size=3, down_stride=2, down_pad=1, in_patch_size=4, in_stride=4, in_pad=0)


def coc_small_cfg(width=1.0):
    """coc_small, backbone/fusion/vr_coc.py:760-782."""
    cfg = dict(COC_SMALL)
    cfg["embed_dims"] = [int(64 * width), int(128 * width), int(320 * width), int(512 * width)]
    return cfg


def vrcoc_forward(x, x_radar, sd, cfg, prefix="", training=False, update=None):
    """VRCoC.forward = forward_embeddings + forward_tokens, backbone/fusion/vr_coc.py:575-704.
    Returns (outs[4], outs_radar[4])."""
    p = prefix
    x = point_reducer(x, sd, p + "image_initial.", 1, 0)
    r = point_reducer(x_radar, sd, p + "radar_initial.", 1, 0)
    x = image_enhance_by_radar(x, r, sd, p + "image_enhance_by_radar1.", training, update)
    r = radar_enhance_by_image(x, r, sd, p + "radar_enhance_by_image1.", True, training, update)
    pos = sd[p + "fea_pos"].permute(2, 0, 1).unsqueeze(0).expand(x.shape[0], -1, -1, -1)
    x = point_reducer(torch.cat([x, pos], 1), sd, p + "patch_embed.", cfg["in_stride"], cfg["in_pad"])
    r = point_reducer(torch.cat([r, pos], 1), sd, p + "patch_embed_radar.", cfg["in_stride"], cfg["in_pad"])
    outs, outs_r = [], []
    nstage = len(cfg["layers"])
    for i in range(nstage):
        kw = dict(heads=cfg["heads"][i], fold_w=cfg["fold_w"][i], fold_h=cfg["fold_h"][i],
                  proposal_w=cfg["proposal_w"][i], proposal_h=cfg["proposal_h"][i])
        for j in range(cfg["layers"][i]):
            x = cluster_block(x, sd, f"{p}network.{3 * i}.{j}.", **kw)
        for j in range(cfg["layers"][i]):
            r = cluster_block(r, sd, f"{p}network_radar.{3 * i}.{j}.", **kw)
        x = image_enhance_by_radar(x, r, sd, f"{p}network.{3 * i + 1}.", training, update)
        r = radar_enhance_by_image(x, r, sd, f"{p}network_radar.{3 * i + 1}.", False, training, update)
        if i in (0, nstage - 1):
            outs.append(x)
            outs_r.append(r)
        if i < nstage - 1:
            x = point_reducer(x, sd, f"{p}network.{3 * i + 2}.", cfg["down_stride"], cfg["down_pad"])
            r = point_reducer(r, sd, f"{p}network_radar.{3 * i + 2}.", cfg["down_stride"], cfg["down_pad"])
            if i in (0, 1):
                outs.append(x)
                outs_r.append(r)
    return outs, outs_r


# --------------------------------------------------------------------------------------------
# neck + head (the callers either side of the path; needed for the whole-model frames/s baseline)
# --------------------------------------------------------------------------------------------
def coc_upsample(x, sd, prefix, scale, training=False, update=None):
    """CoCUpsample, neck/coc_fpn_dual.py:15-26."""
    y = base_conv(x, sd, prefix + "upsample.0.", 1, training=training, update=update)
    return F.interpolate(y, scale_factor=scale, mode="bilinear", align_corners=True)


def aspp(x, sd, prefix, training=False, update=None):
    """ASPP, neck/coc_fpn_dual.py:46-104 (rate 1, bn momentum 0.1, eps 1e-5)."""
    b, c, row, col = x.shape

    def branch(name, dil):
        w = sd[f"{prefix}{name}.0.weight"]
        y = F.conv2d(x, w, sd[f"{prefix}{name}.0.bias"], padding=dil if w.shape[-1] == 3 else 0, dilation=dil)
        return F.relu(batch_norm(y, sd, f"{prefix}{name}.1.", training, 1e-5, 0.1, update))

    outs = [branch("branch1", 1), branch("branch2", 6), branch("branch3", 12), branch("branch4", 18)]
    gf = x.mean(dim=2, keepdim=True).mean(dim=3, keepdim=True)
    gf = F.conv2d(gf, sd[prefix + "branch5_conv.weight"], sd[prefix + "branch5_conv.bias"])
    gf = F.relu(batch_norm(gf, sd, prefix + "branch5_bn.", training, 1e-5, 0.1, update))
    gf = F.interpolate(gf, (row, col), mode="bilinear", align_corners=True)
    cat = torch.cat(outs + [gf], dim=1)
    y = F.conv2d(cat, sd[prefix + "conv_cat.0.weight"], sd[prefix + "conv_cat.0.bias"])
    return F.relu(batch_norm(y, sd, prefix + "conv_cat.1.", training, 1e-5, 0.1, update))


def coc_conv(x, sd, prefix, training=False, update=None):
    """CoC_Conv, neck/coc_fpn_dual.py:29-39: default-arg ClusterBlock (heads 4, head_dim 24, fold 2x2) + 1x1 BaseConv."""
    y = cluster_block(x, sd, prefix + "coc.", heads=4)
    return base_conv(y, sd, prefix + "conv_att.", 1, training=training, update=update)


def neck_forward(x, x_radar, sd, cfg, prefix="", training=False, update=None):
    """CoCFpnDual.forward, neck/coc_fpn_dual.py:184-224."""
    p = prefix
    (s2, s3, s4, s5), (r2, r3, r4, r5) = vrcoc_forward(x, x_radar, sd, cfg, p + "backbone.", training, update)
    s5 = aspp(s5, sd, p + "aspp.", training, update)
    t = torch.cat([s4, coc_upsample(s5, sd, p + "upsample5_4.", 2, training, update)], 1)
    t = shuffle_attention(shuffle_channels(t), sd, p + "sc_attn_seg4.", G=8)
    t = torch.cat([coc_upsample(t, sd, p + "upsample4_3.", 2, training, update), s3], 1)
    t = shuffle_attention(shuffle_channels(t), sd, p + "sc_attn_seg3.", G=8)
    t = torch.cat([coc_upsample(t, sd, p + "upsample3_2.", 2, training, update), s2], 1)
    t = shuffle_attention(shuffle_channels(t), sd, p + "sc_attn_seg2.", G=8)
    seg = coc_upsample(t, sd, p + "upsample2_0.", 4, training, update)
    p5 = coc_conv(r5, sd, p + "p5_out_det.", training, update)
    p4 = coc_conv(torch.cat([r4, coc_upsample(p5, sd, p + "p5_4_det.", 2, training, update)], 1),
                  sd, p + "p4_out_det.", training, update)
    p3 = coc_conv(torch.cat([r3, coc_upsample(p4, sd, p + "p4_3_det.", 2, training, update)], 1),
                  sd, p + "p3_out_det.", training, update)
    return (p3, p4, p5), seg


def head_forward(feats, sd, prefix="", training=False, update=None):
    """DecoupleHead.forward, head/decouplehead.py:42-88 (depthwise=True as built by EfficientVRNet:22 is ignored by the
    reference ctor: ds_conv=True is hard-wired for the 3x3 towers)."""
    outs = []
    for k, x in enumerate(feats):
        x = base_conv(x, sd, f"{prefix}stems.{k}.", 1, training=training, update=update)
        c = x
        for j in range(2):
            c = base_conv(c, sd, f"{prefix}cls_convs.{k}.{j}.", 3, training=training, ds_conv=True, update=update)
        cls = F.conv2d(c, sd[f"{prefix}cls_preds.{k}.weight"], sd[f"{prefix}cls_preds.{k}.bias"])
        g = x
        for j in range(2):
            g = base_conv(g, sd, f"{prefix}reg_convs.{k}.{j}.", 3, training=training, ds_conv=True, update=update)
        reg = F.conv2d(g, sd[f"{prefix}reg_preds.{k}.weight"], sd[f"{prefix}reg_preds.{k}.bias"])
        obj = F.conv2d(g, sd[f"{prefix}obj_preds.{k}.weight"], sd[f"{prefix}obj_preds.{k}.bias"])
        outs.append(torch.cat([reg, obj, cls], 1))
    return outs


PHI_WIDTH = {"nano": 0.25, "tiny": 0.375, "s": 0.50, "m": 0.75, "l": 1.00}


def efficient_vrnet_forward(x, x_radar, sd, phi="l", training=False, update=None):
    """EfficientVRNet.forward, nets/efficient_vrnet.py:13-27 -> (det[3], seg)."""
    cfg = coc_small_cfg(PHI_WIDTH[phi])
    fpn, seg = neck_forward(x, x_radar, sd, cfg, "backbone.", training, update)
    return head_forward(fpn, sd, "head.", training, update), seg


def decode_outputs(outputs, input_shape):
    """utils/utils_bbox.py:32-84 (decode_outputs) without the device plumbing: [B, 5+nc, h, w] x 3 -> [B, sum hw, 5+nc] with
    xy = (xy + cell) * stride, wh = exp(wh) * stride, both normalised by the input size, sigmoid on objectness / classes.
    The reference uses stride = input_shape[0] / h for both axes (:63)."""
    hw = [o.shape[-2:] for o in outputs]
    out = torch.cat([o.flatten(start_dim=2) for o in outputs], dim=2).permute(0, 2, 1).clone()
    out[:, :, 4:] = torch.sigmoid(out[:, :, 4:])
    grids, strides = [], []
    for h, w in hw:
        gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        grids.append(torch.stack((gx, gy), 2).view(1, -1, 2))
        strides.append(torch.full((1, h * w, 1), input_shape[0] / h))
    grids = torch.cat(grids, dim=1).to(out.dtype)
    strides = torch.cat(strides, dim=1).to(out.dtype)
    out[..., :2] = (out[..., :2] + grids) * strides
    out[..., 2:4] = torch.exp(out[..., 2:4]) * strides
    out[..., [0, 2]] = out[..., [0, 2]] / input_shape[1]
    out[..., [1, 3]] = out[..., [1, 3]] / input_shape[0]
    return out

