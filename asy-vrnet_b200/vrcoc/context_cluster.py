"""Context-cluster (CoC) modules with the reference's constructor / forward signatures and state-dict layout,
running on the hand-written sm_100a kernels.

Mirrors reference backbone/fusion/vr_coc.py:83-300 (identical to backbone/vision/context_cluster.py:55-273 and
backbone/radar/context_cluster.py): PointRecuder, GroupNorm, pairwise_cos_sim, Cluster, Mlp, ClusterBlock,
basic_blocks.  Children that hold weights stay real nn.Conv2d / nn.GroupNorm modules so that `weights_init`
(reference nets/yolo_training.py:482-500), the optimizer grouping (reference train.py:460-473), `load_state_dict`,
`deepcopy` (EMA) and nn.DataParallel behave exactly as with the reference; only `forward` is replaced.
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_GELU, ACT_NONE, VrcocError


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class DropPath(nn.Module):
    """Stochastic depth (timm.models.layers.DropPath semantics; reference vr_coc.py:10,258)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


def pairwise_cos_sim(x1: torch.Tensor, x2: torch.Tensor):
    """reference vr_coc.py:114-125.  Stand-alone helper kept for API parity; inside Cluster the similarity is computed
    on chip by vrcoc_cluster_core_fwd."""
    x1 = x1 / x1.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    x2 = x2 / x2.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    return torch.matmul(x1, x2.transpose(-2, -1))


class PointRecuder(nn.Module):
    """reference vr_coc.py:83-102: the point reducer is a strided conv; here one implicit-GEMM launch."""

    def __init__(self, patch_size=16, stride=16, padding=0, in_chans=3, embed_dim=768, norm_layer=None):
        super().__init__()
        patch_size, stride, padding = to_2tuple(patch_size), to_2tuple(stride), to_2tuple(padding)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride, padding=padding)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x, extra=None, extra_bstride=None):
        """`extra` (optional, [C1,H,W] or [B,C1,H,W]) is concatenated after x along the channels inside the kernel
        (the cat([x, pos]) of reference vr_coc.py:582-586 without materialising it)."""
        from .fusion import _conv_launch, conv2d_native
        if x.is_cuda and not torch.is_grad_enabled() and isinstance(self.norm, nn.Identity):
            # gradient-free: the reducer feeds the first ClusterBlock of a stage, whose GroupNorm needs the per-sample sums of
            # this output - they leave the convolution's epilogue with it instead of costing a pass (and a launch) of their own
            sums = ops.new_sample_sums(x.shape[0], x.device)
            y = _conv_launch(x, self.proj.weight, ops._f32(self.proj.bias), self.proj.stride[0], self.proj.padding[0], extra,
                             extra_bstride, ACT_NONE, None, None, None, out_sample_sums=sums)
            return ops.attach_sums(y, sums)
        y = conv2d_native(x, self.proj.weight, self.proj.bias, stride=self.proj.stride[0], pad=self.proj.padding[0],
                          extra=extra, extra_bstride=extra_bstride)
        return self.norm(y)


class GroupNorm(nn.GroupNorm):
    """reference vr_coc.py:105-111: GroupNorm with one group.  Inside ClusterBlock it is folded into the prologue of the
    following projection; called on its own it runs the same kernel with an identity weight."""

    def __init__(self, num_channels, **kwargs):
        super().__init__(1, num_channels, **kwargs)

    def forward(self, x):
        if not x.is_cuda:
            raise VrcocError("vrcoc GroupNorm needs a CUDA tensor (no CPU fallback exists)")
        if not torch.is_grad_enabled():
            # one streaming pass: statistics, then the normalisation as the stand-alone prologue kernel (no contraction)
            x = x.contiguous()
            out = torch.empty_like(x)
            d = ops.conv_desc(x, x, out, gn=(ops.sample_sums_of(x), ops._f32(self.weight), ops._f32(self.bias), self.eps))
            ops.check(ops.lib.vrcoc_table_apply(d, ops._stream()), "table_apply")
            return out
        C = x.shape[1]
        eye = torch.eye(C, device=x.device, dtype=x.dtype)          # with autograd: the projection Function with an identity weight
        return ops.GNProjFn.apply(x, ops.sample_sums_of(x), self.weight, self.bias, self.eps, eye, None, ACT_NONE, 0)


class Cluster(nn.Module):
    """reference vr_coc.py:128-192."""

    def __init__(self, dim, out_dim, proposal_w=2, proposal_h=2, fold_w=2, fold_h=2, heads=4, head_dim=24,
                 return_center=False):
        super().__init__()
        self.heads = heads
        self.head_dim = head_dim
        self.fc1 = nn.Conv2d(dim, heads * head_dim, kernel_size=1)
        self.fc2 = nn.Conv2d(heads * head_dim, out_dim, kernel_size=1)
        self.fc_v = nn.Conv2d(dim, heads * head_dim, kernel_size=1)
        self.sim_alpha = nn.Parameter(torch.ones(1))
        self.sim_beta = nn.Parameter(torch.zeros(1))
        self.centers_proposal = nn.AdaptiveAvgPool2d((proposal_w, proposal_h))
        self.fold_w = fold_w
        self.fold_h = fold_h
        self.return_center = return_center
        if return_center:
            raise VrcocError("return_center=True is deprecated in the reference (vr_coc.py:140) and not supported")

    # -- pieces shared by the stand-alone forward and the fused ClusterBlock path ---------------------------------
    def _proposal(self):
        pw, ph = self.centers_proposal.output_size
        return int(pw), int(ph)

    def _project_in(self, x, gn):
        """feat, value = fc1(x'), fc_v(x') in ONE launch (concatenated weights); x' = GroupNorm(x) when gn is given."""
        ED = self.heads * self.head_dim
        w = torch.cat([self.fc1.weight, self.fc_v.weight], dim=0)
        b = torch.cat([self.fc1.bias, self.fc_v.bias], dim=0)
        if gn is None:
            y = ops.GNProjFn.apply(x, None, None, None, 0.0, w, b, ACT_NONE, ED)
        else:
            sums, norm = gn
            y = ops.GNProjFn.apply(x, sums, norm.weight, norm.bias, norm.eps, w, b, ACT_NONE, ED)       # folds GN when it can
        if isinstance(y, tuple):          # bf16 storage: similarity operand stays fp32
            return y
        return y[:, :ED], y[:, ED:]

    def _core(self, feat, value):
        pw, ph = self._proposal()
        return ops.ClusterCoreFn.apply(feat, value, self.sim_alpha, self.sim_beta, self.heads, self.fold_w, self.fold_h, pw, ph)

    def forward(self, x):  # [b,c,w,h]
        if not x.is_cuda:
            raise VrcocError("vrcoc Cluster needs a CUDA tensor (no CPU fallback exists)")
        feat, value = self._project_in(x, None)
        out = self._core(feat, value)
        return ops.ProjFn.apply(out, self.fc2.weight, self.fc2.bias, ACT_NONE)


SPLIT_WIDE_GN = True        # C > 576: GroupNorm as its own streaming pass in front of mlp.fc1 (False: transform-on-load GEMM; A/B switch)


class Mlp(nn.Module):
    """reference vr_coc.py:195-223 (1x1-conv MLP, exact-erf GELU)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Conv2d(in_features, hidden_features, 1)
        self.act = act_layer()
        self.fc2 = nn.Conv2d(hidden_features, out_features, 1)
        self.drop = nn.Dropout(drop)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Conv2d):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def _fusable(self):
        return isinstance(self.act, nn.GELU) and getattr(self.act, "approximate", "none") == "none" and \
            (self.drop.p == 0.0 or not self.training)

    def forward(self, x):
        if not x.is_cuda:
            raise VrcocError("vrcoc Mlp needs a CUDA tensor (no CPU fallback exists)")
        if self._fusable():
            h = ops.ProjFn.apply(x, self.fc1.weight, self.fc1.bias, ACT_GELU)
            return ops.ProjFn.apply(h, self.fc2.weight, self.fc2.bias, ACT_NONE)
        h = self.drop(self.act(ops.ProjFn.apply(x, self.fc1.weight, self.fc1.bias, ACT_NONE)))
        return self.drop(ops.ProjFn.apply(h, self.fc2.weight, self.fc2.bias, ACT_NONE))


class ClusterBlock(nn.Module):
    """reference vr_coc.py:226-275.  forward =  x += ls1 * Cluster(GN(x));  x += ls2 * Mlp(GN(x))  in five launches:
    [GN+fc1|fc_v] -> [cluster core] -> [fc2 + ls1 + residual + stats] -> [GN+mlp.fc1+GELU] -> [mlp.fc2 + ls2 + residual + stats]."""

    def __init__(self, dim, mlp_ratio=4., act_layer=nn.GELU, norm_layer=GroupNorm, drop=0., drop_path=0.,
                 use_layer_scale=True, layer_scale_init_value=1e-5,
                 proposal_w=2, proposal_h=2, fold_w=2, fold_h=2, heads=4, head_dim=24, return_center=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.token_mixer = Cluster(dim=dim, out_dim=dim, proposal_w=proposal_w, proposal_h=proposal_h,
                                   fold_w=fold_w, fold_h=fold_h, heads=heads, head_dim=head_dim, return_center=False)
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.use_layer_scale = use_layer_scale
        if use_layer_scale:
            self.layer_scale_1 = nn.Parameter(layer_scale_init_value * torch.ones((dim)), requires_grad=True)
            self.layer_scale_2 = nn.Parameter(layer_scale_init_value * torch.ones((dim)), requires_grad=True)

    def _fused_ok(self):
        gn1 = isinstance(self.norm1, nn.GroupNorm) and self.norm1.num_groups == 1 and self.norm1.affine
        gn2 = isinstance(self.norm2, nn.GroupNorm) and self.norm2.num_groups == 1 and self.norm2.affine
        dp = isinstance(self.drop_path, nn.Identity) or not self.training
        return gn1 and gn2 and dp and self.mlp._fusable()

    def _forward_infer(self, x):
        """Gradient-free path: the same five launches, fed from memoised parameter views (concatenated fc1|fc_v weight,
        fp32 copies of the small vectors), no autograd bookkeeping."""
        tm, mlp, n1, n2 = self.token_mixer, self.mlp, self.norm1, self.norm2
        f32 = ops._f32
        x = x.contiguous()
        B, C, H, W = x.shape
        ED = tm.heads * tm.head_dim
        w_in = ops.cached(self, "w_in", [tm.fc1.weight, tm.fc_v.weight],
                          lambda: torch.cat([tm.fc1.weight, tm.fc_v.weight], 0).reshape(2 * ED, C).contiguous())
        b_in = ops.cached(self, "b_in", [tm.fc1.bias, tm.fc_v.bias], lambda: torch.cat([tm.fc1.bias, tm.fc_v.bias], 0).float().contiguous())
        ls1 = f32(self.layer_scale_1) if self.use_layer_scale else None
        ls2 = f32(self.layer_scale_2) if self.use_layer_scale else None
        dev, dt = x.device, x.dtype
        sums0 = ops.sample_sums_of(x)
        pw, ph = tm._proposal()
        fold_w = None
        if ops.gn_fold_ok(x, 2 * ED, ED) and tm.fc1.weight.dtype == dt:
            fold_w = ops.cached(self, "w_fold", [tm.fc1.weight, tm.fc1.bias, tm.fc_v.weight, tm.fc_v.bias, n1.weight, n1.bias],
                                lambda: ops.fold_gn_weights(tm.fc1.weight, tm.fc1.bias, tm.fc_v.weight, tm.fc_v.bias, n1.weight, n1.bias))
        x1 = o_fused = None
        if fold_w is not None and tm.fc2.weight.dtype == dt and ops.token_mixer_fused_ok(x, tm.heads, tm.head_dim, tm.fold_w, tm.fold_h, pw, ph):
            # stages 1 and 2: the whole token-mixer half in one persistent kernel (feat / value / core output stay on chip)
            sums = ops.new_sample_sums(B, dev, n=2)
            x1, _, _ = ops.token_mixer_fused_fwd(x, sums0, n1.eps, fold_w[0], fold_w[1], fold_w[2], f32(tm.sim_alpha), f32(tm.sim_beta),
                                                 tm.fc2.weight.detach().reshape(C, ED), f32(tm.fc2.bias), ls1, sums[0],
                                                 tm.heads, tm.head_dim, tm.fold_w, tm.fold_h)
        elif fold_w is not None and ops.token_mixer_core_ok(x, tm.heads, tm.head_dim, tm.fold_w, tm.fold_h, pw, ph):
            # stage 3: projection + cluster core in one launch (one CTA per region and group of four heads); fc2 follows below
            o_fused, _, _ = ops.token_mixer_core_fwd(x, sums0, n1.eps, fold_w[0], fold_w[1], fold_w[2], f32(tm.sim_alpha), f32(tm.sim_beta),
                                                     tm.heads, tm.head_dim, tm.fold_w, tm.fold_h)
        elif fold_w is not None:
            w_fold, k0, k1 = fold_w
            feat = torch.empty(B, ED, H, W, device=dev, dtype=torch.float32)
            value = torch.empty(B, ED, H, W, device=dev, dtype=dt)
            ops.conv_fwd(ops.conv_desc(x, w_fold, feat, gn_fold=(sums0, k1, n1.eps), e_shift=k0, out2=value))
        elif dt == torch.float32:
            y = torch.empty(B, 2 * ED, H, W, device=dev, dtype=dt)
            ops.conv_fwd(ops.conv_desc(x, w_in, y, gn=(sums0, f32(n1.weight), f32(n1.bias), n1.eps), e_shift=b_in))
            feat, value = y[:, :ED], y[:, ED:]
        else:
            feat = torch.empty(B, ED, H, W, device=dev, dtype=torch.float32)
            value = torch.empty(B, ED, H, W, device=dev, dtype=dt)
            ops.conv_fwd(ops.conv_desc(x, w_in, feat, gn=(sums0, f32(n1.weight), f32(n1.bias), n1.eps), e_shift=b_in, out2=value))
        if x1 is None:
            if o_fused is not None:
                o = o_fused
            else:
                o, _, _ = ops.cluster_core_fwd(feat, value, f32(tm.sim_alpha), f32(tm.sim_beta), tm.heads, tm.fold_w, tm.fold_h, pw, ph,
                                               out_dtype=dt)
            sums = ops.new_sample_sums(B, dev, n=2)
            x1 = torch.empty_like(x)
            ops.conv_fwd(ops.conv_desc(o, tm.fc2.weight.detach().reshape(C, ED), x1, e_shift=f32(tm.fc2.bias), post_scale=ls1, res=x,
                                       out_sample_sums=sums[0]))
        hid = mlp.fc1.weight.shape[0]
        if ops.mlp_fused_ok(x1, hid) and mlp.fc1.weight.dtype == dt and mlp.fc2.weight.dtype == dt:
            # both large stages: the hidden activation (8 x C channels) stays on chip
            x2 = ops.mlp_fused_fwd(x1, sums[0], f32(n2.weight), f32(n2.bias), n2.eps, mlp.fc1.weight.detach().reshape(hid, C),
                                   f32(mlp.fc1.bias), mlp.fc2.weight.detach().reshape(C, hid), f32(mlp.fc2.bias), ls2, sums[1])
            return ops.attach_sums(x2, ops.tag_like(sums[1], sums))
        h = torch.empty(B, hid, H, W, device=dev, dtype=dt)
        if SPLIT_WIDE_GN and dt == torch.bfloat16 and C > 576 and C % 64 == 0 and mlp.fc1.weight.dtype == dt:
            # more than nine k-slabs (neck N4, C = 640): the in-place-prologue GEMM cannot keep X resident and the transform-on-load
            # kernel redoes the normalisation in every one of the 20 output-tile CTAs of a point tile (91 us).  One streaming pass
            # writes GN(x) (the same bf16 operand), then the projection runs on the TMA-only kernel.
            xn = torch.empty_like(x1)
            ops.check(ops.lib.vrcoc_table_apply(ops.conv_desc(x1, x1, xn, gn=(sums[0], f32(n2.weight), f32(n2.bias), n2.eps)), ops._stream()),
                      "table_apply")
            ops.conv_fwd(ops.conv_desc(xn, mlp.fc1.weight.detach().reshape(hid, C), h, e_shift=f32(mlp.fc1.bias), act=ACT_GELU))
        else:
            ops.conv_fwd(ops.conv_desc(x1, mlp.fc1.weight.detach().reshape(hid, C), h, gn=(sums[0], f32(n2.weight), f32(n2.bias), n2.eps),
                                       e_shift=f32(mlp.fc1.bias), act=ACT_GELU))
        x2 = torch.empty_like(x)
        ops.conv_fwd(ops.conv_desc(h, mlp.fc2.weight.detach().reshape(C, hid), x2, e_shift=f32(mlp.fc2.bias), post_scale=ls2, res=x1,
                                   out_sample_sums=sums[1]))
        return ops.attach_sums(x2, ops.tag_like(sums[1], sums))

    def forward(self, x):
        if not x.is_cuda:
            raise VrcocError("vrcoc ClusterBlock needs a CUDA tensor (no CPU fallback exists)")
        if not self._fused_ok():
            return self._forward_composed(x)
        if not torch.is_grad_enabled():
            return self._forward_infer(x)
        ls1 = self.layer_scale_1 if self.use_layer_scale else None
        ls2 = self.layer_scale_2 if self.use_layer_scale else None
        tm, mlp = self.token_mixer, self.mlp
        # token-mixer half
        feat, value = tm._project_in(x, (ops.sample_sums_of(x), self.norm1))
        o = tm._core(feat, value)
        x1, sums1 = ops.ProjResidualFn.apply(o, tm.fc2.weight, tm.fc2.bias, ls1, x)
        # channel-MLP half
        h = ops.GNProjFn.apply(x1, sums1, self.norm2.weight, self.norm2.bias, self.norm2.eps,
                               mlp.fc1.weight, mlp.fc1.bias, ACT_GELU, 0)
        x2, sums2 = ops.ProjResidualFn.apply(h, mlp.fc2.weight, mlp.fc2.bias, ls2, x1)
        return ops.attach_sums(x2, sums2)        # rides along to the next block's norm1 (same python object through nn.Sequential)

    def _forward_composed(self, x):
        """Non-default configurations (BatchNorm norm_layer, active DropPath/Dropout, other activations): same
        kernels, composed module by module as the reference does (vr_coc.py:264-275)."""
        if self.use_layer_scale:
            x = x + self.drop_path(self.layer_scale_1.unsqueeze(-1).unsqueeze(-1) * self.token_mixer(self.norm1(x)))
            x = x + self.drop_path(self.layer_scale_2.unsqueeze(-1).unsqueeze(-1) * self.mlp(self.norm2(x)))
        else:
            x = x + self.drop_path(self.token_mixer(self.norm1(x)))
            x = x + self.drop_path(self.mlp(self.norm2(x)))
        return x


def basic_blocks(dim, index, layers, mlp_ratio=4., act_layer=nn.GELU, norm_layer=GroupNorm, drop_rate=.0,
                 drop_path_rate=0., use_layer_scale=True, layer_scale_init_value=1e-5,
                 proposal_w=2, proposal_h=2, fold_w=2, fold_h=2, heads=4, head_dim=24, return_center=False):
    """reference vr_coc.py:278-300."""
    blocks = []
    for block_idx in range(layers[index]):
        block_dpr = drop_path_rate * (block_idx + sum(layers[:index])) / (sum(layers) - 1)
        blocks.append(ClusterBlock(
            dim, mlp_ratio=mlp_ratio, act_layer=act_layer, norm_layer=norm_layer, drop=drop_rate, drop_path=block_dpr,
            use_layer_scale=use_layer_scale, layer_scale_init_value=layer_scale_init_value,
            proposal_w=proposal_w, proposal_h=proposal_h, fold_w=fold_w, fold_h=fold_h,
            heads=heads, head_dim=head_dim, return_center=False))
    return nn.Sequential(*blocks)
