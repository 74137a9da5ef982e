"""Dual-branch context-cluster backbone with asymmetric fusion: reference backbone/fusion/vr_coc.py:362-808
(VRCoC, coc_tiny, coc_tiny2, coc_small, coc_medium), same constructor arguments, forward signature, module tree and
state-dict keys (incl. the persistent `fea_pos` / `fea_pos_r` buffers and the zero-size tensors under
`radar_enhance_by_image1.image_attn.*`).
"""
import copy

import torch
import torch.nn as nn

from .context_cluster import (Cluster, ClusterBlock, GroupNorm, Mlp, PointRecuder, basic_blocks,  # noqa: F401
                              pairwise_cos_sim)
from .fusion import (BaseConv, DWConv, ImageEnhanceByRadar, RadarEnhanceByImage, ShuffleAttention,  # noqa: F401
                     data_normal, eca_block, shuffle_channels)

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def _cfg(url='', **kwargs):
    return {'url': url, 'num_classes': 1000, 'input_size': (3, 224, 224), 'crop_pct': .95, 'interpolation': 'bicubic',
            'mean': IMAGENET_DEFAULT_MEAN, 'std': IMAGENET_DEFAULT_STD, 'classifier': 'head', **kwargs}


default_cfgs = {'model_small': _cfg(crop_pct=0.9), 'model_medium': _cfg(crop_pct=0.95)}


def register_model(fn):
    return fn


class VRCoC(nn.Module):
    """reference vr_coc.py:362-704."""

    def __init__(self, layers, embed_dims=None, mlp_ratios=None, downsamples=None, norm_layer=nn.BatchNorm2d,
                 act_layer=nn.GELU, num_classes=1000, in_patch_size=4, in_stride=4, in_pad=0,
                 down_patch_size=2, down_stride=2, down_pad=0, drop_rate=0., drop_path_rate=0.,
                 use_layer_scale=True, layer_scale_init_value=1e-5, fork_feat=True, init_cfg=None, pretrained=None,
                 img_w=512, img_h=512,
                 proposal_w=[2, 2, 2, 2], proposal_h=[2, 2, 2, 2], fold_w=[8, 4, 2, 1], fold_h=[8, 4, 2, 1],
                 heads=[2, 4, 6, 8], head_dim=[16, 16, 32, 32], **kwargs):
        super().__init__()
        if not fork_feat:
            self.num_classes = num_classes
        self.fork_feat = fork_feat

        # positional grids (reference :402-413): value i/(n-1) - 0.5, meshgrid 'ij'
        range_w = torch.arange(0, img_w, step=1) / (img_w - 1.0)
        range_h = torch.arange(0, img_h, step=1) / (img_h - 1.0)
        fea_pos = torch.stack(torch.meshgrid(range_w, range_h, indexing='ij'), dim=-1).float() - 0.5
        self.register_buffer('fea_pos', fea_pos)
        self.register_buffer('fea_pos_r', fea_pos.clone())

        self.image_initial = PointRecuder(patch_size=1, stride=1, padding=0, in_chans=3, embed_dim=3)
        self.radar_initial = PointRecuder(patch_size=1, stride=1, padding=0, in_chans=4, embed_dim=4)
        self.radar_enhance_by_image1 = RadarEnhanceByImage(image_in_channels=3, radar_in_channels=4, initial=True)
        self.image_enhance_by_radar1 = ImageEnhanceByRadar(image_in_channels=3, radar_in_channels=4)
        self.patch_embed = PointRecuder(patch_size=in_patch_size, stride=in_stride, padding=in_pad, in_chans=5,
                                        embed_dim=embed_dims[0])
        self.patch_embed_radar = PointRecuder(patch_size=in_patch_size, stride=in_stride, padding=in_pad, in_chans=6,
                                              embed_dim=embed_dims[0])

        network, network_radar = [], []
        for i in range(len(layers)):
            kw = dict(mlp_ratio=mlp_ratios[i], act_layer=act_layer, norm_layer=norm_layer, drop_rate=drop_rate,
                      drop_path_rate=drop_path_rate, use_layer_scale=use_layer_scale,
                      layer_scale_init_value=layer_scale_init_value, proposal_w=proposal_w[i], proposal_h=proposal_h[i],
                      fold_w=fold_w[i], fold_h=fold_h[i], heads=heads[i], head_dim=head_dim[i], return_center=False)
            network.append(basic_blocks(embed_dims[i], i, layers, **kw))
            network_radar.append(basic_blocks(embed_dims[i], i, layers, **kw))
            network.append(ImageEnhanceByRadar(image_in_channels=embed_dims[i], radar_in_channels=embed_dims[i]))
            network_radar.append(RadarEnhanceByImage(image_in_channels=embed_dims[i], radar_in_channels=embed_dims[i]))
            if i >= len(layers) - 1:
                break
            if downsamples[i] or embed_dims[i] != embed_dims[i + 1]:
                network.append(PointRecuder(patch_size=down_patch_size, stride=down_stride, padding=down_pad,
                                            in_chans=embed_dims[i], embed_dim=embed_dims[i + 1]))
                network_radar.append(PointRecuder(patch_size=down_patch_size, stride=down_stride, padding=down_pad,
                                                  in_chans=embed_dims[i], embed_dim=embed_dims[i + 1]))
        self.network = nn.ModuleList(network)
        self.network_radar = nn.ModuleList(network_radar)

        if self.fork_feat:
            self.out_indices = [0, 3, 6, 9]
        else:
            self.norm = norm_layer(embed_dims[-1])
            self.head = nn.Linear(embed_dims[-1], num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self.cls_init_weights)
        self.init_cfg = copy.deepcopy(init_cfg)
        if self.fork_feat and (self.init_cfg is not None or pretrained is not None):
            self.init_weights()

    def cls_init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def init_weights(self, pretrained=None):
        """reference :534-564 loads mmcv checkpoints; here: a plain torch checkpoint path in init_cfg['checkpoint']."""
        if self.init_cfg is None and pretrained is None:
            return
        path = self.init_cfg['checkpoint'] if self.init_cfg is not None else pretrained
        ckpt = torch.load(path, map_location='cpu')
        sd = ckpt.get('state_dict', ckpt.get('model', ckpt)) if isinstance(ckpt, dict) else ckpt
        self.load_state_dict(sd, False)

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def forward_embeddings(self, x, x_radar):
        """reference :575-587.  The cat([x, pos]) tensors are never built: the patch-embed kernel reads the position
        grid as a second, batch-broadcast source."""
        from . import ops
        pair = ops.run_pair if x.is_cuda else (lambda f, g: (f(), g()))
        x_in, r_in = x, x_radar
        x, x_radar = pair(lambda: self.image_initial(x_in), lambda: self.radar_initial(r_in))
        x = self.image_enhance_by_radar1(x, x_radar)
        x_radar = self.radar_enhance_by_image1(x, x_radar)
        pos = ops.cached(self, "pos_" + str(x.dtype), [self.fea_pos], lambda: self.fea_pos.permute(2, 0, 1).to(x.dtype).contiguous())
        if pos.shape[-2:] != x.shape[-2:]:
            raise RuntimeError(f"Sizes of tensors must match: fea_pos is {tuple(pos.shape[-2:])}, input is {tuple(x.shape[-2:])}")
        x_e, r_e = x, x_radar
        return pair(lambda: self.patch_embed(x_e, extra=pos), lambda: self.patch_embed_radar(r_e, extra=pos))

    def forward_tokens(self, x, x_radar):
        """reference :589-675."""
        outs, outs_radar = [], []
        nstage = (len(self.network) + 1) // 3
        from . import ops
        pair = ops.run_pair if x.is_cuda else (lambda f, g: (f(), g()))
        for i in range(nstage):
            # the two modality stacks are independent between fusion points: side by side on two streams
            x, x_radar = pair(lambda x=x: self.network[3 * i](x), lambda r=x_radar: self.network_radar[3 * i](r))
            x = self.network[3 * i + 1](x, x_radar)
            x_radar = self.network_radar[3 * i + 1](x, x_radar)
            if i in (0, nstage - 1):
                outs.append(x)
                outs_radar.append(x_radar)
            if i < nstage - 1:
                x, x_radar = pair(lambda x=x: self.network[3 * i + 2](x), lambda r=x_radar: self.network_radar[3 * i + 2](r))
                if i in (0, 1):
                    outs.append(x)
                    outs_radar.append(x_radar)
        return outs, outs_radar

    def forward(self, x, x_radar):
        from . import ops
        with ops.sums_arena(x.shape[0], x.device):
            return self._forward(x, x_radar)

    def _forward(self, x, x_radar):
        x, x_radar = self.forward_embeddings(x, x_radar)
        x, x_radar = self.forward_tokens(x, x_radar)
        if self.fork_feat:
            return x, x_radar
        x = self.norm(x)
        return self.head(x.mean([-2, -1]))


def replace_pos_buffers(model, img_h, img_w=None):
    """Frames other than 512x512 (BASELINE.json configs[4]: 1024x1024): the reference bakes the positional grid into two fixed
    [512, 512, 2] buffers (vr_coc.py:402-413) and its forward then fails at the cat of :582-583 for any other size.  This is the
    buffer replacement of SURVEY 8c.5, applied to every module under `model` that owns a `fea_pos` buffer (the product's VRCoC or
    the reference's own): same formula, new size.  The kernels themselves are size-agnostic (regions become 32x32 / 64x64)."""
    img_w = img_h if img_w is None else img_w
    n = 0
    for m in model.modules():
        if isinstance(getattr(m, "fea_pos", None), torch.Tensor) and "fea_pos" in m._buffers:
            old = m._buffers["fea_pos"]
            rw = torch.arange(0, img_w, step=1) / (img_w - 1.0)
            rh = torch.arange(0, img_h, step=1) / (img_h - 1.0)
            pos = (torch.stack(torch.meshgrid(rw, rh, indexing='ij'), dim=-1).float() - 0.5).to(device=old.device, dtype=old.dtype)
            m._buffers["fea_pos"] = pos
            if "fea_pos_r" in m._buffers:
                m._buffers["fea_pos_r"] = pos.clone()
            m.__dict__.pop("_vrcoc_cache", None)          # derived views of the old grid
            n += 1
    if n == 0:
        raise ValueError("replace_pos_buffers: no module with a fea_pos buffer found")
    return model


def _build(layers, embed_dims, heads, head_dim, proposal, fold, cfg_name, **kwargs):
    model = VRCoC(layers, embed_dims=embed_dims, norm_layer=GroupNorm, mlp_ratios=[8, 8, 4, 4],
                  downsamples=[True, True, True, True], down_patch_size=3, down_pad=1,
                  proposal_w=proposal, proposal_h=proposal, fold_w=fold, fold_h=fold, heads=heads, head_dim=head_dim, **kwargs)
    model.default_cfg = default_cfgs[cfg_name]
    return model


@register_model
def coc_tiny(pretrained=False, **kwargs):
    """reference :707-730"""
    return _build([3, 4, 5, 2], [32, 64, 196, 320], [4, 4, 8, 8], [24, 24, 24, 24], [2, 2, 2, 2], [8, 4, 2, 1], 'model_small', **kwargs)


@register_model
def coc_tiny2(pretrained=False, **kwargs):
    """reference :733-756"""
    return _build([3, 4, 5, 2], [32, 64, 196, 320], [4, 4, 8, 8], [24, 24, 24, 24], [4, 2, 7, 4], [8, 8, 1, 1], 'model_small', **kwargs)


@register_model
def coc_small(pretrained=False, width=1.0, **kwargs):
    """reference :759-782"""
    dims = [int(64 * width), int(128 * width), int(320 * width), int(512 * width)]
    return _build([2, 2, 6, 2], dims, [4, 4, 8, 8], [32, 32, 32, 32], [2, 2, 2, 2], [8, 4, 2, 1], 'model_small', **kwargs)


@register_model
def coc_medium(pretrained=False, width=1.0, **kwargs):
    """reference :785-808"""
    dims = [int(64 * width), int(128 * width), int(320 * width), int(512 * width)]
    return _build([4, 4, 12, 4], dims, [6, 6, 12, 12], [32, 32, 32, 32], [2, 2, 2, 2], [8, 4, 2, 1], 'model_small', **kwargs)
