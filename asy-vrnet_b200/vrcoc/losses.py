"""Training-step losses without host round trips ("training-step host overheads").

`YOLOLoss` is the reference's detection loss (nets/yolo_training.py:60-436: YOLOX decode, SimOTA dynamic-k label assignment,
IoU / objectness / class terms) with the SAME constructor, forward signature and arithmetic, re-expressed over the whole batch:
the reference walks the images one by one, compacts the candidate anchors with boolean masks, calls `.item()` per ground-truth box
and `torch.cuda.empty_cache()` per image (30 of the 104 ms of a training step on a B200, profiles/r02_prof_train.txt — it is
launch- and sync-bound, ~100 small kernels and ~10 host syncs per image).  Here every tensor has a static shape
([B, G, A]: images x padded ground-truth boxes x anchors), masks replace compaction, and there is no `.item()`, `nonzero` or
boolean indexing anywhere: ~120 launches per step whatever the batch, no host sync, CUDA-graph capturable.

Host-side PyTorch (this file has no kernels); device-agnostic, so `tests/test_losses.py` pins it on the CPU against the
reference's own class (loss value and gradients) whenever the reference tree is present.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _iou_cxcywh(a, b):
    """pairwise-broadcast IoU of boxes in (cx, cy, w, h); reference bboxes_iou(xyxy=False), yolo_training.py:266-289"""
    tl = torch.max(a[..., :2] - a[..., 2:] / 2, b[..., :2] - b[..., 2:] / 2)
    br = torch.min(a[..., :2] + a[..., 2:] / 2, b[..., :2] + b[..., 2:] / 2)
    area_a = a[..., 2] * a[..., 3]
    area_b = b[..., 2] * b[..., 3]
    en = (tl < br).to(tl.dtype).prod(dim=-1)
    area_i = (br - tl).prod(dim=-1) * en
    return area_i / (area_a + area_b - area_i)


class YOLOLoss(nn.Module):
    """reference nets/yolo_training.py:60 (same arguments, same `log_vars` parameter, same result)."""

    def __init__(self, num_classes, fp16=False, strides=(8, 16, 32)):
        super().__init__()
        self.num_classes = num_classes
        self.strides = list(strides)
        self.fp16 = fp16                      # kept for the signature: the assignment cost below is always evaluated in float32
        self.log_vars = nn.Parameter(torch.zeros(3))
        self._grids = {}

    # -- decode (reference :77-113), out of place ---------------------------------------------------------------------
    def _grid(self, k, h, w, device):
        key = (k, h, w, str(device))
        g = self._grids.get(key)
        if g is None:
            yv, xv = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
            g = torch.stack((xv, yv), 2).reshape(1, -1, 2).float()
            self._grids[key] = g
        return g

    def forward(self, inputs, labels=None):
        outs, xs, ys, ss = [], [], [], []
        for k, (stride, o) in enumerate(zip(self.strides, inputs)):
            h, w = o.shape[-2:]
            grid = self._grid(k, h, w, o.device).to(o.dtype)
            o = o.flatten(start_dim=2).permute(0, 2, 1)
            o = torch.cat([(o[..., :2] + grid) * stride, torch.exp(o[..., 2:4]) * stride, o[..., 4:]], dim=-1)
            outs.append(o)
            xs.append(grid[0, :, 0])
            ys.append(grid[0, :, 1])
            ss.append(torch.full_like(grid[0, :, 0], float(stride)))
        return self.get_losses(torch.cat(xs), torch.cat(ys), torch.cat(ss), labels, torch.cat(outs, 1))

    @staticmethod
    def _pad_labels(labels, device, dtype):
        """list of [n_i, 5] -> boxes [B, G, 4], classes [B, G] (int64), valid [B, G]; G = max n_i (>= 1).  Host-side shapes only."""
        B = len(labels)
        G = max(1, max((int(l.shape[0]) for l in labels), default=1))
        if all(int(l.shape[0]) == G for l in labels):
            t = torch.stack([l.to(device=device, dtype=dtype) for l in labels])
            return t[..., :4], t[..., 4].to(torch.int64), torch.ones(B, G, dtype=torch.bool, device=device)
        t = torch.zeros(B, G, 5, device=device, dtype=dtype)
        valid = torch.zeros(B, G, dtype=torch.bool, device=device)
        for i, l in enumerate(labels):
            n = int(l.shape[0])
            if n:
                t[i, :n] = l.to(device=device, dtype=dtype)
                valid[i, :n] = True
        return t[..., :4], t[..., 4].to(torch.int64), valid

    # -- SimOTA over the whole batch (reference get_assignments / get_in_boxes_info / dynamic_k_matching, :193-436) -------------
    @torch.no_grad()
    def assign(self, gt, gcls, valid, bbox, cls_logit, obj_logit, xs, ys, ss, center_radius=2.5):
        """gt [B,G,4] cxcywh, gcls [B,G], valid [B,G]; bbox [B,A,4], cls_logit [B,A,nc], obj_logit [B,A,1]; xs, ys, ss [A].
        Returns fg [B,A] bool, matched ground-truth index [B,A] int64, IoU of the matched pair [B,A]."""
        xc = ((xs + 0.5) * ss)[None, None, :]                            # [1,1,A]
        yc = ((ys + 0.5) * ss)[None, None, :]
        gx, gy, gw, gh = (gt[..., i, None] for i in range(4))            # [B,G,1]
        v = valid[..., None]
        in_box = (torch.stack([xc - (gx - 0.5 * gw), yc - (gy - 0.5 * gh), (gx + 0.5 * gw) - xc, (gy + 0.5 * gh) - yc], -1).min(-1).values > 0.0) & v
        r = center_radius * ss[None, None, :]
        in_ctr = (torch.stack([xc - (gx - r), yc - (gy - r), (gx + r) - xc, (gy + r) - yc], -1).min(-1).values > 0.0) & v
        cand = in_box.any(1) | in_ctr.any(1)                             # [B,A]  the reference's first fg_mask
        in_both = in_box & in_ctr
        candg = cand[:, None, :] & v                                     # [B,G,A]

        ious = _iou_cxcywh(gt[:, :, None, :], bbox[:, None, :, :].float())                       # [B,G,A]
        iou_cost = -torch.log(ious + 1e-8)
        with torch.autocast(device_type=bbox.device.type, enabled=False):
            p = (cls_logit.float().sigmoid() * obj_logit.float().sigmoid()).sqrt()               # [B,A,nc]
            # binary_cross_entropy(p, onehot).sum(-1) with its log clamp at -100: the one-hot picks log p at the class, log(1-p) elsewhere
            lp = torch.log(p).clamp_min(-100.0)
            l1p = torch.log(1.0 - p).clamp_min(-100.0)
            tot = l1p.sum(-1)                                                                     # [B,A]
            idx = gcls.clamp(0, self.num_classes - 1)[..., None].expand(-1, -1, p.shape[1])       # [B,G,A]
            lp_g = torch.gather(lp.transpose(1, 2), 1, idx)                                       # [B,G,A] log p[a, class_g]
            l1p_g = torch.gather(l1p.transpose(1, 2), 1, idx)
            cls_cost = -(lp_g + (tot[:, None, :] - l1p_g))
        cost = cls_cost + 3.0 * iou_cost + 100000.0 * (~in_both).to(cls_cost.dtype)
        inf = torch.full_like(cost, float("inf"))
        cost_c = torch.where(candg, cost, inf)                           # only candidate anchors of real boxes compete

        A = cost.shape[-1]
        kc = min(10, A)
        topk_ious = torch.topk(torch.where(candg, ious, torch.zeros_like(ious)), kc, dim=-1).values
        dyn_k = topk_ious.sum(-1).int().clamp(min=1)                     # [B,G]
        vals, pos = torch.topk(cost_c, kc, dim=-1, largest=False)        # [B,G,kc] ascending
        take = (torch.arange(kc, device=cost.device)[None, None, :] < dyn_k[..., None]) & torch.isfinite(vals)
        matching = torch.zeros_like(cost, dtype=torch.bool)
        matching.scatter_(-1, pos, take)
        multi = matching.sum(1) > 1                                      # [B,A] anchors claimed by several boxes -> the cheapest one
        best = torch.argmin(cost_c, dim=1)                               # [B,A]
        only = F.one_hot(best, cost.shape[1]).to(torch.bool).transpose(1, 2)                      # [B,G,A]
        matching = torch.where(multi[:, None, :], only & candg, matching)
        fg = matching.any(1)
        matched = matching.to(torch.uint8).argmax(1)                     # [B,A]
        pred_iou = (matching.to(ious.dtype) * ious).sum(1)
        return fg, matched, pred_iou

    def get_losses(self, xs, ys, ss, labels, outputs):
        bbox, obj, cls = outputs[..., :4], outputs[..., 4:5], outputs[..., 5:]
        B, A = outputs.shape[:2]
        gt, gcls, valid = self._pad_labels(labels, outputs.device, outputs.dtype)
        fg, matched, pred_iou = self.assign(gt, gcls, valid, bbox.detach(), cls.detach(), obj.detach(), xs.to(outputs.dtype),
                                            ys.to(outputs.dtype), ss.to(outputs.dtype))
        fgf = fg.to(outputs.dtype)
        num_fg = fgf.sum().clamp(min=1.0)
        reg_t = torch.gather(gt, 1, matched[..., None].expand(-1, -1, 4))                         # [B,A,4]
        cls_t = F.one_hot(torch.gather(gcls, 1, matched).clamp(0, self.num_classes - 1), self.num_classes).to(outputs.dtype) * pred_iou[..., None]
        # IOUloss(reduction none, "iou") on the foreground anchors (reference :14-57); background rows get their own target as the
        # prediction so that nothing non-finite can leak through the mask
        pb = torch.where(fg[..., None], bbox, reg_t)
        tl = torch.max(pb[..., :2] - pb[..., 2:] / 2, reg_t[..., :2] - reg_t[..., 2:] / 2)
        br = torch.min(pb[..., :2] + pb[..., 2:] / 2, reg_t[..., :2] + reg_t[..., 2:] / 2)
        area_p, area_g = pb[..., 2] * pb[..., 3], reg_t[..., 2] * reg_t[..., 3]
        en = (tl < br).to(tl.dtype).prod(dim=-1)
        area_i = (br - tl).prod(dim=-1) * en
        iou = area_i / (area_p + area_g - area_i + 1e-16)
        loss_iou = ((1 - iou ** 2) * fgf).sum()
        loss_obj = F.binary_cross_entropy_with_logits(obj, fgf[..., None], reduction="none").sum()
        loss_cls = (F.binary_cross_entropy_with_logits(cls, cls_t, reduction="none") * fgf[..., None]).sum()
        return (loss_iou + 2 * loss_obj + 2 * loss_cls) / num_fg
