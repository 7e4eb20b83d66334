"""Mask R-CNN forward around the B200 PanoSwin backbone (BASELINE.json configs[2]).

The hot path of this repository is the backbone; the detector's neck and heads are its CALLERS
(mmdet/models/detectors/two_stage.py:91-96 `extract_feat = neck(backbone(img))`).  They are restated here in
plain torch + torchvision ops -- the reference's own heads need mmcv-full CUDA ops (`roi_align`, `batched_nms`), which
do not exist offline -- so that a config-3 forward (backbone -> FPN -> RPN -> RoI heads) runs end to end on the GPU:

  FPN        mmdet/models/necks/fpn.py:66-221 as configured by configs/_base_/models/mask_rcnn_swin_fpn.py:21-25
             (in_channels [E,2E,4E,8E], out 256, 5 levels, no norm / activation, nearest top-down, stride-2 extra level)
  RPNHead    mask_rcnn_swin_fpn.py:26-42: 3x3 conv + ReLU, 1x1 objectness (3 anchors) and deltas; anchors scale 8,
             ratios 0.5 / 1 / 2, strides 4..64; test-time proposals per :110-115 (top 1000 per level, NMS 0.7, 1000 kept)
  RoI heads  :43-75 StandardRoIHead: RoIAlign 7x7 -> Shared2FCBBoxHead (1024) -> 81 scores + 80x4 deltas
             (stds .1 .1 .2 .2); RoIAlign 14x14 -> FCNMaskHead (4 convs, deconv, 1x1 -> 80 masks of 28x28);
             test-time: score_thr 0.05, class-wise NMS 0.5, 100 detections per image (:116-121)

Parameter names follow mmdet (`neck.lateral_convs.{i}.conv`, `rpn_head.rpn_conv`, `roi_head.bbox_head.fc_cls`, ...), so a
reference checkpoint's state_dict loads with strict=True.  The FPN is parity-tested against the unmodified reference
neck (tests/golden/fpn_*.npz); the heads have no executable reference here and are covered structurally.
Inference only: the detection losses (assigners, samplers, targets: mmdet/core) are outside the hot path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .backbone import SimplePanoSwinTransformer

__all__ = ["FPN", "RPNHead", "StandardRoIHead", "PanoSwinMaskRCNN"]


class _Conv(nn.Module):
    """mmcv ConvModule without norm / activation: the Conv2d lives under `.conv` (state_dict names)."""

    def __init__(self, cin, cout, k, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding)

    def forward(self, x):
        return self.conv(x)


class FPN(nn.Module):
    """fpn.py:170-221 for the shipped configuration (start_level 0, add_extra_convs False)."""

    def __init__(self, in_channels: Sequence[int], out_channels: int = 256, num_outs: int = 5):
        super().__init__()
        self.in_channels, self.out_channels, self.num_outs = list(in_channels), out_channels, num_outs
        self.lateral_convs = nn.ModuleList(_Conv(c, out_channels, 1) for c in in_channels)
        self.fpn_convs = nn.ModuleList(_Conv(out_channels, out_channels, 3, padding=1) for _ in in_channels)
        for m in self.modules():                              # fpn.py:163-168
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, feats: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
        assert len(feats) == len(self.in_channels)
        lat = [conv(f) for conv, f in zip(self.lateral_convs, feats)]
        for i in range(len(lat) - 1, 0, -1):                  # top-down: additions accumulate down the pyramid
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
        outs = [conv(x) for conv, x in zip(self.fpn_convs, lat)]
        while len(outs) < self.num_outs:                      # max_pool2d(kernel 1, stride 2) = plain sub-sampling
            outs.append(F.max_pool2d(outs[-1], 1, stride=2))
        return tuple(outs)


def _delta2bbox(rois, deltas, stds, max_shape, wh_ratio_clip=16 / 1000):
    """DeltaXYWHBBoxCoder.decode (mmdet/core/bbox/coder/delta_xywh_bbox_coder.py), means 0."""
    d = deltas * deltas.new_tensor(stds)
    max_ratio = abs(math.log(wh_ratio_clip))
    dw, dh = d[..., 2].clamp(-max_ratio, max_ratio), d[..., 3].clamp(-max_ratio, max_ratio)
    px, py = (rois[..., 0] + rois[..., 2]) * 0.5, (rois[..., 1] + rois[..., 3]) * 0.5
    pw, ph = rois[..., 2] - rois[..., 0], rois[..., 3] - rois[..., 1]
    gw, gh = pw * dw.exp(), ph * dh.exp()
    gx, gy = px + pw * d[..., 0], py + ph * d[..., 1]
    x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
    H, W = max_shape
    return torch.stack([x1.clamp(0, W), y1.clamp(0, H), x2.clamp(0, W), y2.clamp(0, H)], -1)


class RPNHead(nn.Module):
    def __init__(self, in_channels=256, feat_channels=256, scales=(8,), ratios=(0.5, 1.0, 2.0), strides=(4, 8, 16, 32, 64),
                 nms_pre=1000, max_per_img=1000, nms_thr=0.7):
        super().__init__()
        self.strides, self.scales, self.ratios = strides, scales, ratios
        self.nms_pre, self.max_per_img, self.nms_thr = nms_pre, max_per_img, nms_thr
        self.num_anchors = len(scales) * len(ratios)
        self.rpn_conv = nn.Conv2d(in_channels, feat_channels, 3, padding=1)
        self.rpn_cls = nn.Conv2d(feat_channels, self.num_anchors, 1)
        self.rpn_reg = nn.Conv2d(feat_channels, self.num_anchors * 4, 1)
        for m in (self.rpn_conv, self.rpn_cls, self.rpn_reg):
            nn.init.normal_(m.weight, std=0.01)
            nn.init.zeros_(m.bias)

    def _anchors(self, stride, h, w, device):
        """AnchorGenerator (mmdet/core/anchor/anchor_generator.py): base size = stride, centre offset 0."""
        r = torch.tensor(self.ratios, device=device).sqrt()
        s = torch.tensor(self.scales, device=device, dtype=torch.float32)
        ws = (stride * (1 / r)[:, None] * s[None, :]).reshape(-1)
        hs = (stride * r[:, None] * s[None, :]).reshape(-1)
        base = torch.stack([-0.5 * ws, -0.5 * hs, 0.5 * ws, 0.5 * hs], -1)                # [A, 4]
        sx = torch.arange(w, device=device, dtype=torch.float32) * stride
        sy = torch.arange(h, device=device, dtype=torch.float32) * stride
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack([xx, yy, xx, yy], -1).reshape(-1, 1, 4)
        return (shifts + base[None]).reshape(-1, 4)                                      # [(h, w, A), 4]

    @torch.no_grad()
    def forward(self, feats: Sequence[torch.Tensor], img_shape: Tuple[int, int]) -> List[torch.Tensor]:
        """-> per image [n <= max_per_img, 5] proposals (x1, y1, x2, y2, score)."""
        from torchvision.ops import batched_nms
        B = feats[0].shape[0]
        per_lvl = []
        for lvl, (f, stride) in enumerate(zip(feats, self.strides)):
            t = F.relu(self.rpn_conv(f))
            score = self.rpn_cls(t).permute(0, 2, 3, 1).reshape(B, -1).float().sigmoid()
            delta = self.rpn_reg(t).permute(0, 2, 3, 1).reshape(B, -1, 4).float()
            anchors = self._anchors(stride, f.shape[2], f.shape[3], f.device)
            k = min(self.nms_pre, score.shape[1])
            top, idx = score.topk(k, dim=1)
            boxes = _delta2bbox(anchors[idx], torch.gather(delta, 1, idx[..., None].expand(-1, -1, 4)), (1.0, 1.0, 1.0, 1.0), img_shape)
            per_lvl.append((boxes, top, torch.full_like(idx, lvl)))
        out = []
        for b in range(B):
            boxes = torch.cat([p[0][b] for p in per_lvl])
            scores = torch.cat([p[1][b] for p in per_lvl])
            lvls = torch.cat([p[2][b] for p in per_lvl])
            keep = batched_nms(boxes, scores, lvls, self.nms_thr)[: self.max_per_img]
            out.append(torch.cat([boxes[keep], scores[keep, None]], -1))
        return out


class _BBoxHead(nn.Module):
    """Shared2FCBBoxHead (mmdet/models/roi_heads/bbox_heads/convfc_bbox_head.py)."""

    def __init__(self, in_channels=256, fc_out=1024, roi_feat=7, num_classes=80):
        super().__init__()
        self.num_classes = num_classes
        self.shared_fcs = nn.ModuleList([nn.Linear(in_channels * roi_feat * roi_feat, fc_out), nn.Linear(fc_out, fc_out)])
        self.fc_cls = nn.Linear(fc_out, num_classes + 1)
        self.fc_reg = nn.Linear(fc_out, 4 * num_classes)
        nn.init.normal_(self.fc_cls.weight, std=0.01)
        nn.init.normal_(self.fc_reg.weight, std=0.001)
        for m in (self.fc_cls, self.fc_reg):
            nn.init.zeros_(m.bias)

    def forward(self, x):
        x = x.flatten(1)
        for fc in self.shared_fcs:
            x = F.relu(fc(x))
        return self.fc_cls(x), self.fc_reg(x)


class _MaskHead(nn.Module):
    """FCNMaskHead (mmdet/models/roi_heads/mask_heads/fcn_mask_head.py): 4 x (conv3x3 + ReLU), deconv 2x2 / 2 + ReLU, 1x1."""

    def __init__(self, in_channels=256, conv_out=256, num_convs=4, num_classes=80):
        super().__init__()
        self.convs = nn.ModuleList(_Conv(in_channels if i == 0 else conv_out, conv_out, 3, padding=1) for i in range(num_convs))
        self.upsample = nn.ConvTranspose2d(conv_out, conv_out, 2, stride=2)
        self.conv_logits = nn.Conv2d(conv_out, num_classes, 1)

    def forward(self, x):
        for c in self.convs:
            x = F.relu(c(x))
        return self.conv_logits(F.relu(self.upsample(x)))


class StandardRoIHead(nn.Module):
    def __init__(self, num_classes=80, strides=(4, 8, 16, 32), score_thr=0.05, nms_thr=0.5, max_per_img=100, finest_scale=56):
        super().__init__()
        self.strides, self.score_thr, self.nms_thr, self.max_per_img, self.finest_scale = strides, score_thr, nms_thr, max_per_img, finest_scale
        self.bbox_head = _BBoxHead(num_classes=num_classes)
        self.mask_head = _MaskHead(num_classes=num_classes)

    def _extract(self, feats, rois, out_size):
        """SingleRoIExtractor (mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py): level by box scale."""
        from torchvision.ops import roi_align
        scale = ((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2])).clamp_min(0).sqrt()
        lvls = torch.floor(torch.log2(scale / self.finest_scale + 1e-6)).clamp(0, len(self.strides) - 1).long()
        out = feats[0].new_zeros((rois.shape[0], feats[0].shape[1], out_size, out_size))
        for i, stride in enumerate(self.strides):
            sel = (lvls == i).nonzero(as_tuple=True)[0]
            if sel.numel():
                out[sel] = roi_align(feats[i].float(), rois[sel].float(), out_size, 1.0 / stride, sampling_ratio=0, aligned=True).to(out.dtype)
        return out

    @torch.no_grad()
    def forward(self, feats, proposals: List[torch.Tensor], img_shape) -> List[Dict[str, torch.Tensor]]:
        from torchvision.ops import batched_nms
        rois = torch.cat([torch.cat([p.new_full((p.shape[0], 1), b), p[:, :4]], 1) for b, p in enumerate(proposals)])
        cls, reg = self.bbox_head(self._extract(feats, rois, 7))
        scores = cls.float().softmax(-1)
        n_cls = self.bbox_head.num_classes
        boxes = _delta2bbox(rois[:, None, 1:].expand(-1, n_cls, -1), reg.float().view(-1, n_cls, 4), (0.1, 0.1, 0.2, 0.2), img_shape)
        results, det_rois = [], []
        for b in range(len(proposals)):
            m = rois[:, 0] == b
            sc, bx = scores[m][:, :n_cls], boxes[m]              # multiclass_nms: background is the last column
            keep = sc > self.score_thr
            idx = keep.nonzero()
            bsel, ssel, lsel = bx[idx[:, 0], idx[:, 1]], sc[keep], idx[:, 1]
            k = batched_nms(bsel, ssel, lsel, self.nms_thr)[: self.max_per_img]
            results.append({"boxes": bsel[k], "scores": ssel[k], "labels": lsel[k]})
            det_rois.append(torch.cat([bsel.new_full((k.numel(), 1), b), bsel[k]], 1))
        det = torch.cat(det_rois)
        if det.shape[0]:
            logits = self.mask_head(self._extract(feats, det, 14))
            labels = torch.cat([r["labels"] for r in results])
            masks = logits[torch.arange(det.shape[0], device=det.device), labels].float().sigmoid()      # [n, 28, 28]
        else:
            masks = det.new_zeros((0, 28, 28))
        o = 0
        for r in results:
            n = r["boxes"].shape[0]
            r["masks"] = masks[o:o + n]
            o += n
        return results


class PanoSwinMaskRCNN(nn.Module):
    """MaskRCNN(TwoStageDetector) inference forward (mmdet/models/detectors/two_stage.py:91-96, :169-186) with the B200
    PanoSwin backbone.  `heads_dtype`: the neck / heads run in this dtype, channels-last (cuDNN / cuBLAS library calls:
    they are callers of the hot path, not part of it)."""

    def __init__(self, backbone: dict = None, num_classes: int = 80, heads_dtype=torch.bfloat16):
        super().__init__()
        cfg = dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7, ape=True, pano_mode=True,
                   drop_path_rate=0.2)
        cfg.update(backbone or {})
        self.backbone = SimplePanoSwinTransformer(**cfg)
        self.neck = FPN(self.backbone.num_features, 256, 5)
        self.rpn_head = RPNHead()
        self.roi_head = StandardRoIHead(num_classes=num_classes)
        self.heads_dtype = heads_dtype

    def _autocast(self):
        return torch.autocast("cuda", dtype=self.heads_dtype, enabled=self.heads_dtype != torch.float32)

    def extract_feat(self, img):
        feats = self.backbone(img)
        with self._autocast():
            return self.neck([f.to(self.heads_dtype).contiguous(memory_format=torch.channels_last) for f in feats])

    @torch.no_grad()
    def forward(self, img) -> List[Dict[str, torch.Tensor]]:
        """img [B, 3, H, W] -> per image {boxes [n,4], scores [n], labels [n], masks [n,28,28]} (n <= 100)."""
        shape = (img.shape[2], img.shape[3])
        feats = self.extract_feat(img)
        with self._autocast():
            proposals = self.rpn_head(feats, shape)
            return self.roi_head(feats[:4], proposals, shape)
