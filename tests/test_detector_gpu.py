"""Config-3 forward (Mask R-CNN around the PanoSwin backbone) on a B200: the product-side FPN against
the unmodified reference neck's golden features, and the RPN / RoI-head forward structurally (no executable reference
exists offline for the heads: they need mmcv-full ops)."""
import pytest
import torch

from _expect import rel_l2
from conftest import load_golden
from oracle import fpn_oracle as FO
from oracle import panoswin_oracle as O

DEV = "cuda:0"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_pano", "panoswin_t_512"])
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_product_fpn_on_our_backbone_matches_reference_fpn_features(name, mode, tol):
    """North star: parity "on block outputs and FPN features": panoswintransformerobjectdetection_b200.detector.FPN fed by
    our backbone vs the reference backbone -> reference FPN (tests/golden/fpn_*.npz)."""
    import panoswintransformerobjectdetection_b200 as P
    meta, _ = load_golden(name)
    fmeta, fz = load_golden("fpn_" + name)
    cfg = meta["cfg"]
    m = P.SimplePanoSwinTransformer(embed_dim=cfg["embed_dim"], depths=list(cfg["depths"]), num_heads=list(cfg["num_heads"]),
                                    ape=True, out_indices=tuple(cfg["out_indices"]))
    m.load_state_dict(O.make_state_dict(cfg, meta["param_seed"]), strict=True)
    m.to(DEV)
    m.eval()
    m.set_compute_dtype(mode)
    neck = P.FPN(fmeta["in_channels"], fmeta["out_channels"], fmeta["num_outs"])
    neck.load_state_dict(FO.make_fpn_state(fmeta["in_channels"], fmeta["out_channels"], fmeta["fpn_seed"]), strict=True)
    neck.to(DEV)
    img = O.make_image(meta["shape"], meta["image_seed"], meta["kind"]).to(DEV)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            levels = neck(m(img))                                # the neck itself in fp32: the tolerance is the backbone's
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    if name == "tiny_pano" and mode == "fp32":
        tol = 2e-4                                               # token maps lower than one window (fp32_tol in test_backbone_gpu)
    assert len(levels) == 5
    for i, o in enumerate(levels):
        assert list(o.shape) == list(fz[f"out{i}_shape"])
        got = o if fmeta["full"] else o.reshape(-1)[::fmeta["stride"]]
        assert rel_l2(got, torch.from_numpy(fz[f"out{i}"])) <= tol, (name, mode, i)


@pytest.mark.gpu
@pytest.mark.parametrize("heads_dtype", [torch.float32, torch.bfloat16])
def test_mask_rcnn_forward_structure(heads_dtype):
    import panoswintransformerobjectdetection_b200 as P
    torch.manual_seed(0)
    det = P.PanoSwinMaskRCNN(backbone=dict(embed_dim=32, depths=[2, 2, 2, 2], num_heads=[1, 2, 4, 8]), num_classes=80,
                             heads_dtype=heads_dtype)
    det.roi_head.bbox_head.fc_cls.weight.data.normal_(0, 0.5)    # random scores, so that detections exist
    det.to(DEV)
    det.eval()
    img = torch.rand(2, 3, 128, 256, device=DEV)
    feats = det.extract_feat(img)
    assert [tuple(f.shape) for f in feats] == [(2, 256, 32, 64), (2, 256, 16, 32), (2, 256, 8, 16), (2, 256, 4, 8), (2, 256, 2, 4)]
    with det._autocast():
        props = det.rpn_head(feats, (128, 256))
    assert len(props) == 2 and all(p.shape[1] == 5 and 0 < p.shape[0] <= 1000 for p in props)
    for p in props:
        assert (p[:, 0] >= 0).all() and (p[:, 2] <= 256).all() and (p[:, 1] >= 0).all() and (p[:, 3] <= 128).all()
        assert (p[:-1, 4] >= p[1:, 4] - 1e-6).all()            # NMS keeps score order
    out = det(img)
    assert len(out) == 2
    for r in out:
        n = r["boxes"].shape[0]
        assert 0 < n <= 100 and r["scores"].shape == (n,) and r["labels"].shape == (n,) and r["masks"].shape == (n, 28, 28)
        assert (r["scores"] > 0.05).all() and (r["labels"] >= 0).all() and (r["labels"] < 80).all()
        assert ((r["masks"] >= 0) & (r["masks"] <= 1)).all() and torch.isfinite(r["boxes"]).all()
    # deterministic, and images do not interact: image 1 alone gives the same detections
    again = det(img[1:])
    assert again[0]["boxes"].shape == out[1]["boxes"].shape
    if heads_dtype == torch.float32:
        assert torch.allclose(again[0]["boxes"], out[1]["boxes"], atol=1e-2)


def test_mask_rcnn_state_dict_names_follow_mmdet():
    import panoswintransformerobjectdetection_b200 as P
    det = P.PanoSwinMaskRCNN(backbone=dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], out_indices=(0, 1)))
    keys = set(det.state_dict().keys())
    for k in ("neck.lateral_convs.0.conv.weight", "neck.fpn_convs.1.conv.bias", "rpn_head.rpn_conv.weight", "rpn_head.rpn_cls.bias",
              "rpn_head.rpn_reg.weight", "roi_head.bbox_head.shared_fcs.0.weight", "roi_head.bbox_head.fc_cls.weight",
              "roi_head.bbox_head.fc_reg.bias", "roi_head.mask_head.convs.3.conv.weight", "roi_head.mask_head.upsample.weight",
              "roi_head.mask_head.conv_logits.bias", "backbone.layers.0.blocks.0.attn.qkv.weight"):
        assert k in keys, k
