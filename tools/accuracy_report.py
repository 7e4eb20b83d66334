"""Accuracy of the CUDA path against the CPU oracle on PanoSwin-T @512x1024 (one image): rel-L2 of the four
stage features for fp32 mode, bf16 mode (fp32 residual stream) and bf16 mode with a bf16 residual stream."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import panoswin_oracle as O  # noqa: E402
import panoswintransformerobjectdetection_b200 as P  # noqa: E402

cfg = O.make_config()
sd = O.make_state_dict(cfg, 1)
for kind in ("rand", "randn"):
    img = O.make_image((1, 3, 512, 1024), 2, kind)
    want = O.backbone_forward(sd, cfg, img)
    m = P.SimplePanoSwinTransformer(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], ape=True)
    m.load_state_dict(sd)
    m.to("cuda:0")
    m.eval()
    for mode, res in (("fp32", "fp32"), ("bf16", "fp32"), ("bf16", "bf16")):
        m.set_compute_dtype(mode)
        m.set_residual_dtype(res)
        outs = m(img.to("cuda:0"))
        errs = [float((o.cpu().double() - w.double()).norm() / w.double().norm()) for o, w in zip(outs, want)]
        mx = [float((o.cpu() - w).abs().max()) for o, w in zip(outs, want)]
        print(f"{kind:5s} compute={mode} residual={res}: rel-L2 " + " ".join(f"{e:.2e}" for e in errs) +
              "  max-abs " + " ".join(f"{e:.2e}" for e in mx), flush=True)
