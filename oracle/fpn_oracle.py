"""CPU oracle for the FPN neck that consumes the backbone's four feature maps.

TEST INFRASTRUCTURE ONLY (same rules as oracle/panoswin_oracle.py).  The neck is a CALLER of the
drop-in boundary, not part of the product: it is restated here so that the parity tests can check
"FPN features" (BASELINE.json north_star) computed from our backbone's outputs against FPN features
computed from the reference backbone's outputs.

Parity status: PINNED.  `oracle/make_golden.py --fpn` executes the unmodified reference neck
(`/root/reference/mmdet/models/necks/fpn.py`, loaded by `oracle/ref_loader.load_reference_fpn`
with stand-ins for the absent mmcv) on the reference backbone's outputs and stores the five pyramid
levels in `tests/golden/fpn_*.npz`; `tests/test_oracle_golden.py` checks this file against them.

Restated configuration: the one the shipped detector configs use
(`configs/_base_/models/mask_rcnn_swin_fpn.py:21-25`, `faster_rcnn_panoswin_fpn.py`):
`FPN(in_channels=[E, 2E, 4E, 8E], out_channels=256, num_outs=5)` — start_level 0, no extra convs,
no norm, no activation, nearest up-sampling.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def make_fpn_state(in_channels: Sequence[int], out_channels: int = 256, seed: int = 3) -> Dict[str, Tensor]:
    """Deterministic parameters with the reference module's key names (`lateral_convs.{i}.conv.*`,
    `fpn_convs.{i}.conv.*`: ConvModule holds its Conv2d as `.conv`, fpn.py:120-141).  Xavier-uniform
    weights like `FPN.init_weights` (fpn.py:163-168) drawn from numpy's legacy RandomState; biases are
    randomised (the reference initialises them to 0) so that the bias adds are exercised."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, Tensor] = {}
    def conv(name, cin, cout, k):
        bound = math.sqrt(6.0 / (cin * k * k + cout * k * k))
        sd[name + ".weight"] = torch.from_numpy(rs.uniform(-bound, bound, (cout, cin, k, k)).astype(np.float32))
        sd[name + ".bias"] = torch.from_numpy((rs.standard_normal((cout,)) * 0.05).astype(np.float32))
    for i, cin in enumerate(in_channels):
        conv(f"lateral_convs.{i}.conv", cin, out_channels, 1)
        conv(f"fpn_convs.{i}.conv", out_channels, out_channels, 3)
    return sd


def fpn_forward(p: Dict[str, Tensor], feats: Sequence[Tensor], num_outs: int = 5) -> List[Tensor]:
    """feats: the backbone's NCHW maps, finest first -> `num_outs` pyramid levels.

    fpn.py:175-178  lateral 1x1 convolutions;
    fpn.py:182-191  top-down pathway: each coarser lateral is resized (nearest) to the next finer one's
                    size and added to it, coarsest first, so additions accumulate down the pyramid;
    fpn.py:195-197  one 3x3 convolution (padding 1) per level on the merged laterals;
    fpn.py:199-204  extra levels without extra convs: `max_pool2d(kernel 1, stride 2)` of the last output,
                    i.e. plain 2x sub-sampling.
    """
    n = len(feats)
    lat = [F.conv2d(feats[i], p[f"lateral_convs.{i}.conv.weight"], p[f"lateral_convs.{i}.conv.bias"]) for i in range(n)]
    for i in range(n - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    outs = [F.conv2d(lat[i], p[f"fpn_convs.{i}.conv.weight"], p[f"fpn_convs.{i}.conv.bias"], padding=1) for i in range(n)]
    while len(outs) < num_outs:
        outs.append(outs[-1][:, :, ::2, ::2].contiguous())
    return outs
