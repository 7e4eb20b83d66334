"""panoswintransformerobjectdetection_b200 — B200-native (sm_100a) implementation of PanoSwin's pano-style
shifted-window attention path behind the reference's backbone API.

    from panoswintransformerobjectdetection_b200 import SimplePanoSwinTransformer   # drop-in nn.Module
    from panoswintransformerobjectdetection_b200 import ops                         # C-ABI op wrappers

The CUDA library (libpanoswin_b200.so, C ABI in include/panoswin_b200.h) is built in-tree by
`__graft_entry__.build()` / `python -m panoswintransformerobjectdetection_b200._build`.
"""
from . import _build, _lib, ops  # noqa: F401
from .backbone import (SimplePanoSwinTransformer, make_relative_position_index, make_uv_hw2,  # noqa: F401
                       planar_attention_mask)
from .detector import FPN, PanoSwinMaskRCNN  # noqa: F401
from .registry import BACKBONES, build_backbone  # noqa: F401

__version__ = "0.1.0"
