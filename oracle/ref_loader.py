"""Load the UNMODIFIED reference hot-path module from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY (see oracle/panoswin_oracle.py header).  Nothing is copied: the reference
file is executed in place via importlib after the few third-party modules it imports at module
level (timm, mmcv, mmcv_custom, mmdet.utils, fvcore, thop, lzx.utils, lzx.pano_rotate) are replaced
by minimal stand-ins, because none of them is installed in this image
(simple_panoswin_transformer.py:25-41, :776, :986-987, :1286, :1332).  `/root/reference` does not
exist on the GPU box, so only `oracle/make_golden.py` and container-side tests call this.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PANOSWIN_REFERENCE_ROOT", "/root/reference")
_HOT = "mmdet/models/backbones/simple_panoswin_transformer.py"
_GC = "lzx/models/great_circle.py"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _HOT))


class _StochasticDepth(nn.Module):
    """timm.models.layers.DropPath semantics (per-sample Bernoulli keep mask, rescaled)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class _Registry:
    """The two members of mmcv.utils.Registry the hot-path file touches (:36-40, :779)."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self):
        def deco(cls):
            self.module_dict[cls.__name__] = cls
            return cls
        return deco


def _install(name: str, **members):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in members.items():
        setattr(mod, k, v)
    return mod


def _exec(modname: str, relpath: str):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_cached = None


def load_reference():
    """Returns the executed reference module (attributes: SimplePanoSwinTransformer,
    WindowTransition, make_uv_hw2, make_relative_position_index, ...)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    as_pair = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    _install("timm"); _install("timm.models")
    _install("timm.models.layers", DropPath=_StochasticDepth, to_2tuple=as_pair,
             trunc_normal_=nn.init.trunc_normal_)
    _install("mmcv"); _install("mmcv.utils", Registry=_Registry)
    _install("mmcv_custom", load_checkpoint=None)
    _install("mmdet"); _install("mmdet.utils", get_root_logger=None)
    _install("fvcore"); _install("fvcore.nn", FlopCountAnalysis=None, parameter_count_table=None)
    _install("thop", profile=None)
    _install("lzx"); _install("lzx.models"); _install("lzx.utils", cv_show1=None)
    _install("lzx.pano_rotate", pano_rotate_image=None, pano_rotate=None)   # PitchAttentionModule only
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _exec("lzx.models.great_circle", _GC)
        _cached = _exec("_reference_simple_panoswin", _HOT)
    return _cached


def build_reference_model(cfg: dict, state_dict=None):
    """Instantiate the reference backbone for an oracle config dict, optionally loading a
    state_dict with strict=True (which pins the parameter names)."""
    ref = load_reference()
    kw = dict(patch_size=cfg["patch_size"], in_chans=cfg["in_chans"], embed_dim=cfg["embed_dim"],
              depths=list(cfg["depths"]), num_heads=list(cfg["num_heads"]), window_size=cfg["window_size"],
              mlp_ratio=cfg["mlp_ratio"], qkv_bias=cfg["qkv_bias"], qk_scale=cfg["qk_scale"],
              ape=cfg["ape"], patch_norm=cfg["patch_norm"], out_indices=tuple(cfg["out_indices"]),
              pano_mode=cfg["pano_mode"], drop_path_rate=0.0)
    model = ref.SimplePanoSwinTransformer(**kw)
    model.init_weights(None)
    if state_dict is not None:
        # alpha/beta alias one storage in the reference (:145-147); give each its own before loading
        for m in model.modules():
            if hasattr(m, "sphere_position_beta_table_Te"):
                m.sphere_position_beta_table_Te.data = m.sphere_position_beta_table_Te.data.clone()
        model.load_state_dict(state_dict, strict=True)
    model.eval()            # returns None in the reference (:981-983): never chain
    return model


@torch.no_grad()
def reference_forward(model, img, return_blocks=False):
    """Run the reference; optionally capture every block's output tokens (uv channels dropped)."""
    blocks, hooks = [], []
    if return_blocks:
        for layer in model.layers:
            for blk in layer.blocks:
                hooks.append(blk.register_forward_hook(lambda m, i, o: blocks.append(o[..., :-2].clone())))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        outs = model(img)
    for h in hooks:
        h.remove()
    return (outs, blocks) if return_blocks else outs


# ----------------------------------------------------------------------------------------------
# the FPN neck (a caller of the backbone; used only to pin oracle/fpn_oracle.py)
# ----------------------------------------------------------------------------------------------
_FPN = "mmdet/models/necks/fpn.py"


class _PlainConvModule(nn.Module):
    """mmcv.cnn.ConvModule for the only configuration FPN builds in the shipped configs: no norm, no
    activation (conv_cfg / norm_cfg / act_cfg all None, fpn.py:120-141) -> a biased Conv2d held as `.conv`."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, conv_cfg=None, norm_cfg=None,
                 act_cfg=None, inplace=False):
        super().__init__()
        assert conv_cfg is None and norm_cfg is None and act_cfg is None, "stand-in covers the plain-conv FPN only"
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding)

    def forward(self, x):
        return self.conv(x)


def _xavier_init(module, gain=1, bias=0, distribution="normal"):
    (nn.init.xavier_uniform_ if distribution == "uniform" else nn.init.xavier_normal_)(module.weight, gain=gain)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


_cached_fpn = None


def load_reference_fpn():
    """Executes the unmodified mmdet/models/necks/fpn.py and returns its FPN class."""
    global _cached_fpn
    if _cached_fpn is not None:
        return _cached_fpn
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, _FPN)):
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    identity_decorator = lambda *a, **k: (lambda fn: fn)
    _install("mmcv"); _install("mmcv.cnn", ConvModule=_PlainConvModule, xavier_init=_xavier_init)
    _install("mmcv.runner", auto_fp16=identity_decorator)
    for pkg in ("mmdet", "mmdet.models", "mmdet.models.necks"):           # parents of the relative import at fpn.py:8
        _install(pkg).__path__ = []
    _install("mmdet.models.builder", NECKS=_Registry("neck"))
    _cached_fpn = _exec("mmdet.models.necks.fpn", _FPN).FPN
    return _cached_fpn


def build_reference_fpn(in_channels, state_dict, out_channels=256, num_outs=5):
    fpn = load_reference_fpn()(in_channels=list(in_channels), out_channels=out_channels, num_outs=num_outs)
    fpn.init_weights()
    fpn.load_state_dict(state_dict, strict=True)                          # pins the parameter names
    fpn.eval()
    return fpn
