"""Drop-in `SimplePanoSwinTransformer` whose forward runs on libpanoswin_b200 (sm_100a CUDA).

Mirrors the reference module's public surface (mmdet/models/backbones/simple_panoswin_transformer.py):
same constructor keywords (:781-801), `forward(x, pano_ratio_v=None) -> tuple of fp32 NCHW maps` (:940-979),
`init_weights(pretrained)` (:885-907), `train(mode)` returning None (:981-983), `set_pano_mode` /
`switch_pano_mode` (:880-883, :207-208), `num_features`, `out_indices`, the `BACKBONES` registry entry
(:779-780) and — so that checkpoints and optimizer `paramwise_cfg` keys keep working — exactly the same
parameter / buffer names.  The sub-modules below are parameter containers; the compute is
the kernel sequence in `_forward_tokens`, which never materialises the shifted / padded / partitioned
copies, the uv channels or the [nW,49,49,heads] bias tensor of the reference.

Compute modes (`set_compute_dtype`): "bf16" (default; bf16 activations + tcgen05 GEMM/attention, fp32
residual stream, fp32 softmax / LayerNorm statistics) and "fp32" (CUDA-core kernels, <=1e-5 of the reference).
In train mode with gradients enabled the forward runs through torch.autograd Functions whose backward is a set of
hand-written kernels behind the same C ABI (autograd.py, csrc/psw_train.cu, csrc/psw_attn_bwd.cu).
"""
from __future__ import annotations

import math
import warnings
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import autograd as AG
from . import ops
from .registry import BACKBONES

__all__ = ["SimplePanoSwinTransformer", "make_uv_hw2", "make_relative_position_index", "planar_attention_mask"]


# ------------------------------------------------------------------------------------------------
# host-side constants (tiny; computed once per resolution on the CPU exactly like the reference)
# ------------------------------------------------------------------------------------------------
def make_relative_position_index(window_size) -> torch.Tensor:
    """idx[i,j] = (dr + wh-1)*(2*ww-1) + (dc + ww-1) for tokens i, j of a window (reference :95-129)."""
    wh, ww = (window_size, window_size) if isinstance(window_size, int) else window_size
    t = torch.arange(wh * ww)
    r, c = torch.div(t, ww, rounding_mode="floor"), t % ww
    return (r[:, None] - r[None, :] + wh - 1) * (2 * ww - 1) + (c[:, None] - c[None, :] + ww - 1)


def make_uv_hw2(H: int, W: int, device="cpu") -> torch.Tensor:
    """Token-centre longitude/latitude grid [H, W, 2], u in [-pi, pi), v in [-pi/2, pi/2); the angular
    step is pi/H on both axes and the fp32 rounding order is the reference's (:173-188)."""
    if W < H:
        raise ValueError(f"make_uv_hw2 needs W >= H, got H={H} W={W}")
    gap = math.pi / H
    u = (torch.arange(W) * gap - math.pi) + 0.5 * gap
    v = (torch.arange(H) * gap - math.pi * 0.5) + 0.5 * gap
    uv = torch.stack([u[None, :].expand(H, W), v[:, None].expand(H, W)], dim=-1).contiguous()
    return uv.to(device)


def planar_attention_mask(H: int, W: int, window_size: int, shift_size: int) -> torch.Tensor:
    """0 / -100 SW-MSA mask [nW, ws*ws, ws*ws] of planar mode (reference :664-688)."""
    ws = window_size
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    band_h = torch.zeros(Hp)
    band_h[Hp - ws:Hp - shift_size] = 1
    band_h[Hp - shift_size:] = 2
    band_w = torch.zeros(Wp)
    band_w[Wp - ws:Wp - shift_size] = 1
    band_w[Wp - shift_size:] = 2
    region = band_h[:, None] * 3 + band_w[None, :]
    rw = region.view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = rw[:, None, :] - rw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff)).contiguous()


def _trunc_normal_(t: torch.Tensor, std: float = 0.02):
    return nn.init.trunc_normal_(t, std=std)


class DropPath(nn.Module):
    """Stochastic depth (timm semantics).  Identity in eval mode / rate 0; kept for API parity."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class DoubleModeModule(object):
    """pano / planar switch shared by every level of the backbone (reference :192-208)."""

    def __init__(self, pano_mode=True):
        super().__init__()
        self.pano_mode = pano_mode
        self.set_pano_mode(pano_mode=pano_mode)

    def set_pano_mode(self, pano_mode: bool):
        self.pano_mode = pano_mode

    def switch_pano_mode(self):
        self.set_pano_mode(not self.pano_mode)


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's names
# ------------------------------------------------------------------------------------------------
class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("libpanoswin_b200 fuses the exact-erf GELU of the reference Mlp only")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


class WindowAttention(nn.Module, DoubleModeModule):
    """Holds qkv / proj / the two great-circle tables / the index buffer (reference :211-239, :315-323)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 pano_mode=True):
        nn.Module.__init__(self)
        DoubleModeModule.__init__(self, pano_mode=pano_mode)
        self.dim = dim
        self.window_size = (window_size, window_size) if isinstance(window_size, int) else tuple(window_size)
        if self.window_size[0] != self.window_size[1]:
            raise NotImplementedError("square windows only (as every reference config)")
        assert dim % num_heads == 0, "channels should be divisible by heads, but we get channel {} and heads{}".format(dim, num_heads)
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.register_buffer("relative_position_index_OO", make_relative_position_index(self.window_size))
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        tsize = (2 * self.window_size[0] - 1) * (2 * self.window_size[1] - 1)
        # the reference aliases alpha and beta at init (:145-147); they are independent tensors here
        # (checkpoints store both) but start from the same draw so a fresh model matches its statistics
        first = _trunc_normal_(torch.zeros(tsize, num_heads), std=0.02)
        self.sphere_position_alpha_table_Te = nn.Parameter(first.clone())
        self.sphere_position_beta_table_Te = nn.Parameter(first.clone())
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)


class WindowTransition(nn.Module, DoubleModeModule):
    """Shift description only: the pano shift / un-shift is address arithmetic inside the attention
    kernel (psw::source_token), so this module carries `shift_size` and the mode flag, no compute."""

    def __init__(self, shift_size=0, pano_mode=False):
        nn.Module.__init__(self)
        DoubleModeModule.__init__(self, pano_mode=pano_mode)
        self.shift_size = shift_size


class PanoSwinTransformerBlock(nn.Module, DoubleModeModule):
    def __init__(self, dim, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, pano_mode=True):
        nn.Module.__init__(self)
        self.dim = dim
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("libpanoswin_b200 implements nn.LayerNorm only")
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim=dim, window_size=window_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop, pano_mode=pano_mode)
        self.window_transition = WindowTransition(shift_size=shift_size, pano_mode=pano_mode)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.H = None
        self.W = None
        DoubleModeModule.__init__(self, pano_mode=pano_mode)

    def set_pano_mode(self, pano_mode):
        self.attn.set_pano_mode(pano_mode=pano_mode)
        self.window_transition.set_pano_mode(pano_mode=pano_mode)
        self.pano_mode = pano_mode


class PitchAttentionModule(nn.Module, DoubleModeModule):
    """Parameter container of the reference's PitchAttentionModule (:990-1022; BasicWindowAttention members :224-239):
    the trailing block of a stage with an odd depth.  In planar mode its "rotated" map is the map itself (:1177-1179):
    every un-shifted window attends to itself through separate q / k / v linears -- that forward runs on the same
    kernels as a regular block (backbone._pitch_block).  In pano mode the reference itself cannot execute it
    (:1038 calls pano_rotate_image(..., with_uv=True), which lzx/pano_rotate.py:169 does not accept), so there is no
    behaviour to reproduce and the forward raises."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, np_v=-0.0001,
                 norm_layer=nn.LayerNorm, drop_path=0.0, mlp_ratio=4.0, drop=0.0, act_layer=nn.GELU, pano_mode=True):
        nn.Module.__init__(self)
        DoubleModeModule.__init__(self, pano_mode=pano_mode)
        self.dim = dim
        self.window_size = (window_size, window_size) if isinstance(window_size, int) else tuple(window_size)
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.shift_size = 0
        self.register_buffer("relative_position_index_OO", make_relative_position_index(self.window_size))
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(drop)
        tsize = (2 * self.window_size[0] - 1) * (2 * self.window_size[1] - 1)
        first = _trunc_normal_(torch.zeros(tsize, num_heads), std=0.02)
        self.sphere_position_alpha_table_Te = nn.Parameter(first.clone())
        self.sphere_position_beta_table_Te = nn.Parameter(first.clone())
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.norm2 = norm_layer(dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm1 = norm_layer(dim)
        self.q_linear = nn.Linear(dim, dim, bias=qkv_bias)
        self.k_linear = nn.Linear(dim, dim, bias=qkv_bias)
        self.v_linear = nn.Linear(dim, dim, bias=qkv_bias)
        self.register_buffer("np_uv", torch.tensor([1.0, np_v]) * math.pi)
        self.H = None
        self.W = None


class PatchMerging(nn.Module):
    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)


class BasicLayer(nn.Module, DoubleModeModule):
    def __init__(self, dim, depth, num_heads, window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 pano_mode=True):
        nn.Module.__init__(self)
        self.window_size = window_size
        self.shift_size = window_size // 2
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        blocks = [
            PanoSwinTransformerBlock(
                dim=dim, num_heads=num_heads, window_size=window_size,
                shift_size=0 if (i % 2 == 0) else window_size // 2, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path, norm_layer=norm_layer,
                pano_mode=pano_mode)
            for i in range(depth - depth % 2)]
        if depth % 2:                                       # reference :636-647
            blocks.append(PitchAttentionModule(dim=dim, num_heads=num_heads, window_size=window_size, qkv_bias=qkv_bias,
                                               qk_scale=qk_scale, attn_drop=attn_drop, mlp_ratio=mlp_ratio,
                                               norm_layer=norm_layer, drop=drop))
        self.blocks = nn.ModuleList(blocks)
        self.downsample = downsample(dim=dim, norm_layer=norm_layer) if downsample is not None else None
        DoubleModeModule.__init__(self, pano_mode=pano_mode)

    def set_pano_mode(self, pano_mode=True):
        self.pano_mode = pano_mode
        for block in self.blocks:
            block.set_pano_mode(pano_mode)


class PatchEmbed(nn.Module):
    """Stem parameters (reference :727-773): conv3x3-BN-ReLU, conv3x3-BN-ReLU, conv(patch)/stride(patch), LN."""

    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.patch_size = patch_size
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        d3 = embed_dim // 3
        self.proj = nn.Sequential(
            nn.Conv2d(in_chans, d3, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(d3), nn.ReLU(inplace=True),
            nn.Conv2d(d3, d3 * 2, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(d3 * 2), nn.ReLU(inplace=True),
            nn.Conv2d(d3 * 2, embed_dim, kernel_size=patch_size, stride=patch_size))
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None


# ------------------------------------------------------------------------------------------------
# the backbone
# ------------------------------------------------------------------------------------------------
@BACKBONES.register_module()
class SimplePanoSwinTransformer(nn.Module, DoubleModeModule):
    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, depths=[2, 2, 7, 2], num_heads=[3, 6, 12, 24],
                 window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.2, norm_layer=nn.LayerNorm, ape=False, patch_norm=True, out_indices=(0, 1, 2, 3),
                 frozen_stages=-1, use_checkpoint=False, pano_mode=True):
        nn.Module.__init__(self)
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape = ape
        self.patch_norm = patch_norm
        self.out_indices = out_indices
        self.frozen_stages = frozen_stages
        self.window_size = window_size
        self.patch_embed = PatchEmbed(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        if self.ape:
            self.abs_encoder = nn.Linear(5, embed_dim)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i_layer in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i_layer), depth=depths[i_layer], num_heads=num_heads[i_layer],
                window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                attn_drop=attn_drop_rate, drop_path=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])],
                norm_layer=norm_layer, downsample=PatchMerging if (i_layer < self.num_layers - 1) else None,
                use_checkpoint=use_checkpoint, pano_mode=pano_mode))
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        for i_layer in out_indices:
            self.add_module(f"norm{i_layer}", norm_layer(self.num_features[i_layer]))
        self._compute_dtype = torch.bfloat16
        self._residual_dtype = torch.float32
        self._fused_conv_relu = None
        # Resolution-dependent constants (uv grids, planar masks, position rows, the per-block bias tables of the bf16
        # attention kernel: ~150 MB per resolution for PanoSwin-T at 512x1024) live in an LRU keyed by the input
        # resolution: multi-scale inputs (the shipped configs train at 480-800) recycle the oldest entry instead of
        # growing without bound.  Converted weights (bf16 copies, folded stem) are resolution-independent.
        self.max_cached_resolutions = 2
        self._res_cache: "OrderedDict[tuple, Dict[tuple, object]]" = OrderedDict()
        self._cur_res: Dict[tuple, object] = {}
        self._weight_cache: Dict[tuple, tuple] = {}
        DoubleModeModule.__init__(self, pano_mode=pano_mode)

    # ---- reference API ------------------------------------------------------------------------
    def set_pano_mode(self, pano_mode=True):
        self.pano_mode = pano_mode
        for layer in self.layers:
            layer.set_pano_mode(pano_mode)

    def init_weights(self, pretrained=None):
        def _init_weights(m):
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

        if isinstance(pretrained, str):
            self.apply(_init_weights)
            _load_checkpoint(self, pretrained)
        elif pretrained is None:
            self.apply(_init_weights)
        else:
            raise TypeError("pretrained must be a str or None")

    def train(self, mode=True):
        """Like the reference (:981-983) this returns None — do not chain `.eval()`."""
        super(SimplePanoSwinTransformer, self).train(mode)

    # ---- B200 additions -----------------------------------------------------------------------
    def set_compute_dtype(self, dtype):
        """'bf16' (throughput path, default) or 'fp32' (parity path)."""
        table = {"bf16": torch.bfloat16, "fp32": torch.float32, torch.bfloat16: torch.bfloat16,
                 torch.float32: torch.float32}
        if dtype not in table:
            raise ValueError("compute dtype must be 'bf16' or 'fp32'")
        self._compute_dtype = table[dtype]
        return self

    @property
    def compute_dtype(self):
        return self._compute_dtype

    def set_residual_dtype(self, dtype):
        """Storage type of the residual stream in bf16 mode: 'fp32' (default, what the reference keeps under
        autocast) or 'bf16' (halves the residual traffic; measured accuracy in profiles/)."""
        table = {"bf16": torch.bfloat16, "fp32": torch.float32}
        if dtype not in table:
            raise ValueError("residual dtype must be 'bf16' or 'fp32'")
        self._residual_dtype = table[dtype]
        return self

    # ---- cached constants / converted weights -------------------------------------------------
    def _enter_resolution(self, H, W, device):
        """Select (creating / evicting as needed) the constants of one input resolution; called once per forward."""
        key = (int(H), int(W), bool(self.pano_mode), str(device))
        ent = self._res_cache.get(key)
        if ent is None:
            ent = {}
            self._res_cache[key] = ent
            while len(self._res_cache) > max(1, int(self.max_cached_resolutions)):
                self._res_cache.popitem(last=False)        # a CUDA graph that captured these tensors keeps its own references
        else:
            self._res_cache.move_to_end(key)
        self._cur_res = ent
        return ent

    def _const(self, key, builder, device):
        t = self._cur_res.get(key)
        if t is None:
            t = builder().to(device).contiguous()
            self._cur_res[key] = t
        return t

    def cache_signature(self):
        """Everything a captured CUDA graph of the forward bakes in besides the input buffer: parameter / buffer
        versions and addresses, the compute and residual dtypes, pano mode and train / eval.  runtime.GraphedForward
        compares it before every replay and re-captures when it changed."""
        sig = [self._compute_dtype, self._residual_dtype, bool(self.pano_mode), bool(self.training)]
        for t in list(self.parameters()) + list(self.buffers()):
            sig.append(t._version)
            sig.append(t.data_ptr())
        return tuple(sig)

    def _w(self, p: torch.Tensor, dtype) -> torch.Tensor:
        """Parameter in the compute dtype (bf16 copies are cached until the parameter changes)."""
        if p.dtype == dtype:
            return p.detach().contiguous()
        k = (id(p), dtype)
        hit = self._weight_cache.get(k)
        if hit is not None and hit[0] == p._version and hit[1] == p.data_ptr():
            return hit[2]
        conv = ops.cast(p.detach().contiguous(), dtype)
        self._weight_cache[k] = (p._version, p.data_ptr(), conv)
        return conv

    def _bias_full(self, attn, uv, mask, H, W, ws, shift):
        """Full additive-bias table of one block at one resolution for the bf16 attention kernel (great-circle +
        relative-position bias, planar shift mask), cached until alpha / beta change."""
        al, be = attn.sphere_position_alpha_table_Te, attn.sphere_position_beta_table_Te
        k = ("bfull", id(attn), H, W, ws, shift)
        ver = (al._version, be._version, al.data_ptr(), be.data_ptr())
        hit = self._cur_res.get(k)
        if hit is None or hit[0] != ver:
            hit = (ver, ops.window_bias_full(al.detach().contiguous(), be.detach().contiguous(), uv, mask, H, W, ws, shift,
                                             self.pano_mode))
            self._cur_res[k] = hit
        return hit[1]

    @staticmethod
    def _f(p: Optional[torch.Tensor]):
        return None if p is None else p.detach().contiguous()

    # ---- stem ---------------------------------------------------------------------------------
    def _stem(self, x: torch.Tensor):
        """conv-BN-ReLU x2 + patch conv -> NHWC tokens [B, Hs, Ws, E] (reference PatchEmbed.forward :756-773).
        bf16 mode: eval-mode BatchNorm folded into the convolutions, all three on libpanoswin_b200 (tcgen05) for the
        stem widths the kernels are instantiated for; fp32 mode and other widths: see the branches below."""
        pe = self.patch_embed
        ph, pw = pe.patch_size
        _, _, H, W = x.shape
        if W % pw != 0:
            x = F.pad(x, (0, pw - W % pw))
        if H % ph != 0:
            x = F.pad(x, (0, 0, 0, ph - H % ph))
        conv1, bn1, _, conv2, bn2, _, conv3 = pe.proj
        if self._compute_dtype == torch.float32:
            # parity path: direct fp32 convolutions on CUDA cores, BatchNorm (running statistics) + ReLU in the epilogue
            def bn_affine(bn):
                s = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
                return s, (bn.bias - bn.running_mean * s).detach().float().contiguous()
            s1, t1 = bn_affine(bn1)
            s2, t2 = bn_affine(bn2)
            y = ops.conv2d_f32(x.contiguous(), self._f(conv1.weight), self._f(conv1.bias), s1, t1, 1, 1, relu=True)
            y = ops.conv2d_f32(y, self._f(conv2.weight), self._f(conv2.bias), s2, t2, 1, 1, relu=True)
            if ph != pw:
                raise NotImplementedError("square patches only")
            return ops.conv2d_f32(y, self._f(conv3.weight), self._f(conv3.bias), None, None, ph, 0, out_nhwc=True)
        # bf16: fold the eval-mode BatchNorm into the convolution, channels-last
        def folded(conv, bn):
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = (conv.weight * s[:, None, None, None]).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            b = ((conv.bias - bn.running_mean) * s + bn.bias).to(torch.bfloat16)
            return w, b
        key = ("stem", tuple(p._version for p in pe.proj.parameters()), bn1.running_mean._version, bn2.running_mean._version)
        if self._weight_cache.get("stem_key") != key:
            s1 = bn1.weight / torch.sqrt(bn1.running_var + bn1.eps)
            self._weight_cache["stem"] = (folded(conv1, bn1), folded(conv2, bn2),
                                          (conv3.weight.to(torch.bfloat16).contiguous(memory_format=torch.channels_last),
                                           conv3.bias.to(torch.bfloat16)),
                                          ((conv1.weight * s1[:, None, None, None]).reshape(conv1.out_channels, -1).float().contiguous(),
                                           ((conv1.bias - bn1.running_mean) * s1 + bn1.bias).float().contiguous()))
            s2 = bn2.weight / torch.sqrt(bn2.running_var + bn2.eps)           # conv2 for psw_stem_conv3x3_c32_relu_fwd:
            self._weight_cache["stem2"] = (                                    # [tap = ky*3+kx][out][in] bf16, fp32 bias
                (conv2.weight * s2[:, None, None, None]).permute(2, 3, 0, 1).reshape(9, conv2.out_channels, -1)
                .to(torch.bfloat16).contiguous(), ((conv2.bias - bn2.running_mean) * s2 + bn2.bias).float().contiguous())
            self._weight_cache["stem_key"] = key
        (w1, b1), (w2, b2), (w3, b3), (w1f, b1f) = self._weight_cache["stem"]
        own_conv1 = (conv1.in_channels == 3 and conv1.out_channels == 32      # libpanoswin_b200 tcgen05 conv (E = 96)
                     and x.shape[3] % 4 == 0)                                 # (its patch loader moves 16-byte groups)
        own_conv2 = own_conv1 and conv2.in_channels == 32 and conv2.out_channels in (32, 64) and conv2.kernel_size == (3, 3)
        if own_conv2:
            w2t, b2f = self._weight_cache["stem2"]
            y = ops.stem_conv3x3_c32_relu(ops.stem_conv3x3_relu(x.contiguous(), w1f, b1f), w2t, b2f)       # NHWC bf16
            if (pw * conv3.in_channels) % 64 == 0 and conv3.out_channels % 16 == 0:
                k3 = ("stem3", conv3.weight._version, conv3.weight.data_ptr(), None if conv3.bias is None else conv3.bias._version)
                hit = self._weight_cache.get("stem3")
                if hit is None or hit[0] != k3:              # [cout, cin, ph, pw] -> [cout, ph, pw, cin] bf16
                    hit = (k3, conv3.weight.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous(),
                           None if conv3.bias is None else conv3.bias.detach().float().contiguous())
                    self._weight_cache["stem3"] = hit
                return ops.patch_conv(y, hit[1], hit[2], pe.patch_size)                                  # [B, Hs, Ws, E]
            y = F.conv2d(y.permute(0, 3, 1, 2), w3, b3, stride=pe.patch_size)
            return y.permute(0, 2, 3, 1).contiguous()      # no copy when the conv output is channels-last
        generic = (conv1.in_channels == 3 and conv1.out_channels <= 64 and conv2.in_channels == conv1.out_channels
                   and conv2.out_channels <= 256 and conv1.kernel_size == (3, 3) and conv2.kernel_size == (3, 3)
                   and ph == pw and conv3.out_channels % 16 == 0 and x.shape[3] % 4 == 0)
        if generic:
            # other stem widths (PanoSwin-B: 42 / 84 channels): the same three tcgen05 kernels families on zero-padded
            # channels -- conv1 -> 64, conv2 as a GEMM over shifted TMA views -> a multiple of 16 with pw * c % 64 == 0,
            # patch conv on the padded input; the padding channels carry exact zeros
            hit = self._weight_cache.get("stem_generic")
            if hit is None or hit[0] != key:
                c1, c2, E = conv1.out_channels, conv2.out_channels, conv3.out_channels
                c2p = -(-c2 // 16) * 16
                while (pw * c2p) % 64:
                    c2p += 16
                s1 = bn1.weight / torch.sqrt(bn1.running_var + bn1.eps)
                s2 = bn2.weight / torch.sqrt(bn2.running_var + bn2.eps)
                dev = x.device
                w1p = torch.zeros(64, 27, device=dev)
                w1p[:c1] = (conv1.weight * s1[:, None, None, None]).reshape(c1, -1).float()
                b1p = torch.zeros(64, device=dev)
                b1p[:c1] = ((conv1.bias - bn1.running_mean) * s1 + bn1.bias).float()
                w2p = torch.zeros(c2p, 3, 3, 64, device=dev)
                w2p[:c2, :, :, :c1] = (conv2.weight * s2[:, None, None, None]).permute(0, 2, 3, 1).float()
                b2p = torch.zeros(c2p, device=dev)
                b2p[:c2] = ((conv2.bias - bn2.running_mean) * s2 + bn2.bias).float()
                w3p = torch.zeros(E, ph, pw, c2p, device=dev)
                w3p[..., :c2] = conv3.weight.permute(0, 2, 3, 1).float()
                b3p = None if conv3.bias is None else conv3.bias.detach().float().contiguous()
                hit = (key, w1p.detach().contiguous(), b1p.detach().contiguous(), w2p.detach().to(torch.bfloat16).contiguous(),
                       b2p.detach().contiguous(), w3p.detach().to(torch.bfloat16).contiguous(), b3p)
                self._weight_cache["stem_generic"] = hit
            _, w1p, b1p, w2p, b2p, w3p, b3p = hit
            y = ops.stem_conv3x3_relu(x.contiguous(), w1p, b1p)                      # [B, H, W, 64] bf16 NHWC
            y = ops.conv3x3_nhwc(y, w2p, b2p, relu=True)                             # [B, H, W, c2p]
            return ops.patch_conv(y, w3p, b3p, pe.patch_size)                        # [B, Hs, Ws, E]
        if own_conv1:
            # fp32 NCHW image -> bf16 NHWC, seen by the next convolution as a channels-last NCHW tensor (no copy)
            y = ops.stem_conv3x3_relu(x.contiguous(), w1f, b1f).permute(0, 3, 1, 2)
        else:
            y = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        if self._fused_conv_relu is None:                   # cuDNN's fused conv+bias+ReLU, probed once
            try:
                pw, pb = (w2, b2) if own_conv1 else (w1, b1)
                t = torch.cudnn_convolution_relu(y[:1, :, :8, :8].contiguous(memory_format=torch.channels_last), pw, pb,
                                                 (1, 1), (1, 1), (1, 1), 1)
                ref = F.relu(F.conv2d(y[:1, :, :8, :8], pw, pb, padding=1))
                self._fused_conv_relu = bool(torch.allclose(t.float(), ref.float(), atol=2e-2, rtol=2e-2))
            except (RuntimeError, AttributeError):
                self._fused_conv_relu = False
        if self._fused_conv_relu:
            if not own_conv1:
                y = torch.cudnn_convolution_relu(y, w1, b1, (1, 1), (1, 1), (1, 1), 1)
            y = torch.cudnn_convolution_relu(y, w2, b2, (1, 1), (1, 1), (1, 1), 1)
        else:
            if not own_conv1:
                y = F.relu_(F.conv2d(y, w1, b1, padding=1))
            y = F.relu_(F.conv2d(y, w2, b2, padding=1))
        y = F.conv2d(y, w3, b3, stride=pe.patch_size)
        return y.permute(0, 2, 3, 1).contiguous()          # no copy when the conv output is channels-last

    def _abs_position(self, Hs, Ws, device) -> torch.Tensor:
        """fp32 [Hs*Ws, E] = abs_encoder(x, y, z, u, v) (reference :925-934); cached per resolution."""
        ver = (self.abs_encoder.weight._version, self.abs_encoder.bias._version, self.abs_encoder.weight.data_ptr())
        def build():
            uv = make_uv_hw2(Hs, Ws)
            u, v = uv[..., 0], uv[..., 1]
            return torch.stack([torch.sin(u) * torch.sin(v), torch.cos(u) * torch.sin(v), torch.cos(v), u, v], -1).reshape(-1, 5)
        xyzuv = self._const(("xyzuv", Hs, Ws), build, device)
        k = ("pos", Hs, Ws)
        hit = self._cur_res.get(k)
        if hit is None or hit[0] != ver:
            pos = ops.linear(xyzuv, self._f(self.abs_encoder.weight), self._f(self.abs_encoder.bias))
            self._cur_res[k] = hit = (ver, pos)
        return hit[1]

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, x_bchw, pano_ratio_v=None):
        if pano_ratio_v is not None:
            warnings.warn("Parameter pano_ratio_v for is deprecated! Please set it to None!")
        if self.pano_mode and x_bchw.shape[3] != x_bchw.shape[2] * 2:
            warnings.warn("PanoSwin is configured in Pano mode, expecting channel3 == 2 * channel2, but get {} and {}, "
                          "probably cause an error".format(x_bchw.shape[3], x_bchw.shape[2]))
        self._check_forward(x_bchw)
        if self.training and torch.is_grad_enabled():
            return self._forward_train(x_bchw.float())      # autograd path: libpanoswin_b200 forward + backward kernels
        with torch.no_grad():
            return self._forward_tokens(x_bchw.float())

    def forward_streamed(self, x_bchw, on_output=None, on_block=None):
        """forward() that additionally calls `on_output(k, feature_map)` as soon as the k-th output map has been
        enqueued, so a caller can start consuming (e.g. copying out) early stages while later ones still run, and
        `on_block(n, tokens)` after the n-th PanoSwinTransformerBlock with its output tokens [B, H*W, C] (the residual
        stream; it is updated in place by the following kernels, so a consumer must copy what it keeps) -- the
        equivalent of a forward hook on `layers[i].blocks[j]` of the reference (:493-536) minus the uv channels."""
        self._check_forward(x_bchw)
        with torch.no_grad():
            return self._forward_tokens(x_bchw.float(), on_output, on_block)

    def _check_forward(self, x_bchw):
        if not x_bchw.is_cuda:
            raise ops.PanoSwinB200Error("SimplePanoSwinTransformer (B200) needs a CUDA input; there is no CPU fallback")
        if self.pano_mode and not self.ape:
            raise AttributeError("pano_mode=True requires ape=True: the reference builds abs_encoder only when ape is set "
                                 "(simple_panoswin_transformer.py:841-842) and always calls it in pano mode (:934)")

    # ---- training forward (autograd) -----------------------------------------------------------
    def _forward_train(self, img: torch.Tensor) -> Tuple[torch.Tensor, ...]:
        """forward() in train mode with gradients (reference :940-979 under autograd; callers
        mmdet/apis/train.py:91-99, mmdet/utils/optimizer.py:22-33).  Every op of the block path runs on a
        libpanoswin_b200 kernel forward AND backward (autograd.py); the stem's convolutions and the 5-input abs_encoder go
        through torch, its train-mode BatchNorm + ReLU (batch statistics) through libpanoswin_b200 in bf16 mode.  compute dtype fp32: CUDA-core kernels,
        gradients within 1e-4 of torch autograd; bf16: bf16 activations, fp32 residual stream and parameters, tcgen05
        forward / input-gradient GEMMs.  DropPath (:533-534) and `use_checkpoint` (:705-706) are honoured."""
        cd = self._compute_dtype
        dev = img.device
        ws = self.window_size
        pe = self.patch_embed
        ph, pw = pe.patch_size
        _, _, H0, W0 = img.shape
        if W0 % pw != 0:
            img = F.pad(img, (0, pw - W0 % pw))
        if H0 % ph != 0:
            img = F.pad(img, (0, 0, 0, ph - H0 % ph))
        if cd == torch.bfloat16:
            # throughput mode: the stem's convolutions forward and backward in bf16 channels-last (cuDNN under autocast,
            # BatchNorm statistics in fp32) -- what the reference does under apex O1 (mmdet/apis/train.py:82-88)
            conv1, bn1, _, conv2, bn2, _, conv3 = pe.proj
            own_bn = all(isinstance(b, nn.BatchNorm2d) and b.training and b.affine and b.track_running_stats
                         and b.momentum is not None and b.num_features % 8 == 0 and b.num_features <= 256 for b in (bn1, bn2))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                x0 = img.contiguous(memory_format=torch.channels_last)
                if own_bn:                                        # train-mode BatchNorm + ReLU on libpanoswin_b200 (NHWC bf16)
                    y = x0
                    for conv, bn in ((conv1, bn1), (conv2, bn2)):
                        y = AG.BatchNormReluFn.apply(conv(y), bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                                     bn.momentum, bn.eps)
                        with torch.no_grad():
                            bn.num_batches_tracked.add_(1)
                    tok = conv3(y).float()
                else:
                    tok = pe.proj(x0).float()
        else:
            prev = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            try:
                tok = pe.proj(img)                                # conv-BN-ReLU x2 + patch conv (torch / cuDNN, fp32)
            finally:
                torch.backends.cudnn.allow_tf32 = prev
        B, E, Hs, Ws = tok.shape
        self._enter_resolution(Hs, Ws, dev)
        x = tok.permute(0, 2, 3, 1).reshape(B, Hs * Ws, E)
        if pe.norm is not None:
            x = AG.LayerNormFn.apply(x, pe.norm.weight, pe.norm.bias, pe.norm.eps, torch.float32)
        if self.pano_mode and self.ape:
            def build():
                uv = make_uv_hw2(Hs, Ws)
                u, v = uv[..., 0], uv[..., 1]
                return torch.stack([torch.sin(u) * torch.sin(v), torch.cos(u) * torch.sin(v), torch.cos(v), u, v], -1).reshape(-1, 5)
            xyzuv = self._const(("xyzuv", Hs, Ws), build, dev)
            x = x + self.abs_encoder(xyzuv)[None]
        x = self.pos_drop(x)

        def wc(p):                                                # weight in the compute dtype (no gradient of its own)
            return p.detach() if cd == torch.float32 else self._w(p, cd)

        H, W = Hs, Ws
        outs = []
        for i, layer in enumerate(self.layers):
            C = self.num_features[i]
            uv = self._const(("uv", H, W), lambda: make_uv_hw2(H, W), dev) if self.pano_mode else None
            for blk in layer.blocks:
                if isinstance(blk, PitchAttentionModule):
                    raise NotImplementedError("PitchAttentionModule (odd stage depth) has no training path; use even depths")
                shift = blk.shift_size
                mask = None
                if not self.pano_mode and shift > 0:
                    mask = self._const(("mask", H, W, ws, shift), lambda: planar_attention_mask(H, W, ws, shift), dev)

                def block_fn(x, blk=blk, shift=shift, mask=mask, uv=uv, H=H, W=W, C=C):
                    a = blk.attn
                    xn = AG.LayerNormFn.apply(x, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, cd)
                    qkv = AG.LinearFn.apply(xn, a.qkv.weight, a.qkv.bias, wc(a.qkv.weight), cd)
                    att = AG.WindowAttentionFn.apply(qkv.view(B, H, W, 3 * C), a.sphere_position_alpha_table_Te,
                                                     a.sphere_position_beta_table_Te, a.qkv.bias, uv, mask, a.num_heads, ws,
                                                     shift, self.pano_mode, a.scale)
                    y = AG.LinearFn.apply(att.view(B, H * W, C), a.proj.weight, a.proj.bias, wc(a.proj.weight), torch.float32)
                    x = x + blk.drop_path(a.proj_drop(y))
                    xn2 = AG.LayerNormFn.apply(x, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, cd)
                    h = AG.LinearFn.apply(xn2, blk.mlp.fc1.weight, blk.mlp.fc1.bias, wc(blk.mlp.fc1.weight), cd)
                    h = blk.mlp.drop(AG.GeluFn.apply(h))
                    y2 = AG.LinearFn.apply(h, blk.mlp.fc2.weight, blk.mlp.fc2.bias, wc(blk.mlp.fc2.weight), torch.float32)
                    return x + blk.drop_path(blk.mlp.drop(y2))

                if layer.use_checkpoint:                          # reference :705-706
                    from torch.utils.checkpoint import checkpoint
                    x = checkpoint(block_fn, x, use_reentrant=False)
                else:
                    x = block_fn(x)
            if i in self.out_indices:
                n = getattr(self, f"norm{i}")
                o = AG.LayerNormFn.apply(x, n.weight, n.bias, n.eps, torch.float32)
                outs.append(o.view(B, H, W, C).permute(0, 3, 1, 2).contiguous())
            if layer.downsample is not None:
                d = layer.downsample
                xm = AG.PatchMergeLayerNormFn.apply(x, d.norm.weight, d.norm.bias, H, W, d.norm.eps, cd)
                x = AG.LinearFn.apply(xm, d.reduction.weight, None, wc(d.reduction.weight), torch.float32)
                H, W = (H + 1) // 2, (W + 1) // 2
        return tuple(outs)

    def _pitch_block(self, blk, x, xn, B, H, W, C, cd, rd):
        """PitchAttentionModule.forward (:1143-1209) in planar mode: un-shifted windows attend to themselves with
        separate q / k / v linears (one GEMM on the concatenated weights), beta-only bias, no mask.  The reference takes
        its shortcut as a view BEFORE writing norm1 back in place (:1163-1164), so the residual is norm1(x): reproduced."""
        if self.pano_mode:
            raise NotImplementedError(
                "PitchAttentionModule (odd stage depth) cannot run in pano mode: the reference itself fails there "
                "(simple_panoswin_transformer.py:1038 passes with_uv=True to lzx.pano_rotate.pano_rotate_image, which "
                "lzx/pano_rotate.py:169 does not accept), so there is no behaviour to match; use even depths (every "
                "shipped config does) or planar mode")
        ws = blk.window_size[0]
        ver = tuple((p._version, p.data_ptr()) for l in (blk.q_linear, blk.k_linear, blk.v_linear) for p in l.parameters())
        hit = self._weight_cache.get(("pam_qkv", id(blk), cd))
        if hit is None or hit[0] != ver:
            w = torch.cat([blk.q_linear.weight, blk.k_linear.weight, blk.v_linear.weight], 0).detach()
            b = None
            if blk.q_linear.bias is not None:
                b = torch.cat([blk.q_linear.bias, blk.k_linear.bias, blk.v_linear.bias], 0).detach().float().contiguous()
            hit = (ver, ops.cast(w.contiguous(), cd) if cd != w.dtype else w.contiguous(), b)
            self._weight_cache[("pam_qkv", id(blk), cd)] = hit
        _, wqkv, bqkv = hit
        xr = ops.layernorm(x, self._f(blk.norm1.weight), self._f(blk.norm1.bias), blk.norm1.eps, rd)     # the new residual
        xn = xr if rd == cd else ops.cast(xr, cd)
        qkv = ops.linear(xn, wqkv, bqkv)
        if cd == torch.bfloat16 and ops.window_attention_full_supported(ws, C // blk.num_heads):
            k = ("bfull_pam", id(blk), H, W, ws)
            al, be = blk.sphere_position_alpha_table_Te, blk.sphere_position_beta_table_Te
            ver = (al._version, be._version, al.data_ptr(), be.data_ptr())
            bf = self._cur_res.get(k)
            if bf is None or bf[0] != ver:
                bf = (ver, ops.window_bias_full(self._f(al), self._f(be), None, None, H, W, ws, 0, False))
                self._cur_res[k] = bf
            att = ops.window_attention_full(qkv.view(B, H, W, 3 * C), bf[1], bqkv, blk.num_heads, ws, 0, False, blk.scale)
        else:
            att = ops.window_attention(qkv.view(B, H, W, 3 * C), self._f(blk.sphere_position_alpha_table_Te),
                                       self._f(blk.sphere_position_beta_table_Te), bqkv, None, None, blk.num_heads, ws, 0,
                                       False, blk.scale)
        x = ops.linear(att.view(B, H * W, C), self._w(blk.proj.weight, cd), self._f(blk.proj.bias), residual=xr, out=xr)
        xn2 = ops.layernorm(x, self._f(blk.norm2.weight), self._f(blk.norm2.bias), blk.norm2.eps, cd)
        hid = ops.linear(xn2, self._w(blk.mlp.fc1.weight, cd), self._f(blk.mlp.fc1.bias), gelu=True)
        return ops.linear(hid, self._w(blk.mlp.fc2.weight, cd), self._f(blk.mlp.fc2.bias), residual=x, out=x)

    def _forward_tokens(self, img: torch.Tensor, on_output=None, on_block=None) -> Tuple[torch.Tensor, ...]:
        cd = self._compute_dtype
        rd = torch.float32 if cd == torch.float32 else self._residual_dtype      # residual-stream storage
        dev = img.device
        ws = self.window_size
        tok = self._stem(img)                                     # [B, Hs, Ws, E] in the compute dtype
        B, Hs, Ws, E = tok.shape
        self._enter_resolution(Hs, Ws, dev)
        pos = self._abs_position(Hs, Ws, dev) if (self.pano_mode and self.ape) else None
        xn_first = None                                       # norm1 of the very first block when the stem LN produces it
        if self.patch_embed.norm is not None and cd == torch.bfloat16 and rd == torch.float32 and len(self.layers[0].blocks) > 0:
            # patch_norm (+ position add) and the first block's norm1 in one pass over the rows
            n, n1 = self.patch_embed.norm, self.layers[0].blocks[0].norm1
            x, xn_first = ops.layernorm2(tok.view(B, Hs * Ws, E), self._f(n.weight), self._f(n.bias), n.eps, pos,
                                         self._f(n1.weight), self._f(n1.bias), n1.eps)
        elif self.patch_embed.norm is not None:
            n = self.patch_embed.norm
            x = ops.layernorm(tok.view(B, Hs * Ws, E), self._f(n.weight), self._f(n.bias), n.eps, rd, pos)
        else:
            x = tok.view(B, Hs * Ws, E).float()
            if pos is not None:
                x = x + pos[None]
            x = x.to(rd).contiguous()
        H, W = Hs, Ws
        outs = []
        n_block = 0
        for i, layer in enumerate(self.layers):
            C = self.num_features[i]
            uv = self._const(("uv", H, W), lambda: make_uv_hw2(H, W), dev) if self.pano_mode else None
            # LayerNorm fused into the producing GEMM's epilogue where one tile holds complete rows (C <= 256)
            fuse_ln = cd == torch.bfloat16 and rd == torch.float32 and C % 32 == 0 and C <= 256
            xn = xn_first if i == 0 else None                 # norm1(x) of the current block when already computed
            stage_map = None                                  # the stage's NCHW output when the last fc2 produced it
            for j, blk in enumerate(layer.blocks):
                if isinstance(blk, PitchAttentionModule):
                    x = self._pitch_block(blk, x, xn, B, H, W, C, cd, rd)
                    xn = None
                    if on_block is not None:
                        on_block(n_block, x)
                    n_block += 1
                    continue
                a = blk.attn
                shift = blk.shift_size
                mask = None
                if not self.pano_mode and shift > 0:
                    mask = self._const(("mask", H, W, ws, shift), lambda: planar_attention_mask(H, W, ws, shift), dev)
                if xn is None:
                    xn = ops.layernorm(x, self._f(blk.norm1.weight), self._f(blk.norm1.bias), blk.norm1.eps, cd)
                qkv = ops.linear(xn, self._w(a.qkv.weight, cd), self._f(a.qkv.bias))
                if cd == torch.bfloat16 and ops.window_attention_full_supported(ws, C // a.num_heads):
                    # tcgen05 kernel, all additive logit terms precomputed
                    att = ops.window_attention_full(qkv.view(B, H, W, 3 * C), self._bias_full(a, uv, mask, H, W, ws, shift),
                                                    self._f(a.qkv.bias), a.num_heads, ws, shift, self.pano_mode, a.scale)
                else:
                    # generic CUDA-core kernel: the fp32 parity path, and bf16 for other window sizes / head dims
                    att = ops.window_attention(qkv.view(B, H, W, 3 * C), self._f(a.sphere_position_alpha_table_Te),
                                               self._f(a.sphere_position_beta_table_Te), self._f(a.qkv.bias), uv, mask,
                                               a.num_heads, ws, shift, self.pano_mode, a.scale)
                if fuse_ln:                                   # proj + shortcut -> norm2 in one kernel
                    x, xn2 = ops.linear_layernorm(att.view(B, H * W, C), self._w(a.proj.weight, cd), self._f(a.proj.bias), x,
                                                  self._f(blk.norm2.weight), self._f(blk.norm2.bias), blk.norm2.eps, out=x)
                else:
                    x = ops.linear(att.view(B, H * W, C), self._w(a.proj.weight, cd), self._f(a.proj.bias), residual=x, out=x)
                    xn2 = ops.layernorm(x, self._f(blk.norm2.weight), self._f(blk.norm2.bias), blk.norm2.eps, cd)
                nxt = layer.blocks[j + 1] if j + 1 < len(layer.blocks) else None
                if fuse_ln and C == 96 and blk.mlp.fc1.out_features == 384:
                    # fc1 + GELU + fc2 + shortcut in one kernel: the 4C hidden activation never leaves the SM
                    x = ops.mlp_fused(xn2, self._w(blk.mlp.fc1.weight, cd), self._f(blk.mlp.fc1.bias),
                                      self._w(blk.mlp.fc2.weight, cd), self._f(blk.mlp.fc2.bias), x)
                    xn = None
                    if on_block is not None:
                        on_block(n_block, x)
                    n_block += 1
                    continue
                hid = ops.linear(xn2, self._w(blk.mlp.fc1.weight, cd), self._f(blk.mlp.fc1.bias), gelu=True)
                if fuse_ln and nxt is not None:               # fc2 + shortcut -> the next block's norm1
                    x, xn = ops.linear_layernorm(hid, self._w(blk.mlp.fc2.weight, cd), self._f(blk.mlp.fc2.bias), x,
                                                 self._f(nxt.norm1.weight), self._f(nxt.norm1.bias), nxt.norm1.eps, out=x)
                elif fuse_ln and i in self.out_indices:       # last fc2 of the stage + shortcut -> the stage's output map
                    n = getattr(self, f"norm{i}")
                    x, stage_map = ops.linear_layernorm_nchw(hid, self._w(blk.mlp.fc2.weight, cd), self._f(blk.mlp.fc2.bias), x,
                                                             self._f(n.weight), self._f(n.bias), n.eps, H, W, out=x)
                    xn = None
                else:
                    x = ops.linear(hid, self._w(blk.mlp.fc2.weight, cd), self._f(blk.mlp.fc2.bias), residual=x, out=x)
                    xn = None
                if on_block is not None:
                    on_block(n_block, x)
                n_block += 1
            if i in self.out_indices:
                n = getattr(self, f"norm{i}")
                outs.append(stage_map if stage_map is not None else
                            ops.layernorm_nchw(x, self._f(n.weight), self._f(n.bias), H, W, n.eps))
                if on_output is not None:
                    on_output(len(outs) - 1, outs[-1])
            if layer.downsample is not None:
                d = layer.downsample
                xm = ops.patch_merge_layernorm(x, self._f(d.norm.weight), self._f(d.norm.bias), H, W, d.norm.eps, cd)
                x = ops.linear(xm, self._w(d.reduction.weight, cd), None, out_dtype=rd)
                H, W = (H + 1) // 2, (W + 1) // 2
        return tuple(outs)


def _load_checkpoint(model: nn.Module, path: str):
    """mmcv_custom.load_checkpoint when the reference tree is importable (it also resizes bias tables,
    mmcv_custom/checkpoint.py:286-356); otherwise a plain non-strict state_dict load with the same
    `state_dict` / `model` unwrapping and `module.` / `encoder.` prefix stripping."""
    try:
        from mmcv_custom import load_checkpoint            # type: ignore
        from mmdet.utils import get_root_logger            # type: ignore
        return load_checkpoint(model, path, strict=False, logger=get_root_logger())
    except ImportError:
        pass
    ckpt = torch.load(path, map_location="cpu")
    sd = ckpt.get("state_dict", ckpt.get("model", ckpt)) if isinstance(ckpt, dict) else ckpt
    clean = {}
    for k, v in sd.items():
        for prefix in ("module.", "encoder.", "backbone."):
            if k.startswith(prefix):
                k = k[len(prefix):]
        clean[k] = v
    return model.load_state_dict(clean, strict=False)
