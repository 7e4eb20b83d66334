"""CPU oracle for the PanoSwin pano-style shifted-window attention path.

TEST INFRASTRUCTURE ONLY.  This file is a from-scratch fp32 CPU restatement of the reference
algorithm; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product (`panoswintransformerobjectdetection_b200`) never does.

Parity status: PINNED.  `oracle/make_golden.py` executes the unmodified reference
(`/root/reference/mmdet/models/backbones/simple_panoswin_transformer.py`, loaded by
`oracle/ref_loader.py`) in the build container and stores its outputs in `tests/golden/`;
`tests/test_oracle_golden.py` checks this file against those vectors and against the reference's
five in-file known answers.

Style: purely functional, driven by a reference-compatible ``state_dict`` (same key names as the
reference module), closed-form gather indices instead of roll/flip/cat chains.  All citations are
relative to /root/reference/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------
def make_config(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=7,
                mlp_ratio=4.0, patch_size=4, in_chans=3, ape=True, patch_norm=True, qkv_bias=True,
                qk_scale=None, out_indices=(0, 1, 2, 3), pano_mode=True) -> dict:
    """Hyper-parameters of SimplePanoSwinTransformer.__init__
    (mmdet/models/backbones/simple_panoswin_transformer.py:781-801)."""
    return dict(embed_dim=embed_dim, depths=tuple(depths), num_heads=tuple(num_heads),
                window_size=window_size, mlp_ratio=mlp_ratio, patch_size=patch_size,
                in_chans=in_chans, ape=ape, patch_norm=patch_norm, qkv_bias=qkv_bias,
                qk_scale=qk_scale, out_indices=tuple(out_indices), pano_mode=pano_mode)


PANOSWIN_T = make_config()
PANOSWIN_B = make_config(embed_dim=128, depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32))


# ----------------------------------------------------------------------------------------------
# small index / coordinate helpers
# ----------------------------------------------------------------------------------------------
def relative_position_index(ws: int) -> Tensor:
    """idx[i, j] = (row_i - row_j + ws-1) * (2 ws - 1) + (col_i - col_j + ws-1), int64 [ws², ws²].
    Follows make_relative_position_index (simple_panoswin_transformer.py:95-129)."""
    t = torch.arange(ws * ws)
    r, c = t // ws, t % ws
    return (r[:, None] - r[None, :] + ws - 1) * (2 * ws - 1) + (c[:, None] - c[None, :] + ws - 1)


def uv_grid(H: int, W: int) -> Tensor:
    """Equirectangular (u, v) centre of every token, fp32 [H, W, 2].

    Follows make_uv_hw2 (simple_panoswin_transformer.py:153-189): gap = pi / H for BOTH axes,
    value = fl(fl(index * gap) - offset) + gap / 2, evaluated in fp32 in that order."""
    if W < H:
        raise ValueError("make_uv_hw2 requires W >= H (reference :174-175)")
    gap = math.pi / H
    col = torch.arange(W, dtype=torch.int64)
    row = torch.arange(H, dtype=torch.int64)
    u = (col * gap) - math.pi             # int64 * python float -> fp32 product, as in the reference
    v = (row * gap) - math.pi * 0.5
    u = u + 0.5 * gap
    v = v + 0.5 * gap
    assert u.dtype == torch.float32
    return torch.stack([u[None, :].expand(H, W), v[:, None].expand(H, W)], dim=-1).contiguous()


def haversine(uv: Tensor) -> Tensor:
    """Pairwise great-circle distance on the unit sphere, [..., N, 2] -> [..., N, N].
    Follows haversine22 (lzx/models/great_circle.py:71-86) with uv1 = uv2 = uv:
    d[i, j] = 2 asin( sqrt( sin²(|v_j - v_i| / 2) + cos v_j cos v_i sin²((u_j - u_i) / 2) ) )."""
    u, v = uv[..., 0], uv[..., 1]
    dv = v[..., None, :] - v[..., :, None]
    du = u[..., None, :] - u[..., :, None]
    a = torch.sin(0.5 * dv.abs()) ** 2 + torch.cos(v)[..., None, :] * torch.cos(v)[..., :, None] * \
        torch.sin(0.5 * du) ** 2
    return torch.arcsin(a ** 0.5) * 2


def pano_source_index(H: int, W: int, shift: int) -> Tensor:
    """Closed-form gather map of the forward pano shift.

    Returns int64 [2H, ceil(W/2)]: flat source index h*W + w of the token that lands at each cell
    of the north-south layout, or -1 where the cell is the zero column added for odd W.
    Restates WindowTransition.forward(reverse=False) in pano mode
    (simple_panoswin_transformer.py:399-406 with ew2ns :337-353):
      roll(+s) along W  ->  [pad odd W] split halves, right half flipped in both axes and stacked
      above the left half  ->  roll(+s) along the new (2H) axis."""
    We = W + (W & 1)
    half = We // 2
    i = torch.arange(2 * H)[:, None].expand(2 * H, half)
    j = torch.arange(half)[None, :].expand(2 * H, half)
    i0 = (i - shift) % (2 * H)                 # undo the vertical roll
    top = i0 < H                               # rows that came from the flipped right half
    h = torch.where(top, H - 1 - i0, i0 - H)
    w1 = torch.where(top, We - 1 - j, j)       # column in the rolled (and odd-padded) map
    w = (w1 - shift) % W                       # undo the horizontal roll
    return torch.where(w1 >= W, torch.full_like(h, -1), h * W + w)


def planar_shift_mask(H: int, W: int, ws: int, shift: int) -> Tensor:
    """Planar-mode SW-MSA mask, fp32 [nW, ws², ws²] with values 0 / -100.
    Follows BasicLayer._get_attention_mask (simple_panoswin_transformer.py:664-688)."""
    Hp = -(-H // ws) * ws
    Wp = -(-W // ws) * ws
    def band(n):
        r = torch.zeros(n, dtype=torch.int64)
        r[n - ws:n - shift] = 1
        r[n - shift:] = 2
        return r
    region = (band(Hp)[:, None] * 3 + band(Wp)[None, :]).to(torch.float32)
    reg_w = region.view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = reg_w[:, None, :] - reg_w[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


# ----------------------------------------------------------------------------------------------
# block pieces
# ----------------------------------------------------------------------------------------------
def _windows(x: Tensor, ws: int) -> Tensor:
    """[B, Hp, Wp, C] -> [B * nW, ws², C]; windows row-major per image, batch-major
    (window_partition, simple_panoswin_transformer.py:64-75)."""
    B, Hp, Wp, C = x.shape
    return x.view(B, Hp // ws, ws, Wp // ws, ws, C).transpose(2, 3).reshape(-1, ws * ws, C)


def _unwindows(xw: Tensor, ws: int, B: int, Hp: int, Wp: int) -> Tensor:
    """Inverse of `_windows` (window_reverse, simple_panoswin_transformer.py:78-92)."""
    C = xw.shape[-1]
    return xw.view(B, Hp // ws, Wp // ws, ws, ws, C).transpose(2, 3).reshape(B, Hp, Wp, C)


def window_attention(xw: Tensor, uvw: Tensor, p: Dict[str, Tensor], prefix: str, heads: int, ws: int,
                     scale: float, pano: bool, mask: Tensor | None) -> Tensor:
    """W-MSA on partitioned windows (BasicWindowAttention.forward,
    simple_panoswin_transformer.py:274-311; bias from _sphere_bias :241-260).

    xw  [n, N, c] features, uvw [n, N, 2] token coordinates, returns [n, N, c]."""
    n, N, c = xw.shape
    hd = c // heads
    qkv = F.linear(xw, p[prefix + "qkv.weight"], p.get(prefix + "qkv.bias"))
    qkv = qkv.view(n, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    logits = q @ k.transpose(-1, -2)                                     # [n, heads, N, N]
    idx = relative_position_index(ws).reshape(-1)
    beta = p[prefix + "sphere_position_beta_table_Te"][idx].view(N, N, heads)
    if pano:
        alpha = p[prefix + "sphere_position_alpha_table_Te"][idx].view(N, N, heads)
        bias = haversine(uvw)[..., None] * alpha[None] + beta            # [n, N, N, heads]
    else:
        bias = beta[None]
    logits = logits + bias.permute(0, 3, 1, 2)
    if mask is not None:
        nW = mask.shape[0]
        logits = (logits.view(n // nW, nW, heads, N, N) + mask[None, :, None]).view(n, heads, N, N)
    prob = torch.softmax(logits, dim=-1)
    out = (prob @ v).transpose(1, 2).reshape(n, N, c)
    return F.linear(out, p[prefix + "proj.weight"], p[prefix + "proj.bias"])


def attention_branch(xn: Tensor, uv_hw2: Tensor, p, prefix, H, W, heads, ws, shift, scale, pano) -> Tensor:
    """shift -> pad -> partition -> W-MSA -> reverse -> crop -> un-shift on LayerNorm'ed tokens.

    xn [B, H*W, c], uv_hw2 [H, W, 2] (zeros in planar mode).  Returns [B, H*W, c].
    Pano branch: PanoSwinTransformerBlock.forward (simple_panoswin_transformer.py:507-519);
    planar branch :520-528.  Padded cells carry zero features AND zero uv (:486-491, :344-347)."""
    B, S, c = xn.shape
    feat = torch.cat([xn, uv_hw2.reshape(1, S, 2).expand(B, S, 2)], dim=-1)       # [B, S, c+2]
    if pano:
        src = pano_source_index(H, W, shift)                                       # [2H, W'/2]
        SH, SW = src.shape
        flat = src.reshape(-1)
        zero_row = feat.new_zeros(B, 1, c + 2)
        gathered = torch.cat([feat, zero_row], dim=1)[:, torch.where(flat < 0, S, flat)]
        ns = gathered.view(B, SH, SW, c + 2)
        Hp, Wp = -(-SH // ws) * ws, -(-SW // ws) * ws
        ns = F.pad(ns, (0, 0, 0, Wp - SW, 0, Hp - SH))
        xw = _windows(ns, ws)
        yw = window_attention(xw[..., :c], xw[..., c:], p, prefix, heads, ws, scale, True, None)
        y_ns = _unwindows(yw, ws, B, Hp, Wp)[:, :SH, :SW].reshape(B, SH * SW, c)
        out = xn.new_zeros(B, S + 1, c)
        out[:, torch.where(flat < 0, S, flat)] = y_ns       # bijection on real cells; pads -> slot S
        return out[:, :S]
    # planar Swin: pad, roll(-s, -s), mask, roll(+s, +s), crop
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    img = F.pad(feat.view(B, H, W, c + 2), (0, 0, 0, Wp - W, 0, Hp - H))
    ii = (torch.arange(Hp) + shift) % Hp
    jj = (torch.arange(Wp) + shift) % Wp
    shifted = img[:, ii][:, :, jj]
    mask = planar_shift_mask(H, W, ws, shift) if shift else None
    xw = _windows(shifted, ws)
    yw = window_attention(xw[..., :c], xw[..., c:], p, prefix, heads, ws, scale, False, mask)
    y = _unwindows(yw, ws, B, Hp, Wp)
    back = torch.empty_like(y)
    back[:, ii[:, None], jj[None, :]] = y
    return back[:, :H, :W].reshape(B, S, c)


def mlp(x: Tensor, p, prefix) -> Tensor:
    """fc2(GELU_erf(fc1(x))) (Mlp, simple_panoswin_transformer.py:44-61)."""
    h = F.gelu(F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def layer_norm(x: Tensor, p, prefix, eps=1e-5) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), p[prefix + "weight"], p[prefix + "bias"], eps)


def block(x: Tensor, uv_hw2: Tensor, p, prefix, H, W, heads, ws, shift, scale, pano) -> Tensor:
    """One PanoSwinTransformerBlock in eval mode (simple_panoswin_transformer.py:493-536)."""
    y = attention_branch(layer_norm(x, p, prefix + "norm1."), uv_hw2, p, prefix + "attn.",
                         H, W, heads, ws, shift, scale, pano)
    x = x + y
    return x + mlp(layer_norm(x, p, prefix + "norm2."), p, prefix + "mlp.")


def pitch_attention_block_planar(x: Tensor, p, prefix, H, W, heads, ws, scale) -> Tensor:
    """PitchAttentionModule.forward with pano_mode=False (simple_panoswin_transformer.py:1143-1209, _attention
    :1212-1237): the "rotated" map is the map itself (:1177-1179), so every un-shifted window attends to itself with
    separate q / k / v linears, the beta-only bias (:257-258) and no mask.  Reference quirk reproduced on purpose: the
    shortcut is a VIEW of the input taken before norm1 is written back in place (:1163-1164), so the residual added
    after the attention is norm1(x), not x."""
    B, S, c = x.shape
    xn = layer_norm(x, p, prefix + "norm1.")
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    xw = _windows(F.pad(xn.view(B, H, W, c), (0, 0, 0, Wp - W, 0, Hp - H)), ws)          # zero features on padded cells
    n, N, _ = xw.shape
    hd = c // heads
    def lin(name):
        return F.linear(xw, p[prefix + name + ".weight"], p.get(prefix + name + ".bias")).view(n, N, heads, hd).transpose(1, 2)
    q, k, v = lin("q_linear") * scale, lin("k_linear"), lin("v_linear")
    idx = relative_position_index(ws).reshape(-1)
    beta = p[prefix + "sphere_position_beta_table_Te"][idx].view(N, N, heads).permute(2, 0, 1)
    prob = torch.softmax(q @ k.transpose(-1, -2) + beta[None], dim=-1)
    yw = F.linear((prob @ v).transpose(1, 2).reshape(n, N, c), p[prefix + "proj.weight"], p[prefix + "proj.bias"])
    y = _unwindows(yw, ws, B, Hp, Wp)[:, :H, :W].reshape(B, S, c)
    x = xn + y                                               # the in-place norm1 made the shortcut norm1(x)
    return x + mlp(layer_norm(x, p, prefix + "norm2."), p, prefix + "mlp.")


def patch_merging(x: Tensor, p, prefix, H, W) -> Tensor:
    """2x2 gather (order: (0,0), (1,0), (0,1), (1,1)) -> LN(4c) -> Linear(4c -> 2c, no bias)
    (PatchMerging.forward, simple_panoswin_transformer.py:551-576)."""
    B, S, c = x.shape
    img = F.pad(x.view(B, H, W, c), (0, 0, 0, W % 2, 0, H % 2))
    quad = torch.cat([img[:, 0::2, 0::2], img[:, 1::2, 0::2], img[:, 0::2, 1::2], img[:, 1::2, 1::2]], -1)
    quad = quad.reshape(B, -1, 4 * c)
    return F.linear(layer_norm(quad, p, prefix + "norm."), p[prefix + "reduction.weight"])


def stem(img: Tensor, p, cfg) -> Tensor:
    """PatchEmbed in eval mode (simple_panoswin_transformer.py:727-773): pad to a multiple of the
    patch size, conv3x3-BN-ReLU, conv3x3-BN-ReLU, conv(patch, stride patch), LayerNorm over channels.
    Returns [B, E, Hs, Ws]."""
    ps = cfg["patch_size"]
    _, _, H, W = img.shape
    img = F.pad(img, (0, (-W) % ps, 0, (-H) % ps))
    pre = "patch_embed.proj."
    def bn(x, i):
        return F.batch_norm(x, p[f"{pre}{i}.running_mean"], p[f"{pre}{i}.running_var"],
                            p[f"{pre}{i}.weight"], p[f"{pre}{i}.bias"], False, 0.0, 1e-5)
    x = F.relu(bn(F.conv2d(img, p[pre + "0.weight"], p[pre + "0.bias"], padding=1), 1))
    x = F.relu(bn(F.conv2d(x, p[pre + "3.weight"], p[pre + "3.bias"], padding=1), 4))
    x = F.conv2d(x, p[pre + "6.weight"], p[pre + "6.bias"], stride=ps)
    if cfg["patch_norm"]:
        x = layer_norm(x.permute(0, 2, 3, 1), p, "patch_embed.norm.").permute(0, 3, 1, 2)
    return x


def abs_position(p, H, W) -> Tensor:
    """Pano absolute position encoding Linear(5 -> E) on (x, y, z, u, v), [H, W, E]
    (SimplePanoSwinTransformer._pano_abs_position, simple_panoswin_transformer.py:925-934)."""
    uv = uv_grid(H, W)
    u, v = uv[..., 0], uv[..., 1]
    xyzuv = torch.stack([torch.sin(u) * torch.sin(v), torch.cos(u) * torch.sin(v), torch.cos(v), u, v], -1)
    return F.linear(xyzuv, p["abs_encoder.weight"], p["abs_encoder.bias"])


# ----------------------------------------------------------------------------------------------
# whole backbone
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def backbone_forward(p: Dict[str, Tensor], cfg: dict, img: Tensor, return_blocks: bool = False):
    """SimplePanoSwinTransformer.forward in eval mode (simple_panoswin_transformer.py:940-979).

    Returns the tuple of NCHW fp32 stage features; with ``return_blocks`` also the list of every
    block's output tokens [B, H*W, c] (uv channels never materialised here)."""
    pano = cfg["pano_mode"]
    ws = cfg["window_size"]
    if pano and not cfg["ape"]:
        raise AttributeError("pano_mode=True needs ape=True (reference :934 vs :841-842)")
    x = stem(img.float(), p, cfg)
    B, E, H, W = x.shape
    x = x.permute(0, 2, 3, 1)
    if pano and cfg["ape"]:
        x = x + abs_position(p, H, W)[None]
    x = x.reshape(B, H * W, E)
    outs, blocks = [], []
    n_layers = len(cfg["depths"])
    for li in range(n_layers):
        c = cfg["embed_dim"] * 2 ** li
        heads = cfg["num_heads"][li]
        scale = cfg["qk_scale"] or (c // heads) ** -0.5
        depth = cfg["depths"][li]
        if depth % 2 and pano:
            raise NotImplementedError("odd depths add PitchAttentionModule, which the reference cannot "
                                      "execute in pano mode (simple_panoswin_transformer.py:1038)")
        uv = uv_grid(H, W) if pano else torch.zeros(H, W, 2)
        for bi in range(depth - depth % 2):
            shift = 0 if bi % 2 == 0 else ws // 2
            x = block(x, uv, p, f"layers.{li}.blocks.{bi}.", H, W, heads, ws, shift, scale, pano)
            blocks.append(x)
        if depth % 2:                                        # trailing PitchAttentionModule (:636-647), planar mode
            x = pitch_attention_block_planar(x, p, f"layers.{li}.blocks.{depth - 1}.", H, W, heads, ws, scale)
            blocks.append(x)
        if li in cfg["out_indices"]:
            o = layer_norm(x, p, f"norm{li}.")
            outs.append(o.view(B, H, W, c).permute(0, 3, 1, 2).contiguous())
        if li < n_layers - 1:
            x = patch_merging(x, p, f"layers.{li}.downsample.", H, W)
            H, W = (H + 1) // 2, (W + 1) // 2
    return (tuple(outs), blocks) if return_blocks else tuple(outs)


# ----------------------------------------------------------------------------------------------
# deterministic, torch-version independent parameters and inputs
# ----------------------------------------------------------------------------------------------
def make_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """A full reference-compatible state_dict drawn from numpy's legacy RandomState (stable across
    numpy/torch versions).  Key names are those of the reference module (SURVEY.md §5 checkpoint
    row); `oracle/make_golden.py` loads it into the real reference with strict=True, which pins
    the names.  Biases, LayerNorm affine, alpha/beta tables (independent draws) and BatchNorm running
    statistics are all randomised so that padding tokens and BN folding are exercised
    (SURVEY.md §0.3, §8d config 1)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, Tensor] = {}
    def normal(shape, std, mean=0.0):
        return torch.from_numpy((rs.standard_normal(shape) * std + mean).astype(np.float32))
    def uniform(shape, lo, hi):
        return torch.from_numpy(rs.uniform(lo, hi, shape).astype(np.float32))
    def linear(name, fin, fout, bias=True):
        sd[name + ".weight"] = normal((fout, fin), 1.0 / math.sqrt(fin))
        if bias:
            sd[name + ".bias"] = normal((fout,), 0.1)
    def norm(name, n):
        sd[name + ".weight"] = normal((n,), 0.1, 1.0)
        sd[name + ".bias"] = normal((n,), 0.1)
    E, ws, ps = cfg["embed_dim"], cfg["window_size"], cfg["patch_size"]
    c1, c2 = E // 3, 2 * (E // 3)
    chans = [(cfg["in_chans"], c1, 3), (c1, c2, 3), (c2, E, ps)]
    for (cin, cout, k), ci, bi in zip(chans, (0, 3, 6), (1, 4, None)):
        sd[f"patch_embed.proj.{ci}.weight"] = normal((cout, cin, k, k), 1.0 / math.sqrt(cin * k * k))
        sd[f"patch_embed.proj.{ci}.bias"] = normal((cout,), 0.1)
        if bi is not None:
            sd[f"patch_embed.proj.{bi}.weight"] = normal((cout,), 0.1, 1.0)
            sd[f"patch_embed.proj.{bi}.bias"] = normal((cout,), 0.1)
            sd[f"patch_embed.proj.{bi}.running_mean"] = normal((cout,), 0.1)
            sd[f"patch_embed.proj.{bi}.running_var"] = uniform((cout,), 0.5, 1.5)
            sd[f"patch_embed.proj.{bi}.num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
    if cfg["patch_norm"]:
        norm("patch_embed.norm", E)
    if cfg["ape"]:
        linear("abs_encoder", 5, E)
    for li, (depth, heads) in enumerate(zip(cfg["depths"], cfg["num_heads"])):
        c = E * 2 ** li
        hidden = int(c * cfg["mlp_ratio"])
        if depth % 2:                                        # PitchAttentionModule parameters (:990-1022), last block
            pre = f"layers.{li}.blocks.{depth - 1}."
            sd[pre + "relative_position_index_OO"] = relative_position_index(ws)
            sd[pre + "np_uv"] = torch.tensor([1.0, -0.0001]) * math.pi
            linear(pre + "proj", c, c)
            sd[pre + "sphere_position_alpha_table_Te"] = normal(((2 * ws - 1) ** 2, heads), 0.5)
            sd[pre + "sphere_position_beta_table_Te"] = normal(((2 * ws - 1) ** 2, heads), 0.5)
            linear(pre + "mlp.fc1", c, hidden)
            linear(pre + "mlp.fc2", hidden, c)
            norm(pre + "norm2", c)
            norm(pre + "norm1", c)
            for nm in ("q_linear", "k_linear", "v_linear"):
                linear(pre + nm, c, c, bias=cfg["qkv_bias"])
        for bi in range(depth - depth % 2):
            pre = f"layers.{li}.blocks.{bi}."
            norm(pre + "norm1", c)
            sd[pre + "attn.relative_position_index_OO"] = relative_position_index(ws)
            linear(pre + "attn.proj", c, c)
            sd[pre + "attn.sphere_position_alpha_table_Te"] = normal(((2 * ws - 1) ** 2, heads), 0.5)
            sd[pre + "attn.sphere_position_beta_table_Te"] = normal(((2 * ws - 1) ** 2, heads), 0.5)
            linear(pre + "attn.qkv", c, 3 * c, bias=cfg["qkv_bias"])
            norm(pre + "norm2", c)
            linear(pre + "mlp.fc1", c, hidden)
            linear(pre + "mlp.fc2", hidden, c)
        if li < len(cfg["depths"]) - 1:
            linear(f"layers.{li}.downsample.reduction", 4 * c, 2 * c, bias=False)
            norm(f"layers.{li}.downsample.norm", 4 * c)
    for li in cfg["out_indices"]:
        norm(f"norm{li}", E * 2 ** li)
    return sd


def make_image(shape: Sequence[int], seed: int = 2, kind: str = "rand") -> Tensor:
    """Synthetic input image batch: 'rand' = U[0,1) like the reference's own smoke tests
    (simple_panoswin_transformer.py:1255, :1314); 'randn' = ImageNet-normalised range."""
    rs = np.random.RandomState(seed)
    a = rs.uniform(0.0, 1.0, shape) if kind == "rand" else rs.standard_normal(shape)
    return torch.from_numpy(a.astype(np.float32))
