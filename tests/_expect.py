"""Expected values for op-level GPU tests, built from the CPU oracle's primitives (fp32, CPU)."""
import torch
import torch.nn.functional as F

from oracle import panoswin_oracle as O


def attention_core(qkv, alpha, beta, qkv_bias, uv, H, W, heads, ws, shift, pano, scale):
    """What psw_window_attn_fwd must produce: qkv [B, H, W, 3C] (bias already applied on real tokens)
    -> [B, H, W, C].  Padding cells carry q/k/v = qkv_bias and uv = (0, 0) and take part as keys/values
    (reference simple_panoswin_transformer.py:486-491, :344-347, :507-519)."""
    B, _, _, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    S = H * W
    bias_row = (qkv_bias if qkv_bias is not None else torch.zeros(C3)).view(1, 1, C3)
    feat = torch.cat([qkv.reshape(B, S, C3), uv.reshape(1, S, 2).expand(B, S, 2)], -1)
    pad_row = torch.cat([bias_row, torch.zeros(1, 1, 2)], -1).expand(B, 1, C3 + 2)
    src_ext = torch.cat([feat, pad_row], 1)                             # index S = padding token
    if pano:
        src = O.pano_source_index(H, W, shift)
        mask = None
    else:
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        base = torch.full((Hp, Wp), -1, dtype=torch.int64)
        base[:H, :W] = torch.arange(S).view(H, W)
        ii = (torch.arange(Hp) + shift) % Hp
        jj = (torch.arange(Wp) + shift) % Wp
        src = base[ii][:, jj]
        mask = O.planar_shift_mask(H, W, ws, shift) if shift else None
    SH, SW = src.shape
    Hp, Wp = -(-SH // ws) * ws, -(-SW // ws) * ws
    full = torch.full((Hp, Wp), -1, dtype=torch.int64)
    full[:SH, :SW] = src
    flat = full.reshape(-1)
    gathered = src_ext[:, torch.where(flat < 0, S, flat)].view(B, Hp, Wp, C3 + 2)
    xw = O._windows(gathered, ws)                                       # [n, N, 3C+2]
    n, N, _ = xw.shape
    qkv_w = xw[..., :C3].view(n, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv_w[0] * scale, qkv_w[1], qkv_w[2]
    logits = q @ k.transpose(-1, -2)
    idx = O.relative_position_index(ws).reshape(-1)
    bt = beta[idx].view(N, N, heads)
    if pano:
        bias = O.haversine(xw[..., C3:])[..., None] * alpha[idx].view(N, N, heads)[None] + bt
    else:
        bias = bt[None]
    logits = logits + bias.permute(0, 3, 1, 2)
    if mask is not None:
        nW = mask.shape[0]
        logits = (logits.view(n // nW, nW, heads, N, N) + mask[None, :, None]).view(n, heads, N, N)
    o = (torch.softmax(logits, -1) @ v).transpose(1, 2).reshape(n, N, C)
    o_map = O._unwindows(o, ws, B, Hp, Wp).reshape(B, Hp * Wp, C)
    out = torch.zeros(B, S + 1, C, dtype=o_map.dtype)
    out[:, torch.where(flat < 0, S, flat)] = o_map
    return out[:, :S].view(B, H, W, C)


def rel_l2(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
