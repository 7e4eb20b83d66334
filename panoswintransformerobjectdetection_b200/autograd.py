"""torch.autograd bindings of the backward entry points of libpanoswin_b200 (training path).

The reference trains through plain torch autograd (mmdet/apis/train.py:91-99 wraps the detector in DDP and
mmdet/utils/optimizer.py:22-33 calls loss.backward()).  Here every op of the PanoSwin block that runs on a
libpanoswin_b200 kernel in the forward has a hand-written backward kernel behind the same C ABI; these Functions only
save what the kernels need (inputs, never attention probabilities) and route gradients.  They work in both compute
modes: fp32 tensors (CUDA-core kernels, the gradient-parity path) and bf16 activations with fp32 parameters (tcgen05
forward GEMMs / attention, tcgen05 input gradients, fp32-accumulated weight gradients).
"""
from __future__ import annotations

import torch

from . import ops


class LayerNormFn(torch.autograd.Function):
    """y = LayerNorm(x) * gamma + beta, x fp32 or bf16, y in `out_dtype` (reference :504, :534, :768-772, :975-976)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, out_dtype):
        x = x.contiguous()
        ctx.save_for_backward(x, gamma)
        ctx.eps = eps
        return ops.layernorm(x, gamma.detach().contiguous(), beta.detach().contiguous(), eps, out_dtype)

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        need_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dg, db = ops.layernorm_bwd(x, dy.contiguous(), gamma.detach().contiguous(), ctx.eps, need_params=need_p)
        return dx, dg, db, None, None


class PatchMergeLayerNormFn(torch.autograd.Function):
    """PatchMerging front half: 2x2 gather + LayerNorm(4C) (reference :563-574); x [B, H*W, C] -> [B, H2*W2, 4C]."""

    @staticmethod
    def forward(ctx, x, gamma, beta, H, W, eps, out_dtype):
        x = x.contiguous()
        ctx.save_for_backward(x, gamma)
        ctx.args = (H, W, eps)
        return ops.patch_merge_layernorm(x, gamma.detach().contiguous(), beta.detach().contiguous(), H, W, eps, out_dtype)

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        H, W, eps = ctx.args
        need_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dg, db = ops.patch_merge_layernorm_bwd(x, dy.contiguous(), gamma.detach().contiguous(), H, W, eps, need_params=need_p)
        return dx, dg, db, None, None, None, None


class LinearFn(torch.autograd.Function):
    """y = x @ w.T + b in `out_dtype` (nn.Linear of qkv :287, proj :309, fc1 / fc2 :55-61, reduction :575).
    x is in the compute dtype; `w` / `b` are the fp32 parameters; `w_c` is the weight in the compute dtype (the
    parameter itself in fp32 mode, a cached bf16 copy otherwise) and carries no gradient of its own."""

    @staticmethod
    def forward(ctx, x, w, b, w_c, out_dtype):
        x = x.contiguous()
        ctx.save_for_backward(x, w_c)
        ctx.has_bias = b is not None
        ctx.x_dtype = x.dtype
        return ops.linear(x, w_c, None if b is None else b.detach().contiguous(), out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, dy):
        x, w_c = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != w_c.dtype:                           # fp32 gradient of an fp32-output GEMM on the bf16 path
            dy = ops.cast(dy, w_c.dtype)
        dx, dw, db = ops.linear_bwd(x, w_c, dy, need_dx=ctx.needs_input_grad[0], need_dw=ctx.needs_input_grad[1],
                                    need_db=ctx.has_bias and ctx.needs_input_grad[2], dx_dtype=ctx.x_dtype)
        return dx, dw, db, None, None


class GeluFn(torch.autograd.Function):
    """Exact (erf) GELU keeping the pre-activation (Mlp.act, reference :51, :57)."""

    @staticmethod
    def forward(ctx, h):
        h = h.contiguous()
        ctx.save_for_backward(h)
        return ops.gelu(h)

    @staticmethod
    def backward(ctx, dy):
        (h,) = ctx.saved_tensors
        return ops.gelu_bwd(h, dy.contiguous())


class BatchNormReluFn(torch.autograd.Function):
    """Train-mode BatchNorm2d + ReLU of the stem on a channels-last bf16 NCHW tensor (reference PatchEmbed.proj[1:3] /
    [4:6], :743-748, with batch statistics).  Updates the module's running statistics like nn.BatchNorm2d (momentum,
    unbiased variance) when they are given.  x [B, C, H, W] bf16 channels_last -> same."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps):
        xh = x.permute(0, 2, 3, 1)                           # NHWC view of the channels-last tensor
        if not xh.is_contiguous():
            xh = xh.contiguous()
        y, mean, rstd, var = ops.bn_relu_train(xh, gamma.detach().float().contiguous(), beta.detach().float().contiguous(), eps)
        if running_mean is not None:
            n = xh.numel() // xh.shape[-1]
            with torch.no_grad():
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                running_var.mul_(1 - momentum).add_((var * (n / max(n - 1, 1))).to(running_var.dtype), alpha=momentum)
        ctx.save_for_backward(xh, y, mean, rstd, gamma)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xh, y, mean, rstd, gamma = ctx.saved_tensors
        dyh = dy.permute(0, 2, 3, 1)
        if dyh.dtype != torch.bfloat16 or not dyh.is_contiguous():
            dyh = dyh.to(torch.bfloat16).contiguous()
        dx, dg, db = ops.bn_relu_bwd(xh, y, dyh, mean, rstd, gamma.detach().float().contiguous())
        return dx.permute(0, 3, 1, 2), dg.to(gamma.dtype), db.to(gamma.dtype), None, None, None, None


class WindowAttentionFn(torch.autograd.Function):
    """Fused pano / planar (shifted-)window attention core (reference :274-311 minus the two linears, with the shift,
    padding, partition, reverse and crop of :376-409, :473-519 folded in).  Saves qkv only; the backward kernel
    recomputes the probabilities.  The gradient returned for `qkv_bias` is the part that reaches the bias through the
    padding cells; autograd adds it to the bias gradient of the qkv linear."""

    @staticmethod
    def forward(ctx, qkv, alpha, beta, qkv_bias, uv, mask, heads, window, shift, pano_mode, scale):
        qkv = qkv.contiguous()
        a, b = alpha.detach().contiguous(), beta.detach().contiguous()
        qb = None if qkv_bias is None else qkv_bias.detach().contiguous()
        ctx.save_for_backward(qkv, a, b, qb, uv, mask)
        ctx.args = (heads, window, shift, pano_mode, scale)
        B, H, W, C3 = qkv.shape
        if qkv.dtype == torch.bfloat16 and ops.window_attention_full_supported(window, C3 // 3 // heads):
            table = ops.window_bias_full(a, b, uv, mask, H, W, window, shift, pano_mode)   # alpha / beta change every step
            return ops.window_attention_full(qkv, table, qb, heads, window, shift, pano_mode, scale)
        return ops.window_attention(qkv, a, b, qb, uv, mask, heads, window, shift, pano_mode, scale)

    @staticmethod
    def backward(ctx, dout):
        qkv, a, b, qb, uv, mask = ctx.saved_tensors
        heads, window, shift, pano_mode, scale = ctx.args
        dqkv, da, db, dqb = ops.window_attention_bwd(qkv, dout.contiguous(), a, b, qb, uv, mask, heads, window, shift,
                                                     pano_mode, scale)
        return dqkv, (da if pano_mode else None), db, dqb, None, None, None, None, None, None, None
