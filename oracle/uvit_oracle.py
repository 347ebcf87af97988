"""ORACLE (test infrastructure, not product code): CPU restatement of the
reference network ``libs/uvit_t2i.py:UViT.forward``.

A *functional* re-statement: it takes a plain ``state_dict`` (the reference's
key layout) plus the constructor kwargs and evaluates the
forward with explicit tensor algebra in torch on the CPU (float32 by default,
float64 on request for an error yardstick).  No ``nn.Module`` is involved, so
it shares no code with either the reference or the product.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks it against
fixtures produced by executing the real reference (``tests/golden/make_golden.py``),
and, when ``/root/reference`` is present, against the reference directly.

Reference lines followed (``/root/reference``):
  timestep_embedding       libs/uvit_t2i.py:20-38
  unpatchify               libs/uvit_t2i.py:46-51
  Attention.forward        libs/uvit_t2i.py:66-92   (flash branch: fp32 SDPA, scale 1/sqrt(hd))
  Block._forward           libs/uvit_t2i.py:177-226 (panoptic branch dead: ``and False``)
  Mlp.forward              libs/timm.py:105-111     (GELU exact erf)
  PatchEmbed.forward       libs/uvit_t2i.py:237-244
  zeroconv.forward         libs/uvit_t2i.py:253-257
  UViT.forward             libs/uvit_t2i.py:378-525
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    # libs/uvit_t2i.py:30-38 -- frequencies are always built in float32.
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def _gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _patch_embed(img, w, b, p):
    # Conv2d(k=s=p) == per-patch dot product; weight (D, C, p, p): patch order (C, p1, p2).
    B, C, H, W = img.shape
    h, wd = H // p, W // p
    patches = img.reshape(B, C, h, p, wd, p).permute(0, 2, 4, 1, 3, 5).reshape(B, h * wd, C * p * p)
    return patches @ w.reshape(w.shape[0], -1).t() + b


def _unpatchify(x, channels):
    # 'B (h w) (p1 p2 C) -> B C (h p1) (w p2)'   (libs/uvit_t2i.py:50)
    B, P, F_ = x.shape
    p = int((F_ // channels) ** 0.5)
    h = w = int(P ** 0.5)
    assert h * w == P and p * p * channels == F_
    x = x.reshape(B, h, w, p, p, channels).permute(0, 5, 1, 3, 2, 4)
    return x.reshape(B, channels, h * p, w * p)


def _attention(x, sd, pre, num_heads):
    B, L, C = x.shape
    hd = C // num_heads
    qkv = _linear(x, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias"))
    qkv = qkv.reshape(B, L, 3, num_heads, hd).permute(2, 0, 3, 1, 4)  # K B H L D
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, L, C)
    return _linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def _block(x, sd, pre, num_heads, skip=None):
    if (pre + "skip_linear.weight") in sd:
        x = _linear(torch.cat([x, skip], dim=-1), sd[pre + "skip_linear.weight"], sd[pre + "skip_linear.bias"])
    x = x + _attention(_layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]), sd, pre + "attn.", num_heads)
    h = _layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    h = _gelu_erf(_linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
    return x + _linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def uvit_forward(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, timesteps: torch.Tensor,
                 context: torch.Tensor, mask_token: Optional[torch.Tensor] = None,
                 dtype: torch.dtype = torch.float32, use_ground_truth: bool = False):
    """Evaluate the network.  ``cfg`` holds the reference ctor kwargs
    (img_size, patch_size, in_chans, embed_dim, depth, num_heads, num_clip_token,
    num_panoptic_class, separate).  Returns ``noise`` or ``(noise, y)``."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    D = cfg["embed_dim"]
    depth = cfg["depth"]
    H = cfg["num_heads"]
    p = cfg["patch_size"]
    in_ch = cfg["in_chans"]
    ncls = cfg.get("num_panoptic_class", 8)
    separate = bool(cfg.get("separate", False))
    extras = 1 + cfg.get("num_clip_token", 77)

    x = _patch_embed(x.to(dtype), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], p)
    L = x.shape[1]
    time_token = timestep_embedding(timesteps, D).to(dtype).unsqueeze(1)
    ctx_token = _linear(context.to(dtype), sd["context_embed.weight"], sd["context_embed.bias"])
    two = separate and mask_token is not None
    m = None
    if mask_token is not None:
        me = _patch_embed(mask_token.to(dtype), sd["mask_embed.proj.weight"], sd["mask_embed.proj.bias"], p)
        if not separate:
            x = torch.cat((time_token, ctx_token, x, me), dim=1) + sd["pos_embed"]
        else:
            x = torch.cat((time_token, ctx_token, x), dim=1) + sd["pos_embed"]
            m = me + sd["pos_embed_mask"]
    else:
        x = torch.cat((time_token, ctx_token, x), dim=1) + sd["pos_embed"][:, :extras + L, :]

    skips, skips_m = [], []
    li = 0
    for i in range(depth // 2):
        if two:
            mx = torch.cat((x, m), dim=1)
        x = _block(x, sd, f"in_blocks.{i}.", H)
        if two:
            mx = _block(mx, sd, f"in_blocks_mask.{i}.", H)
            xa, m = mx[:, :extras + L], mx[:, extras + L:]
            zc = f"zero_convs.{2 * li + 1}.conv."
            x = x + _linear(xa, sd[zc + "weight"][:, :, 0], sd[zc + "bias"])
            skips_m.append(mx)
        skips.append(x)
        li += 1
    if two:
        mx = torch.cat((x, m), dim=1)
    x = _block(x, sd, "mid_block.", H)
    if two:
        mx = _block(mx, sd, "mid_block_mask.", H)
        xa, m = mx[:, :extras + L], mx[:, extras + L:]
        zc = f"zero_convs.{2 * li + 1}.conv."
        x = x + _linear(xa, sd[zc + "weight"][:, :, 0], sd[zc + "bias"])
        li += 1
    for j in range(depth // 2):
        if two:
            mx = torch.cat((x, m), dim=1)
        x = _block(x, sd, f"out_blocks.{j}.", H, skips.pop())
        if two:
            mx = _block(mx, sd, f"out_blocks_mask.{li - 1 - depth // 2}.", H, skips_m.pop())
            xa, m = mx[:, :extras + L], mx[:, extras + L:]
            zc = f"zero_convs.{2 * li + 1}.conv."
            x = x + _linear(xa, sd[zc + "weight"][:, :, 0], sd[zc + "bias"])
        li += 1
    x = _layer_norm(x, sd["norm.weight"], sd["norm.bias"])

    y = None
    if mask_token is not None and use_ground_truth:
        # libs/uvit_t2i.py:486-496: decode image feature + mask feature, hand the given mask back as the "prediction"
        mask_feature = x[:, extras + L:] if not separate else m
        noise = _linear(x[:, extras:extras + L] + mask_feature, sd["decoder_pred.weight"], sd["decoder_pred.bias"])
        y = mask_token
    elif mask_token is not None:
        if not separate:
            noise = _linear(x[:, extras:extras + L], sd["decoder_pred.weight"], sd["decoder_pred.bias"])
            y = _linear(x[:, extras + L:], sd["decoder_pred_mask.weight"], sd["decoder_pred_mask.bias"])
        else:
            noise = _linear(x[:, extras:], sd["decoder_pred.weight"], sd["decoder_pred.bias"])
            y = _linear(m, sd["decoder_pred_mask.weight"], sd["decoder_pred_mask.bias"])  # un-normed (uvit_t2i.py:507)
        y = _unpatchify(y, ncls)
        y = torch.tanh(F.conv2d(y, sd["final_layer_mask.weight"], sd["final_layer_mask.bias"], padding=1))
    else:
        noise = _linear(x[:, extras:extras + L], sd["decoder_pred.weight"], sd["decoder_pred.bias"])
    noise = _unpatchify(noise, in_ch)
    noise = F.conv2d(noise, sd["final_layer.weight"], sd["final_layer.bias"], padding=1)
    return noise if y is None else (noise, y)
