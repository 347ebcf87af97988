"""ORACLE (test infrastructure, not product code): CPU restatement of the reference VAE decode path
``FrozenAutoencoderKL.decode`` (``libs/autoencoder.py:446-450``) = ``post_quant_conv`` + ``Decoder.forward``
(``:376-409``) with ``ResnetBlock.forward`` (``:114-134``), ``AttnBlock.forward`` (``:171-195``), ``Upsample.forward``
(``:46-50``), ``Normalize`` = GroupNorm(32, eps 1e-6) (``:31-32``) and swish (``:26-28``).

Functional: takes the reference ``state_dict`` and the ``ddconfig``; no ``nn.Module``.  Parity status: PINNED --
``tests/test_oracle_golden.py::test_vae_decode_matches_reference`` checks it against a fixture produced by executing the
real reference (``tests/golden/make_vae.py``)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _gn(x, sd, pre):
    return F.group_norm(x, 32, sd[pre + "weight"], sd[pre + "bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(x, sd, pre, padding):
    return F.conv2d(x, sd[pre + "weight"], sd[pre + "bias"], stride=1, padding=padding)


def _resnet(x, sd, pre):
    h = _conv(_swish(_gn(x, sd, pre + "norm1.")), sd, pre + "conv1.", 1)
    h = _conv(_swish(_gn(h, sd, pre + "norm2.")), sd, pre + "conv2.", 1)
    if (pre + "nin_shortcut.weight") in sd:
        x = _conv(x, sd, pre + "nin_shortcut.", 0)
    return x + h


def _attn(x, sd, pre):
    h = _gn(x, sd, pre + "norm.")
    q, k, v = (_conv(h, sd, pre + n + ".", 0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, sd, pre + "proj_out.", 0)


def vae_decode(sd: Dict[str, torch.Tensor], ddconfig: dict, z: torch.Tensor, scale_factor: float) -> torch.Tensor:
    nlev = len(ddconfig["ch_mult"])
    z = (1.0 / scale_factor) * z
    z = _conv(z, sd, "post_quant_conv.", 0)
    h = _conv(z, sd, "decoder.conv_in.", 1)
    h = _resnet(h, sd, "decoder.mid.block_1.")
    h = _attn(h, sd, "decoder.mid.attn_1.")
    h = _resnet(h, sd, "decoder.mid.block_2.")
    for lev in reversed(range(nlev)):
        for i in range(ddconfig["num_res_blocks"] + 1):
            h = _resnet(h, sd, f"decoder.up.{lev}.block.{i}.")
        if lev != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, f"decoder.up.{lev}.upsample.conv.", 1)
    h = _swish(_gn(h, sd, "decoder.norm_out."))
    return _conv(h, sd, "decoder.conv_out.", 1)
