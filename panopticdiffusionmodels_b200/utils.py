"""Boundary helpers mirroring the reference ``utils.py`` for the sampling path:
``get_nnet`` (utils.py:291-299), the analog-bit codec (utils.py:475-518) and ``amortize`` (utils.py:452-455)."""
from __future__ import annotations

import torch

from . import _lib


def get_nnet(name, **kwargs):
    if name == "uvit_t2i":
        from .libs.uvit_t2i import UViT
        return UViT(**kwargs)
    raise NotImplementedError(name)


def amortize(n_samples, batch_size):
    k, r = n_samples // batch_size, n_samples % batch_size
    return k * [batch_size] if r == 0 else k * [batch_size] + [r]


def int2bits(x: torch.Tensor, n: int = 8, out_dtype=None) -> torch.Tensor:
    """ids (b,1,h,w) -> bits (b,n,h,w) in {0,1}, MSB first (utils.py:475-488)."""
    if not x.is_cuda:
        raise RuntimeError("libpdm has no CPU path")
    b, c, h, w = x.shape
    assert c == 1
    ids = x.to(torch.int32).contiguous()
    out = torch.empty(b, n, h, w, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().pdm_int2bits(_lib.ptr(ids), _lib.ptr(out), b, n, h * w, _lib.current_stream()))
    out = (out + 1.0) * 0.5  # kernel emits analog bits in {-1,+1}
    return out.to(out_dtype) if out_dtype else out.to(torch.int32)


def bits2int(x: torch.Tensor, out_dtype=torch.int, n: int = 8, c: int = 1) -> torch.Tensor:
    """bits (b,n,h,w) (bool or {0,1}) -> ids (b,1,h,w) float on the CPU, like the reference (utils.py:490-518)."""
    if not x.is_cuda:
        raise RuntimeError("libpdm has no CPU path")
    b, nb, h, w = x.shape
    analog = (x.to(torch.float32) * 2.0 - 1.0).contiguous()  # >0 <=> bit set
    labels = torch.empty(b, h, w, device=x.device, dtype=torch.int32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().pdm_bits2int(_lib.ptr(analog), _lib.ptr(labels), b, n, h * w, _lib.current_stream()))
    return labels.unsqueeze(1).float().cpu()


def labels_from_pred_mask(pred_mask: torch.Tensor) -> torch.Tensor:
    """``bits2int(pred_mask > 0)`` (utils.py:596) without the round trip: int32 labels (b,h,w) on the device."""
    b, n, h, w = pred_mask.shape
    pm = pred_mask.to(torch.float32).contiguous()
    labels = torch.empty(b, h, w, device=pm.device, dtype=torch.int32)
    with torch.cuda.device(pm.device):
        _lib.check(_lib.lib().pdm_bits2int(_lib.ptr(pm), _lib.ptr(labels), b, n, h * w, _lib.current_stream()))
    return labels
