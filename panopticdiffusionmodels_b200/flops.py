"""Algorithmic FLOP count of one U-ViT t2i forward; bookkeeping for bench.py's roofline numbers.

2 flop per multiply-accumulate; per sample per forward:
    F = (depth + 1) (24 L D^2 + 4 L^2 D) + (depth / 2) 4 L D^2 + F_io
with L = 78 + 2P (single-stream) or the block terms once for L1 = 78 + P and once for L2 = 78 + 2P plus the zero-conv
bridges (two-stream); F_io = patch embeds + context embed + decoders + the two 3x3 convs (< 0.2 %)."""
from __future__ import annotations


def flops_per_forward(cfg: dict, with_mask: bool = True) -> float:
    D, depth = cfg["embed_dim"], cfg["depth"]
    P = (cfg["img_size"] // cfg["patch_size"]) ** 2
    ext = 1 + cfg.get("num_clip_token", 77)
    clip = cfg.get("clip_dim", 768)

    def blocks(L):
        return (depth + 1) * (24 * L * D * D + 4 * L * L * D) + (depth // 2) * 4 * L * D * D

    f_io = 2 * P * D * (16 + 32) * 2 + 2 * (ext - 1) * clip * D + 18 * 4 * P * (16 + 64)
    if not with_mask:
        return blocks(ext + P) + f_io
    if cfg.get("separate", False):
        L1, L2 = ext + P, ext + 2 * P
        return blocks(L1) + blocks(L2) + (depth + 1) * 2 * L1 * D * D + f_io
    return blocks(ext + 2 * P) + f_io
