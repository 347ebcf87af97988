"""Stress the bf16 attention kernel against a torch fp64 reference; reports where mismatches sit (development tool)."""
import sys, os, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panopticdiffusionmodels_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.lib()
shapes = [(24, 590, 4), (40, 334, 8), (2, 130, 3), (3, 590, 2), (64, 590, 8), (2, 2126, 2), (3, 256, 3), (5, 128, 2)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for nb, L, H in shapes:
    g = torch.Generator().manual_seed(nb * L + H)
    qkv = torch.randn(nb, L, 3 * H * 64, generator=g).to(dev)
    src = qkv.bfloat16().float()
    q, k, v = src.reshape(nb, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(nb, L, H * 64)
    for r in range(reps):
        out = torch.full((nb, L, H * 64), float("nan"), device=dev)
        _lib.check(lib.pdm_debug_attention(_lib.ptr(qkv), _lib.ptr(out), nb, L, H, 0, 0, None, _lib.current_stream()))
        torch.cuda.synchronize()
        err = (out - ref).abs()
        err = torch.nan_to_num(err, nan=1e9)
        bad = err > 1.2e-2
        msg = ""
        if bad.any():
            idx = bad.nonzero()
            bs, ls, ds = idx[:, 0], idx[:, 1], idx[:, 2] // 64
            combos = sorted(set((int(b), int(h), int(l) // 128) for b, l, h in zip(bs, ls, ds)))
            msg = f" BAD n={int(bad.sum())} (b,h,qtile)={combos[:12]} rows={sorted(set(int(l) for l in ls))[:8]}"
        print(f"nb={nb} L={L} H={H} rep={r} maxerr={float(err.max()):.4g}{msg}", flush=True)
