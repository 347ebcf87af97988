"""Per-kernel timing through the diagnostic C-ABI entry points (CUDA events inside libpdm around the
kernel alone).  Prints one JSON line per kernel: achieved TFLOP/s or GB/s against MEASURED_PEAKS.json.

    python tools/kernel_bench.py [--config small|mid|large|small_512] [--batch 256]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panopticdiffusionmodels_b200 import _lib  # noqa: E402

SHAPES = {"small": (512, 8, 590), "mid": (768, 12, 590), "large": (1024, 16, 590), "small_512": (512, 8, 2126)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--prec", type=int, default=0)
    ap.add_argument("--only", default="", help="comma list: gemm,layernorm,attention")
    ap.add_argument("--attn-nb", type=int, default=0, help="attention batch rows (default min(2*batch, 64))")
    ap.add_argument("--attn-L", type=int, default=0)
    a = ap.parse_args()
    D, H, L = SHAPES[a.config]
    M = 2 * a.batch * L
    dev = torch.device("cuda:0")
    hbm, tf, src = peaks()
    lib = _lib.lib()
    ms = C.c_float(0)
    s = _lib.current_stream()
    out = []

    def gemm(name, N, K, K2=0, gelu=0, resid=False, bytes_alg=None):
        A = torch.randn(M, K, device=dev)
        A2 = torch.randn(M, K2, device=dev) if K2 else None
        W = torch.randn(N, K + K2, device=dev) * 0.02
        b = torch.randn(N, device=dev)
        o = torch.empty(M, N, device=dev)
        r = torch.randn(M, N, device=dev) if resid else None
        _lib.check(lib.pdm_debug_linear(_lib.ptr(A), _lib.ptr(A2), _lib.ptr(W), _lib.ptr(b), _lib.ptr(r), _lib.ptr(o),
                                        M, N, K, K2, a.prec, gelu, a.iters, C.byref(ms), s))
        fl = 2.0 * M * N * (K + K2)
        out.append(dict(kernel=name, M=M, N=N, K=K + K2, ms=round(ms.value, 4), tflops=round(fl / ms.value / 1e9, 1),
                        frac_of_bf16_peak=round(fl / ms.value / 1e9 / tf, 3), peak_src=src))
        print(json.dumps(out[-1]), flush=True)
        del A, A2, W, o, r

    def ln_gemm(name, N, gelu):
        x = torch.randn(M, D, device=dev)
        gam, bet = torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev) * 0.1
        W = torch.randn(N, D, device=dev) * 0.02
        b = torch.randn(N, device=dev)
        o = torch.empty(M, N, device=dev)
        _lib.check(lib.pdm_debug_ln_chain(None, None, None, _lib.ptr(x), _lib.ptr(gam), _lib.ptr(bet), _lib.ptr(W),
                                          _lib.ptr(b), _lib.ptr(o), None, M, N, D, 0, gelu, a.iters, C.byref(ms), s))
        fl = 2.0 * M * N * D
        print(json.dumps(dict(kernel=name, M=M, N=N, K=D, ms=round(ms.value, 4), tflops=round(fl / ms.value / 1e9, 1),
                              frac_of_bf16_peak=round(fl / ms.value / 1e9 / tf, 3), peak_src=src)), flush=True)

    only = set(a.only.split(",")) if a.only else {"gemm", "layernorm", "attention"}
    if "gemm" in only:
        ln_gemm("LN+gemm_qkv(bf16 out)", 3 * D, 0)
        ln_gemm("LN+gemm_fc1(+gelu, bf16 out)", 4 * D, 1)
        gemm("gemm_proj(+resid, +bf16 copy, +row sums)", D, D, gelu=4, resid=True)
        gemm("gemm_fc2(+resid, +bf16 copy, +row sums)", D, 4 * D, gelu=4, resid=True)
        gemm("gemm_skip(+bf16 copy, +row sums)", D, D, K2=D, gelu=4)
    if "gemm" in only:
        gemm("gemm_qkv(bf16 out)", 3 * D, D, gelu=2)
        gemm("gemm_proj(+resid)", D, D, resid=True)
        gemm("gemm_fc1(+gelu, bf16 out)", 4 * D, D, gelu=3)
        gemm("gemm_fc2(+resid)", D, 4 * D, resid=True)
        gemm("gemm_skip", D, D, K2=D)

    if "layernorm" in only:
        x = torch.randn(M, D, device=dev)
        w = torch.randn(D, device=dev)
        o = torch.empty(M, D, device=dev)
        _lib.check(lib.pdm_debug_layernorm(_lib.ptr(x), _lib.ptr(w), _lib.ptr(w), _lib.ptr(o), M, D, a.prec, a.iters,
                                           C.byref(ms), s))
        by = M * D * (4 + (2 if a.prec == 0 else 4))
        print(json.dumps(dict(kernel="layernorm", rows=M, D=D, ms=round(ms.value, 4), gbs=round(by / ms.value / 1e6, 1),
                              frac_of_hbm_peak=round(by / ms.value / 1e6 / hbm, 3), peak_src=src)), flush=True)
        del x, o
    if "attention" not in only:
        return

    nb = a.attn_nb or min(2 * a.batch, 64)
    L = a.attn_L or L
    qkv = torch.randn(nb, L, 3 * D, device=dev)
    o = torch.empty(nb, L, D, device=dev)
    _lib.check(lib.pdm_debug_attention(_lib.ptr(qkv), _lib.ptr(o), nb, L, H, a.prec, max(1, a.iters // 3), C.byref(ms), s))
    fl = 4.0 * nb * H * L * L * 64
    print(json.dumps(dict(kernel="attention", nb=nb, L=L, H=H, ms=round(ms.value, 4), tflops=round(fl / ms.value / 1e9, 1),
                          frac_of_bf16_peak=round(fl / ms.value / 1e9 / tf, 3), peak_src=src)), flush=True)


if __name__ == "__main__":
    main()
