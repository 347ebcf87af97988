"""VAE decoder throughput (latents -> images) on one B200: libpdm's decoder (csrc/vae.cu) next to the reference's own
`FrozenAutoencoderKL.decode` in eager PyTorch on the same GPU (when baseline/_ref/ holds libs/autoencoder.py; staged by
__graft_entry__.build() in the dev container).  Random weights (the checkpoint is not in the tree).

    python tools/vae_bench.py [--batch 50] [--latent 32] [--iters 5]
Algorithmic FLOPs: 2 * MACs of every convolution / attention GEMM of the decoder (SD layout: ch 128, mult 1-2-4-4)."""
import argparse
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def decoder_flops(s, ch=128, mult=(1, 2, 4, 4), nrb=2):
    """per image, latent side s"""
    f = 0
    c = ch * mult[-1]
    side = s
    f += side * side * 9 * 4 * c                       # conv_in
    res = lambda cin, cout, px: px * (9 * cin * cout + 9 * cout * cout + (cin * cout if cin != cout else 0))  # noqa: E731
    f += 2 * res(c, c, side * side)                    # mid blocks
    f += side * side * 4 * c * c + 2 * (side * side) ** 2 * c   # attention: q, k, v, proj + QK^T + PV
    cin = c
    for lev in reversed(range(len(mult))):
        cout = ch * mult[lev]
        for _ in range(nrb + 1):
            f += res(cin, cout, side * side)
            cin = cout
        if lev:
            side *= 2
            f += side * side * 9 * cin * cin           # upsample conv
    f += side * side * 9 * cin * 3                     # conv_out
    return 2.0 * f


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=50)   # decode mini-batch of the reference (eval_t2i_discrete.py:75)
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    from panopticdiffusionmodels_b200.libs.autoencoder import get_model
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    vae = get_model(None, 0.23010).to(dev)
    z = torch.randn(a.batch, 4, a.latent, a.latent, device=dev)
    ms = timed(lambda: vae.decode(z, max_batch=a.batch), a.iters)
    fl = decoder_flops(a.latent) * a.batch
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    out = {"kernel": "vae_decode (libpdm)", "batch": a.batch, "latent": a.latent, "ms": round(ms, 2),
           "images_per_s": round(a.batch / ms * 1e3, 1), "tflops": round(fl / ms / 1e9, 1),
           "frac_of_bf16_peak": round(fl / ms / 1e9 / peaks.get("bf16_tflops", 1590.0), 3)}
    print(json.dumps(out), flush=True)
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.exists(os.path.join(ref_dir, "libs", "autoencoder.py")):
        sys.path.insert(0, ref_dir)
        import libs.autoencoder as ref_ae
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "ae.pth")
            torch.save(vae.state_dict(), p)
            ref = ref_ae.get_model(p, 0.23010).to(dev)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        nb = min(a.batch, 16)
        with torch.no_grad():
            ms32 = timed(lambda: ref.decode(z[:nb]), 2)
            with torch.autocast("cuda", dtype=torch.float16):
                ms16 = timed(lambda: ref.decode(z[:nb]), 2)
            err = float((ref.decode(z[:2]) - vae.decode(z[:2])).abs().max())
        print(json.dumps({"kernel": "vae_decode (reference, eager torch on this GPU)", "batch": nb,
                          "fp32_images_per_s": round(nb / ms32 * 1e3, 1), "fp16_autocast_images_per_s": round(nb / ms16 * 1e3, 1),
                          "max_abs_diff_vs_libpdm": err}), flush=True)


if __name__ == "__main__":
    main()
