"""BASELINE config 1 as a golden fixture: the REAL reference (``/root/reference``, CPU, fp32) runs the joint
image+mask sample of ``mscoco_uvit_small`` (U-ViT-S/2 as shipped: separate=True) -- random-init weights, batch 4,
DPM-Solver++ 'fast' order 3, 20 NFE, CFG scale 2.0 -- and the final latents / mask prediction are stored.

    python tests/golden/make_config1.py          (dev container only; ~2-3 minutes of CPU)

The weights are NOT stored: both sides rebuild them deterministically (``panopticdiffusionmodels_b200`` UViT
constructed on the CPU under ``torch.manual_seed(1234)``, zero-conv bridges randomised), the generator loads that
state_dict into the reference model with ``strict=True``.  Inputs come from ``torch.Generator().manual_seed(1234)``.
Wiring = train_t2i_discrete.py:387-439, :480-546 (same harness as make_golden.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def build_ours():
    """State dict + ctor kwargs, reproducible anywhere (CPU RNG only)."""
    from panopticdiffusionmodels_b200 import configs
    from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT
    kw = dict(configs.get_config("mscoco_uvit_small").nnet)
    kw.pop("name")
    torch.manual_seed(1234)
    net = UViT(**kw)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.startswith("zero_convs"):
                p.copy_(torch.randn_like(p) * 0.02)
    return net, kw


def inputs(B=4):
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, 4, 32, 32, generator=g)
    m = torch.randn(B, 8, 32, 32, generator=g)
    ctx = torch.randn(B, 77, 768, generator=g)
    empty = torch.randn(77, 768, generator=g)
    return x, m, ctx, empty


def main():
    from make_golden import import_reference, sd_betas
    net_ours, kw = build_ours()
    sd = {k: v.clone() for k, v in net_ours.state_dict().items()}
    UViT, dpm = import_reference()
    kw_ref = {k: v for k, v in kw.items() if k != "patch_factor"}
    ref = UViT(**kw_ref).eval()
    ref.load_state_dict(sd, strict=True)
    x, m, ctx, empty = inputs()
    scale, steps = 2.0, 20
    ns = dpm.NoiseScheduleVP("discrete", betas=sd_betas())

    def model_fn(xx, t_cont, panoptic=None, mask_token=None, use_ground_truth=False, enable_panoptic=False):
        tt = t_cont * 1000
        ec = empty.unsqueeze(0).expand(xx.shape[0], -1, -1)
        c, pc = ref(xx, tt, context=ctx, mask_token=mask_token)
        u, pu = ref(xx, tt, context=ec, mask_token=mask_token)
        return c + scale * (c - u), pc + scale * (pc - pu)

    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        z, pm = dpm.DPM_Solver(model_fn, ns, predict_x0=True, thresholding=False).sample(
            x.clone(), steps=steps, eps=1e-3, T=1.0, order=3, mask_token=m.clone(), enable_mask_opt=True, enable_panoptic=True)
    np.savez_compressed(os.path.join(HERE, "config1_small_two.npz"), z=z.numpy(), pm=pm.numpy(), scale=np.float32(scale),
                        steps=np.int32(steps))
    print("written", z.shape, pm.shape, float(z.abs().max()), float(pm.abs().max()))


if __name__ == "__main__":
    main()
