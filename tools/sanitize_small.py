"""Development aid: one small pass over every product kernel family (fp32 and bf16 forwards of both topologies, a short
joint sample with and without the CUDA graph, multistep, the codec, a VAE decode) for `compute-sanitizer`:

    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panopticdiffusionmodels_b200 import utils  # noqa: E402
from panopticdiffusionmodels_b200.libs.autoencoder import FrozenAutoencoderKL  # noqa: E402
from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT  # noqa: E402
from panopticdiffusionmodels_b200.sampling import JointSampler  # noqa: E402

dev = torch.device("cuda:0")
kw = dict(img_size=16, patch_size=2, in_chans=4, embed_dim=128, depth=4, num_heads=2, mlp_ratio=4, qkv_bias=False,
          mlp_time_embed=False, clip_dim=64, num_clip_token=7, enable_panoptic=True, use_ground_truth=False, num_panoptic_class=8)
g = torch.Generator().manual_seed(0)
for separate in (False, True):
    torch.manual_seed(1)
    net = UViT(separate=separate, **kw)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.startswith("zero_convs"):
                p.copy_(torch.randn_like(p) * 0.02)
    net = net.to(dev).eval()
    B = 3
    x, m = torch.randn(B, 4, 16, 16, generator=g).to(dev), torch.randn(B, 8, 16, 16, generator=g).to(dev)
    ctx, ec = torch.randn(B, 7, 64, generator=g).to(dev), torch.randn(7, 64, generator=g).to(dev)
    t = torch.tensor([999.0, 500.5, 21.0], device=dev)
    for prec in ("fp32", "bf16"):
        net.precision = prec
        n, y = net(x, t, ctx, mask_token=m)
        n2 = net(x, t, ctx)
        assert torch.isfinite(n).all() and torch.isfinite(y).all() and torch.isfinite(n2).all()
        for method in ("fast", "multistep"):
            s = JointSampler(net, z_shape=(4, 16, 16), scale=2.0, sample_steps=7, method=method)
            for graph in (True, False):
                z, pm = s.sample(ctx, ec, x, m, use_graph=graph)
                assert torch.isfinite(z).all() and torch.isfinite(pm).all()
    labels = utils.labels_from_pred_mask(pm)
    bits = utils.int2bits(labels.unsqueeze(1))
    assert torch.isfinite(bits.float()).all()
    print("ok", "two-stream" if separate else "single-stream", flush=True)
dd = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2], num_res_blocks=1,
          attn_resolutions=[], dropout=0.0)
torch.manual_seed(2)
vae = FrozenAutoencoderKL(dd, 4, None, 0.2).to(dev)
img = vae.decode(torch.randn(2, 4, 16, 16, generator=g).to(dev))
assert torch.isfinite(img).all() and tuple(img.shape) == (2, 3, 32, 32)
torch.cuda.synchronize()
print("ok vae", flush=True)
