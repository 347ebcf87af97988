"""On-disk formats either side of the path: checkpoint directory layout, partial (strict=False)
loads, the extracted-feature cache and the mask colour map.  CPU only."""
import os

import numpy as np
import torch

from conftest import TINY
from panopticdiffusionmodels_b200 import checkpoint as ck
from panopticdiffusionmodels_b200.libs.uvit_t2i import UViT


def test_resolve_checkpoint_like_trainstate_resume(tmp_path):
    root = tmp_path / "ckpts"
    assert ck.resolve_checkpoint(str(root)) is None                      # utils.py:387-388
    root.mkdir()
    assert ck.resolve_checkpoint(str(root)) is None                      # no *.ckpt entries (utils.py:391-392)
    for s in (5000, 20000, 10000):
        (root / f"{s}.ckpt").mkdir()
    assert ck.resolve_checkpoint(str(root)) == str(root / "20000.ckpt")  # highest step (utils.py:399-400)
    assert ck.resolve_checkpoint(str(root), step=5000) == str(root / "5000.ckpt")
    best = tmp_path / "b"
    best.mkdir()
    (best / "best.ckpt").mkdir()
    assert ck.resolve_checkpoint(str(best)) == str(best / "best.ckpt")   # utils.py:393-397


def test_checkpoint_roundtrip_and_partial_load(tmp_path):
    torch.manual_seed(0)
    a = UViT(separate=True, **TINY)
    path = str(tmp_path / "100.ckpt")
    ck.save_nnet(path, a, step=100)
    assert sorted(os.listdir(path)) == ["nnet.pth", "nnet_ema.pth", "step.pth"]
    assert ck.load_step(path) == 100
    torch.manual_seed(1)
    b = UViT(separate=True, **TINY)
    missing, unexpected = ck.load_nnet(b, path, which="nnet_ema")
    assert not missing and not unexpected
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    # image-only pretrained checkpoint (train_t2i_discrete.py:300-301): strict=False fills the image stream only
    sd = {k: v for k, v in a.state_dict().items() if "mask" not in k and not k.startswith("zero_convs")}
    torch.save(sd, str(tmp_path / "pretrained.pth"))
    torch.manual_seed(2)
    c = UViT(separate=True, **TINY)
    before = {k: v.clone() for k, v in c.state_dict().items()}
    missing, unexpected = ck.load_pretrained(c, str(tmp_path / "pretrained.pth"))
    assert not unexpected and missing and all(("mask" in k) or k.startswith("zero_convs") for k in missing)
    after = c.state_dict()
    assert all(torch.equal(after[k], sd[k]) for k in sd)
    assert all(torch.equal(after[k], before[k]) for k in missing)


def test_feature_cache_layout(tmp_path):
    split = tmp_path / "val2017"
    split.mkdir()
    rng = np.random.default_rng(0)
    for i, ncap in enumerate((2, 1, 3)):
        np.save(split / f"{i}.npy", rng.standard_normal((8, 32, 32)).astype(np.float32))
        for k in range(ncap):
            np.save(split / f"{i}_{k}.npy", rng.standard_normal((77, 768)).astype(np.float32))
        np.save(split / f"{i}_seg.npy", rng.integers(0, 200, (3, 128, 128)).astype(np.float32))
    np.save(tmp_path / "empty_context.npy", rng.standard_normal((77, 768)).astype(np.float32))
    n, caps = ck.get_feature_dir_info(str(split))
    assert n == 3 and caps == {0: 2, 1: 1, 2: 3}                  # datasets.py:551-561 (the _seg files are not captions)
    fc = ck.FeatureCache(str(split))
    z, c, s, idx = fc.__getitem__(2, k=1)
    assert z.shape == (8, 32, 32) and c.shape == (77, 768) and idx == 2
    seg = np.load(split / "2_seg.npy")
    assert s.shape == (1, 32, 32)                                  # block_reduce (3,4,4) with np.min (datasets.py:589)
    assert s[0, 3, 5] == seg[:, 12:16, 20:24].min()
    assert fc.contexts([0, 2], k=0).shape == (2, 77, 768)
    assert ck.load_empty_context(str(tmp_path)).shape == (77, 768)


def test_color_map_and_png(tmp_path):
    cm_path = str(tmp_path / "colormap.pt")
    cm = ck.get_colormap(cm_path)
    assert cm.shape == (256, 3) and os.path.isfile(cm_path)
    assert torch.equal(ck.get_colormap(cm_path), cm)               # persisted, reused (utils.py:522-523)
    labels = torch.randint(0, 256, (2, 1, 8, 8)).float()
    rgb = ck.color_map(labels, cm)
    assert rgb.shape == (2, 3, 8, 8)
    assert torch.equal(rgb[1, :, 2, 3], cm[int(labels[1, 0, 2, 3])])
    out = str(tmp_path / "m.png")
    ck.save_mask_png(labels[0, 0], out, cm)
    from PIL import Image
    im = np.array(Image.open(out))
    assert im.shape == (8, 8, 3) and tuple(im[4, 5]) == tuple(int(v) for v in cm[int(labels[0, 0, 4, 5])])
