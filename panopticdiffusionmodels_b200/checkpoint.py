"""On-disk formats either side of the sampling path -- host-side Python only.

* checkpoint directory layout of the reference ``TrainState`` (``utils.py:366-405``):
  ``<ckpt_root>/<step>.ckpt/{step.pth, nnet.pth, nnet_ema.pth, optimizer.pth, lr_scheduler.pth}`` (or ``best.ckpt``);
  the ``nnet*`` files are bare ``state_dict``s without prefix, loaded with ``strict=False`` (``utils.py:381-382``);
* the pretrained image-only U-ViT partial load of ``train_t2i_discrete.py:300-301`` (``strict=False``: the panoptic
  tensors keep their initialisation);
* the feature cache written by the reference's ``scripts/extract_*`` and read by ``datasets.py:551-613, 629``:
  ``<split>/{i}.npy`` (latent moments), ``{i}_{k}.npy`` (CLIP context of caption k), ``{i}_seg.npy`` /
  ``{i}_encode_p.npy`` (panoptic map) and ``empty_context.npy`` one level up.
Nothing here touches the GPU: the tensors these functions return are what ``UViT.load_state_dict`` /
``JointSampler.sample`` take.
"""
from __future__ import annotations

import glob
import os
import random
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


# ------------------------------------------------------------------------------------------ checkpoints
def resolve_checkpoint(ckpt_root: str, step: Optional[int] = None) -> Optional[str]:
    """Directory ``TrainState.resume`` would load (``utils.py:386-405``): the given step, else ``best.ckpt`` when the
    entries are not numbered, else the highest numbered ``<step>.ckpt``; ``None`` if there is nothing to resume."""
    if not os.path.exists(ckpt_root):
        return None
    if step is None:
        ckpts = [x for x in os.listdir(ckpt_root) if ".ckpt" in x]
        if not ckpts:
            return None
        if not ckpts[0].split(".")[0].isnumeric():
            return os.path.join(ckpt_root, "best.ckpt")
        step = max(int(x.split(".")[0]) for x in ckpts)
    return os.path.join(ckpt_root, f"{step}.ckpt")


def load_nnet(nnet: torch.nn.Module, ckpt_path: str, which: str = "nnet_ema", strict: bool = False):
    """``val.load_state_dict(torch.load(<ckpt>/<which>.pth, map_location='cpu'), strict=False)`` (``utils.py:381-382``).
    ``which`` is ``'nnet'`` or ``'nnet_ema'`` (evaluation / sampling use the EMA weights, ``train_t2i_discrete.py:480``).
    Returns the ``(missing_keys, unexpected_keys)`` of ``load_state_dict``."""
    sd = torch.load(os.path.join(ckpt_path, f"{which}.pth"), map_location="cpu")
    return nnet.load_state_dict(sd, strict=strict)


def load_step(ckpt_path: str) -> int:
    return int(torch.load(os.path.join(ckpt_path, "step.pth"), map_location="cpu"))


def save_nnet(ckpt_path: str, nnet: torch.nn.Module, nnet_ema: Optional[torch.nn.Module] = None, step: int = 0) -> None:
    """The sampling-relevant subset of ``TrainState.save`` (``utils.py:366-371``): ``step.pth``, ``nnet.pth``,
    ``nnet_ema.pth`` (optimizer / scheduler belong to the training loop, out of scope)."""
    os.makedirs(ckpt_path, exist_ok=True)
    torch.save(step, os.path.join(ckpt_path, "step.pth"))
    torch.save(nnet.state_dict(), os.path.join(ckpt_path, "nnet.pth"))
    torch.save((nnet_ema if nnet_ema is not None else nnet).state_dict(), os.path.join(ckpt_path, "nnet_ema.pth"))


def load_pretrained(nnet: torch.nn.Module, path: str):
    """``nnet.load_state_dict(torch.load(config.pretrained), strict=False)`` (``train_t2i_discrete.py:300-301``): an
    image-only U-ViT checkpoint fills the image stream; mask-stream / panoptic tensors keep their initialisation."""
    return nnet.load_state_dict(torch.load(path, map_location="cpu"), strict=False)


# ------------------------------------------------------------------------------------------ feature cache
def get_feature_dir_info(root: str) -> Tuple[int, Dict[int, int]]:
    """``datasets.py:551-561``: number of samples and captions per sample from the file names."""
    files = glob.glob(os.path.join(root, "*.npy"))
    files_caption = glob.glob(os.path.join(root, "*_*.npy"))
    num_data = len(files) - len(files_caption)
    n_captions = {k: 0 for k in range(num_data)}
    for f in files_caption:
        k1, k2 = os.path.splitext(os.path.split(f)[-1])[0].split("_", 1)
        if k2.isnumeric():
            n_captions[int(k1)] += 1
    return num_data, n_captions


class FeatureCache:
    """Reader of one split of the extracted-feature directory (``MSCOCOFeatureDataset``, ``datasets.py:564-613``):
    ``(z_moments, context, panoptic, index)`` per item.  ``pool`` mirrors the reference's
    ``skimage.measure.block_reduce(s, (3, 4, 4), np.min)`` of the category-id map (done with numpy here)."""

    def __init__(self, root: str, use_category_id: bool = True):
        self.root = root
        self.num_data, self.n_captions = get_feature_dir_info(root)
        self.use_category_id = use_category_id

    def __len__(self) -> int:
        return self.num_data

    @staticmethod
    def pool(s: np.ndarray, block=(3, 4, 4)) -> np.ndarray:
        c, h, w = s.shape
        bc, bh, bw = block
        pad = lambda n, b: (-n) % b
        s = np.pad(s, ((0, pad(c, bc)), (0, pad(h, bh)), (0, pad(w, bw))), mode="constant", constant_values=0)
        c2, h2, w2 = s.shape
        return s.reshape(c2 // bc, bc, h2 // bh, bh, w2 // bw, bw).min(axis=(1, 3, 5))

    def __getitem__(self, index: int, k: Optional[int] = None):
        z = np.load(os.path.join(self.root, f"{index}.npy"))
        if k is None:
            k = random.randint(0, self.n_captions[index] - 1)
        c = np.load(os.path.join(self.root, f"{index}_{k}.npy"))
        if self.use_category_id:
            s = self.pool(np.load(os.path.join(self.root, f"{index}_seg.npy")))
        else:
            s = np.load(os.path.join(self.root, f"{index}_encode_p.npy"))
        return z, c, s, index

    def contexts(self, indices: List[int], k: int = 0) -> torch.Tensor:
        """Stack the CLIP contexts of caption ``k`` for a batch: ``(B, 77, 768)`` float32, the ``context`` argument of
        ``JointSampler.sample``."""
        return torch.from_numpy(np.stack([np.load(os.path.join(self.root, f"{i}_{k}.npy")) for i in indices])).float()


def load_empty_context(path: str) -> torch.Tensor:
    """``np.load(<path>/empty_context.npy)`` (``datasets.py:629``): the unconditional context of classifier-free guidance."""
    return torch.from_numpy(np.load(os.path.join(path, "empty_context.npy"))).float()


# ------------------------------------------------------------------------------------------ mask visualisation
def get_colormap(path: str, force: bool = False) -> torch.Tensor:
    """``utils.py:521-531``: a (256, 3) random colour table persisted next to the run (created on first use)."""
    if os.path.isfile(path) and not force:
        return torch.load(path)
    colormap = torch.randint(0, 255, (256, 3))
    torch.save(colormap, path)
    return colormap


def color_map(x: torch.Tensor, colormap: Optional[torch.Tensor] = None, path: str = "colormap.pt") -> torch.Tensor:
    """``utils.py:533-543``: label ids ``(B, 1, H, W)`` or ``(B, H, W)`` -> colours ``(B, 3, H, W)``."""
    if x.dim() > 3:
        x = x.squeeze(1)
    if colormap is None:
        colormap = get_colormap(path)
    return colormap[x.to(torch.long).cpu()].permute(0, 3, 1, 2)


def save_mask_png(labels: torch.Tensor, path: str, colormap: Optional[torch.Tensor] = None) -> None:
    """One panoptic label map ``(H, W)`` / ``(1, H, W)`` -> RGB PNG, the per-sample write of ``utils.py:627-632``."""
    from PIL import Image
    if labels.dim() == 2:
        labels = labels.unsqueeze(0)
    rgb = color_map(labels, colormap)[0].permute(1, 2, 0).to("cpu", torch.uint8).numpy()
    Image.fromarray(rgb).save(path)


def save_image_png(img: torch.Tensor, path: str) -> None:
    """One image ``(3, H, W)`` in [0, 1] -> PNG, the rounding of ``torchvision.utils.save_image`` (``utils.py:634``):
    ``img * 255 + 0.5`` clamped to [0, 255], truncated to uint8."""
    from PIL import Image
    arr = img.detach().float().mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8).numpy()
    Image.fromarray(arr).save(path)
