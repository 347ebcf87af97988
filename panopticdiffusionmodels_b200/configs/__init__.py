"""The four MSCOCO U-ViT t2i configurations BASELINE.json names as one table (``get_config()`` -> ConfigDict).  Every key
of the reference's ``configs/mscoco_uvit_{small,mid,large,small_512}.py`` is present with the reference's value, except the
three machine-specific absolute paths (``dataset.path``, ``sample.path`` and the small config's ``pretrained``), which are
relative / ``None`` here (``tests/test_host_cpu.py::test_config_table_matches_reference_files`` pins this against
``tests/golden/reference_configs.json``).  ``load_config_file`` loads a reference-style config *file* unchanged
(ml_collections shim) when a user has one."""
from __future__ import annotations

import importlib.util
import sys
import types

from ..ml_collections_shim import ConfigDict

# name -> (embed_dim, depth, heads, img_size, extra nnet kwargs, train batch, mini_batch, n_samples, data dir)
_TABLE = {
    "mscoco_uvit_small": (512, 12, 8, 32, dict(enable_panoptic=True, use_ground_truth=False, separate=True,
                                               num_panoptic_class=8, patch_factor=2), 64, 32, 10000, "coco256_features"),
    "mscoco_uvit_mid": (768, 16, 12, 32, dict(enable_panoptic=False, use_ground_truth=False, separate=False,
                                              num_panoptic_class=8, patch_factor=1), 32, 32, 30000, "coco256_features"),
    "mscoco_uvit_large": (1024, 20, 16, 32, {}, 64, 32, 30000, "coco256_features"),
    "mscoco_uvit_small_512": (512, 12, 8, 64, {}, 8, 10, 30000, "coco512_features"),
}
NAMES = tuple(_TABLE)


def get_config(name: str) -> ConfigDict:
    D, depth, heads, img, extra, train_bs, mini_bs, n_samples, data = _TABLE[name]
    c = ConfigDict()
    c.seed = 1234
    c.z_shape = (4, img, img)
    c.autoencoder = ConfigDict(dict(pretrained_path="assets/stable-diffusion/autoencoder_kl.pth", scale_factor=0.23010))
    c.train = ConfigDict(dict(n_steps=2000000 if name in ("mscoco_uvit_small", "mscoco_uvit_small_512") else 1000000, batch_size=train_bs,
                              log_interval=20 if name in ("mscoco_uvit_small", "mscoco_uvit_mid") else 10, eval_interval=5000,
                              save_interval=50000))
    c.optimizer = ConfigDict(dict(name="adamw", lr=0.0002, weight_decay=0.03, betas=(0.9, 0.9)))
    c.lr_scheduler = ConfigDict(dict(name="customized", warmup_steps=5000))
    c.nnet = ConfigDict(dict(name="uvit_t2i", img_size=img, in_chans=4, patch_size=2, embed_dim=D, depth=depth,
                             num_heads=heads, mlp_ratio=4, qkv_bias=False, mlp_time_embed=False, clip_dim=768,
                             num_clip_token=77, **extra))
    c.dataset = ConfigDict(dict(name="mscoco256_features", path=data, cfg=True, p_uncond=0.1))
    c.sample = ConfigDict(dict(sample_steps=50, n_samples=n_samples, mini_batch_size=mini_bs, cfg=True, scale=1.0,
                               path="sample"))
    if name in ("mscoco_uvit_small", "mscoco_uvit_mid"):
        c.use_unet = False
        c.mask_channel = 1
        c.pretrained = None
    return c


def load_config_file(path: str) -> ConfigDict:
    """Execute a reference-style ``configs/*.py`` (``import ml_collections`` + ``get_config()``) unchanged."""
    if "ml_collections" not in sys.modules:
        shim = types.ModuleType("ml_collections")
        from .. import ml_collections_shim as m
        shim.ConfigDict, shim.FrozenConfigDict = m.ConfigDict, m.FrozenConfigDict
        sys.modules["ml_collections"] = shim
    spec = importlib.util.spec_from_file_location("_pdm_user_config", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_config()
