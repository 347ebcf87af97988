"""End-to-end joint sampling driver: the step either side of the hot path.

Mirrors the live wiring of the reference (``train_t2i_discrete.py:480-546`` evaluation branch + ``utils.sample2dir``
``utils.py:552-637``; the stale ``sample_t2i_discrete.py`` is not reused):

    config file  ->  nnet from ``<ckpt_root>/<step>.ckpt/nnet_ema.pth``  ->  contexts  ->  JointSampler (libpdm)
                 ->  all-gather over ranks  ->  latents ``.pt`` + panoptic label PNGs (``bits2int`` + ``color_map``)

Either side of the loop (both run once per sample, outside it):
  * contexts: ``--prompts FILE`` (one caption per line) through ``libs.clip.FrozenCLIPEmbedder`` (``libs/clip.py:13-38``: the
    Hugging Face CLIP text model, as in the reference; ``--clip`` names a hub id or a local directory), or the
    extracted-feature cache (``--features``, ``datasets.py:577-613``), or ``--synthetic``;
  * images: ``--autoencoder autoencoder_kl.pth`` decodes the latents on the GPU with libpdm's VAE decoder
    (``libs.autoencoder``, csrc/vae.cu; the reference's ``decode_large_batch``, ``eval_t2i_discrete.py:74-84``) and writes
    ``<out>/{i}.png`` after ``unpreprocess`` (``0.5 (x + 1)`` clamped, ``datasets.py``).  The weights of both models are not part
    of this tree (no network): without them the driver still writes latents and panoptic label maps.

    python -m panopticdiffusionmodels_b200.sample_t2i --config mscoco_uvit_small --ckpt-root ckpts \\
           --features assets/datasets/coco256_features --autoencoder assets/stable-diffusion/autoencoder_kl.pth \\
           --out samples --n-samples 64
    torchrun --nproc-per-node 8 -m panopticdiffusionmodels_b200.sample_t2i ...     # one rank per GPU, batch sharded
"""
from __future__ import annotations

import argparse
import os
from typing import Callable, Optional

import torch

from . import checkpoint as ck
from . import configs, utils
from .distributed import rank_seed, sample_all, world
from .sampling import JointSampler


def build_nnet(config, device, ckpt_root: Optional[str] = None, nnet_path: Optional[str] = None, precision: str = "bf16"):
    """``utils.get_nnet(**config.nnet)`` + weights: an explicit ``nnet_path`` (``eval_t2i_discrete.py:50-51``), else the
    newest checkpoint directory under ``ckpt_root`` (EMA weights), else the reference initialisation."""
    nnet = utils.get_nnet(**config.nnet)
    if nnet_path:
        nnet.load_state_dict(torch.load(nnet_path, map_location="cpu"))
    elif ckpt_root:
        path = ck.resolve_checkpoint(ckpt_root)
        if path is not None:
            ck.load_nnet(nnet, path, which="nnet_ema")
    nnet = nnet.to(device).eval()
    nnet.precision = precision
    return nnet


def sample_to_dir(nnet, config, contexts_fn: Callable[[int], torch.Tensor], empty_context: torch.Tensor, out_dir: str,
                  n_samples: int, mini_batch_size: int, use_panoptic: bool = True,
                  decode: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, generator=None):
    """``utils.sample2dir``: draw ``n_samples`` over all ranks in mini-batches, gather, write.  Rank 0 writes
    ``<out>/latents.pt`` (z and pred_mask), ``<out>/mask/{i}.png`` (colour-mapped panoptic labels) and, when ``decode`` is
    given, ``<out>/{i}.png`` images."""
    rank, n = world()
    js = JointSampler(nnet, z_shape=tuple(config.z_shape), scale=config.sample.scale, cfg=config.sample.cfg,
                      sample_steps=config.sample.sample_steps)

    def one(b):
        ctx = contexts_fn(b)
        return js.sample(ctx, empty_context if config.sample.cfg else None, use_panoptic=use_panoptic, generator=generator)

    z, pm = sample_all(one, n_samples, mini_batch_size)
    labels = None
    if pm is not None:
        labels = utils.labels_from_pred_mask(pm)          # == bits2int(pred_mask > 0) (utils.py:596), on the device
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
        torch.save({"z": z.cpu(), "pred_mask": None if pm is None else pm.cpu()}, os.path.join(out_dir, "latents.pt"))
        if labels is not None:
            mask_dir = os.path.join(out_dir, "mask")
            os.makedirs(mask_dir, exist_ok=True)
            cm = ck.get_colormap(os.path.join(out_dir, "colormap.pt"))
            lab = labels.cpu()
            for i in range(lab.shape[0]):
                ck.save_mask_png(lab[i], os.path.join(mask_dir, f"{i}.png"), cm)
        if decode is not None:
            imgs = unpreprocess(decode(z))
            for i in range(imgs.shape[0]):
                ck.save_image_png(imgs[i], os.path.join(out_dir, f"{i}.png"))
    return z, pm, labels


def unpreprocess(v: torch.Tensor) -> torch.Tensor:
    """[-1, 1] -> [0, 1], clamped (datasets.py ``DatasetFactory.unpreprocess``)."""
    return (0.5 * (v + 1.0)).clamp_(0.0, 1.0)


def context_indices(batch_index: int, mini_batch_size: int, rank: int, n_ranks: int, total: int):
    """Dataset rows for this rank's mini-batch number ``batch_index``: CONTIGUOUS per-rank blocks inside each global batch, so
    that after the rank-major all-gather sample ``i`` of the output belongs to caption ``i`` of the dataset."""
    base = batch_index * mini_batch_size * n_ranks + rank * mini_batch_size
    return [(base + i) % total for i in range(mini_batch_size)]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default="mscoco_uvit_small", help="config name or path to a reference config file")
    ap.add_argument("--ckpt-root", default=None)
    ap.add_argument("--nnet-path", default=None)
    ap.add_argument("--features", default=None, help="extracted-feature directory (val2017/ + empty_context.npy)")
    ap.add_argument("--synthetic", action="store_true", help="random contexts (no feature cache at hand)")
    ap.add_argument("--prompts", default=None, help="text file, one caption per line -> CLIP text encoder (libs.clip)")
    ap.add_argument("--clip", default="openai/clip-vit-large-patch14", help="hub id or local directory of the CLIP text model")
    ap.add_argument("--autoencoder", default=None, help="autoencoder_kl.pth: decode the latents to PNGs with libpdm's VAE decoder")
    ap.add_argument("--out", default="samples")
    ap.add_argument("--n-samples", type=int, default=None)
    ap.add_argument("--mini-batch-size", type=int, default=None)
    ap.add_argument("--scale", type=float, default=None)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    a = ap.parse_args(argv)

    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    config = configs.load_config_file(a.config) if os.path.isfile(a.config) else configs.get_config(a.config)
    if a.scale is not None:
        config.sample.scale = a.scale
    n_samples = a.n_samples or config.sample.n_samples
    mbs = a.mini_batch_size or config.sample.mini_batch_size
    rank, _ = world()
    gen = torch.Generator(device=dev).manual_seed(rank_seed(config.seed))   # seed + rank (train_t2i_discrete.py:237)
    nnet = build_nnet(config, dev, a.ckpt_root, a.nnet_path, a.precision)
    use_panoptic = bool(config.nnet.get("enable_panoptic", True))
    calls = [0]
    if a.prompts:
        from .libs.clip import FrozenCLIPEmbedder
        clip = FrozenCLIPEmbedder(a.clip, device=dev).to(dev)
        prompts = [ln.strip() for ln in open(a.prompts) if ln.strip()]
        n_samples = a.n_samples or len(prompts)
        empty = clip.encode([""])[0].float()           # the reference's empty_context (scripts/extract_empty_feature.py)

        def contexts_fn(b):
            idx = context_indices(calls[0], b, rank, world()[1], len(prompts))
            calls[0] += 1
            return clip.encode([prompts[i] for i in idx]).float()
    elif a.features and not a.synthetic:
        cache = ck.FeatureCache(os.path.join(a.features, "val2017"))
        empty = ck.load_empty_context(a.features).to(dev)

        def contexts_fn(b):
            idx = context_indices(calls[0], b, rank, world()[1], len(cache))
            calls[0] += 1
            return cache.contexts(idx).to(dev)
    else:
        T, cd = int(config.nnet.get("num_clip_token", 77)), int(config.nnet.get("clip_dim", 768))
        empty = torch.randn(T, cd, device=dev, generator=gen)
        contexts_fn = lambda b: torch.randn(b, T, cd, device=dev, generator=gen)  # noqa: E731
    decode = None
    if a.autoencoder:
        from .libs import autoencoder as ae
        vae = ae.get_model(a.autoencoder, scale_factor=float(config.autoencoder.get("scale_factor", 0.18215))).to(dev)
        decode = vae.decode                              # chunked inside (decode_large_batch, eval_t2i_discrete.py:74-84)
    sample_to_dir(nnet, config, contexts_fn, empty, a.out, n_samples, mbs, use_panoptic=use_panoptic, decode=decode,
                  generator=gen)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
