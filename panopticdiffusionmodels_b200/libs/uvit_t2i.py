"""Drop-in ``UViT`` for the joint image+mask text-to-image network.

Same constructor keywords, ``forward`` signature and ``state_dict`` layout as the reference
(``libs/uvit_t2i.py:258-525``), so ``utils.get_nnet(**config.nnet)``, ``load_state_dict`` of a
reference ``nnet.pth`` and ``nnet(x, timesteps, context, mask_token=...)`` keep working -- but the
forward itself is executed by libpdm.so (hand-written sm_100a kernels behind the C ABI of
``include/pdm.h``).  The torch sub-modules below only *hold* parameters under the reference's names;
their own ``forward`` is never used and there is no PyTorch/CPU fallback.

Documented deviations: ``patch_factor`` is accepted and ignored so the shipped config
files load; ``mlp_time_embed=True``, ``qkv_bias=True``, ``qk_scale``, ``conv=False`` and ``skip=False`` are rejected (no
MSCOCO t2i config uses them).  ``forward(..., use_ground_truth=True)`` (``libs/uvit_t2i.py:380, 486-496``) IS
implemented (``pdm_nnet_forward_ex`` with ``PDM_FWD_GROUND_TRUTH``).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib


def _trunc_normal_(t: torch.Tensor, std: float) -> None:
    # reference init: libs/timm.py:44-62 -> N(0, std) truncated to [-2, 2] (absolute bounds)
    nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class _Holder(nn.Module):
    """A named bag of sub-modules / parameters (never called)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the network runs inside libpdm.so")


def _attention_params(dim: int) -> nn.Module:
    m = _Holder()
    m.qkv = nn.Linear(dim, dim * 3, bias=False)
    m.proj = nn.Linear(dim, dim)
    return m


def _mlp_params(dim: int, hidden: int) -> nn.Module:
    m = _Holder()
    m.fc1 = nn.Linear(dim, hidden)
    m.fc2 = nn.Linear(hidden, dim)
    return m


def _block_params(dim: int, mlp_ratio: float, skip: bool) -> nn.Module:
    m = _Holder()
    m.norm1 = nn.LayerNorm(dim)
    m.attn = _attention_params(dim)
    m.norm2 = nn.LayerNorm(dim)
    m.mlp = _mlp_params(dim, int(dim * mlp_ratio))
    if skip:
        m.skip_linear = nn.Linear(2 * dim, dim)
    return m


def _patch_params(patch: int, chans: int, dim: int) -> nn.Module:
    m = _Holder()
    m.proj = nn.Conv2d(chans, dim, kernel_size=patch, stride=patch)
    return m


def _bridge_params(dim: int) -> nn.Module:
    m = _Holder()
    m.conv = nn.Conv1d(dim, dim, 1, padding=0)
    return m


class UViT(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.,
                 qkv_bias=False, qk_scale=None, norm_layer=nn.LayerNorm, mlp_time_embed=False, use_checkpoint=False,
                 clip_dim=768, num_clip_token=77, conv=True, skip=True, num_panoptic_class=8, enable_panoptic=True,
                 use_ground_truth=False, separate=False, patch_factor=None):
        super().__init__()
        if mlp_time_embed or qkv_bias or qk_scale is not None or not conv or not skip or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("libpdm implements the configuration the MSCOCO t2i configs use: "
                                      "mlp_time_embed=False, qkv_bias=False, qk_scale=None, conv=True, skip=True")
        if embed_dim % 64 or embed_dim // num_heads != 64:
            raise NotImplementedError("libpdm kernels are specialised for head_dim == 64")
        if int(mlp_ratio) != mlp_ratio:
            raise NotImplementedError("mlp_ratio must be integral")
        self.num_features = self.embed_dim = embed_dim
        self.in_chans = in_chans
        self.enable_panoptic = bool(enable_panoptic)
        self.separate = bool(separate)
        self.depth = depth
        self.num_heads = num_heads
        self.img_size, self.patch_size = img_size, patch_size
        self.clip_dim, self.num_clip_token = clip_dim, num_clip_token
        self.mlp_ratio = int(mlp_ratio)
        self.num_panoptic_class = num_panoptic_class
        self.use_ground_truth = use_ground_truth
        self.extras = 1 + num_clip_token
        self.patch_dim = patch_size ** 2 * in_chans
        num_patches = (img_size // patch_size) ** 2
        D = embed_dim

        self.patch_embed = _patch_params(patch_size, in_chans, D)
        self.time_embed = nn.Identity()
        self.context_embed = nn.Linear(clip_dim, D)
        ntok = self.extras + num_patches * (2 if (self.enable_panoptic and not self.separate) else 1)
        self.pos_embed = nn.Parameter(torch.zeros(1, ntok, D))
        if self.enable_panoptic and self.separate:
            self.pos_embed_mask = nn.Parameter(torch.zeros(1, num_patches, D))

        def stack(n, with_skip):
            return nn.ModuleList([_block_params(D, mlp_ratio, with_skip) for _ in range(n)])

        self.in_blocks = stack(depth // 2, False)
        self.mid_block = _block_params(D, mlp_ratio, False)
        self.out_blocks = stack(depth // 2, True)
        if self.separate:
            self.in_blocks_mask = stack(depth // 2, False)
            self.mid_block_mask = _block_params(D, mlp_ratio, False)
            self.out_blocks_mask = stack(depth // 2, True)
            self.zero_convs = nn.ModuleList([_bridge_params(D) for _ in range(depth * 2 + 2)])
        self.norm = nn.LayerNorm(D)
        self.decoder_pred = nn.Linear(D, self.patch_dim, bias=True)
        self.final_layer = nn.Conv2d(in_chans, in_chans, 3, padding=1)
        if self.enable_panoptic:
            self.mask_embed = _patch_params(patch_size, num_panoptic_class, D)
            self.mask_embed_0 = _patch_params(patch_size, num_panoptic_class, D)  # dead weights kept for the layout
            self.decoder_pred_mask = nn.Linear(D, patch_size ** 2 * num_panoptic_class, bias=True)
            self.final_layer_mask = nn.Conv2d(num_panoptic_class, num_panoptic_class, 3, padding=1)
            self.final_act = nn.Tanh()
        self._init_like_reference()

        self.precision = "bf16"      # "bf16" (tcgen05 tensor cores) | "fp32" (parity anchor)
        self._handle: Optional[C.c_void_p] = None
        self._handle_device = None
        self._fingerprint = None

    # ---- init: Linear -> trunc_normal(.02)/zero bias, LayerNorm -> 1/0, Conv1d bridges -> 0 (uvit_t2i.py:352-372)
    def _init_like_reference(self):
        if hasattr(self, "pos_embed_mask"):
            _trunc_normal_(self.pos_embed_mask, .02)
        _trunc_normal_(self.pos_embed, .02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, .02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Conv1d):
                nn.init.zeros_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed"}

    # ---- engine management -------------------------------------------------------------------
    def pdm_config(self) -> _lib.PdmConfig:
        return _lib.PdmConfig(self.img_size, self.patch_size, self.in_chans, self.embed_dim, self.depth,
                              self.num_heads, self.mlp_ratio, self.clip_dim, self.num_clip_token,
                              self.num_panoptic_class, int(self.enable_panoptic), int(self.separate))

    def _release(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            # plain dict write: nn.Module.__setattr__ is not usable any more when this runs from __del__ at interpreter exit
            self.__dict__["_handle"] = None
            try:
                _lib.lib().pdm_destroy(h)
            except Exception:
                pass

    def __del__(self):
        self._release()

    # The engine handle is a per-object device resource: copies (copy.copy / copy.deepcopy, pickle, torch.save of the
    # whole module, e.g. an EMA twin of the network) must not alias it (double pdm_destroy) nor try to pickle
    # a ctypes pointer.  A copy lazily creates its own engine on first use.
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"] = state["_handle_device"] = state["_fingerprint"] = None
        return state

    def engine(self) -> C.c_void_p:
        """Create the device engine (once per device) and upload parameters whenever they changed."""
        dev = self.pos_embed.device
        if dev.type != "cuda":
            raise RuntimeError("UViT (libpdm) has no CPU path: move the module to a CUDA device")
        L = _lib.lib()
        with torch.cuda.device(dev):
            if self._handle is None or self._handle_device != dev:
                self._release()
                h = C.c_void_p()
                cfg = self.pdm_config()
                _lib.check(L.pdm_create(C.byref(cfg), C.byref(h)))
                self._handle, self._handle_device, self._fingerprint = h, dev, None
            fp = tuple((p.data_ptr(), p._version) for p in self.parameters())
            if fp != self._fingerprint:
                stream = _lib.current_stream()
                for key, t in self.state_dict().items():
                    t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(L.pdm_set_param(self._handle, key.encode(), t.data_ptr(), shape, t.dim(), stream))
                half = self.embed_dim // 2
                freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(dev)
                shape = (C.c_int64 * 1)(half)
                _lib.check(L.pdm_set_param(self._handle, b"__timestep_freqs__", freqs.data_ptr(), shape, 1, stream))
                _lib.check(L.pdm_finalize_params(self._handle, stream))
                self._fingerprint = fp
        return self._handle

    def prec_code(self) -> int:
        if self.precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision!r}")
        return _lib.PREC_BF16 if self.precision == "bf16" else _lib.PREC_FP32

    # ---- forward (libs/uvit_t2i.py:378-525) ---------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, timesteps, context, mask_token=None, mask_0=None, use_ground_truth=False,
                enable_panoptic=False):
        # the reference stores the per-call flag on the module (libs/uvit_t2i.py:380); it only matters with a mask_token
        self.use_ground_truth = bool(use_ground_truth)
        if mask_token is not None and not self.enable_panoptic:
            raise TypeError("this UViT was built with enable_panoptic=False and cannot take mask_token")
        if x.dim() == 3:
            x = x.unsqueeze(1)
        B, Cc, Hh, Ww = x.shape
        assert Hh % self.patch_size == 0 and Ww % self.patch_size == 0
        if Cc != self.in_chans or Hh != self.img_size or Ww != self.img_size:
            raise ValueError(f"expected x of shape (B,{self.in_chans},{self.img_size},{self.img_size}), got {tuple(x.shape)}")
        h = self.engine()
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        x = x.to(**f32).contiguous()
        t = torch.as_tensor(timesteps, **f32).reshape(-1)
        if t.numel() == 1 and B > 1:
            t = t.expand(B)
        t = t.contiguous()
        context = context.to(**f32).contiguous()
        if tuple(context.shape) != (B, self.num_clip_token, self.clip_dim):
            raise ValueError(f"expected context of shape {(B, self.num_clip_token, self.clip_dim)}, got {tuple(context.shape)}")
        noise = torch.empty_like(x)
        y = None
        if mask_token is not None:
            mask_token = mask_token.to(**f32).contiguous()
            if tuple(mask_token.shape) != (B, self.num_panoptic_class, self.img_size, self.img_size):
                raise ValueError("mask_token must be (B, num_panoptic_class, img_size, img_size) (SURVEY F4)")
            y = torch.empty_like(mask_token)
        if B == 0:  # an empty batch flows through the reference's torch ops as empty tensors; nothing to launch
            return noise if y is None else (noise, y)
        with torch.cuda.device(dev):
            flags = _lib.FWD_GROUND_TRUTH if (use_ground_truth and mask_token is not None) else 0
            _lib.check(_lib.lib().pdm_nnet_forward_ex(
                h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(context), _lib.ptr(mask_token), _lib.ptr(noise), _lib.ptr(y),
                B, self.prec_code(), flags, _lib.current_stream()))
        return noise if y is None else (noise, y)
