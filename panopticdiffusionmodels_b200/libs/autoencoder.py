"""Drop-in ``FrozenAutoencoderKL`` / ``get_model`` for the step AFTER the sampling loop: latents -> images.

Same constructor arguments, ``state_dict`` layout (``encoder.* / decoder.* / quant_conv.* / post_quant_conv.*``) and
``decode(z)`` / ``forward(inputs, fn='decode')`` as the reference (``libs/autoencoder.py:412-485``), so a reference
``autoencoder_kl.pth`` loads with ``strict=True`` -- but ``decode`` is executed by libpdm.so (``pdm_vae_decode``:
3x3 convolutions as tcgen05 implicit GEMMs, GroupNorm / swish / upsample kernels, csrc/vae.cu).  The torch sub-modules below
only *hold* parameters under the reference's names.  The sampling path never encodes: ``encode`` / ``encode_moments`` raise
``NotImplementedError`` (their weights are kept so that checkpoints round-trip).  There is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
import torch.nn as nn

from .. import _lib


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the decoder runs inside libpdm.so")


def _norm(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


def _resnet(cin, cout):
    m = _Holder()
    m.norm1 = _norm(cin)
    m.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
    m.norm2 = _norm(cout)
    m.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
    if cin != cout:
        m.nin_shortcut = nn.Conv2d(cin, cout, 1, 1, 0)
    return m


def _attn(c):
    m = _Holder()
    m.norm = _norm(c)
    for n in ("q", "k", "v", "proj_out"):
        setattr(m, n, nn.Conv2d(c, c, 1, 1, 0))
    return m


def _mid(c):
    m = _Holder()
    m.block_1 = _resnet(c, c)
    m.attn_1 = _attn(c)
    m.block_2 = _resnet(c, c)
    return m


def _encoder(ch, ch_mult, num_res_blocks, in_channels, z_channels, double_z):
    enc = _Holder()
    enc.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
    in_mult = (1,) + tuple(ch_mult)
    enc.down = nn.ModuleList()
    block_in = ch
    for lev in range(len(ch_mult)):
        block_in, block_out = ch * in_mult[lev], ch * ch_mult[lev]
        d = _Holder()
        d.block = nn.ModuleList()
        d.attn = nn.ModuleList()
        for _ in range(num_res_blocks):
            d.block.append(_resnet(block_in, block_out))
            block_in = block_out
        if lev != len(ch_mult) - 1:
            d.downsample = _Holder()
            d.downsample.conv = nn.Conv2d(block_in, block_in, 3, 2, 0)
        enc.down.append(d)
    enc.mid = _mid(block_in)
    enc.norm_out = _norm(block_in)
    enc.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, 1, 1)
    return enc


def _decoder(ch, ch_mult, num_res_blocks, out_ch, z_channels):
    dec = _Holder()
    nlev = len(ch_mult)
    block_in = ch * ch_mult[nlev - 1]
    dec.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
    dec.mid = _mid(block_in)
    ups = []
    for lev in reversed(range(nlev)):
        block_out = ch * ch_mult[lev]
        u = _Holder()
        u.block = nn.ModuleList()
        u.attn = nn.ModuleList()
        for _ in range(num_res_blocks + 1):
            u.block.append(_resnet(block_in, block_out))
            block_in = block_out
        if lev != 0:
            u.upsample = _Holder()
            u.upsample.conv = nn.Conv2d(block_in, block_in, 3, 1, 1)
        ups.insert(0, u)
    dec.up = nn.ModuleList(ups)
    dec.norm_out = _norm(block_in)
    dec.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)
    return dec


class FrozenAutoencoderKL(nn.Module):
    def __init__(self, ddconfig, embed_dim, pretrained_path: Optional[str], scale_factor=0.18215):
        super().__init__()
        if ddconfig.get("attn_resolutions"):
            raise NotImplementedError("libpdm implements the SD autoencoder layout: attention in the mid block only")
        if not ddconfig.get("double_z", True):
            raise NotImplementedError("double_z=False")
        self.ddconfig = dict(ddconfig)
        ch, ch_mult, nrb = ddconfig["ch"], tuple(ddconfig["ch_mult"]), ddconfig["num_res_blocks"]
        zc = ddconfig["z_channels"]
        self.encoder = _encoder(ch, ch_mult, nrb, ddconfig["in_channels"], zc, True)
        self.decoder = _decoder(ch, ch_mult, nrb, ddconfig["out_ch"], zc)
        self.quant_conv = nn.Conv2d(2 * zc, 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, zc, 1)
        self.embed_dim = embed_dim
        self.scale_factor = scale_factor
        if pretrained_path is not None:  # reference: load, assert no missing / unexpected keys (autoencoder.py:423-424)
            m, u = self.load_state_dict(torch.load(pretrained_path, map_location="cpu"))
            assert len(m) == 0 and len(u) == 0
        self.eval()
        self.requires_grad_(False)
        self._handle: Optional[C.c_void_p] = None
        self._handle_device = None
        self._fingerprint = None

    # ---- engine management (same scheme as libs.uvit_t2i.UViT) ----
    def _release(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            # plain dict write: nn.Module.__setattr__ is not usable any more when this runs from __del__ at interpreter exit
            self.__dict__["_handle"] = None
            try:
                _lib.lib().pdm_vae_destroy(h)
            except Exception:
                pass

    def __del__(self):
        self._release()

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"] = state["_handle_device"] = state["_fingerprint"] = None
        return state

    def engine(self) -> C.c_void_p:
        dev = self.post_quant_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("FrozenAutoencoderKL (libpdm) has no CPU path: move the module to a CUDA device")
        L = _lib.lib()
        with torch.cuda.device(dev):
            if self._handle is None or self._handle_device != dev:
                self._release()
                dd = self.ddconfig
                mult = (C.c_int32 * 8)(*([int(v) for v in dd["ch_mult"]] + [0] * (8 - len(dd["ch_mult"]))))
                cfg = _lib.PdmVaeConfig(int(dd["ch"]), len(dd["ch_mult"]), mult, int(dd["num_res_blocks"]), int(dd["z_channels"]),
                                        int(self.embed_dim), int(dd["out_ch"]), float(self.scale_factor))
                h = C.c_void_p()
                _lib.check(L.pdm_vae_create(C.byref(cfg), C.byref(h)))
                self._handle, self._handle_device, self._fingerprint = h, dev, None
            params = [(k, v) for k, v in self.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))]
            fp = tuple((v.data_ptr(), v._version) for _, v in params)
            if fp != self._fingerprint:
                stream = _lib.current_stream()
                for key, t in params:
                    t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(L.pdm_vae_set_param(self._handle, key.encode(), t.data_ptr(), shape, t.dim(), stream))
                _lib.check(L.pdm_vae_finalize_params(self._handle, stream))
                self._fingerprint = fp
        return self._handle

    # ---- reference surface ----
    def encode_moments(self, x):
        raise NotImplementedError("the sampling path only decodes; the VAE encoder is outside libpdm's scope")

    def encode(self, x):
        raise NotImplementedError("the sampling path only decodes; the VAE encoder is outside libpdm's scope")

    @torch.no_grad()
    def decode(self, z: torch.Tensor, max_batch: int = 32) -> torch.Tensor:
        """``(1 / scale_factor) z -> post_quant_conv -> Decoder`` (libs/autoencoder.py:446-450), in chunks of ``max_batch``."""
        if not z.is_cuda:
            raise RuntimeError("libpdm has no CPU path: z must be a CUDA tensor")
        n, c, h, w = z.shape
        if c != self.ddconfig["z_channels"] or h != w:
            raise ValueError(f"expected z of shape (n, {self.ddconfig['z_channels']}, s, s), got {tuple(z.shape)}")
        hnd = self.engine()
        up = 2 ** (len(self.ddconfig["ch_mult"]) - 1)
        z = z.to(torch.float32).contiguous()
        out = torch.empty(n, self.ddconfig["out_ch"], h * up, w * up, device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device):
            for i in range(0, n, max_batch):
                zi, oi = z[i:i + max_batch], out[i:i + max_batch]
                _lib.check(_lib.lib().pdm_vae_decode(hnd, _lib.ptr(zi), _lib.ptr(oi), zi.shape[0], h, _lib.current_stream()))
        return out

    def forward(self, inputs, fn):
        if fn == "decode":
            return self.decode(inputs)
        if fn in ("encode", "encode_moments"):
            return self.encode(inputs)
        raise NotImplementedError


SD_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                   num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def get_model(pretrained_path, scale_factor=0.18215):
    """libs/autoencoder.py:471-485: the Stable-Diffusion KL autoencoder (f = 8, 4 latent channels)."""
    return FrozenAutoencoderKL(dict(SD_DDCONFIG), 4, pretrained_path, scale_factor)
