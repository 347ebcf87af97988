"""Text encoder of the t2i path: prompts -> CLIP ViT-L/14 last hidden states [B, 77, 768] (the ``context`` of ``UViT.forward``).

Mirror of the reference ``libs/clip.py:13-38`` (``FrozenCLIPEmbedder``): like the reference it is a thin wrapper over the
Hugging Face ``transformers`` CLIP text model -- library code that runs ONCE per sample, outside the timed denoising loop
(north star: "VAE decode and CLIP text encoding run once per sample outside the timed loop"), so it is not re-implemented
in libpdm.  ``version`` may be a hub id or a local directory; without network access the weights must already be on disk.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class AbstractEncoder(nn.Module):
    def encode(self, *args, **kwargs):
        raise NotImplementedError


class FrozenCLIPEmbedder(AbstractEncoder):
    def __init__(self, version: str = "openai/clip-vit-large-patch14", device="cuda", max_length: int = 77):
        super().__init__()
        from transformers import CLIPTextModel, CLIPTokenizer  # imported lazily: only this class needs transformers
        self.tokenizer = CLIPTokenizer.from_pretrained(version)
        self.transformer = CLIPTextModel.from_pretrained(version).eval()
        self.device, self.max_length = device, max_length
        for p in self.parameters():
            p.requires_grad_(False)

    @torch.no_grad()
    def forward(self, text):
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, padding="max_length", return_tensors="pt")
        return self.transformer(input_ids=enc["input_ids"].to(self.device)).last_hidden_state

    def encode(self, text):
        return self(text)
