"""RoPE tables of the UniDisc DiT (host-side, computed once at construction).

text : 1-D, angle = pos * 10000^(-2i/hd), i < hd/2                      (reference models/dit.py:307-330, 1228-1239)
image: 2-D "Lumina" layout, per axis hd/4 frequencies 10000^(-2i/(hd/2)) / linear_factor, angles interleaved
       [row f0, col f0, row f1, col f1, ...] over a row-major side x side grid (reference models/dit.py:1046-1061 calling
       diffusers 0.32.2 get_2d_rotary_pos_embed_lumina).
"""
from __future__ import annotations

import math

import torch


# data.require_sample_ids: (image-block size, linear_factor) of the per-size 2-D tables (reference models/dit.py:1209)
INTERLEAVED_IMG_TABLES = ((256, 1.0), (1024, 2.0), (2304, 3.0), (4096, 4.0))


def rope_1d(head_dim: int, seq_len: int):
    inv_freq = 1.0 / (10000 ** (torch.arange(0, head_dim, 2).float() / head_dim))
    ang = torch.einsum("i,j->ij", torch.arange(seq_len).float(), inv_freq)
    return ang.cos(), ang.sin()


def rope_2d(head_dim: int, img_len: int, linear_factor: float = 1.0):
    side = int(math.sqrt(img_len))
    if side * side != img_len:
        raise ValueError(f"img_length must be a square number, got {img_len}")
    half = head_dim // 2
    freqs = 1.0 / (10000.0 ** (torch.arange(0, half, 2, dtype=torch.float32)[: half // 2] / half)) / linear_factor
    ang = torch.outer(torch.arange(side), freqs).float()
    ang_h = ang.view(side, 1, half // 2, 1).repeat(1, side, 1, 1)
    ang_w = ang.view(1, side, half // 2, 1).repeat(side, 1, 1, 1)
    a = torch.cat([ang_h, ang_w], dim=-1).flatten(2).flatten(0, 1)
    c = torch.polar(torch.ones_like(a), a)   # same arithmetic as the reference (.real / .imag of the polar form)
    return c.real.contiguous(), c.imag.contiguous()


def token_tables(modality: torch.Tensor, cos_txt, sin_txt, cos_img, sin_img, img_length: int):
    """Per-token cos/sin [B*N, hd/2] fp32 (reference models/dit.py:1419-1458, multimodal non-sample-ids branch):
    text tokens take the 1-D table at their absolute position, image tokens the 2-D table right-aligned to the
    end of the sequence."""
    B, N = modality.shape
    pos = torch.arange(N, device=modality.device)
    pad = max(N - img_length, 0)
    ipos = (pos - pad).clamp(min=0, max=img_length - 1)
    is_txt = (modality == 0)[..., None]
    cos = torch.where(is_txt, cos_txt[pos][None], cos_img[ipos][None])
    sin = torch.where(is_txt, sin_txt[pos][None], sin_img[ipos][None])
    return cos.reshape(B * N, -1).contiguous(), sin.reshape(B * N, -1).contiguous()
