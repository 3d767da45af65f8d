"""TEST / BASELINE INFRASTRUCTURE — not part of the product path (only tests/ and bench.py's reference arms import it).

A device-agnostic `nn.Module` restatement of the reference backbone + training step built from the SAME library ops the
reference issues in its default large-scale configuration (`use_spda_attn`, `qk_norm`, `sandwich_normalization`,
`norm_type=rms`, `rope_2d`, `modality_embed`, no time conditioning), so that it can be TIMED as "the reference's own
torch-SDPA path" on whatever device it is placed on (the reference itself cannot travel to the GPU box):

  nn.Linear / F.layer_norm / F.scaled_dot_product_attention / F.gelu(tanh) / F.dropout under torch.autocast(bf16)
  (reference models/dit.py:77-100 RMSNorm, :616-887 Attention.forward sdpa branch, :948-1033 DDiTBlock.forward,
   :1063-1092 DDitFinalLayer, :1324-1500 DIT.forward; models/standalone_rotary.py:14-31 rotary),
  SUBS parameterisation + weighted NLL done the reference's way, i.e. materialising the [B,N,V] log-prob tensor
  (reference model.py:621-658, 797-1173), torch.optim.AdamW(fused) + clip_grad_norm_ (model_setup.py:385-424,
  model.py:1518-1537).

Deviation (documented in SURVEY.md §8c): the reference's in-place q/k-norm write (dit.py:680-682) raises in eager
autograd; the out-of-place `torch.cat` form of the reference's own XLA branch (dit.py:675-678) is used instead.
State-dict keys equal the reference's, so parameters from `oracle.restated.init_params` load directly; parity with
`oracle.restated.dit_forward` (itself pinned against the unmodified reference) is asserted in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import restated as R


class RMSNorm(nn.Module):                                             # dit.py:77-100
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        x32 = x.float()
        out = (x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x)
        return out * self.weight


def rotate_half(x):                                                   # standalone_rotary.py:5-11
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary(x, cos, sin):                                        # standalone_rotary.py:14-31; x [B,N,G,hd], cos [B,N,hd/2]
    cos = torch.cat([cos, cos], dim=-1)[:, :, None, :]
    sin = torch.cat([sin, sin], dim=-1)[:, :, None, :]
    return x * cos + rotate_half(x) * sin


class Attention(nn.Module):                                           # dit.py:562-571, 616-887
    def __init__(self, dim, n_heads):
        super().__init__()
        self.n_heads, self.head_dim = n_heads, dim // n_heads
        self.attn_qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.attn_out = nn.Linear(dim, dim, bias=False)
        self.q_norm = nn.LayerNorm(dim)
        self.k_norm = nn.LayerNorm(dim)

    def forward(self, x, cos, sin):
        B, N, D = x.shape
        H, hd = self.n_heads, self.head_dim
        qkv = self.attn_qkv(x)                                                        # :642
        qkv = torch.cat([self.q_norm(qkv[:, :, :D]), self.k_norm(qkv[:, :, D:2 * D]), qkv[:, :, 2 * D:]], dim=-1)  # :675-682
        qkv = qkv.view(B, N, 3, H, hd)                                                # :699
        orig = qkv.dtype
        with torch.autocast(x.device.type, enabled=False):                            # :703
            qk = apply_rotary(qkv[:, :, :2].reshape(B, N, 2 * H, hd), cos, sin)      # :724-726 (bf16 x fp32 -> fp32)
        qk = qk.to(orig).view(B, N, 2, H, hd)
        q, k, v = qk[:, :, 0], qk[:, :, 1], qkv[:, :, 2]
        q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)             # :782
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=False, scale=1.0 / math.sqrt(hd))   # :826/829
        o = o.transpose(1, 2).reshape(B, N, D)                                        # :846
        return self.attn_out(o)                                                       # :887 (sandwich: no residual here)


class DDiTBlock(nn.Module):                                           # dit.py:890-1033
    def __init__(self, dim, n_heads, dropout):
        super().__init__()
        self.attention = Attention(dim, n_heads)
        self.norm1, self.norm2 = RMSNorm(dim), RMSNorm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, 4 * dim), nn.GELU(approximate="tanh"), nn.Linear(4 * dim, dim))
        self.post_ff_norm, self.pre_residual_norm = RMSNorm(dim), RMSNorm(dim)
        self.dropout = dropout

    def forward(self, x, cos, sin):
        x_skip = x
        a = self.attention(self.norm1(x), cos, sin)
        x = x_skip + self.pre_residual_norm(a)                                        # :993-994
        br = self.post_ff_norm(self.mlp(self.norm2(x)))                               # :1025
        if self.dropout > 0.0:
            br = F.dropout(br, p=self.dropout, training=self.training)                # :218-222,239
        return x + br                                                                 # :250-251


class _Embedding(nn.Module):                                          # dit.py:1036-1043
    def __init__(self, dim, vocab):
        super().__init__()
        self.embedding = nn.Parameter(torch.empty(vocab, dim))
        nn.init.kaiming_uniform_(self.embedding, a=math.sqrt(5))

    def forward(self, idx):
        return self.embedding[idx]


class _FinalLayer(nn.Module):                                         # dit.py:1063-1092
    def __init__(self, dim, vocab):
        super().__init__()
        self.norm_final = RMSNorm(dim)
        self.linear = nn.Linear(dim, vocab)

    def forward(self, x):
        return self.linear(self.norm_final(x))


class EagerDIT(nn.Module):
    def __init__(self, cfg: R.OracleConfig, dropout=0.0):
        super().__init__()
        self.cfg = cfg
        D = cfg.hidden_size
        self.vocab_embed = _Embedding(D, cfg.vocab_size)
        self.modality_embed = _Embedding(D, 2)
        self.blocks = nn.ModuleList([DDiTBlock(D, cfg.n_heads, dropout) for _ in range(cfg.n_blocks)])
        self.output_layer = _FinalLayer(D, cfg.vocab_size)
        ct, st = R.rope_table_1d(cfg.head_dim, cfg.txt_length + cfg.img_length)
        ci, si = R.rope_table_2d(cfg.head_dim, cfg.img_length)
        for n, t in (("rotary_cos_emb_txt", ct), ("rotary_sin_emb_txt", st), ("rotary_cos_emb_img", ci), ("rotary_sin_emb_img", si)):
            self.register_buffer(n, t.float().contiguous(), persistent=False)

    def _cos_sin(self, modality):                                                     # dit.py:1419-1458
        B, N = modality.shape
        pos = torch.arange(N, device=modality.device)
        il = self.cfg.img_length
        ipos = (pos - (N - il)).clamp(0, il - 1)
        m0 = (modality == 0)[..., None]
        cos = torch.where(m0, self.rotary_cos_emb_txt[pos][None], self.rotary_cos_emb_img[ipos][None])
        sin = torch.where(m0, self.rotary_sin_emb_txt[pos][None], self.rotary_sin_emb_img[ipos][None])
        return cos, sin

    def forward(self, indices, modality):
        x = self.vocab_embed(indices)                                                 # :1375
        me = self.modality_embed.embedding
        x = x + torch.where((modality == 0)[..., None], me[0], me[1])                 # :1406
        cos, sin = self._cos_sin(modality)
        with torch.autocast(x.device.type, dtype=torch.bfloat16, enabled=torch.is_autocast_enabled(x.device.type)):  # :1484
            for blk in self.blocks:
                x = blk(x, cos, sin)
            return self.output_layer(x)                                               # :1495


def reference_style_loss(model: EagerDIT, x0, modality, attention_mask, mask_index, text_vocab_size, *, img_loss_weight=0.6,
                         text_loss_weight=1.0, autocast=True, generator=None):
    """One `Diffusion.compute_loss` (model.py:797-1173) the way the reference executes it: `_sample_t`, `q_xt` with
    torch.rand, backbone under autocast(bf16), SUBS on the materialised [B,N,V] tensor in the backbone's dtype, gather, loss
    weighting.  Returns the scalar loss."""
    B, N = x0.shape
    dev = x0.device
    u = torch.rand(B, device=dev, generator=generator)
    t = R.sample_t(u)
    sigma, dsigma = R.loglinear_noise(t)
    move_chance = 1 - torch.exp(-sigma[:, None])
    xt, _, _ = R.q_xt(x0, move_chance, torch.rand(B, N, device=dev, generator=generator), mask_index)
    with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast):            # model.py:693-696
        logits = model(xt, modality)
        logp = R.subs_parameterization(logits, xt, modality, mask_index, text_vocab_size)   # model.py:784-789 (bf16 under autocast)
    logp = logp.float() if logp.dtype != torch.float32 else logp                      # model.py:924-925
    out = R.diffusion_loss(logp, x0, t, modality, attention_mask, text_loss_weight=text_loss_weight, img_loss_weight=img_loss_weight)
    return out["loss"]
