"""TEST INFRASTRUCTURE — CPU restatement (pure torch / numpy) of the UniDisc hot path.

This file is the parity ORACLE.  It is imported only by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`; the product (`unidisc_b200`) never imports it.

Every function cites the reference file:line it follows (paths relative to the reference repo,
commit 01b6125c).  Pinning: `oracle/gen_golden.py` runs the *unmodified reference code* in the build
container (via `oracle/ref_loader.py`) on seeded inputs, asserts that this restatement reproduces it,
and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` re-checks the restatement against those
committed vectors everywhere (no /root/reference needed).  The reference itself ships no tests or
golden vectors for this path; the one piece whose pin is a restatement-of-a-dependency
is the 2-D Lumina RoPE table (diffusers 0.32.2, absent here) — "parity unpinned" for that table only.

Two numerics modes for the backbone:
  mode="fp32" : everything fp32 (reference with `DIT(dtype=torch.float32)`, no outer autocast).
  mode="bf16" : CUDA-autocast(bf16) semantics restated op by op with the rounding points of
                SURVEY.md §8(a'): GEMM inputs/outputs bf16 with fp32 accumulate, fp32 residual stream,
                fp32 norm maths, LayerNorm(q,k)->bf16, RoPE fp32 -> bf16, GELU on the bf16 GEMM output.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

NEG_INF = -1_000_000.0  # `self.neg_infinity`, model_setup.py (used at model.py:626)


# --------------------------------------------------------------------------------------
# configuration of the restated backbone
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    hidden_size: int
    n_heads: int
    n_blocks: int
    txt_length: int          # model.txt_length
    img_length: int          # model.img_length (square)
    vocab_size: int
    text_vocab_size: int
    mask_index: int
    linear_factor: float = 1.0
    rms_eps: float = 1e-6    # dit.py:78
    ln_eps: float = 1e-5     # nn.LayerNorm default, dit.py:569-571
    time_conditioning: bool = False
    cond_dim: int = 128
    require_sample_ids: bool = False   # data.require_sample_ids: interleaved batches (dit.py:1208-1216,1421-1443)

    @property
    def length(self):
        return self.txt_length + self.img_length

    @property
    def head_dim(self):
        return self.hidden_size // self.n_heads


# --------------------------------------------------------------------------------------
# RoPE tables — dit.py:307-330 (Rotary), dit.py:1046-1061 (get_2d_rope), dit.py:1203-1239
# --------------------------------------------------------------------------------------
def rope_table_1d(head_dim: int, seq_len: int):
    """cos/sin [seq_len, head_dim/2]: angle = pos * 10000^(-2i/head_dim)  (dit.py:310,320-324,1228-1230)."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(seq_len).float()
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    return freqs.cos(), freqs.sin()


def rope_table_2d(head_dim: int, img_len: int, linear_factor: float = 1.0):
    """cos/sin [img_len, head_dim/2], Lumina 2-D layout [row f0, col f0, row f1, col f1, ...]
    (dit.py:1046-1061 + diffusers 0.32.2 get_2d_rotary_pos_embed_lumina, restated)."""
    side = int(math.sqrt(img_len))
    assert side * side == img_len
    half = head_dim // 2
    freqs = 1.0 / (10000.0 ** (torch.arange(0, half, 2, dtype=torch.float32)[: half // 2] / half)) / linear_factor
    ang = torch.outer(torch.arange(side), freqs).float()        # [side, hd/4]
    ang_h = ang.view(side, 1, half // 2, 1).repeat(1, side, 1, 1)
    ang_w = ang.view(1, side, half // 2, 1).repeat(side, 1, 1, 1)
    a = torch.cat([ang_h, ang_w], dim=-1).flatten(2).flatten(0, 1)  # [img_len, hd/2]
    # reference takes .real/.imag of polar(1, angle)
    c = torch.polar(torch.ones_like(a), a)
    return c.real.contiguous(), c.imag.contiguous()


def token_cos_sin(cfg: OracleConfig, modality: torch.Tensor):
    """Per-token cos/sin [B,N,hd/2] (dit.py:1419-1458, non-sample-ids branch).

    Text tokens use the 1-D table at their absolute position; image tokens use the 2-D table
    right-aligned to the end of the sequence (NaN padding in the reference is never selected).
    """
    B, N = modality.shape
    hd = cfg.head_dim
    cos_t, sin_t = rope_table_1d(hd, cfg.length)
    cos_i, sin_i = rope_table_2d(hd, cfg.img_length, cfg.linear_factor)
    pos = torch.arange(N)
    pad = max(N - cfg.img_length, 0)
    ipos = (pos - pad).clamp(min=0, max=cfg.img_length - 1)
    is_txt = (modality == 0)[..., None]
    cos = torch.where(is_txt, cos_t[pos][None], cos_i[ipos][None])
    sin = torch.where(is_txt, sin_t[pos][None], sin_i[ipos][None])
    return cos.to(modality.device), sin.to(modality.device)


# interleaved batches (data.require_sample_ids): per-image-block 2-D tables, text position = offset within the sample
INTERLEAVED_IMG_TABLES = ((256, 1.0), (1024, 2.0), (2304, 3.0), (4096, 4.0))      # dit.py:1209


def _runs(flag_row):
    """[(start, end)) maximal runs of equal values in a 1-D tensor -> list of (start, end, value)."""
    v = flag_row.tolist()
    out, s = [], 0
    for i in range(1, len(v) + 1):
        if i == len(v) or v[i] != v[s]:
            out.append((s, i, v[s]))
            s = i
    return out


def interleaved_token_tables(cfg: OracleConfig, modality: torch.Tensor, sample_ids: torch.Tensor):
    """dit.py:1421-1443 with add_img_data_to_blocks (dit.py:122-178) and add_txt_data_to_blocks (dit.py:181-191).

    Returns (cos, sin) [B,N,hd/2] and `ordinal` int64 [B,N]: for tokens of an image block whose size has a table,
    the number of earlier image blocks of the same packed sample in that row (index into `img_count_embedding`),
    else -1.  Image blocks = maximal runs of modality==1 (unidisc/utils/tensor_utils.py:4-22, regardless of sample
    boundaries); sizes without a table (e.g. 64) keep cos=sin=0 and get no count embedding.  Text tokens of a run of
    equal sample_id >= 0 (tensor_utils.py:24-44) take the 1-D table at (position - run start); pad runs (-1) stay 0."""
    B, N = modality.shape
    hd = cfg.head_dim
    cos_t, sin_t = rope_table_1d(hd, cfg.length)
    tabs = {size: rope_table_2d(hd, size, lf) for size, lf in INTERLEAVED_IMG_TABLES}
    cos = torch.zeros(B, N, hd // 2)
    sin = torch.zeros(B, N, hd // 2)
    ordinal = torch.full((B, N), -1, dtype=torch.int64)
    for b in range(B):
        seen = []                                   # sample ids at the start of earlier image blocks of this row
        for s, e, m in _runs(modality[b] != 0):
            if not m:
                continue
            sid0 = int(sample_ids[b, s])
            size = e - s
            if size in tabs:                        # dit.py:131 (blocks of other sizes are skipped entirely)
                cos[b, s:e] = tabs[size][0][:size]
                sin[b, s:e] = tabs[size][1][:size]
                ordinal[b, s:e] = sum(1 for x in seen if x == sid0)   # dit.py:134-142
            seen.append(sid0)
        for s, e, sid in _runs(sample_ids[b]):
            if sid < 0:                             # tensor_utils.py:42
                continue
            txt = modality[b, s:e] == 0
            cos[b, s:e] = torch.where(txt[:, None], cos_t[: e - s], cos[b, s:e])       # dit.py:190
            sin[b, s:e] = torch.where(txt[:, None], sin_t[: e - s], sin[b, s:e])
    return cos, sin, ordinal


# --------------------------------------------------------------------------------------
# time conditioning (config.time_conditioning, off in every shipped config) — dit.py:415-449, 266-268, 301-304, 229-253
# --------------------------------------------------------------------------------------
def timestep_embedding(t, dim=256, max_period=10000):
    """TimestepEmbedder.timestep_embedding dit.py:426-444 (t: [B])."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def conditioning_vector(P, sigma, mode):
    """c = silu(sigma_map(sigma)) (dit.py:446-449, 1378-1379); under autocast every Linear / SiLU output is bf16."""
    e = timestep_embedding(sigma)
    h = _linear(e, P["sigma_map.mlp.0.weight"], P["sigma_map.mlp.0.bias"], mode)
    h = torch.nn.functional.silu(h.float())
    if mode == "bf16":
        h = _bf(h)
    h = _linear(h, P["sigma_map.mlp.2.weight"], P["sigma_map.mlp.2.bias"], mode)
    c = torch.nn.functional.silu(h.float())
    return _bf(c) if mode == "bf16" else c


def modulate_fused(x, shift, scale, modality, mode):
    """dit.py:258-268, 301-304: image tokens only — unless the batch holds no image token at all, in which case the
    reference modulates EVERY token (`modality.any()` false -> plain `modulate`).  `1 + scale` is evaluated in the dtype of
    `scale` (bf16 under autocast), the product / sum promote to fp32."""
    one_plus = (1 + scale)
    if mode == "bf16":
        one_plus = _bf(one_plus)
    y = x.float() * one_plus.float() + shift.float()
    if modality is not None and bool(modality.any()):
        return torch.where((modality == 1)[..., None], y, x.float())
    return y


# --------------------------------------------------------------------------------------
# backbone
# --------------------------------------------------------------------------------------
def _bf(x):
    return x.to(torch.bfloat16)


def _rms(x32, eps):
    return x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + eps)


def _linear(x, w, b, mode):
    """nn.Linear under autocast: inputs & weight rounded to bf16, fp32 accumulate, bf16 output."""
    if mode == "fp32":
        y = x.float() @ w.float().t()
        return y + b.float() if b is not None else y
    y = _bf(x).float() @ _bf(w).float().t()
    if b is not None:
        y = y + _bf(b).float()
    return _bf(y)


def _rmsnorm(x, w, eps, mode):
    """RMSNorm.forward dit.py:95-100: `_norm(x.float()).type_as(x) * weight`."""
    out = _rms(x.float(), eps)
    if mode == "bf16" and x.dtype == torch.bfloat16:
        out = _bf(out)            # `.type_as(x)` rounding point for bf16 inputs
    return out.float() * w.float()


def _rope(x, cos, sin):
    """standalone_rotary.py:14-31 (non-interleaved): x [B,N,H,hd], cos/sin [B,N,hd/2]."""
    cos2 = torch.cat([cos, cos], dim=-1)[:, :, None, :]
    sin2 = torch.cat([sin, sin], dim=-1)[:, :, None, :]
    x1, x2 = x.chunk(2, dim=-1)
    rot = torch.cat((-x2, x1), dim=-1)
    return x * cos2 + rot * sin2


def attention_core(q, k, v, scale):
    """softmax(q k^T * scale) v, bidirectional, fp32 math. q,k,v: [B,H,N,hd] (dit.py:826/829)."""
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    return p @ v.float()


def block_forward(cfg: OracleConfig, P: Dict[str, torch.Tensor], i: int, x, cos, sin, mode, sample_ids=None,
                  taps: Optional[dict] = None, drop_scale: Optional[torch.Tensor] = None, c=None, modality=None,
                  attn_mask=None, kv_cache: Optional[dict] = None, cache_op: Optional[str] = None, update_slice=None,
                  attend_cache: bool = False):
    """DDiTBlock.forward dit.py:948-1033 (rms, sandwich, qk_norm, no time-conditioning) with
    Attention.forward dit.py:616-887 (sdpa branch)."""
    pre = f"blocks.{i}."
    B, N, D = x.shape
    H, hd = cfg.n_heads, cfg.head_dim
    h = _rmsnorm(x, P[pre + "norm1.weight"], cfg.rms_eps, mode)                       # step 1
    if c is not None:                                                                 # dit.py:966-967, 973-974
        cond = _linear(c, P[pre + "adaLN_modulation.weight"], P[pre + "adaLN_modulation.bias"], mode)[:, None, :]
        shift_msa, scale_msa, _gate_msa, shift_mlp, scale_mlp, gate_mlp = cond.chunk(6, dim=2)
        h = modulate_fused(h, shift_msa, scale_msa, modality, mode)
    qkv = _linear(h, P[pre + "attention.attn_qkv.weight"], None, mode)               # step 2  [B,N,3D]
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    ln = lambda t, w, b: torch.nn.functional.layer_norm(t.float(), (D,), w.float(), b.float(), cfg.ln_eps)
    q = ln(q, P[pre + "attention.q_norm.weight"], P[pre + "attention.q_norm.bias"])   # step 3
    k = ln(k, P[pre + "attention.k_norm.weight"], P[pre + "attention.k_norm.bias"])
    if mode == "bf16":
        q, k = _bf(q), _bf(k)
    q = _rope(q.reshape(B, N, H, hd).float(), cos, sin)                               # step 4
    k = _rope(k.reshape(B, N, H, hd).float(), cos, sin)
    v = v.reshape(B, N, H, hd)
    if mode == "bf16":
        q, k = _bf(q), _bf(k)
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    if kv_cache is not None and cache_op is not None:
        # inference image-K/V cache of the FlexAttention path, dit.py:793-806.  "store" (step 1 of a caching cycle): the cache
        # becomes this step's full K, V.  "update" (steps 2..): the text-only forward writes its K, V into the cache's text
        # slice; the reference then calls flex_attention with the LOCAL k, v (dit.py:812 — the cache is written, not read);
        # attend_cache=True attends to the updated cache instead (what its comment at dit.py:795-797 describes).
        if cache_op == "store":
            kv_cache[i] = dict(k=kh.clone(), v=vh.clone())
        elif cache_op == "update":
            kv_cache[i]["k"][:, :, update_slice] = kh.to(kv_cache[i]["k"].dtype)
            kv_cache[i]["v"][:, :, update_slice] = vh.to(kv_cache[i]["v"].dtype)
            if attend_cache:
                kh, vh = kv_cache[i]["k"], kv_cache[i]["v"]
    if attn_mask is not None:
        s = (qh.float() @ kh.float().transpose(-1, -2)) / math.sqrt(hd)
        s = s.masked_fill(~attn_mask, float("-inf"))
        o = torch.softmax(s, dim=-1) @ vh.float()
    elif sample_ids is None:
        o = attention_core(qh, kh, vh, 1.0 / math.sqrt(hd))                          # step 5
    else:
        # document mask (model_utils.py:740-771): (sid[q]==sid[kv]) & (sid[q] != -1)
        s = (qh.float() @ kh.float().transpose(-1, -2)) / math.sqrt(hd)
        m = (sample_ids[:, :, None] == sample_ids[:, None, :]) & (sample_ids[:, :, None] != -1)
        s = s.masked_fill(~m[:, None], float("-inf"))
        p = torch.softmax(s, dim=-1)
        p = torch.nan_to_num(p, nan=0.0)
        o = p @ vh.float()
    if mode == "bf16":
        o = _bf(o)
    o = o.permute(0, 2, 1, 3).reshape(B, N, D)
    a = _linear(o, P[pre + "attention.attn_out.weight"], None, mode)                  # step 6
    x1 = x + _rmsnorm(a, P[pre + "pre_residual_norm.weight"], cfg.rms_eps, mode)      # step 7
    h2 = _rmsnorm(x1, P[pre + "norm2.weight"], cfg.rms_eps, mode)                     # step 8
    if c is not None:
        h2 = modulate_fused(h2, shift_mlp, scale_mlp, modality, mode)                 # dit.py:1017
    u = _linear(h2, P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"], mode)             # step 9
    g = torch.nn.functional.gelu(u.float(), approximate="tanh")
    if mode == "bf16":
        g = _bf(g)
    d = _linear(g, P[pre + "mlp.2.weight"], P[pre + "mlp.2.bias"], mode)
    br = _rmsnorm(d, P[pre + "post_ff_norm.weight"], cfg.rms_eps, mode)
    plain = br
    if drop_scale is not None:
        # F.dropout(training=True) of the branch (dit.py:218-222,239,1024-1031) with an explicit keep-scale tensor
        # (0 or 1/(1-p)) standing in for the Bernoulli draw
        br = br * drop_scale
    if c is not None:
        # bias_dropout_add_scale with modality (dit.py:239-249): image tokens get gate * dropout(branch), text tokens the
        # plain branch (no gate, no dropout)
        br = torch.where((modality == 1)[..., None], gate_mlp.float() * br, plain)
    x2 = x1 + br                                                                       # step 10
    if taps is not None:
        taps[f"b{i}"] = dict(h=h, qkv=qkv, q=q, k=k, o=o, a=a, x1=x1, h2=h2, u=u, g=g, d=d, x2=x2)
    return x2


def caching_step_mask(txt_length: int, N: int, device=None):
    """model_utils.py:721-738 `_attn_mask(txt_batch_dropout=False, img_batch_dropout=True)` — the mask of step 1 of an
    attention-caching cycle (model_eval.py:2329-2338): image queries see image keys only, text queries see everything."""
    qi = torch.arange(N, device=device)[:, None]
    ki = torch.arange(N, device=device)[None, :]
    return ((qi >= txt_length) & (ki >= txt_length)) | (qi < txt_length)


def dit_forward(cfg: OracleConfig, P: Dict[str, torch.Tensor], indices, modality, mode="fp32", sample_ids=None,
                taps: Optional[dict] = None, return_hidden=False, drop_scales=None, sigma=None, attn_mask=None,
                kv_cache: Optional[dict] = None, cache_op: Optional[str] = None, update_slice=None, attend_cache=False):
    """DIT.forward dit.py:1324-1500 (discrete, multimodal_batches, modality_embed, rope_2d, no time-cond).

    indices, modality: int64 [B,N].  Returns logits [B,N,V] (bf16 in mode="bf16", fp32 otherwise).
    """
    x = P["vocab_embed.embedding"].float()[indices]                                   # dit.py:1375
    me = P["modality_embed.embedding"].float()
    x = x + torch.where((modality == 0)[..., None], me[0][None, None], me[1][None, None])  # dit.py:1406
    if cfg.require_sample_ids:                                                        # dit.py:1421-1443
        cos, sin, ordinal = interleaved_token_tables(cfg, modality.cpu(), sample_ids.cpu())
        ice = P["img_count_embedding"].float()
        x = x + torch.where((ordinal >= 0)[..., None].to(x.device), ice[ordinal.clamp(min=0)], torch.zeros_like(x))  # dit.py:163-167
    else:
        cos, sin = token_cos_sin(cfg, modality.cpu())
    cos, sin = cos.to(x.device), sin.to(x.device)
    c = conditioning_vector(P, sigma, mode) if cfg.time_conditioning else None       # dit.py:1378-1379
    for i in range(cfg.n_blocks):
        x = block_forward(cfg, P, i, x, cos, sin, mode, sample_ids=sample_ids, taps=taps,
                          drop_scale=None if drop_scales is None else drop_scales[i], c=c, modality=modality,
                          attn_mask=attn_mask, kv_cache=kv_cache, cache_op=cache_op, update_slice=update_slice,
                          attend_cache=attend_cache)
    if return_hidden:
        return x
    hf = _rmsnorm(x, P["output_layer.norm_final.weight"], cfg.rms_eps, mode)          # dit.py:1089
    if c is not None:                                                                 # dit.py:1083-1087
        cond = _linear(c, P["output_layer.adaLN_modulation.weight"], P["output_layer.adaLN_modulation.bias"], mode)[:, None, :]
        shift, scale = cond.chunk(2, dim=2)
        hf = modulate_fused(hf, shift, scale, modality, mode)
    return _linear(hf, P["output_layer.linear.weight"], P["output_layer.linear.bias"], mode)  # dit.py:1091


# --------------------------------------------------------------------------------------
# noise schedule / time sampling — noise_schedule.py:128-157, model.py:589-619
# --------------------------------------------------------------------------------------
def loglinear_noise(t: torch.Tensor, eps: float = 1e-3):
    """(total_noise, rate_noise): sigma = -log1p(-(1-eps) t), dsigma = (1-eps)/(1-(1-eps) t)."""
    return -torch.log1p(-(1 - eps) * t), (1 - eps) / (1 - (1 - eps) * t)


def sample_t(u: torch.Tensor, sampling_eps: float = 1e-3, antithetic: bool = True):
    """model.py:589-619 given the uniform draw `u = torch.rand(n)`."""
    n = u.shape[0]
    e = u
    if antithetic:
        offset = torch.arange(n, device=u.device) / n
        e = (e / n + offset) % 1
    t = (1 - sampling_eps) * e + sampling_eps
    return t.to(torch.float32)


# --------------------------------------------------------------------------------------
# q_xt — model.py:424-587 (absorbing, multimodal non-interleaved, optional mask_entire_modality)
# --------------------------------------------------------------------------------------
def q_xt(x0: torch.Tensor, move_chance: torch.Tensor, rand: torch.Tensor, mask_index: int,
         modality_mask: Optional[torch.Tensor] = None, mask_entire_modality: Optional[float] = None,
         rand_txt: Optional[torch.Tensor] = None, rand_img: Optional[torch.Tensor] = None):
    """`rand` is the tensor the reference draws at model.py:439 (`torch.rand(*x.shape)`); `rand_txt`,
    `rand_img` [B,1] are the two draws at model.py:479-480.  Returns (xt, move_indices, ignore_mask)."""
    move = rand < move_chance                                                          # model.py:439
    ignore = None
    if mask_entire_modality is not None:
        smt = rand_txt < mask_entire_modality / 2                                      # model.py:479
        smi = rand_img < mask_entire_modality / 2                                      # model.py:480
        both = smt & smi                                                               # model.py:524-526
        smt = torch.where(both, False, smt)
        smi = torch.where(both, False, smi)
        move = torch.where(smt, modality_mask[..., 0], move)                           # model.py:527
        move = torch.where(smi, modality_mask[..., 1], move)                           # model.py:528
        ignore = smi | smt                                                             # model.py:529
    xt = torch.where(move, mask_index, x0)                                             # model.py:579
    return xt, move, ignore


def q_xt_interleaved(x0, move_chance, rand, mask_index, modality, sample_ids, mask_entire_modality, rand_blocks):
    """model.py:439 + 483-522 + 579 (trainer.interleaved with mask_entire_modality).  `rand` [B,N] is the draw at :439,
    `rand_blocks` [M,1] the per-block draw at :510 (M = number of (modality, sample) blocks longer than 4 tokens, in
    row-major order; the two [B,1] draws at :479-480 are consumed by the reference but unused on this branch).
    Returns (xt, move_indices, ignore [B])."""
    B, N = x0.shape
    move = (rand < move_chance).clone()
    blocks = []                                          # tensor_utils.py:46-68 + model.py:486-488
    for b in range(B):
        key = modality[b] * (int(sample_ids.max()) + 3) + (sample_ids[b] + 1)
        for s, e, _ in _runs(key):
            if int(sample_ids[b, s]) >= 0 and e - s > 4:
                blocks.append((b, s, e, int(sample_ids[b, s])))
    assert rand_blocks.shape[0] == len(blocks)
    ignore = torch.zeros(B, dtype=torch.bool)
    for i, (b, s, e, sid) in enumerate(blocks):
        k = sum(1 for (b2, _, _, sid2) in blocks[:i] if b2 == b and sid2 == sid)
        K = sum(1 for (b2, _, _, sid2) in blocks if b2 == b and sid2 == sid)
        if float(rand_blocks[i, 0]) < float(torch.tensor(mask_entire_modality, dtype=torch.float32) * torch.tensor((k + 1) / K, dtype=torch.float32) * 2):
            move[b, s:e] = True
            ignore[b] = True
    return torch.where(move, mask_index, x0), move, ignore


# --------------------------------------------------------------------------------------
# SUBS parameterization + loss — model.py:621-658, 797-1173
# --------------------------------------------------------------------------------------
def subs_parameterization(logits: torch.Tensor, xt: Optional[torch.Tensor], modality: torch.Tensor, mask_index: int,
                          text_vocab_size: int, force_argmax_valid_indices: bool = True):
    """model.py:621-658 (multimodal_batches).  Computed in the dtype of `logits` exactly as the reference
    does (bf16 logits -> bf16 log-softmax); pass fp32 logits for the high-precision variant."""
    logits = logits.clone()
    logits[..., mask_index] += NEG_INF                                                 # model.py:626
    if force_argmax_valid_indices:                                                     # model.py:627-635
        txt = (modality == 0)[..., None]
        img = (modality == 1)[..., None]
        neg = torch.tensor(NEG_INF, dtype=logits.dtype)
        logits[..., text_vocab_size:] = torch.where(txt, neg, logits[..., text_vocab_size:])
        logits[..., :text_vocab_size] = torch.where(img, neg, logits[..., :text_vocab_size])
    logits = logits - torch.logsumexp(logits, dim=-1, keepdim=True)                    # model.py:639
    if xt is not None:                                                                 # model.py:646-656
        unmasked = (xt != mask_index)[..., None]
        logits = torch.where(unmasked, torch.full_like(logits, NEG_INF), logits)
        onehot = torch.arange(logits.size(-1), device=logits.device) == xt[..., None]
        logits = torch.where(unmasked & onehot, torch.zeros_like(logits), logits)
    return logits


def diffusion_loss(model_output: torch.Tensor, x0, t, modality, attention_mask, *, text_loss_weight=1.0,
                   img_loss_weight=0.6, softmin_snr: Optional[float] = None, eps: float = 1e-3):
    """model.py:924-925, 967-993, 1021-1057 (discrete, subs, T=0).  model_output: log-probs [B,N,V].
    Returns dict(loss, txt_loss, img_loss, nlls [B,N], std_loss [B,N])."""
    out = model_output.to(torch.float32)                                               # model.py:924-925
    sigma, dsigma = loglinear_noise(t, eps)
    log_p = torch.gather(out, -1, x0[:, :, None]).squeeze(-1)                          # model.py:967
    std_w = (dsigma / torch.expm1(sigma))[:, None]                                     # model.py:975
    loss = -log_p * std_w
    if softmin_snr is not None:                                                        # model.py:990-993
        w = (dsigma / (torch.expm1(sigma) + (1 / softmin_snr)))[:, None]
        loss = -log_p * w
    std_loss = -log_p * std_w
    modality_mask = torch.stack([modality == 0, modality == 1], dim=-1)
    am = attention_mask.bool()
    txt_mask = modality_mask[..., 0] & am                                              # model.py:1021-1022
    img_mask = modality_mask[..., 1] & am
    txt_count, img_count = txt_mask.sum(), img_mask.sum()
    total = txt_count + img_count
    loss = loss * am
    txt_loss = (loss[txt_mask].sum() / txt_count) * (txt_count / total) * text_loss_weight   # model.py:1038-1040
    img_loss = (loss[img_mask].sum() / img_count) * (img_count / total) * img_loss_weight    # model.py:1041-1043
    txt_loss = torch.nan_to_num(txt_loss, nan=0.0)
    img_loss = torch.nan_to_num(img_loss, nan=0.0)
    return dict(loss=txt_loss + img_loss, txt_loss=txt_loss, img_loss=img_loss, nlls=std_loss * am, std_loss=std_loss,
                log_p=log_p)


# --------------------------------------------------------------------------------------
# samplers — model_utils.py:95-97, model_eval.py:1734-1833, 2042-2104, 2964-3114
# --------------------------------------------------------------------------------------
def sample_categorical(probs: torch.Tensor, u: torch.Tensor):
    """model_utils.py:95-97 given `u = torch.rand_like(probs)`."""
    gumbel_norm = 1e-10 - (u + 1e-10).log()
    return (probs / gumbel_norm).argmax(dim=-1)


def ddpm_caching_update(x, t, dt, p_x0, u, mask_index):
    """model_eval.py:2072-2104 (loglinear): x [B,N] int64, t [B] or [B,1], p_x0 [B,N,V], u = rand_like(q_xs)."""
    if t.ndim > 1:
        t = t.squeeze(-1)
    mc_t = t[:, None, None]
    mc_s = (t - dt)[:, None, None]
    q_xs = p_x0 * (mc_t - mc_s)                                                        # model_eval.py:2091
    q_xs[:, :, mask_index] = mc_s[:, :, 0]                                             # model_eval.py:2092
    _x = sample_categorical(q_xs, u)
    copy_flag = (x != mask_index).to(x.dtype)
    return copy_flag * x + (1 - copy_flag) * _x                                        # model_eval.py:2104


def ddpm_update(x, t, dt, p_x0, u, mask_index, eps=1e-3):
    """model_eval.py:2042-2070: move chances from the noise schedule instead of t directly."""
    if t.ndim > 1:
        t = t.squeeze(-1)
    sigma_t, _ = loglinear_noise(t, eps)
    sigma_s, _ = loglinear_noise(t - dt, eps)
    mc_t = (1 - torch.exp(-sigma_t))[:, None, None]
    mc_s = (1 - torch.exp(-sigma_s))[:, None, None]
    q_xs = p_x0 * (mc_t - mc_s)
    q_xs[:, :, mask_index] = mc_s[:, :, 0]
    _x = sample_categorical(q_xs, u)
    copy_flag = (x != mask_index).to(x.dtype)
    return copy_flag * x + (1 - copy_flag) * _x


def cfg_combine(logit_c, logit_u, t, cfg_scale):
    """model_eval.py:1746,1812: w = cfg (1 - t); logits = (1+w) c - w u."""
    w = (cfg_scale * (1 - t))[:, None]
    if w.ndim == 2 and logit_c.ndim == 3:
        w = w.unsqueeze(-1)
    return (1 + w) * logit_c - w * logit_u


def adap_sche(x, step, mask_index, mode="arccos"):
    """model_eval.py:2964-3001: per-sample number of tokens to unmask at each step."""
    num_masked = (x == mask_index).sum(dim=-1)
    r = torch.linspace(1, 0, step)
    if mode == "root":
        val = 1 - (r ** 0.5)
    elif mode == "linear":
        val = 1 - r
    elif mode == "square":
        val = 1 - (r ** 2)
    elif mode == "cosine":
        val = torch.cos(r * math.pi * 0.5)
    elif mode == "arccos":
        val = torch.arccos(r) / (math.pi * 0.5)
    else:
        return None
    out = []
    for seq_len in num_masked:
        sche = (val / val.sum()) * seq_len
        sche = sche.round()
        sche[sche == 0] = 1
        sche[-1] += seq_len - sche.sum()
        sche[-1] = max(sche[-1], 0)
        out.append(sche.int())
    return torch.stack(out, dim=0)


def maskgit_update(x, t, p_x0, pred_code, gumbel, num_unmask, mask_index, r_temp=10.0):
    """model_eval.py:3045-3114 given the two random draws the reference makes:
    `pred_code` = torch.multinomial(p_x0) [B,N] and `gumbel` = np.random.gumbel [B,N].  t: [B,1]."""
    copy_flag = x != mask_index
    num_unmask = torch.minimum(num_unmask, (~copy_flag).sum(dim=-1))
    if torch.all(num_unmask <= 0):
        return x
    conf = torch.gather(p_x0, -1, pred_code.unsqueeze(-1)).squeeze(-1)
    rand = r_temp * gumbel * t
    conf = torch.log(conf.squeeze()) + rand
    conf = torch.where(copy_flag, -torch.inf, conf)
    k = int(num_unmask.max().item())
    tresh, _ = torch.topk(conf, k=k, dim=-1)
    gi = torch.clamp(num_unmask - 1, min=0)[:, None]
    tresh = tresh.gather(-1, gi)
    tresh = torch.where((num_unmask <= 0)[:, None], torch.inf, tresh)
    sel = conf >= tresh.expand_as(conf)
    return torch.where(sel, pred_code, x)


def multinomial_from_exponential(p_x0, e_noise):
    """`torch.multinomial(p.view(-1, V), 1)` (model_eval.py:3073) given the Exp(1) tensor it draws internally: ATen's
    single-sample path is `q = empty_like(p).exponential_(1); argmax(p / q)` (aten/src/ATen/native/TensorCompare /
    Distributions `multinomial`), on CPU and CUDA alike."""
    V = p_x0.shape[-1]
    return (p_x0.reshape(-1, V) / e_noise.reshape(-1, V)).argmax(dim=-1).view(p_x0.shape[:-1])


def maskgit_update_from_noise(x, t, p_x0, e_noise, gumbel, num_unmask, mask_index, r_temp=10.0):
    """model_eval.py:3045-3114 driven by the two raw noise tensors (Exp(1) [B,N,V] fp32 and Gumbel [B,N] fp64)."""
    return maskgit_update(x, t, p_x0, multinomial_from_exponential(p_x0, e_noise), gumbel, num_unmask, mask_index, r_temp=r_temp)


# --------------------------------------------------------------------------------------
# synthetic inputs and parameter init mirroring the reference module
# --------------------------------------------------------------------------------------
def synthetic_batch(B, txt_len, img_len, text_vocab_size, vocab_size, seed=42):
    g = torch.Generator().manual_seed(seed)
    txt = torch.randint(0, text_vocab_size - 1, (B, txt_len), generator=g)
    img = torch.randint(text_vocab_size, vocab_size, (B, img_len), generator=g)
    ids = torch.cat([txt, img], dim=1)
    modality = torch.cat([torch.zeros(B, txt_len, dtype=torch.int64), torch.ones(B, img_len, dtype=torch.int64)], dim=1)
    return ids, modality


def init_params(cfg: OracleConfig, seed=0) -> Dict[str, torch.Tensor]:
    """Random parameters with the reference's state-dict keys/shapes — values are arbitrary
    (seeded normal/uniform); parity tests copy the SAME tensors into both implementations."""
    g = torch.Generator().manual_seed(seed)
    D, V, L = cfg.hidden_size, cfg.vocab_size, cfg.n_blocks
    u = lambda *s, a: (torch.rand(*s, generator=g) * 2 - 1) * a
    P = {"vocab_embed.embedding": u(V, D, a=1.0 / math.sqrt(D)),
         "modality_embed.embedding": u(2, D, a=1.0 / math.sqrt(D))}
    for i in range(L):
        p = f"blocks.{i}."
        P[p + "attention.attn_qkv.weight"] = u(3 * D, D, a=1.0 / math.sqrt(D))
        P[p + "attention.attn_out.weight"] = u(D, D, a=1.0 / math.sqrt(D))
        for n in ("q_norm", "k_norm"):
            P[p + f"attention.{n}.weight"] = 1.0 + u(D, a=0.2)
            P[p + f"attention.{n}.bias"] = u(D, a=0.1)
        for n in ("norm1", "norm2", "post_ff_norm", "pre_residual_norm"):
            P[p + f"{n}.weight"] = 1.0 + u(D, a=0.2)
        P[p + "mlp.0.weight"] = u(4 * D, D, a=1.0 / math.sqrt(D))
        P[p + "mlp.0.bias"] = u(4 * D, a=0.1)
        P[p + "mlp.2.weight"] = u(D, 4 * D, a=1.0 / math.sqrt(4 * D))
        P[p + "mlp.2.bias"] = u(D, a=0.1)
    P["output_layer.norm_final.weight"] = 1.0 + u(D, a=0.2)
    P["output_layer.linear.weight"] = u(V, D, a=1.0 / math.sqrt(D))
    P["output_layer.linear.bias"] = u(V, a=0.1)
    return P


def training_loss(cfg: OracleConfig, P, x0, modality, attention_mask, u_t, rand_move, mode="fp32", *,
                  img_loss_weight=0.6, text_loss_weight=1.0, softmin_snr=None, fp32_logsoftmax=True, drop_scales=None):
    """q_xt -> DIT -> SUBS -> weighted NLL, i.e. `Diffusion.compute_loss` (model.py:797-1173) for the default
    large-scale config, driven by explicit random draws (u_t [B], rand_move [B,N]).  With time conditioning sigma is handed
    to the backbone exactly as compute_loss does (model.py:858-859)."""
    t = sample_t(u_t)
    sigma, _ = loglinear_noise(t)
    move_chance = 1 - torch.exp(-sigma[:, None])                                       # model.py:858-860
    xt, move, _ = q_xt(x0, move_chance, rand_move, cfg.mask_index)
    logits = dit_forward(cfg, P, xt, modality, mode=mode, drop_scales=drop_scales, sigma=sigma if cfg.time_conditioning else None)
    if fp32_logsoftmax:
        logits = logits.float()
    logp = subs_parameterization(logits, xt, modality, cfg.mask_index, cfg.text_vocab_size)
    out = diffusion_loss(logp, x0, t, modality, attention_mask, text_loss_weight=text_loss_weight,
                         img_loss_weight=img_loss_weight, softmin_snr=softmin_snr)
    out.update(xt=xt, t=t, logits=logits)
    return out
