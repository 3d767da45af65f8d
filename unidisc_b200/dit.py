"""B200-native drop-in for the reference backbone `models.dit.DIT` (reference models/dit.py:1095-1500).

Same constructor signature, `forward` keyword surface, sub-module / state-dict key names (`vocab_embed.embedding`,
`blocks.{i}.attention.attn_qkv.weight`, ... SURVEY.md §8b) and class names the launcher keys on (`DDiTBlock`,
`EmbeddingLayer`).  The sub-modules only HOLD the fp32 master parameters; all maths runs in the CUDA library
(libunidisc_b200.so) through `unidisc_b200.ops`:

  embed+RMSNorm -> L x [ qkv GEMM -> q/k LayerNorm+RoPE -> tcgen05 attention -> out GEMM -> norm+residual+norm2
                         -> MLP GEMM(+bias+GELU) -> MLP GEMM(+bias) -> norm+residual+next-norm ] -> head GEMM(+bias)

with the rounding points of the reference's CUDA bf16-autocast execution (SURVEY.md §8a').  Parameters live in one flat
fp32 buffer (with a bf16 shadow for the tensor cores and a flat fp32 gradient buffer the backward kernels accumulate
into), so the optimizer and the DDP all-reduce work on contiguous memory.  There is no eager / CPU fallback.

Supported configuration = the shipped large/small-scale training configs: norm_type=rms, sandwich_normalization,
qk_norm, rope_2d, modality_embed, multimodal_batches, full_attention; `model.dropout` is applied to the MLP branch in
training mode exactly where the reference does (in-kernel Philox mask, regenerated in backward).  `time_conditioning`
(adaLN shift / scale / gate from sigma; off in every shipped config) is fused into the same norm kernels.
`data.require_sample_ids` (interleaved / packed batches): per-image-block RoPE tables, `img_count_embedding` and the
document mask are derived on the device from `modality` / `sample_ids` (csrc/interleaved.cu).
(see DESIGN.md for what is not covered: KV caches).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops, rope

bf16 = torch.bfloat16


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _wrap_cfg(c):
    if isinstance(c, dict) and not isinstance(c, _AttrDict):
        return _AttrDict({k: _wrap_cfg(v) for k, v in c.items()})
    return c


# ----------------------------------------------------------------------------------------------------------------
# parameter containers (names / shapes / init follow the reference so checkpoints load unchanged)
# ----------------------------------------------------------------------------------------------------------------
class EmbeddingLayer(nn.Module):                                     # reference dit.py:1036-1043
    def __init__(self, dim, vocab_dim):
        super().__init__()
        self.embedding = nn.Parameter(torch.empty((vocab_dim, dim)))
        torch.nn.init.kaiming_uniform_(self.embedding, a=math.sqrt(5))


class RMSNorm(nn.Module):                                            # reference dit.py:77-100
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class Attention(nn.Module):                                          # parameters of reference dit.py:562-571
    def __init__(self, dim, n_heads):
        super().__init__()
        self.n_heads = n_heads
        self.head_dim = dim // n_heads
        self.attn_qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.attn_out = nn.Linear(dim, dim, bias=False)
        self.q_norm = nn.LayerNorm(dim)
        self.k_norm = nn.LayerNorm(dim)


class TimestepEmbedder(nn.Module):                                   # parameters of reference dit.py:415-449
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size


def _zero_linear(i, o):
    lin = nn.Linear(i, o, bias=True)                                  # adaLN_modulation: zero-initialised (dit.py:923-925)
    lin.weight.data.zero_()
    lin.bias.data.zero_()
    return lin


class DDiTBlock(nn.Module):                                          # parameters of reference dit.py:890-934
    def __init__(self, dim, n_heads, mlp_ratio=4, cond_dim=None):
        super().__init__()
        self.attention = Attention(dim, n_heads)
        self.norm1 = RMSNorm(dim)
        self.norm2 = RMSNorm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, mlp_ratio * dim, bias=True), nn.GELU(approximate="tanh"),
                                 nn.Linear(mlp_ratio * dim, dim, bias=True))
        if cond_dim is not None:
            self.adaLN_modulation = _zero_linear(cond_dim, 6 * dim)
        self.post_ff_norm = RMSNorm(dim)
        self.pre_residual_norm = RMSNorm(dim)


class DDitFinalLayer(nn.Module):                                     # parameters of reference dit.py:1063-1092
    def __init__(self, hidden_size, out_channels, zero_linear_init=True, cond_dim=None):
        super().__init__()
        self.norm_final = RMSNorm(hidden_size)
        self.linear = nn.Linear(hidden_size, out_channels)
        if zero_linear_init:
            self.linear.weight.data.zero_()
        self.linear.bias.data.zero_()
        if cond_dim is not None:
            self.adaLN_modulation = _zero_linear(cond_dim, 2 * hidden_size)


# ----------------------------------------------------------------------------------------------------------------
# autograd glue: one Function for the whole backbone.  Parameter gradients are accumulated by the kernels directly
# into the module's flat gradient buffer (p.grad are views of it); autograd only carries the logits gradient.
# ----------------------------------------------------------------------------------------------------------------
class _DiTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, module, indices, modality, sample_ids, save, doc_mask, sigma, head_rows=None, head_split=None):
        logits, saved = module._forward_impl(indices, modality, sample_ids, save=save, doc_mask=doc_mask, sigma=sigma,
                                             head_rows=head_rows, head_split=head_split)
        ctx.module = module
        ctx.saved = saved
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        module, saved = ctx.module, ctx.saved
        if saved is None:
            raise RuntimeError("unidisc_b200.DIT: backward called on a forward that did not save activations")
        module._backward_impl(saved, dlogits)
        ctx.saved = None
        return torch.zeros_like(module._anchor), None, None, None, None, None, None, None, None, None


class TextFullImageSelfMask:
    """Stands in for the FlexAttention BlockMask the reference builds for step 1 of an inference attention-caching cycle
    (`get_block_mask(txt_batch_attn_dropout=False, img_batch_attn_dropout=True)`, model_utils.py:721-738,
    model_eval.py:2329-2338): text queries attend to every token, image queries to image tokens only.  The layout is the static
    one of the shipped configs: the first `txt_length` tokens are text."""

    def __init__(self, txt_length: int):
        self.txt_length = int(txt_length)


class DIT(nn.Module):
    def __init__(self, config, vocab_size: int, text_vocab_size: int, mask_index: int, dtype=None, device=None,
                 static_img_sl=None, static_txt_sl=None, **kwargs):
        super().__init__()
        config = _wrap_cfg(config)
        self.config = config
        self.autocast_dtype = dtype
        self.vocab_size = vocab_size
        self.text_vocab_size = text_vocab_size
        self.mask_index = mask_index
        self.static_img_sl, self.static_txt_sl = static_img_sl, static_txt_sl
        m = config.model
        g = lambda o, k, d=None: getattr(o, k, d) if not isinstance(o, dict) else o.get(k, d)
        self.time_conditioning = bool(config.time_conditioning or g(m, "force_time_conditioning", False))
        unsupported = []
        if g(m, "norm_type", "rms") != "rms":
            unsupported.append(f"norm_type={g(m, 'norm_type')}")
        if not g(m, "sandwich_normalization", False):
            unsupported.append("sandwich_normalization=False")
        if not g(m, "qk_norm", False):
            unsupported.append("qk_norm=False")
        if not g(m, "rope_2d", False) or not g(m, "modality_embed", False) or not g(config.trainer, "multimodal_batches", False):
            unsupported.append("rope_2d / modality_embed / multimodal_batches must be on")
        if not g(m, "full_attention", True):
            unsupported.append("causal attention")
        if g(m, "img_cond", False) or g(m, "use_pretrained_img_emb", False) or g(m, "cond_label", False) \
                or g(config.trainer, "image_mode", "discrete") == "continuous":
            unsupported.append("img_cond / pretrained image embedding / label conditioning / continuous image mode")
        if unsupported:
            raise NotImplementedError("unidisc_b200.DIT does not implement: " + "; ".join(unsupported))

        D, H, nb = m.hidden_size, m.n_heads, m.n_blocks
        self.hidden_size, self.n_heads, self.n_blocks = D, H, nb
        self.head_dim = D // H
        if self.head_dim not in (64, 128):
            raise NotImplementedError(f"head_dim {self.head_dim}: the attention kernels are built for 64 and 128")
        if D % 128 != 0:
            raise NotImplementedError("hidden_size must be a multiple of 128")
        self.dropout = float(g(m, "dropout", 0.0))
        self.dropout_seed = int(g(config, "seed", 42))
        self._dropout_calls = 0
        self._dropout_rank = None
        self.txt_length, self.img_length, self.total_length = m.txt_length, m.img_length, m.length
        self.multimodal_batches = True
        self.rope_2d = True
        self.require_sample_ids = bool(g(config.data, "require_sample_ids", False))
        self.use_gradient_checkpointing = bool(g(config.trainer, "use_gradient_checkpointing", False))

        self.vocab_embed = EmbeddingLayer(D, vocab_size)
        self.modality_embed = EmbeddingLayer(D, 2)
        cond_dim = int(m.cond_dim) if self.time_conditioning else None
        self.cond_dim = cond_dim
        self.blocks = nn.ModuleList([DDiTBlock(D, H, cond_dim=cond_dim) for _ in range(nb)])
        self.output_layer = DDitFinalLayer(D, vocab_size, zero_linear_init=bool(g(m, "zero_linear_init", True)), cond_dim=cond_dim)
        self.sigma_map = TimestepEmbedder(cond_dim) if self.time_conditioning else None      # dit.py:1186-1188

        lf = g(m, "linear_factor", 1.0)
        ct, st = rope.rope_1d(self.head_dim, self.total_length)
        if self.require_sample_ids:
            # interleaved batches: one 2-D table per supported image-block size, count embedding (dit.py:1208-1216)
            cat_c, cat_s, off = [ct], [st], {"txt": 0}
            rows = ct.shape[0]
            for size, factor in rope.INTERLEAVED_IMG_TABLES:
                ci, si = rope.rope_2d(self.head_dim, size, factor)
                self.register_buffer(f"rotary_cos_emb_img_{size}", ci, persistent=False)
                self.register_buffer(f"rotary_sin_emb_img_{size}", si, persistent=False)
                off[f"s{size}"] = rows
                rows += size
                cat_c.append(ci)
                cat_s.append(si)
            self._rope_offsets = off
            self.register_buffer("_rope_cat_cos", torch.cat(cat_c, 0).contiguous(), persistent=False)
            self.register_buffer("_rope_cat_sin", torch.cat(cat_s, 0).contiguous(), persistent=False)
            self.img_count_embedding = nn.Parameter(torch.zeros((16, D)))
        else:
            ci, si = rope.rope_2d(self.head_dim, self.img_length, lf)
            self.register_buffer("rotary_cos_emb_img", ci, persistent=False)
            self.register_buffer("rotary_sin_emb_img", si, persistent=False)
        self.register_buffer("rotary_cos_emb_txt", ct.contiguous(), persistent=False)
        self.register_buffer("rotary_sin_emb_txt", st.contiguous(), persistent=False)

        self.Vp = (vocab_size + 63) // 64 * 64     # logits row pitch (TMA / 16-byte vector stores)
        self.supports_head_rows = True             # forward(head_rows=...): output projection of a subset of the token rows
        self._flat_p = None
        self._flat_g = None
        self._flat_bf16 = None
        self._shadow_dirty = True
        self._shadow_versions = -1
        self._anchor = None
        self.training_graph_enabled = True
        self.grad_ready_hook = None                # thin-DDP / optimizer: hook(block) when that bucket's flat grads are final
        self._param_events = None                  # FusedAdamW(overlap): per-bucket "weights updated" events of the last step
        self._dbg_fwd_events = None                # debug: list collecting a timing event at the start of every block's forward
        self.grad_sumsq_acc = None                 # FusedAdamW (N=1): fp32 scalar the wgrad GEMM epilogues add sum(dW^2) into
        self._last_bwd_fused_sumsq = False
        self._grads_attached = False
        self._kv_cache = None                      # inference attention caching (set_flex_attention_cache)
        # ThinDDP (world > 1): weight gradients are written by the wgrad GEMM epilogue straight into a bf16 staging buffer in the
        # wire format of the bf16 compress hook (attach_grad_stage); _last_bwd_staged tells the DDP hook which path the last
        # backward took
        self._ddp_stage = None
        self._ddp_alpha = None
        self._ddp_sync_fn = None
        self._stage_views = None
        self._last_bwd_staged = False
        # additive key: with eval.attention_caching the reference's text-only steps write the cache but attend to the LOCAL text
        # K/V (dit.py:798-812); True makes them attend to the cached image K/V as its comment describes
        self.cache_attend_cached = bool(g(g(config, "eval"), "attention_caching_attend_cache", False))
        if device is not None:
            self.to(device)

    # ------------------------------------------------------------------------------------------------------------
    # reference API stubs
    # ------------------------------------------------------------------------------------------------------------
    def reset_kv_cache(self, *a, **k):
        raise NotImplementedError("unidisc_b200.DIT: the causal start_pos KV cache (model.use_kv_cache, dit.py:588-600) belongs to the "
                                  "autoregressive parameterisation, which is off the hot path; use set_flex_attention_cache")

    def set_flex_attention_cache(self, batch_size, seq_len, device=None, dtype=None):
        """reference dit.py:610-614 / 1320-1322: allocate the per-block K/V cache of an inference attention-caching cycle
        (model_eval.py:2297-2367).  K is cached AFTER q/k-LayerNorm + RoPE, V as projected: bf16 [B*seq_len, D] per block."""
        dev = device if device is not None else self.vocab_embed.embedding.device
        D = self.hidden_size
        self._kv_cache = dict(B=int(batch_size), N=int(seq_len),
                              k=[torch.zeros((batch_size * seq_len, D), device=dev, dtype=bf16) for _ in range(self.n_blocks)],
                              v=[torch.zeros((batch_size * seq_len, D), device=dev, dtype=bf16) for _ in range(self.n_blocks)])

    def clear_flex_attention_cache(self):
        self._kv_cache = None

    # ------------------------------------------------------------------------------------------------------------
    # flat parameter / gradient storage
    # ------------------------------------------------------------------------------------------------------------
    def _param_order(self):
        """GEMM weights first (their grads may be overwritten instead of accumulated), then everything else."""
        named = dict(self.named_parameters())
        big = [n for n in named if n.endswith(("attn_qkv.weight", "attn_out.weight", "mlp.0.weight", "mlp.2.weight"))
               or n == "output_layer.linear.weight"]
        blocks_big = sorted(big, key=lambda n: (0, int(n.split(".")[1])) if n.startswith("blocks.") else (1, 0))
        rest = [n for n in named if n not in set(big)]
        return blocks_big, rest, named

    def _flatten(self):
        big, rest, named = self._param_order()
        dev = named[big[0]].device
        if dev.type != "cuda":
            raise L.UnidiscB200Error("unidisc_b200.DIT runs on CUDA only (no CPU fallback); move the module to a GPU")
        offs, off = {}, 0
        for n in big + rest:
            offs[n] = off
            off += (named[n].numel() + 63) // 64 * 64      # 256-byte aligned slots
        total = off
        self._big_end = offs[rest[0]]
        flat = torch.empty(total, device=dev, dtype=torch.float32)
        for n in big + rest:
            p = named[n]
            v = flat[offs[n]: offs[n] + p.numel()].view(p.shape)
            v.copy_(p.data.float())
            p.data = v
        self._flat_p = flat
        self._flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self._flat_bf16 = torch.empty(total, device=dev, dtype=bf16)
        self._offs = offs
        self._names = big + rest
        self._shadow_dirty = True
        self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        self._grads_attached = False
        self._views()

    def _views(self):
        named = dict(self.named_parameters())
        o = self._offs
        f32 = lambda n: named[n].data
        b16 = lambda n: self._flat_bf16[o[n]: o[n] + named[n].numel()].view(named[n].shape)
        gr = lambda n: self._flat_g[o[n]: o[n] + named[n].numel()].view(named[n].shape)
        self._blk = []
        for i in range(self.n_blocks):
            p = f"blocks.{i}."
            self._blk.append(dict(
                wqkv=b16(p + "attention.attn_qkv.weight"), wout=b16(p + "attention.attn_out.weight"),
                w1=b16(p + "mlp.0.weight"), b1=b16(p + "mlp.0.bias"), w2=b16(p + "mlp.2.weight"), b2=b16(p + "mlp.2.bias"),
                gq=f32(p + "attention.q_norm.weight"), bq=f32(p + "attention.q_norm.bias"),
                gk=f32(p + "attention.k_norm.weight"), bk=f32(p + "attention.k_norm.bias"),
                n1=f32(p + "norm1.weight"), n2=f32(p + "norm2.weight"), npre=f32(p + "pre_residual_norm.weight"),
                npost=f32(p + "post_ff_norm.weight"),
                d_wqkv=gr(p + "attention.attn_qkv.weight"), d_wout=gr(p + "attention.attn_out.weight"),
                d_w1=gr(p + "mlp.0.weight"), d_b1=gr(p + "mlp.0.bias"), d_w2=gr(p + "mlp.2.weight"), d_b2=gr(p + "mlp.2.bias"),
                d_gq=gr(p + "attention.q_norm.weight"), d_bq=gr(p + "attention.q_norm.bias"),
                d_gk=gr(p + "attention.k_norm.weight"), d_bk=gr(p + "attention.k_norm.bias"),
                d_n1=gr(p + "norm1.weight"), d_n2=gr(p + "norm2.weight"), d_npre=gr(p + "pre_residual_norm.weight"),
                d_npost=gr(p + "post_ff_norm.weight"),
            ))
            if self.time_conditioning:
                self._blk[-1].update(wada=b16(p + "adaLN_modulation.weight"), bada=b16(p + "adaLN_modulation.bias"),
                                     d_wada=gr(p + "adaLN_modulation.weight"), d_bada=gr(p + "adaLN_modulation.bias"))
        self._top = dict(
            E=f32("vocab_embed.embedding"), Emod=f32("modality_embed.embedding"), nf=f32("output_layer.norm_final.weight"),
            wh=b16("output_layer.linear.weight"), bh=b16("output_layer.linear.bias"),
            d_E=gr("vocab_embed.embedding"), d_Emod=gr("modality_embed.embedding"), d_nf=gr("output_layer.norm_final.weight"),
            d_wh=gr("output_layer.linear.weight"), d_bh=gr("output_layer.linear.bias"),
        )
        if self.require_sample_ids:
            self._top["Ecount"], self._top["d_Ecount"] = f32("img_count_embedding"), gr("img_count_embedding")
        if self.time_conditioning:
            for key, n in (("sm0", "sigma_map.mlp.0"), ("sm2", "sigma_map.mlp.2"), ("adaf", "output_layer.adaLN_modulation")):
                self._top["w_" + key], self._top["b_" + key] = b16(n + ".weight"), b16(n + ".bias")
                self._top["d_w_" + key], self._top["d_b_" + key] = gr(n + ".weight"), gr(n + ".bias")
        self._grad_views = {n: gr(n) for n in self._names}

    def _param_versions(self):
        """Sum of the autograd version counters of all parameters: changes whenever a parameter is written in place through
        the Parameter itself (`p.mul_()`, `p.copy_()`, torch.optim steps); writes through `p.data` are invisible to it."""
        return sum(p._version for p in self.parameters())

    def _ensure_ready(self):
        first = self.blocks[0].attention.attn_qkv.weight
        if self._flat_p is None or first.data_ptr() != self._flat_p.data_ptr() or first.device != self._flat_p.device:
            self._flatten()
        # The tensor cores read the bf16 SHADOW of the fp32 master weights.  It is refreshed when (a) a backward ran (an optimizer
        # step is expected), (b) load_state_dict / mark_weights_updated / train() / eval() was called — the reference swaps EMA
        # weights in with `ema.copy_to(params)` (writes through p.data) right before switching to eval and restores them before
        # train (model_eval.py:164-166, model_utils.py:338-345) — or (c) any parameter's version counter moved.
        v = self._param_versions()
        if self._shadow_dirty or v != self._shadow_versions:
            ops.cast_bf16(self._flat_p, self._flat_bf16)
            self._shadow_dirty = False
            self._shadow_versions = v

    def train(self, mode: bool = True):
        if self._flat_p is not None:
            self._shadow_dirty = True          # EMA swap / restore happens around mode switches (see _ensure_ready)
        return super().train(mode)

    def wait_param_events(self, which=None):
        """Order the current stream after the streamed optimizer update (FusedAdamW overlap mode): bucket `which`
        ("pre", block index, "head") or, with None, all of them (and forget them)."""
        ev = self._param_events
        if ev is None:
            return
        if which is None:
            for e in ev.values():
                torch.cuda.current_stream().wait_event(e)
            self._param_events = None
        else:
            torch.cuda.current_stream().wait_event(ev[which])

    def state_dict(self, *a, **k):
        self.wait_param_events()
        return super().state_dict(*a, **k)

    def mark_weights_updated(self, shadow_is_current: bool = False):
        """Call after the fp32 parameters changed behind the module's back — e.g. written through `p.data` outside a mode
        switch (EMA swaps, manual re-initialisation).  FusedAdamW passes shadow_is_current=True: it writes both copies."""
        self._shadow_dirty = not shadow_is_current
        if shadow_is_current:
            self._shadow_versions = self._param_versions()

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._shadow_dirty = True
        return r

    def attach_grad_stage(self, stage, alpha, sync_fn):
        """ThinDDP: `stage` = bf16 buffer covering the GEMM-weight part of the flat gradient layout ([0, _big_end)), `alpha` = device
        fp32 scalar 1 / world, `sync_fn()` -> True when the running backward is the one whose gradients get all-reduced."""
        self._ensure_ready()
        named = dict(self.named_parameters())
        o = self._offs
        sv = lambda n: stage[o[n]: o[n] + named[n].numel()].view(named[n].shape)
        self._stage_views = dict(head=sv("output_layer.linear.weight"), blocks=[
            dict(wqkv=sv(f"blocks.{i}.attention.attn_qkv.weight"), wout=sv(f"blocks.{i}.attention.attn_out.weight"),
                 w1=sv(f"blocks.{i}.mlp.0.weight"), w2=sv(f"blocks.{i}.mlp.2.weight")) for i in range(self.n_blocks)])
        self._ddp_stage, self._ddp_alpha, self._ddp_sync_fn = stage, alpha, sync_fn

    def _attach_grads(self):
        """Make p.grad a view of the flat gradient buffer.  Returns True if the gradients were (re)created — i.e. the
        caller zeroed them with set_to_none=True — in which case this backward overwrites instead of accumulating."""
        named = dict(self.named_parameters())
        fresh = getattr(self, "_force_fresh_grads", False) or any(
            named[n].grad is None or named[n].grad.data_ptr() != self._grad_views[n].data_ptr() for n in self._names[:2])
        self._force_fresh_grads = False
        if fresh:
            self._flat_g[self._big_end:].zero_()
            for n in self._names:
                named[n].grad = self._grad_views[n]
        return fresh

    @property
    def flat_params(self):
        self._ensure_ready()
        self.wait_param_events()
        return self._flat_p

    @property
    def flat_grads(self):
        self._ensure_ready()
        return self._flat_g

    @property
    def flat_params_bf16(self):
        self._ensure_ready()
        self.wait_param_events()
        return self._flat_bf16

    def block_grad_range(self, i):
        """[lo, hi) ranges of the flat gradient buffer owned by block i (GEMM weights, then small params)."""
        p = f"blocks.{i}."
        names = [n for n in self._names if n.startswith(p)]
        big = [n for n in names if self._offs[n] < self._big_end]
        small = [n for n in names if self._offs[n] >= self._big_end]
        rng = lambda ns: (min(self._offs[n] for n in ns), max(self._offs[n] + (dict(self.named_parameters())[n].numel() + 63) // 64 * 64 for n in ns))
        return rng(big), rng(small)

    # ------------------------------------------------------------------------------------------------------------
    # forward / backward
    # ------------------------------------------------------------------------------------------------------------
    @torch.compiler.disable
    def forward(self, indices, sigma=None, label=None, x_cond=None, attention_mask=None, continuous_mode=False,
                x_img_emb=None, modality=None, start_pos=None, block_mask=None, update_cache_slice=None, sample_ids=None,
                head_rows=None, head_split=None):
        """Returns logits [B,N,V] in bf16 (what the reference returns under its outer bf16 autocast, model.py:693-729).
        `sigma` is used only with `time_conditioning` (off in every shipped training config).
        `head_rows` (additive, default None = reference behaviour): int64 indices into the B*N token rows; the output projection
        is then evaluated for those rows only and the result is [1, len(head_rows), V].  The SUBS loss needs logits of MASKED
        positions only (model.py:621-658: an unmasked token's log-probability is exactly 0), so `Diffusion.compute_loss` passes
        the masked rows: same loss, same gradients, about half the head GEMM work.  A zero-argument callable returning
        `(rows or None, split or None)` is resolved right before the output projection.
        `head_split` (with head_rows): the first `head_split` head rows are TEXT tokens, the rest IMAGE tokens; text rows are then
        projected onto the text vocabulary only and image rows onto the image vocabulary only (the other columns of the returned
        rows are NOT written).  Exact whenever the loss restricts each token to its modality's vocabulary
        (model.force_argmax_valid_indices: the reference adds -1e6 to the other logits, model.py:627-640)."""
        if label is not None or x_cond is not None or continuous_mode or x_img_emb is not None or start_pos is not None:
            raise NotImplementedError("unidisc_b200.DIT.forward: label/x_cond/continuous/start_pos arguments are not supported")
        if attention_mask is not None:
            raise NotImplementedError("unidisc_b200.DIT.forward: dense attention_mask is not supported (model.use_attention_mask)")
        cache_op = None
        if self._kv_cache is not None and not self.training:
            # the three steps of the reference's caching cycle, told apart exactly like dit.py:798-810
            if indices.shape[1] != self._kv_cache["N"]:
                if update_cache_slice is None:
                    raise ValueError("attention caching: a shortened sequence needs update_cache_slice")
                cache_op = ("update", update_cache_slice)
            elif isinstance(block_mask, TextFullImageSelfMask):
                if update_cache_slice is None:
                    raise ValueError("attention caching: the cache-filling step needs update_cache_slice")
                cache_op = ("store", block_mask.txt_length)
            if cache_op is not None and torch.is_grad_enabled():
                raise RuntimeError("unidisc_b200.DIT: attention caching is an inference path; call under torch.no_grad()")
            block_mask = None if sample_ids is None else block_mask
        elif update_cache_slice is not None or isinstance(block_mask, TextFullImageSelfMask):
            raise ValueError("unidisc_b200.DIT.forward: update_cache_slice / caching masks need set_flex_attention_cache() and eval mode")
        if block_mask is True and sample_ids is None:
            block_mask = None                          # the reference's "full attention" marker (dit.py:808-811)
        if block_mask is not None and sample_ids is None:
            raise NotImplementedError("unidisc_b200.DIT.forward: FlexAttention block_mask objects are not supported; pass sample_ids")
        if self.require_sample_ids and sample_ids is None:
            raise ValueError("data.require_sample_ids: sample_ids is required")
        if modality is None:
            raise ValueError("modality is required (trainer.multimodal_batches)")
        if self.time_conditioning and sigma is None:
            raise ValueError("time_conditioning: sigma is required")
        if not indices.is_cuda:
            raise L.UnidiscB200Error("unidisc_b200.DIT.forward needs CUDA tensors (no CPU fallback)")
        self._ensure_ready()
        save = torch.is_grad_enabled() and self.training_graph_enabled
        # The reference hands FlexAttention a BlockMask built from sample_ids (model.py:876-878, model_utils.py:740-771);
        # here a non-None `block_mask` switches the attention kernels' document mask on and the mask itself is derived
        # from `sample_ids` on the fly.  Without require_sample_ids, passing sample_ids alone also enables it.
        doc_mask = sample_ids is not None and (block_mask is not None or not self.require_sample_ids)
        if cache_op is not None:
            if head_rows is not None:
                raise ValueError("head_rows is a training-path argument (no attention caching)")
            logits, _ = self._forward_impl(indices, modality, sample_ids, save=False, doc_mask=doc_mask,
                                           sigma=sigma if self.time_conditioning else None, cache_op=cache_op)
            return logits
        if head_rows is not None and not callable(head_rows) \
                and (head_rows.dtype != torch.int64 or head_rows.dim() != 1 or head_rows.numel() == 0):
            raise ValueError("head_rows: a non-empty 1-D int64 tensor of token-row indices (or a callable returning "
                             "(rows, split), resolved right before the output projection) is required")
        if head_split is not None and (head_rows is None or callable(head_rows) or not 0 < int(head_split) < head_rows.numel()):
            raise ValueError("head_split needs head_rows and 0 < head_split < len(head_rows)")
        return _DiTFunction.apply(self._anchor, self, indices, modality, sample_ids, save, doc_mask,
                                  sigma if self.time_conditioning else None, head_rows, head_split)

    # ------------------------------------------------------------------------------------------------------------
    # time conditioning (reference dit.py:415-449, 1378-1379, 966-967, 1083-1087).  The conditioning network acts on one
    # row per SAMPLE ([B, cond_dim], B <= a few dozen): it is evaluated with torch bf16 linears exactly as the reference
    # does under autocast, and its outputs (shift / scale / gate, bf16 [B, L*6D + 2D]) are consumed per TOKEN inside the
    # fused norm kernels (csrc/elementwise.cu *_tc_kernel), which also reduce their gradients per sample.
    # ------------------------------------------------------------------------------------------------------------
    def _cond_forward(self, sigma, mod_flat, B, N):
        import torch.nn.functional as F
        T, D = self._top, self.hidden_size
        half = self.sigma_map.frequency_embedding_size // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=sigma.device) / half)
        args = sigma.reshape(-1)[:, None].float() * freqs[None]
        e = torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(bf16)
        h1 = F.linear(e, T["w_sm0"], T["b_sm0"])
        s1 = F.silu(h1)
        h2 = F.linear(s1, T["w_sm2"], T["b_sm2"])
        c = F.silu(h2)                                                        # bf16 [B, cond_dim]
        cond = torch.cat([F.linear(c, W["wada"], W["bada"]) for W in self._blk] + [F.linear(c, T["w_adaf"], T["b_adaf"])],
                         dim=1).contiguous()                                  # bf16 [B, L*6D + 2D]
        is_img = mod_flat == 1
        # modulate_fused quirk (dit.py:301-304): a batch without any image token modulates EVERY token
        sel = torch.where(mod_flat.any(), is_img, torch.ones_like(is_img)).to(torch.uint8)
        return dict(e=e, h1=h1, s1=s1, h2=h2, c=c, cond=cond, sel=sel, img=is_img.to(torch.uint8), N=N,
                    d_cond=None)

    def _tc(self, C, norm_block=None, gate_block=None, mlp=False, bwd=False):
        """`ud_adaln` for one fused norm kernel: shift/scale of `norm_block`'s norm1 (mlp=False) / norm2 (mlp=True) — block
        index L = the final layer — and the MLP gate of `gate_block` applied to the branch the kernel adds."""
        if C is None:
            return None
        D, cond, ld = self.hidden_size, C["cond"], C["cond"].shape[1]
        kw = dict(ld=ld, tokens_per_sample=C["N"])
        if bwd:
            kw["ld_d"] = ld
        if norm_block is not None:
            base = norm_block * 6 * D + (3 * D if mlp else 0)
            kw.update(shift=cond[:, base:], scale=cond[:, base + D:])
            if bwd:
                kw.update(d_shift=C["d_cond"][:, base:], d_scale=C["d_cond"][:, base + D:])
        if gate_block is not None:
            kw["gate"] = cond[:, gate_block * 6 * D + 5 * D:]
            if bwd:
                kw["d_gate"] = C["d_cond"][:, gate_block * 6 * D + 5 * D:]
        return L.adaln(C["sel"], C["img"], **kw)

    def _adaln_param_grads(self, C, i):
        """adaLN Linear backward for block i (i = L: final layer) once its slice of d_cond is complete."""
        D, T = self.hidden_size, self._top
        lo = i * 6 * D
        hi = lo + (6 * D if i < self.n_blocks else 2 * D)
        dc = C["d_cond"][:, lo:hi]
        if i < self.n_blocks:
            w, dw, db = self._blk[i]["wada"], self._blk[i]["d_wada"], self._blk[i]["d_bada"]
        else:
            w, dw, db = T["w_adaf"], T["d_w_adaf"], T["d_b_adaf"]
        dw.addmm_(dc.t(), C["c"].float())
        db.add_(dc.sum(0))
        C["d_c"].addmm_(dc, w.float())

    def _cond_backward(self, C):
        import torch.nn.functional as F
        T = self._top
        dsilu = lambda pre: (lambda x: torch.sigmoid(x) * (1 + x * (1 - torch.sigmoid(x))))(pre.float())
        dh2 = C["d_c"] * dsilu(C["h2"])
        T["d_w_sm2"].addmm_(dh2.t(), C["s1"].float())
        T["d_b_sm2"].add_(dh2.sum(0))
        dh1 = (dh2 @ T["w_sm2"].float()) * dsilu(C["h1"])
        T["d_w_sm0"].addmm_(dh1.t(), C["e"].float())
        T["d_b_sm0"].add_(dh1.sum(0))

    def _cached_attention(self, i, qk, qkv, B, N, scale, cache_op):
        """Attention of block i inside an inference caching cycle (reference dit.py:793-812)."""
        D, H, hd = self.hidden_size, self.n_heads, self.head_dim
        C = self._kv_cache
        q, k, v = qk[:, :D], qk[:, D:], qkv[:, 2 * D:]
        if cache_op[0] == "store":
            # step 1: text queries see every key, image queries see image keys only; this step's K / V become the cache
            T = cache_op[1]
            o = torch.empty((B * N, D), device=qk.device, dtype=bf16)
            if T > 0:
                ops.attn_fwd_kv(q, k, v, B, T, N, H, hd, scale, o=o, q_bs=N * 2 * D, k_bs=N * 2 * D, v_bs=N * 3 * D, o_bs=N * D)
            if N - T > 0:
                ops.attn_fwd_kv(q[T:], k[T:], v[T:], B, N - T, N - T, H, hd, scale, o=o[T:], q_bs=N * 2 * D, k_bs=N * 2 * D,
                                v_bs=N * 3 * D, o_bs=N * D)
            C["k"][i].copy_(k)
            C["v"][i].copy_(v)
            return o
        # steps 2..: the shortened (text-only) sequence writes its K / V into the cache's slice ...
        sl = cache_op[1]
        Nc = C["N"]
        C["k"][i].view(C["B"], Nc, D)[:B, sl] = k.reshape(B, N, D)
        C["v"][i].view(C["B"], Nc, D)[:B, sl] = v.reshape(B, N, D)
        if self.cache_attend_cached:            # ... and attends to the whole cache (fresh text + cached image K / V)
            return ops.attn_fwd_kv(q, C["k"][i], C["v"][i], B, N, Nc, H, hd, scale)[0]
        return ops.attn_fwd(q, k, v, B, N, H, hd, scale)[0]       # reference dataflow: local K / V (dit.py:812)

    def _block_forward(self, i, x, h, cos, sin, sid, B, N, C, p_drop, drop_base, cache_op=None):
        """DDiTBlock.forward (reference dit.py:948-1033) as 8 launches; returns (x_out, h_next, activations for backward)."""
        D, H, hd = self.hidden_size, self.n_heads, self.head_dim
        W, T = self._blk[i], self._top
        scale = 1.0 / math.sqrt(hd)
        w_next = self._blk[i + 1]["n1"] if i + 1 < self.n_blocks else T["nf"]
        qkv = ops.gemm(h, W["wqkv"])
        qk, stats = ops.qk_ln_rope_fwd(qkv, W["gq"], W["bq"], W["gk"], W["bk"], cos, sin, hd)
        if cache_op is not None:
            o, lse = self._cached_attention(i, qk, qkv, B, N, scale, cache_op), None
        else:
            o, lse = ops.attn_fwd(qk[:, :D], qk[:, D:], qkv[:, 2 * D:], B, N, H, hd, scale, sample_ids=sid)
        a = ops.gemm(o, W["wout"])
        x1, h2, ra, rx1 = ops.norm_residual_fwd(a, x, W["npre"], W["n2"], tc=self._tc(C, norm_block=i, mlp=True))
        u, gl = ops.gemm(h2, W["w1"], epi=L.EPI_BF16_GELU, bias=W["b1"])
        d = ops.gemm(gl, W["w2"], bias=W["b2"])
        x2, h_next, rd, rx2 = ops.norm_residual_fwd(d, x1, W["npost"], w_next, p_drop=p_drop, seed=self.dropout_seed,
                                                    offset=drop_base + i, tc=self._tc(C, norm_block=i + 1, gate_block=i))
        rec = dict(h=h, qkv=qkv, qk=qk, stats=stats, o=o, lse=lse, a=a, ra=ra, x1=x1, rx1=rx1, h2=h2, u=u, g=gl, d=d, rd=rd, x2=x2,
                   rx2=rx2)
        return x2, h_next, rec

    def _forward_impl(self, indices, modality, sample_ids, save, doc_mask=True, sigma=None, cache_op=None, head_rows=None, head_split=None):
        B, N = indices.shape
        M, D, H, hd, V = B * N, self.hidden_size, self.n_heads, self.head_dim, self.vocab_size
        T = self._top
        ids = indices.reshape(-1).contiguous()
        mod = modality.reshape(-1).contiguous()
        sid = sample_ids.contiguous() if (sample_ids is not None and doc_mask) else None
        ordinal = None
        if self.require_sample_ids:
            cos, sin, ordinal = ops.interleaved_prep(modality, sample_ids, self._rope_cat_cos, self._rope_cat_sin, self._rope_offsets)
            if sid is not None:
                # model_utils.py:764-767: a row that is all padding gets one token re-labelled so no query row is empty
                allpad = (sid == -1).all(dim=-1)
                sid = torch.where(allpad[:, None] & (torch.arange(N, device=sid.device) == 0)[None], torch.zeros_like(sid), sid)
        else:
            cos, sin = rope.token_tables(modality, self.rotary_cos_emb_txt, self.rotary_sin_emb_txt, self.rotary_cos_emb_img,
                                         self.rotary_sin_emb_img, self.img_length)
        scale = 1.0 / math.sqrt(hd)
        self.wait_param_events("pre")
        self.wait_param_events(0)              # block 0's norm1 weight is applied by the embedding kernel
        if self.time_conditioning:
            self.wait_param_events()           # the conditioning network reads every block's adaLN weights up front
        C = self._cond_forward(sigma, mod, B, N) if self.time_conditioning else None
        x, h, rstd0 = ops.embed_rmsnorm_fwd(ids, mod, T["E"], T["Emod"], self._blk[0]["n1"], ordinal=ordinal,
                                            Ecount=T.get("Ecount"), tc=self._tc(C, norm_block=0))
        # training-mode dropout of the MLP branch (dit.py:1024-1031): Philox mask keyed by (seed, call counter * L + block)
        p_drop = self.dropout if self.training else 0.0
        if self._dropout_rank is None:
            # the reference seeds every rank with seed + rank (main.py:1062): data-parallel replicas draw independent masks
            import torch.distributed as dist
            self._dropout_rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
            self.dropout_seed += self._dropout_rank
        self._dropout_calls += 1
        drop_base = self._dropout_calls * self.n_blocks
        saved = dict(ids=ids, mod=mod, sid=sid, cos=cos, sin=sin, B=B, N=N, x0=x, rstd0=rstd0, blocks=[], p_drop=p_drop,
                     drop_base=drop_base, ordinal=ordinal, C=C) if save else None
        ckpt = save and self.use_gradient_checkpointing and self.training
        for i in range(self.n_blocks):
            if i + 1 < self.n_blocks:
                self.wait_param_events(i + 1)  # this block's last kernel applies the next block's norm1 weight
            if self._dbg_fwd_events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                self._dbg_fwd_events.append(e)
            x2, h_next, rec = self._block_forward(i, x, h, cos, sin, sid, B, N, C, p_drop, drop_base, cache_op)
            if save:
                # trainer.use_gradient_checkpointing (reference dit.py:1485-1490 wraps every block in torch.utils.checkpoint):
                # keep only the block's inputs; the backward re-runs the block's forward (same Philox dropout mask)
                saved["blocks"].append(dict(ckpt=True, x=x, h=h) if ckpt else rec)
            x, h = x2, h_next
        self.wait_param_events()               # head (last bucket) => everything; the events are dropped
        if callable(head_rows):
            # lazily resolved row selection: the caller's device -> host read of the row counts was enqueued before this forward,
            # so by now it has long completed and the host never waits for the GPU (which is still working through the blocks)
            head_rows, head_split = head_rows()
        if head_rows is not None:
            # output projection of the requested token rows only (the masked positions: the only logits the SUBS loss reads)
            # buffers are allocated at their all-rows size and sliced: the number of head rows changes every step, and a new
            # maximum would send the caching allocator to cudaMalloc in the middle of the step (seen as a ~3 ms slower
            # device-resident loop on some runs)
            Mh = head_rows.numel()
            hq = torch.empty((M, D), device=x.device, dtype=bf16)[:Mh]
            torch.index_select(h, 0, head_rows, out=hq)
            buf = torch.empty((M, self.Vp), device=x.device, dtype=bf16)[:Mh]
            if head_split is None:
                ops.gemm(hq, T["wh"], N=V, out=buf[:, :V], bias=T["bh"])
            else:
                # text rows x text vocabulary, image rows x image vocabulary.  The image block starts at the text vocabulary size
                # rounded DOWN to 8 columns (16-byte aligned output / bias / weight-slice pointers); the <= 7 extra columns are
                # outside the image rows' valid range and ignored by the loss
                Mt, tv, tv8 = int(head_split), self.text_vocab_size, self.text_vocab_size // 8 * 8
                ops.gemm(hq[:Mt], T["wh"][:tv], N=tv, out=buf[:Mt, :tv], bias=T["bh"][:tv])
                ops.gemm(hq[Mt:], T["wh"][tv8:], N=V - tv8, out=buf[Mt:, tv8:V], bias=T["bh"][tv8:])
            if save:
                saved["hf"] = hq
                saved["head_rows"] = head_rows
                saved["head_split"] = head_split
                saved["logits_buf"] = buf
            return buf.view(1, Mh, self.Vp)[:, :, :V], saved
        buf = torch.empty((M, self.Vp), device=x.device, dtype=bf16)
        ops.gemm(h, T["wh"], N=V, out=buf[:, :V], bias=T["bh"])
        if save:
            saved["hf"] = h
            saved["head_rows"] = None
            saved["logits_buf"] = buf
        return buf.view(B, N, self.Vp)[:, :, :V], saved

    def _backward_impl(self, S, dlogits):
        B, N = S["B"], S["N"]
        M, D, H, hd, V = B * N, self.hidden_size, self.n_heads, self.head_dim, self.vocab_size
        T = self._top
        fresh = self._attach_grads()
        wacc = L.EPI_F32 if fresh else L.EPI_F32_ACC
        # gradient-norm fusion: when the weight gradients are written (not accumulated) their squares are summed by the GEMM
        # epilogue that stores them, so the optimizer never re-reads the 5.6 GB of GEMM-weight gradients for clip_grad_norm_
        gacc = self.grad_sumsq_acc if fresh else None
        self._last_bwd_fused_sumsq = gacc is not None
        # data parallel: the synchronising backward writes bf16(bf16(dW) / world) into the all-reduce staging buffer (no fp32
        # store, no compression pass); accumulation micro-steps keep the fp32 path
        staged = fresh and self._ddp_stage is not None and self._ddp_sync_fn()
        self._last_bwd_staged = staged

        def wgrad(dy, x, dst, stage_dst, **shape):
            if staged:
                ops.gemm(dy, x, ta=True, tb=True, epi=L.EPI_BF16_SCALED, out=stage_dst, aux=self._ddp_alpha, **shape)
            else:
                ops.gemm(dy, x, ta=True, tb=True, epi=wacc, out=dst, aux=gacc, **shape)
        SV = self._stage_views
        # logits gradient in the padded [Mh, Vp] layout (produced in place by the fused SUBS-NLL backward when possible);
        # Mh = M, or the number of head rows when the forward projected a subset of the token rows
        rows = S.get("head_rows")
        Mh = M if rows is None else S["hf"].shape[0]
        if dlogits.dtype == bf16 and dlogits.dim() == 3 and dlogits.stride() == (dlogits.shape[1] * self.Vp, self.Vp, 1) \
                and dlogits.shape[0] * dlogits.shape[1] == Mh:
            dl = dlogits.as_strided((Mh, self.Vp), (self.Vp, 1))[:, :V]
        else:
            buf = torch.zeros((Mh, self.Vp), device=dlogits.device, dtype=bf16)
            buf[:, :V].copy_(dlogits.reshape(Mh, V))
            dl = buf[:, :V]
        # head
        split = S.get("head_split") if rows is not None else None
        if split is None:
            wgrad(dl, S["hf"], T["d_wh"], SV["head"] if staged else None, M=V, N=D, K=Mh)
            ops.colsum(dl, T["d_bh"], Mh, V)
            dh = ops.gemm(dl, T["wh"], tb=True, M=Mh, N=D, K=V, out=torch.empty((M, D), device=dl.device, dtype=bf16)[:Mh])
        else:
            # the SUBS-NLL backward wrote every column of every head row (zeros outside a row's vocabulary range), so the two
            # blocks can be multiplied independently.  Weight rows [tv8, tv) are covered by both weight-gradient GEMMs: the image
            # one (whose rows have zero gradient there) runs first and the text one then stores / accumulates the real values.
            Mt, tv, tv8 = int(split), self.text_vocab_size, self.text_vocab_size // 8 * 8
            hq, dwh, st = S["hf"], T["d_wh"], (SV["head"] if staged else None)
            wgrad(dl[Mt:, tv8:], hq[Mt:], dwh[tv8:], st[tv8:] if st is not None else None, M=V - tv8, N=D, K=Mh - Mt)
            wgrad(dl[:Mt, :tv], hq[:Mt], dwh[:tv], st[:tv] if st is not None else None, M=tv, N=D, K=Mt)
            ops.colsum(dl[Mt:, tv8:], T["d_bh"][tv8:], Mh - Mt, V - tv8)
            ops.colsum(dl[:Mt, :tv], T["d_bh"][:tv], Mt, tv)
            dh = torch.empty((M, D), device=dl.device, dtype=bf16)[:Mh]
            ops.gemm(dl[:Mt, :tv], T["wh"][:tv], tb=True, M=Mt, N=D, K=tv, out=dh[:Mt])
            ops.gemm(dl[Mt:, tv8:], T["wh"][tv8:], tb=True, M=Mh - Mt, N=D, K=V - tv8, out=dh[Mt:])
        if rows is not None:                         # rows without logits received no gradient from the head
            dh = torch.zeros((M, D), device=dh.device, dtype=dh.dtype).index_copy_(0, rows, dh)
        S["logits_buf"] = None
        # Gradient buckets that are final are handed to the DDP hook right BEFORE the next attention backward: the
        # all-reduce chain (pack -> NCCL -> unpack, ~0.4 ms) then runs next to the attention / row kernels, whose many
        # small CTAs are load-balanced by the hardware scheduler, instead of next to a persistent GEMM whose statically
        # assigned tiles wait for the SMs the NCCL CTAs slow down (measured: 7.7 ms -> see DESIGN.md §5).
        pending = [self.n_blocks]                    # head weight gradient is final

        def flush_pending():
            if self.grad_ready_hook is not None:
                for b in pending:
                    self.grad_ready_hook(b)
            pending.clear()
        g_res = None       # fp32 gradient flowing down the residual stream
        scale = 1.0 / math.sqrt(hd)
        C = S.get("C")
        if C is not None:
            C["d_cond"] = torch.zeros(C["cond"].shape, device=C["cond"].device, dtype=torch.float32)
            C["d_c"] = torch.zeros(C["c"].shape, device=C["c"].device, dtype=torch.float32)
        for i in range(self.n_blocks - 1, -1, -1):
            W, A = self._blk[i], S["blocks"][i]
            if A.get("ckpt"):      # gradient checkpointing: rebuild this block's activations from its saved inputs
                _, _, A = self._block_forward(i, A["x"], A["h"], S["cos"], S["sin"], S["sid"], B, N, C, S["p_drop"], S["drop_base"])
            w_next = self._blk[i + 1]["n1"] if i + 1 < self.n_blocks else T["nf"]
            d_wnext = self._blk[i + 1]["d_n1"] if i + 1 < self.n_blocks else T["d_nf"]
            # x2 = x1 + rms(d)*w_post ; h_next = rms(x2)*w_next
            g_res, dd = ops.norm_residual_bwd(g_res, dh, A["x2"], A["rx2"], w_next, A["d"], A["rd"], W["npost"], d_wnext, W["d_npost"],
                                              db_a=W["d_b2"],           # also accumulates mlp.2.bias.grad = colsum(dd)
                                              p_drop=S["p_drop"], seed=self.dropout_seed, offset=S["drop_base"] + i,
                                              tc=self._tc(C, norm_block=i + 1, gate_block=i, bwd=True))
            if C is not None and i == self.n_blocks - 1:
                self._adaln_param_grads(C, self.n_blocks)                 # final layer's shift / scale are complete
            # MLP
            SB = SV["blocks"][i] if staged else None
            wgrad(dd, A["g"], W["d_w2"], SB and SB["w2"])
            du = ops.gemm(dd, W["w2"], tb=True, epi=L.EPI_BF16_DGELU, aux=A["u"])
            wgrad(du, A["h2"], W["d_w1"], SB and SB["w1"])
            ops.colsum(du, W["d_b1"])
            dh2 = ops.gemm(du, W["w1"], tb=True)
            # x1 = x + rms(a)*w_pre ; h2 = rms(x1)*w_n2
            g_res, da = ops.norm_residual_bwd(g_res, dh2, A["x1"], A["rx1"], W["n2"], A["a"], A["ra"], W["npre"], W["d_n2"], W["d_npre"],
                                              tc=self._tc(C, norm_block=i, mlp=True, bwd=True))
            if C is not None and i + 1 < self.n_blocks:
                # block i+1's six chunks are final (its norm1 modulation was differentiated by this iteration's first kernel)
                self._adaln_param_grads(C, i + 1)
            # attention
            wgrad(da, A["o"], W["d_wout"], SB and SB["wout"])
            do = ops.gemm(da, W["wout"], tb=True)
            dqk = torch.empty((M, 2 * D), device=do.device, dtype=bf16)
            dqkv = torch.empty((M, 3 * D), device=do.device, dtype=bf16)
            qk, qkv = A["qk"], A["qkv"]
            flush_pending()
            ops.attn_bwd(qk[:, :D], qk[:, D:], qkv[:, 2 * D:], A["o"], do, A["lse"], dqk[:, :D], dqk[:, D:], dqkv[:, 2 * D:],
                         B, N, H, hd, scale, sample_ids=S["sid"])
            ops.qk_ln_rope_bwd(dqk, qkv, A["stats"], W["gq"], W["gk"], S["cos"], S["sin"], dqkv, W["d_gq"], W["d_bq"], W["d_gk"],
                               W["d_bk"], hd)
            wgrad(dqkv, A["h"], W["d_wqkv"], SB and SB["wqkv"])
            dh = ops.gemm(dqkv, W["wqkv"], tb=True)
            S["blocks"][i] = None      # release this block's activations
            pending.append(i)
            del A
        # first norm + embedding
        g0 = ops.rmsnorm_bwd(g_res, dh, S["x0"], S["rstd0"], self._blk[0]["n1"], self._blk[0]["d_n1"],
                             tc=self._tc(C, norm_block=0, bwd=True))
        if C is not None:
            self._adaln_param_grads(C, 0)
            self._cond_backward(C)
        ops.embed_bwd(S["ids"], S["mod"], g0, T["d_E"], T["d_Emod"], hot_id=self.mask_index, ordinal=S["ordinal"],
                      dEcount=T.get("d_Ecount"))
        self._shadow_dirty = True      # an optimizer step is expected to follow
        pending.append(-1)
        flush_pending()
