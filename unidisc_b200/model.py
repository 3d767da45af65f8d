"""B200-native mirror of the hot-path methods of the reference trainer class `Diffusion`
(reference model.py:51 + model_setup.py / model_utils.py / model_eval.py, bound at model.py:54-99).

Method names, argument meaning and return conventions follow the reference so that the parity tests read like the
reference's own call sites:

  _sample_t (model.py:589)  q_xt (model.py:424)  forward (model.py:674)  _subs_parameterization (model.py:621)
  compute_loss (model.py:797)  training_step (model.py:420)
  _sample_prior (model_eval.py:1734)  get_cfg_weight (:1737)  _ddpm_forward (:1761)  _ddpm_update (:2042)
  _ddpm_caching_update (:2072)  _sample (:2107)  adap_sche (:2964)  _maskgit_update (:3045)

All vocabulary-sized maths (SUBS log-softmax, NLL gather, Gumbel arg-max, absorbing update) runs in the CUDA
library; only [B,N]-sized bookkeeping stays in torch.  Out of scope: data loading, tokenizers/VAEs,
evaluation metrics, checkpoint/launcher glue, AR / SEDD / D3PM parameterisations.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import noise_schedule, ops
from .dit import DIT, TextFullImageSelfMask, _wrap_cfg

bf16 = torch.bfloat16


@dataclass
class Loss:                                                           # reference model_utils.py:110-120
    loss: torch.Tensor
    img_loss: torch.Tensor = None
    txt_loss: torch.Tensor = None
    nlls: torch.Tensor = None
    token_mask: torch.Tensor = None
    txt_nlls: torch.Tensor = None
    img_nlls: torch.Tensor = None
    extra_losses: dict = None
    modality_mask: torch.Tensor = None


def _g(o, k, d=None):
    if o is None:
        return d
    if isinstance(o, dict):
        return o.get(k, d)
    return getattr(o, k, d)


def interleaved_block_masking(move_indices, modality, sample_ids, mask_prob):
    """reference model.py:483-522 (trainer.interleaved + mask_entire_modality): every (modality, sample) block longer than 4
    tokens is force-masked with probability mask_prob * 2 * (k+1)/K, k = its index among the K blocks of its packed sample.
    Draws `torch.rand(M, 1)` (M = number of blocks) exactly like the reference.  Returns (move_indices, ignore [B])."""
    B, N = modality.shape
    dev = modality.device
    diff = (modality[:, 1:] != modality[:, :-1]) | (sample_ids[:, 1:] != sample_ids[:, :-1])    # tensor_utils.py:46-68
    diff = torch.nn.functional.pad(diff, (1, 0), mode="constant", value=True)
    starts = diff.nonzero(as_tuple=False)
    ends = torch.nn.functional.pad(diff[:, 1:], (0, 1), mode="constant", value=True).nonzero(as_tuple=False)
    bi, sp, ep = starts[:, 0], starts[:, 1], ends[:, 1] + 1
    keep = (sample_ids[bi, sp] >= 0) & ((ep - sp) > 4)                                            # model.py:486-488
    bi, sp, ep = bi[keep], sp[keep], ep[keep]
    M = bi.shape[0]
    sids = sample_ids[bi, sp]
    same = (bi[:, None] == bi[None, :]) & (sids[:, None] == sids[None, :])
    order = torch.arange(M, device=dev)
    block_counts = (same & (order[None, :] < order[:, None])).sum(dim=1)                          # model.py:494-505
    max_num = same.sum(dim=1)
    block_prob = (block_counts + 1) / max_num
    positions = torch.arange(N, device=dev).unsqueeze(0)
    mask = (positions >= sp.unsqueeze(1)) & (positions < ep.unsqueeze(1))
    mask = mask & (torch.rand(M, 1, device=dev) < (mask_prob * block_prob * 2)[..., None])
    accum = torch.zeros((B, N), dtype=torch.int32, device=dev)
    accum.scatter_add_(0, bi.unsqueeze(1).expand(-1, N), mask.int())
    move_indices = move_indices | accum.to(torch.bool)
    ignore = torch.zeros((B,), dtype=torch.int32, device=dev)
    ignore.scatter_add_(0, bi, mask.any(dim=-1).int())
    return move_indices, ignore.to(torch.bool)


def q_xt_general(x, move_chance, mask_index, trainer_cfg, *, backbone_training, training, batch=None, allow_move_mask=None):
    """reference model.py:439-579 for absorbing diffusion on multimodal batches, in torch on [B,N] / [B,1] booleans, with
    the reference's order of torch.rand draws.  Returns (xt, ignore_batch_mask, should_mask_txt, should_mask_img, move)."""
    move_indices = torch.rand(*x.shape, device=x.device) < move_chance
    ignore = None
    should_mask_txt = should_mask_img = None
    mask_prob = _g(trainer_cfg, "mask_entire_modality", None)
    if mask_prob is not None and backbone_training:
        assert batch is not None
        bsz = x.shape[0]
        if _g(trainer_cfg, "mask_txt_only", False):
            should_mask_txt = torch.rand(bsz, 1, device=x.device) < mask_prob
            should_mask_img = torch.zeros_like(should_mask_txt)
        else:
            should_mask_txt = torch.rand(bsz, 1, device=x.device) < mask_prob / 2
            should_mask_img = torch.rand(bsz, 1, device=x.device) < mask_prob / 2
        if _g(trainer_cfg, "interleaved", False):
            move_indices, ignore = interleaved_block_masking(move_indices, batch["modality"], batch["sample_ids"], mask_prob)
        else:
            modality_mask = batch.get("modality_mask", None)
            if modality_mask is None:
                modality_mask = torch.stack([batch["modality"] == 0, batch["modality"] == 1], dim=-1)
            both = should_mask_txt & should_mask_img
            should_mask_txt = torch.where(both, False, should_mask_txt)
            should_mask_img = torch.where(both, False, should_mask_img)
            move_indices = torch.where(should_mask_txt, modality_mask[..., 0], move_indices)
            move_indices = torch.where(should_mask_img, modality_mask[..., 1], move_indices)
            ignore = should_mask_img | should_mask_txt
    if _g(trainer_cfg, "add_label", False):
        move_indices[:, 0] = False
    ftd = _g(trainer_cfg, "first_token_dropout", None)
    if ftd is not None and training:
        init = torch.rand(x.shape[0], device=x.device) < ftd
        move_indices[:, 0] = torch.where(init, True, move_indices[:, 0])
        ignore = init if ignore is None else (ignore | init)
    if allow_move_mask is not None:
        move_indices = move_indices & allow_move_mask
    xt = torch.where(move_indices, mask_index, x)
    return xt, ignore, should_mask_txt, should_mask_img, move_indices


def select_head_rows(sel_t: torch.Tensor, sel_i: torch.Tensor, n_t: int, n_i: int, split_ok: bool):
    """Token rows the output projection has to be evaluated for, from the flat boolean selections of masked & attended TEXT rows
    (`sel_t`) and IMAGE rows (`sel_i`) and their (host) counts.  Returns `(rows, split)`: `rows` int64, text rows first, or None when
    nothing or everything is selected (the plain all-rows head is used); `split` = number of leading text rows when both
    modalities are present and the loss restricts each row to its own vocabulary block (`split_ok`), else None."""
    total = sel_t.numel()
    if not 0 < n_t + n_i < total:
        return None, None
    if split_ok and n_t > 0 and n_i > 0:
        rows = torch.cat([torch.nonzero_static(sel_t, size=n_t).squeeze(1), torch.nonzero_static(sel_i, size=n_i).squeeze(1)])
        return rows, n_t
    return torch.nonzero_static(sel_t | sel_i, size=n_t + n_i).squeeze(1), None


class _SubsNLL(torch.autograd.Function):
    """log p_theta(x0 | xt) under the SUBS parameterisation, fused: logits are read once, the [B,N,V] log-prob tensor of
    the reference (model.py:621-658 + gather at :967) is never materialised.  Backward writes dlogits IN PLACE over the
    logits buffer (it has no other consumer in the training step)."""

    @staticmethod
    def forward(ctx, logits, xt, x0, modality, V, text_vocab, mask_index):
        B, N = xt.shape
        if logits.dtype != bf16 or logits.stride(-1) != 1 or logits.stride(0) != N * logits.stride(1):
            raise L.UnidiscB200Error("_SubsNLL expects the bf16 logits returned by unidisc_b200.DIT.forward")
        ldv = logits.stride(1)
        l2 = logits.as_strided((B * N, ldv), (ldv, 1))
        xt_f, x0_f, md_f = xt.reshape(-1).contiguous(), x0.reshape(-1).contiguous(), modality.reshape(-1).contiguous()
        logp, lse = ops.subs_nll_fwd(l2, xt_f, x0_f, md_f, V, text_vocab, mask_index)
        ctx.save_for_backward(logits, xt_f, x0_f, md_f, lse)
        ctx.meta = (B, N, V, text_vocab, mask_index, ldv)
        return logp.view(B, N)

    @staticmethod
    def backward(ctx, dlogp):
        logits, xt_f, x0_f, md_f, lse = ctx.saved_tensors
        B, N, V, text_vocab, mask_index, ldv = ctx.meta
        l2 = logits.as_strided((B * N, ldv), (ldv, 1))
        ops.subs_nll_bwd_(l2, xt_f, x0_f, md_f, lse, dlogp.reshape(-1).contiguous().float(), V, text_vocab, mask_index)
        return logits, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------------------------
# batch contract: reference Diffusion.update_batch, model.py:157-395 — pre-tokenised datasets
# ----------------------------------------------------------------------------------------------------------------
def contiguous_blocks(ids: torch.Tensor):
    """Runs of equal values along the sequence: (batch_indices, start_positions, end_positions) of every run whose value is
    >= 0 (reference unidisc/utils/tensor_utils.py:24-44)."""
    diff = ids[:, 1:] != ids[:, :-1]
    diff = torch.nn.functional.pad(diff, (1, 0), mode="constant", value=True)
    starts = diff.nonzero(as_tuple=False)
    ends = torch.nn.functional.pad(diff[:, 1:], (0, 1), mode="constant", value=True).nonzero(as_tuple=False)
    bi, sp, ep = starts[:, 0], starts[:, 1], ends[:, 1] + 1
    valid = ids[bi, sp] >= 0
    return bi[valid], sp[valid], ep[valid]


def update_batch(batch, config, *, text_vocab_size: int, device, training: bool = True):
    """Dataloader batch -> the dict `compute_loss` consumes (reference model.py:157-395), for the tokenised paths of the shipped
    configs: (a) `txt_input_ids` / `txt_attention_mask` / `img_input_ids` (image ids are shifted by `text_vocab_size`,
    model.py:200), (b) already joint `input_ids` + `modality` (trainer.multimodal_batches).  Adds `modality`, `modality_mask`
    (one-hot bool [B,N,2]), `batch_contains_img`, `txt_sl` / `img_sl`, normalises dtypes, applies the `sample_ids` / padding
    rules of data.require_sample_ids and attaches the interleaved block metadata.  Raw-image batches (VAE tokenisation,
    model.py:215-283), class labels and the AR flip are outside the hot path and raise."""
    if batch is None:
        return batch
    g = _g
    tr, md, data = config.trainer, config.model, g(config, "data")
    if g(g(config, "eval"), "big_seq_len_eval", False) or g(tr, "image_mode", "discrete") == "continuous" or g(tr, "add_label", False) \
            or g(md, "img_cond", False) or g(tr, "force_remove_img_tokens", False):
        raise NotImplementedError("unidisc_b200.update_batch: big_seq_len_eval / continuous / add_label / img_cond / "
                                  "force_remove_img_tokens are not on the hot path")
    batch = dict(batch.items())
    if "txt_input_ids" in batch or "img_input_ids" in batch:                           # model.py:183-210
        for key in ("img_input_ids", "txt_input_ids", "sample_ids"):
            if key in batch:
                if isinstance(batch[key], list):
                    batch[key] = torch.stack(batch[key], dim=0)
                batch[key] = batch[key].to(torch.int64)
        img_input_ids = batch.pop("img_input_ids")
        batch["input_ids"] = img_input_ids
        batch["attention_mask"] = torch.ones_like(img_input_ids).to(torch.bool)
        if "txt_input_ids" in batch:
            batch["input_ids"] = torch.cat([batch["txt_input_ids"], batch["input_ids"] + text_vocab_size], dim=-1)
            batch["attention_mask"] = torch.cat([batch["txt_attention_mask"], batch["attention_mask"]], dim=-1)
        batch["input_ids"] = batch["input_ids"].to(torch.int64)
        if "modality" not in batch:
            if g(tr, "ignore_text_in_unified", False):
                modality = torch.ones_like(batch["input_ids"], dtype=torch.int64)
            else:
                assert md.txt_length > 0 and md.img_length > 0
                modality = torch.zeros_like(batch["input_ids"], dtype=torch.int64)
                modality[:, -img_input_ids.shape[-1]:] = 1
            batch["modality"] = modality
    elif g(tr, "multimodal_batches", False) and not g(tr, "use_legacy_update_batch_fn", False):   # model.py:212-256
        if "img" in batch:
            raise NotImplementedError("unidisc_b200.update_batch: raw-image batches need the VAE tokeniser (model.py:215-232)")
        batch["input_ids"] = batch["input_ids"].to(torch.int64)
        if "sample_ids" in batch:
            batch["sample_ids"] = batch["sample_ids"].to(torch.int64)
        if g(tr, "force_shift_image_batches", False):
            batch["input_ids"] = torch.where(batch["modality"] == 1, batch["input_ids"] + text_vocab_size, batch["input_ids"])
    else:
        raise NotImplementedError("unidisc_b200.update_batch: only tokenised batches (txt/img_input_ids or multimodal_batches)")
    if batch["input_ids"].shape[1] != md.length and not g(tr, "ar_inpainting", False):     # model.py:284-287
        raise AssertionError(f"input ids are not the correct length input ids shape: {batch['input_ids'].shape}, model length: {md.length}")
    batch["modality"] = batch["modality"].to(torch.int64)                                   # model.py:293-316
    if g(tr, "multimodal_batches", False) and batch["modality"].ndim == 2 and batch["modality"].shape[-1] == 1:
        batch["modality"] = batch["modality"].repeat(1, md.length)
    batch["modality"][batch["modality"] == -1] = 0
    assert batch["modality"].min() == 0 and batch["modality"].max() == 1
    batch["modality_mask"] = torch.nn.functional.one_hot(batch["modality"], num_classes=2).to(torch.bool)
    batch["batch_contains_img"] = (batch["modality"] == 1).any(dim=-1)
    batch["txt_sl"] = batch["modality_mask"][..., 0]
    batch["img_sl"] = batch["modality_mask"][..., 1]
    for key in batch.keys():                                                                # model.py:338-341
        if isinstance(batch[key], torch.Tensor):
            batch[key] = batch[key].to(device)
    if g(tr, "force_full_attention_mask", False):
        batch["attention_mask"] = torch.ones_like(batch["attention_mask"], dtype=torch.bool)
    batch["attention_mask"] = batch["attention_mask"].to(torch.bool)
    if g(data, "require_sample_ids", False):                                                # model.py:351-354
        assert "sample_ids" in batch
        batch["sample_ids"][~(batch["attention_mask"].bool())] = -1
        batch["attention_mask"][batch["sample_ids"] == -1] = False
    if g(tr, "rand_flip_ar_prob", None) is not None and g(config, "parameterization", "subs") == "ar":
        raise NotImplementedError("unidisc_b200.update_batch: AR flip (model.py:358-374)")
    if g(tr, "interleaved", False):                                                         # model.py:376-393
        if "sample_ids" not in batch:
            batch["sample_ids"] = torch.zeros_like(batch["modality"], dtype=torch.int64)
        bi, sp, ep = contiguous_blocks(batch["modality"])
        batch["interleaved_metadata"] = dict(batch_indices=bi, start_positions=sp, end_positions=ep)
    return batch



def first_hitting_select(x, x_sampled, num_unmask, random_values, mask_index):
    """Selection step of the first-hitting sampler (reference model_eval.py:3028-3043): per row, `num_unmask` of the still
    masked positions — the ones with the largest `random_values` — take their sampled token; everything else is kept.
    Pure tensor logic (any device); `random_values = torch.rand_like(copy_flag, dtype=float32)` in the reference."""
    copy_flag = x != mask_index
    num_unmask = torch.minimum(num_unmask, (~copy_flag).sum(dim=-1))
    if torch.all(num_unmask <= 0):
        return x
    rv = torch.where(~copy_flag, random_values, -1)
    _, indices = torch.sort(rv, dim=-1, descending=True)
    range_tensor = torch.arange(copy_flag.shape[-1], device=copy_flag.device).expand(copy_flag.shape)
    final_mask = range_tensor < num_unmask[:, None]
    result = torch.zeros_like(copy_flag)
    result.scatter_(-1, indices, final_mask)
    return torch.where(result, x_sampled, x)



class Diffusion(nn.Module):
    def __init__(self, config, tokenizer=None, device=None, vocab_size: Optional[int] = None,
                 text_vocab_size: Optional[int] = None, mask_index: Optional[int] = None):
        super().__init__()
        config = _wrap_cfg(config)
        self.config = config
        self.tokenizer = tokenizer
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        prec = str(_g(config.trainer, "precision", "bf16"))
        self.dtype = torch.float32 if ("fp32" in prec or prec == "no") else bf16
        if self.dtype != bf16:
            raise NotImplementedError("unidisc_b200: only trainer.precision=bf16 is built (the reference's CUDA autocast mode)")
        # ---- vocabulary / mask index: reference model_setup.py:90-115 (unified text+image model) ----
        if vocab_size is None:
            tv = _g(config.model, "force_text_vocab_size", None)
            if tv is None:
                if tokenizer is None:
                    raise ValueError("need a tokenizer or explicit vocab sizes")
                tv = len(tokenizer)
            if tokenizer is None or getattr(tokenizer, "mask_token", None) is None:
                mask_index = tv
                tv += 1
            else:
                mask_index = tokenizer.mask_token_id
            text_vocab_size = tv
            vocab_size = tv + int(config.model.image_vocab_size)
        self.vocab_size, self.text_vocab_size, self.mask_index = int(vocab_size), int(text_vocab_size), int(mask_index)
        self.parameterization = _g(config, "parameterization", "subs")
        if self.parameterization != "subs" or _g(config, "backbone", "dit") != "dit":
            raise NotImplementedError("unidisc_b200: only backbone=dit with parameterization=subs is on the hot path")
        self.T = int(_g(config, "T", 0))
        self.time_conditioning = bool(_g(config, "time_conditioning", False))
        self.sampler = _g(_g(config, "sampling"), "predictor", "ddpm_cache")
        self.antithetic_sampling = bool(_g(config.trainer, "antithetic_sampling", True))
        self.importance_sampling = bool(_g(config.trainer, "importance_sampling", False))
        self.change_of_variables = bool(_g(config.trainer, "change_of_variables", False))
        self.sampling_eps = float(_g(config.trainer, "sampling_eps", 1e-3))
        self.neg_infinity = -1_000_000.0
        self.allow_slicing = False
        self.noise = noise_schedule.get_noise(config)
        self.static_txt_sl = slice(None, config.model.txt_length)
        self.static_img_sl = slice(-config.model.img_length, None)
        self.backbone = DIT(config, vocab_size=self.vocab_size, text_vocab_size=self.text_vocab_size, mask_index=self.mask_index,
                            static_txt_sl=self.static_txt_sl, static_img_sl=self.static_img_sl)
        self.backbone.to(self.device)
        self.global_step = 0
        self._backbone_kwargs = {}     # per-step backbone arguments of an attention-caching sampling run (block_mask, update_cache_slice)
        self.fast_rng = bool(_g(config.trainer, "b200_philox_rng", False))   # additive key: in-kernel Philox instead of torch.rand
        self._rng_offset = 0
        self._rng_seed = None

    def _philox_seed(self):
        """in-kernel Philox key: config.seed + rank (the reference seeds every rank with seed + rank, main.py:1062)"""
        if self._rng_seed is None:
            import torch.distributed as dist
            rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
            self._rng_seed = int(_g(self.config, "seed", 42)) + rank
        return self._rng_seed

    # ------------------------------------------------------------------------------------------------------------
    # training side
    # ------------------------------------------------------------------------------------------------------------
    def update_batch(self, batch):                                    # reference model.py:157-395
        return update_batch(batch, self.config, text_vocab_size=self.text_vocab_size, device=self.device, training=self.training)

    def _sample_t(self, n, device):                                   # reference model.py:589-619
        _eps_t = torch.rand(n, device=device)
        if self.antithetic_sampling:
            offset = torch.arange(n, device=device) / n
            _eps_t = (_eps_t / n + offset) % 1
        ft = _g(self.config.trainer, "force_timestep", None)
        if ft is not None:
            _eps_t[:] = ft
        t = (1 - self.sampling_eps) * _eps_t + self.sampling_eps
        return t.to(torch.float32)

    def q_xt(self, x, move_chance, allow_move_mask=None, return_ignore_batch_mask_for_metrics=False, mask_image_square=False,
             mask_text_region=False, batch=None):
        """reference model.py:424-587 (absorbing state; multimodal non-interleaved batches)."""
        if mask_image_square or mask_text_region:
            raise NotImplementedError("unidisc_b200.q_xt: mask_image_square / mask_text_region are eval-time visualisation paths")
        mask_prob = _g(self.config.trainer, "mask_entire_modality", None)
        plain = (mask_prob is None or not self.backbone.training) and allow_move_mask is None \
            and _g(self.config.trainer, "joint_ar_nar_prob", None) is None and not _g(self.config.trainer, "add_label", False) \
            and _g(self.config.trainer, "first_token_dropout", None) is None
        if plain:
            # one kernel: xt = (rand < move_chance) ? mask : x      (model.py:439,579)
            if self.fast_rng:
                self._rng_offset += 1
                xt, move = ops.q_xt(x, move_chance, self.mask_index, rand=None, seed=self._philox_seed(),
                                    offset=self._rng_offset, return_move=True)
            else:
                rnd = torch.rand(*x.shape, device=x.device)
                xt, move = ops.q_xt(x, move_chance, self.mask_index, rand=rnd, return_move=True)
            if return_ignore_batch_mask_for_metrics:
                return xt, None, None, None, None, move
            return xt
        # general path: the whole-modality masking logic works on [B,1] / [B,N] booleans (model.py:470-579)
        xt, ignore, smt, smi, move_indices = q_xt_general(x, move_chance, self.mask_index, self.config.trainer,
                                                          backbone_training=self.backbone.training, training=self.training,
                                                          batch=batch, allow_move_mask=allow_move_mask)
        if return_ignore_batch_mask_for_metrics:
            return xt, ignore, None, smt, smi, move_indices
        return xt

    def _process_sigma(self, sigma):                                  # reference model.py:660-672
        if sigma is None:
            return sigma
        if sigma.ndim > 1:
            sigma = sigma.squeeze(-1)
        return sigma

    def _modality_of(self, batch, kwargs):
        md = kwargs.get("modality", None)
        if md is None and batch is not None:
            md = batch.get("modality", None)
        if md is None:
            raise ValueError("modality is required (trainer.multimodal_batches)")
        return md

    def _subs_parameterization(self, logits, xt, batch=None, modality=None, **kwargs):
        """reference model.py:621-658.  `logits` must be the bf16 tensor returned by the backbone; returns fp32
        log-probs [B,N,V] (the reference keeps bf16 here and upcasts at model.py:924-925; see DESIGN.md numerics)."""
        if modality is None:
            modality = batch["modality"]
        B, N, V = logits.shape
        ldv = logits.stride(1)
        l2 = logits.as_strided((B * N, ldv), (ldv, 1))
        tv = self.text_vocab_size if _g(self.config.model, "force_argmax_valid_indices", False) else -1
        out = ops.subs_logprobs(l2, None if xt is None else xt.reshape(-1).contiguous(), modality.reshape(-1).contiguous(), V, tv,
                                self.mask_index)
        return out.view(B, N, V)

    def forward(self, x, sigma, batch=None, forward_attention_mask=None, return_additional_loss=False, x_img_emb=None,
                disable_ar_shift=False, continuous_mode=False, joint_ar_nar_mask=None, return_logits=False, block_mask=None,
                update_cache_slice=None, **kwargs):
        """reference model.py:674-795 ("Returns log score")."""
        sigma = self._process_sigma(sigma)
        modality = self._modality_of(batch, kwargs)
        logits = self.backbone(x, sigma, modality=modality, sample_ids=kwargs.get("sample_ids", None), block_mask=block_mask)
        if return_logits:
            return logits
        return self._subs_parameterization(logits, xt=x, batch=batch, modality=modality)

    def _log_p_x0(self, logits, xt, x0, modality):
        tv = self.text_vocab_size if _g(self.config.model, "force_argmax_valid_indices", False) else -1
        return _SubsNLL.apply(logits, xt, x0, modality, self.vocab_size, tv, self.mask_index)

    def compute_loss(self, batch, prefix="train", batch_idx=-1):
        """reference model.py:797-1173 (discrete diffusion, subs, T=0, multimodal loss weighting)."""
        x0 = batch["input_ids"]
        attention_mask = batch.get("attention_mask", None)
        if attention_mask is None:
            attention_mask = torch.ones_like(x0, dtype=torch.bool)
        modality = batch["modality"]
        modality_mask = batch.get("modality_mask", None)
        if modality_mask is None:
            modality_mask = torch.stack([modality == 0, modality == 1], dim=-1)
        t = self._sample_t(x0.shape[0], x0.device)                                   # model.py:844
        sigma, dsigma = self.noise(t)                                                # model.py:858
        move_chance = 1 - torch.exp(-sigma[:, None])                                 # model.py:860
        xt, ignore_batch, _, _, _, _ = self.q_xt(x0, move_chance, return_ignore_batch_mask_for_metrics=True, batch=batch)
        # model.py:876-878: packed batches hand the backbone sample_ids + the document BlockMask (here: a flag, the mask is
        # evaluated inside the attention kernels from sample_ids)
        flex = bool(_g(self.config.trainer, "interleaved_training_flex_attention", False))
        # Additive key trainer.b200_masked_head (default true): the SUBS parameterisation gives every UNMASKED token log p = 0
        # exactly (model.py:621-658) and padding is multiplied by the attention mask below, so only the masked, attended rows
        # need logits.  The output projection (7 % of the step's FLOPs over all rows), the fused NLL and their backward then run
        # on those rows only: same loss, same gradients.  Costs one device->host read of the row count per step.
        # With model.force_argmax_valid_indices (each token restricted to its modality's vocabulary, model.py:627-640) the rows are
        # ordered text first and every row is projected onto its own vocabulary block only (`head_split`).
        resolve, picked = None, [None, None]
        if bool(_g(self.config.trainer, "b200_masked_head", True)) and getattr(self.backbone, "supports_head_rows", False):
            sel = ((xt == self.mask_index) & attention_mask.bool()).reshape(-1)
            split_ok = bool(_g(self.config.model, "force_argmax_valid_indices", False)) and bool(_g(self.config.trainer, "b200_split_head", True))
            is_img = modality.reshape(-1) != 0
            sel_t, sel_i = sel & ~is_img, sel & is_img
            # the two row counts travel to the host asynchronously; they are read right before the output projection, when the copy
            # (enqueued ahead of the whole backbone forward) has long finished: the host never stalls on the GPU
            if getattr(self, "_head_counts_host", None) is None:
                self._head_counts_host = torch.empty(2, dtype=torch.int64, pin_memory=True)
            self._head_counts_host.copy_(torch.stack([sel_t.sum(), sel_i.sum()]), non_blocking=True)
            counts_ready = torch.cuda.Event()
            counts_ready.record()

            def resolve():
                counts_ready.synchronize()
                n_t, n_i = self._head_counts_host.tolist()
                picked[0], picked[1] = select_head_rows(sel_t, sel_i, n_t, n_i, split_ok)
                return picked[0], picked[1]
        bb_kwargs = dict(modality=modality,
                         sample_ids=batch.get("sample_ids", None) if (flex or self.backbone.require_sample_ids) else None,
                         block_mask=True if flex else None)
        logits = self.backbone(xt, None if not self.time_conditioning else sigma, **(dict(head_rows=resolve) if resolve else {}),
                               **bb_kwargs)
        head_rows, head_split = picked
        if head_rows is None:
            log_p_theta = self._log_p_x0(logits, xt, x0, modality)                   # model.py:787 + :967, fused
        else:
            take = lambda a: a.reshape(-1).index_select(0, head_rows)[None]
            lp = self._log_p_x0(logits, take(xt), take(x0), take(modality))           # [1, rows]
            log_p_theta = torch.zeros(x0.numel(), device=x0.device, dtype=lp.dtype).index_copy(0, head_rows, lp.reshape(-1))
            log_p_theta = log_p_theta.view(x0.shape)
        self._last_head_rows = None if head_rows is None else int(head_rows.numel())
        self._last_head_split = head_split
        std_weighting = (dsigma / torch.expm1(sigma))[:, None]                       # model.py:975
        loss = -log_p_theta * std_weighting
        gamma = _g(self.config.trainer, "softmin_snr", None)
        if gamma is not None:                                                        # model.py:990-993
            loss = -log_p_theta * (dsigma / (torch.expm1(sigma) + (1 / gamma)))[:, None]
        std_loss = (-log_p_theta * std_weighting).detach()
        am = attention_mask.bool()
        txt_mask = modality_mask[..., 0] & am                                        # model.py:1021-1022
        img_mask = modality_mask[..., 1] & am
        txt_count, img_count = txt_mask.sum(), img_mask.sum()
        total = txt_count + img_count
        extra = {"trainer/img_frac": img_count / total, "trainer/txt_frac": txt_count / total,
                 "trainer/attention_mask_valid_frac": am.sum() / am.numel()}
        tw, iw = _g(self.config.trainer, "text_loss_weight", None), _g(self.config.trainer, "img_loss_weight", None)
        loss = loss * am
        if tw is not None and iw is not None:                                        # model.py:1036-1057
            txt_loss = ((loss * txt_mask).sum() / txt_count) * (txt_count / total) * tw
            img_loss = ((loss * img_mask).sum() / img_count) * (img_count / total) * iw
            ratio = _g(self.config.trainer, "set_max_txt_loss_ratio", None)
            if ratio is not None:
                scale = torch.minimum(torch.ones((), device=loss.device), ratio * img_loss.detach() / (txt_loss.detach() + 1e-8))
                scale = torch.where(torch.isnan(img_loss) | torch.isnan(txt_loss), torch.ones_like(scale), scale)
                txt_loss = txt_loss * scale
            txt_loss = torch.nan_to_num(txt_loss, nan=0.0)
            img_loss = torch.nan_to_num(img_loss, nan=0.0)
            total_loss = txt_loss + img_loss
        else:                                                                        # model.py:1070-1073
            total_loss = torch.nan_to_num(loss.sum() / am.sum(), nan=0.0)
            txt_loss = img_loss = torch.zeros((), device=loss.device)
        token_mask = am
        if ignore_batch is not None:
            token_mask = torch.where(ignore_batch.reshape(-1, 1), torch.zeros_like(am), am)
        return Loss(loss=total_loss, img_loss=img_loss.detach(), txt_loss=txt_loss.detach(), nlls=std_loss * am,
                    txt_nlls=std_loss * modality_mask[..., 0] * am, img_nlls=std_loss * modality_mask[..., 1] * am,
                    token_mask=token_mask, modality_mask=modality_mask, extra_losses=extra)

    def training_step(self, batch, batch_idx=-1):                                    # reference model.py:420-422
        return self.compute_loss(batch, prefix="train", batch_idx=batch_idx)

    # ------------------------------------------------------------------------------------------------------------
    # sampling side
    # ------------------------------------------------------------------------------------------------------------
    def _sample_prior(self, *batch_dims):                                            # reference model_eval.py:1734
        return self.mask_index * torch.ones(*batch_dims, dtype=torch.int64)

    def get_cfg_weight(self, t):                                                     # reference model_eval.py:1737-1759
        cfg = self.config.eval.cfg
        return (cfg * (1 - t))[:, None]

    def _ddpm_forward(self, x, t, sigma_t, x0=None, x0_unmask=None, force_cfg=None, **kwargs):
        """reference model_eval.py:1761-1833: returns p_x0 [B,N,V] (fp32, materialised — the API-parity path; the samplers'
        default paths consume the logits directly)."""
        modality = kwargs.get("modality")
        B, N = x.shape
        lc, lu, w = self._sampling_logits(x, t, x0_unmask, modality, kwargs.get("sample_ids"))
        V = self.vocab_size
        if lu is not None and bool((w > 0).any()):
            out = ((1 + w[:, None, None]) * lc[:, :V].float().view(B, N, V) - w[:, None, None] * lu[:, :V].float().view(B, N, V)).to(bf16)
            buf = torch.zeros((B * N, self.backbone.Vp), device=out.device, dtype=bf16)   # model_eval.py:1812 (then SUBS, xt=None)
            buf[:, :V] = out.reshape(B * N, V)
            return self._subs_parameterization(buf.view(B, N, -1)[:, :, :V], xt=None, modality=modality).exp()
        ldv = lc.stride(0)
        return self._subs_parameterization(lc.as_strided((B, N, V), (N * ldv, ldv, 1)), xt=x, modality=modality).exp()

    def _fused_absorbing_update(self, x, t, mc_t, mc_s, x0=None, x0_unmask=None, modality=None, u=None, sample_ids=None,
                                logits_cache=None, return_cache=False):
        """backbone -> (CFG) -> SUBS softmax -> q_xs -> Gumbel arg-max -> copy-through in ONE vocabulary pass.
        `logits_cache` = the (lc, lu, w) of a previous call on an unchanged x: the backbone forward is skipped (ddpm_cache)."""
        lc, lu, w = logits_cache if logits_cache is not None else self._sampling_logits(x, t, x0_unmask, modality, sample_ids)
        if return_cache:
            return (lc, lu, w), self._fused_absorbing_update(x, t, mc_t, mc_s, modality=modality, u=u, logits_cache=(lc, lu, w))
        self._rng_offset += 1
        tv = self.text_vocab_size if _g(self.config.model, "force_argmax_valid_indices", False) else -1
        return ops.ddpm_update_logits(x, lc, modality.reshape(-1).contiguous(), mc_t.float().contiguous(), mc_s.float().contiguous(),
                                      self.mask_index, tv, self.vocab_size, logits_uncond=lu, cfg_w=w, u=u,
                                      seed=self._philox_seed(), offset=self._rng_offset)

    @torch.no_grad()
    def _ddpm_update(self, x, t, dt, **kwargs):                                      # reference model_eval.py:2042-2070
        tt = t.reshape(-1)
        sigma_t, _ = self.noise(tt)
        sigma_s, _ = self.noise(tt - dt)
        mc_t, mc_s = 1 - torch.exp(-sigma_t), 1 - torch.exp(-sigma_s)
        if kwargs.pop("parity_noise", False):
            p_x0 = self._ddpm_forward(x, t, None, **kwargs)
            u = torch.rand_like(p_x0)
            return ops.ddpm_update_probs(x, p_x0, mc_t.contiguous(), mc_s.contiguous(), self.mask_index, u=u.view(-1, u.shape[-1])), 1
        return self._fused_absorbing_update(x, tt, mc_t, mc_s, **{k: kwargs.get(k) for k in ("x0", "x0_unmask", "modality", "sample_ids")}), 1

    @torch.no_grad()
    def _ddpm_caching_update(self, x, t, dt, p_x0=None, x0=None, x0_unmask=None, modality=None, **kwargs):
        """reference model_eval.py:2072-2104.  With a p_x0 TENSOR given (or parity_noise=True) the update consumes a
        materialised fp32 p_x0 and torch.rand noise exactly like the reference; otherwise the fused single-pass kernel is
        used and the returned cache is the raw bf16 logits of this step's forward (never a materialised p_x0): handing it back
        while x is unchanged skips the backbone exactly where the reference reuses p_x0_cache (CFG weights follow t, as the
        reference's cached p_x0 does not — the cache is only reused without CFG)."""
        tt = t.reshape(-1)
        mc_t, mc_s = tt, tt - dt
        nfe = 0
        if isinstance(p_x0, tuple):                              # logits cache of the fused path
            if p_x0[1] is None:
                return p_x0, self._fused_absorbing_update(x, tt, mc_t, mc_s, modality=modality, logits_cache=p_x0), 0
            p_x0 = None
        if p_x0 is not None or kwargs.get("parity_noise", False):
            if p_x0 is None:
                p_x0 = self._ddpm_forward(x, t, None, x0=x0, x0_unmask=x0_unmask, modality=modality)
                nfe = 1
            u = torch.rand_like(p_x0)
            xn = ops.ddpm_update_probs(x, p_x0, mc_t.float().contiguous(), mc_s.float().contiguous(), self.mask_index,
                                       u=u.view(-1, u.shape[-1]))
            return p_x0, xn, nfe
        cache, xn = self._fused_absorbing_update(x, tt, mc_t, mc_s, x0=x0, x0_unmask=x0_unmask, modality=modality,
                                                 sample_ids=kwargs.get("sample_ids"), return_cache=True)
        return cache, xn, 1

    @staticmethod
    def adap_sche(x, step, mask_index, mode="arccos"):                               # reference model_eval.py:2964-3001
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
        val = val.to(x.device)
        out = []
        for seq_len in num_masked:
            sche = (val / val.sum()) * seq_len
            sche = sche.round()
            sche[sche == 0] = 1
            sche[-1] += seq_len - sche.sum()
            sche[-1] = max(sche[-1], 0)
            out.append(sche.int())
        return torch.stack(out, dim=0)

    def _sampling_logits(self, x, t, x0_unmask, modality, sample_ids=None):
        """backbone forward of a sampling step (+ the CFG pair of model_eval.py:1763-1812): returns (lc, lu, w) as 2-D views of
        the padded bf16 logits buffer; lu / w are None without classifier-free guidance."""
        B, N = x.shape
        sigma = self._sampling_sigma(t)
        cfg = _g(_g(self.config, "eval"), "cfg", None)
        use_cfg = cfg is not None and x0_unmask is not None and bool(x0_unmask.any())
        kw = dict(sample_ids=sample_ids) if sample_ids is not None else {}
        kw.update(self._backbone_kwargs)
        if use_cfg:
            x_uncond = torch.where(x0_unmask, torch.full_like(x, self.mask_index), x)
            if sample_ids is not None:
                kw["sample_ids"] = torch.cat([sample_ids, sample_ids], 0)
            lg = self.backbone(torch.cat([x, x_uncond], 0), None if sigma is None else torch.cat([sigma, sigma], 0),
                               modality=torch.cat([modality, modality], 0), **kw)
            ldv = lg.stride(1)
            l2 = lg.as_strided((2 * B * N, ldv), (ldv, 1))
            return l2[: B * N], l2[B * N:], (cfg * (1 - t.reshape(-1))).float().contiguous()
        lg = self.backbone(x, sigma, modality=modality, **kw)
        ldv = lg.stride(1)
        return lg.as_strided((B * N, ldv), (ldv, 1)), None, None

    def _sampling_sigma(self, t):
        """sigma handed to the backbone while sampling (model_eval.py:3060: sigma_t unless trainer.force_null_sigma); only a
        time-conditioned backbone looks at it."""
        if not self.time_conditioning or _g(self.config.trainer, "force_null_sigma", False):
            return None
        return self.noise(t.reshape(-1))[0]

    @torch.no_grad()
    def _maskgit_update(self, x, t, dt, schedule=None, step=None, parity_noise=None, **kwargs):
        """reference model_eval.py:3045-3114 as two kernels (csrc/loss_sampler.cu): one vocabulary pass per masked row (SUBS
        softmax -> multinomial draw -> confidence) and a per-sample k-th-largest selection; the [B,N,V] probability tensor,
        torch.multinomial and torch.topk of the reference never materialise.  `parity_noise=(E, gumbel)`: the Exp(1) tensor
        torch.multinomial draws internally (fp32 [B,N,V]) and the np.random.gumbel draw (fp64 [B,N]) — bit-exact mode."""
        copy_flag = x != self.mask_index
        r_temp = _g(_g(self.config, "eval"), "maskgit_r_temp", 10)
        sched = schedule[:, step].to(x.device).to(torch.int32).contiguous()
        num_unmask = torch.minimum(sched, (~copy_flag).sum(dim=-1).to(torch.int32))
        if torch.all(num_unmask <= 0):                                               # model_eval.py:3070-3071
            return x, 0
        modality = kwargs.get("modality")
        lc, lu, w = self._sampling_logits(x, t, kwargs.get("x0_unmask"), modality, kwargs.get("sample_ids"))
        self._rng_offset += 1
        tv = self.text_vocab_size if _g(self.config.model, "force_argmax_valid_indices", False) else -1
        e_noise, gumbel = parity_noise if parity_noise is not None else (None, None)
        out, _, _ = ops.maskgit_update(x, lc, modality.reshape(-1).contiguous(), t.reshape(-1).float().contiguous(), sched, self.mask_index,
                                       tv, self.vocab_size, r_temp=float(r_temp), logits_uncond=lu, cfg_w=w,
                                       e_noise=None if e_noise is None else e_noise.reshape(-1, e_noise.shape[-1]), gumbel=gumbel,
                                       seed=self._philox_seed(), offset=self._rng_offset)
        return out, 1

    @torch.no_grad()
    def _first_hitting_update(self, x, t, dt, schedule=None, step=None, **kwargs):   # reference model_eval.py:3004-3043
        copy_flag = x != self.mask_index
        num_unmask = torch.minimum(schedule[:, step].to(x.device), (~copy_flag).sum(dim=-1))
        p_x0 = self._ddpm_forward(x, t, None, **{k: kwargs.get(k) for k in ("x0", "x0_unmask", "modality")})
        x_sampled = ops.sample_categorical(p_x0, u=torch.rand_like(p_x0))            # model_utils.py:95-97, same draw order
        if torch.all(num_unmask <= 0):
            return x, 1
        random_values = torch.rand_like(copy_flag, dtype=torch.float32)
        return first_hitting_select(x, x_sampled, num_unmask, random_values, self.mask_index), 1

    @torch.no_grad()
    def _sample(self, num_steps=None, eps=1e-5, text_only=True, x0=None, x0_unmask=None, batch_size_per_gpu=None,
                example_batch=None, sample_batch_idx=None, sample_modality=None, sample_ids=None, return_raw_data=False,
                return_nfe=False, **kwargs):
        """reference model_eval.py:2107-2454 (token-space part: returns the sampled token ids [B,N])."""
        assert (x0 is None) == (x0_unmask is None)
        B = x0.shape[0] if x0 is not None else (batch_size_per_gpu or _g(_g(self.config, "loader"), "eval_batch_size", 1))
        N = self.config.model.length
        modality = sample_modality if sample_modality is not None else kwargs.get("modality")
        if modality is None:
            raise ValueError("sample_modality is required")
        if num_steps is None:
            num_steps = _g(_g(self.config, "sampling"), "steps", 64)
        x = self._sample_prior(B, N).to(self.device)
        if x0_unmask is None:
            x0_unmask = torch.zeros(B, N, dtype=torch.bool, device=self.device)
            x0 = x.clone()
        num_steps = int(min(num_steps, int((~x0_unmask).sum(dim=-1).min())))
        x = torch.where(x0_unmask, x0, x)
        schedule = None
        if self.sampler in ("maskgit", "first_hitting"):                             # model_eval.py:2274
            schedule = self.adap_sche(x, num_steps, self.mask_index, mode="arccos")
        timesteps = torch.linspace(1, eps, num_steps + 1, device=self.device)
        dt = (1 - eps) / num_steps
        p_cache, nfe = None, 0
        parity = bool(kwargs.get("parity_noise", False))
        # ---- inference attention caching (eval.attention_caching, reference model_eval.py:2297-2367): cycles of `ratio` steps —
        # one full joint step, one step that stores the image K/V (image queries see image keys only), then text-only steps ----
        ev = _g(self.config, "eval")
        caching = bool(_g(ev, "attention_caching", False))
        ratio = int(_g(ev, "attention_caching_txt_to_img_ratio", 10))
        txt_sl = self.static_txt_sl
        full, sliced = {}, False
        if caching:
            if sample_ids is not None:
                raise NotImplementedError("attention caching assumes the static [text | image] layout (no sample_ids)")
            use_cfg = _g(ev, "cfg", None) is not None and bool(x0_unmask.any())
            self.backbone.set_flex_attention_cache(B * (2 if use_cfg else 1), N, self.device)

        def restore(key, new):                                                       # model_eval.py:2314-2317
            full[key][:, txt_sl] = new
            return full[key]

        try:
            for i in range(num_steps):
                t = timesteps[i] * torch.ones(B, 1, device=self.device)
                if caching:
                    if i % ratio == 0:
                        if sliced:
                            x, x0, x0_unmask, modality = restore("x", x), restore("x0", x0), restore("x0_unmask", x0_unmask), full["modality"]
                            full, sliced, p_cache = {}, False, None
                        self._backbone_kwargs = dict(block_mask=True, update_cache_slice=None)
                    elif (i - 1) % ratio == 0:
                        self._backbone_kwargs = dict(block_mask=TextFullImageSelfMask(self.config.model.txt_length),
                                                     update_cache_slice=slice(0, N))
                    else:
                        self._backbone_kwargs = dict(block_mask=True, update_cache_slice=txt_sl)
                        if not sliced:
                            full.update(x=x.clone(), x0=x0.clone(), x0_unmask=x0_unmask.clone(), modality=modality)
                            x, x0, x0_unmask, modality = (v[:, txt_sl].contiguous() for v in (x, x0, x0_unmask, modality))
                            sliced, p_cache = True, None
                if self.sampler == "maskgit":
                    x, n = self._maskgit_update(x, t, dt, x0=x0, x0_unmask=x0_unmask, schedule=schedule, step=i, modality=modality,
                                                sample_ids=sample_ids)
                elif self.sampler == "first_hitting":                                # model_eval.py:2373-2374
                    x, n = self._first_hitting_update(x, t, dt, x0=x0, x0_unmask=x0_unmask, schedule=schedule, step=i, modality=modality)
                elif self.sampler == "ddpm":
                    x, n = self._ddpm_update(x, t, dt, x0=x0, x0_unmask=x0_unmask, modality=modality, parity_noise=parity)
                elif self.sampler == "ddpm_cache":
                    p_cache, x_next, n = self._ddpm_caching_update(x, t, dt, p_x0=p_cache, x0=x0, x0_unmask=x0_unmask,
                                                                   modality=modality, parity_noise=parity, sample_ids=sample_ids)
                    if p_cache is not None and (not torch.equal(x_next, x) or self.time_conditioning):
                        p_cache = None
                    x = x_next
                else:
                    raise NotImplementedError(f"sampling.predictor={self.sampler}")
                nfe += n
                x = torch.where(x0_unmask, x0, x)
            if sliced:                                                               # model_eval.py:2425-2438
                x, x0, x0_unmask, modality = restore("x", x), restore("x0", x0), restore("x0_unmask", x0_unmask), full["modality"]
        finally:
            self._backbone_kwargs = {}
            if caching:
                self.backbone.clear_flex_attention_cache()
        if _g(_g(self.config, "sampling"), "noise_removal", True):                   # model_eval.py:2440-2446
            t_last = timesteps[-1] * torch.ones(B, 1, device=self.device)
            lg = self.backbone(x, self._sampling_sigma(t_last), modality=modality,
                               **(dict(sample_ids=sample_ids) if sample_ids is not None else {}))
            ldv = lg.stride(1)
            l2 = lg.as_strided((B * N, ldv), (ldv, 1))
            # argmax of the SUBS log-probs with carry-over: x where unmasked, else argmax over the valid vocabulary — one
            # vocabulary pass, the fp32 [B,N,V] log-prob tensor is not materialised
            tv = self.text_vocab_size if _g(self.config.model, "force_argmax_valid_indices", False) else -1
            x = ops.subs_argmax(l2, x.reshape(-1).contiguous(), modality.reshape(-1).contiguous(), self.vocab_size, tv,
                                self.mask_index).view(B, N)
        x = torch.where(x0_unmask, x0, x)
        return (x, nfe) if return_nfe else x
