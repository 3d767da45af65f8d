"""Minimal config trees with the reference's key names (configs/config.yaml, configs/model/*.yaml,
configs/experiments/large_scale_train*.yaml).  A hydra/OmegaConf DictConfig composed from the reference's own YAML
tree works equally well — the classes only use attribute access — this helper exists so the package can run without
hydra installed (tests, bench)."""
from __future__ import annotations

from .dit import _wrap_cfg

# configs/model/*.yaml : hidden_size / n_blocks / n_heads
MODEL_PRESETS = {
    "xs": (256, 4, 8), "tiny": (512, 8, 8), "small": (768, 12, 12), "medium": (1024, 24, 16), "large": (1280, 28, 20),
    "extra_large": (2048, 24, 16), "xxl": (4096, 30, 16),
}


def make_config(preset="extra_large", *, txt_length=256, img_length=1024, image_vocab_size=16384, text_vocab_size=32001,
                hidden_size=None, n_blocks=None, n_heads=None, dropout=0.0, img_loss_weight=0.6, text_loss_weight=1.0,
                softmin_snr=None, mask_entire_modality=None, predictor="ddpm_cache", sampling_steps=64, cfg=None, seed=42,
                zero_linear_init=False, time_conditioning=False, **extra):
    d, l, h = MODEL_PRESETS[preset]
    d, l, h = hidden_size or d, n_blocks or l, n_heads or h
    cfgd = dict(
        mode="train", backbone="dit", parameterization="subs", time_conditioning=bool(time_conditioning), T=0, seed=seed,
        noise=dict(type="loglinear"),
        data=dict(require_sample_ids=False),
        model=dict(hidden_size=d, n_blocks=l, n_heads=h, cond_dim=128, dropout=dropout, length=txt_length + img_length,
                   txt_length=txt_length, img_length=img_length, norm_type="rms", qk_norm=True, sandwich_normalization=True,
                   rope_2d=True, modality_embed=True, full_attention=True, use_spda_attn=True, attn_type="flash",
                   force_varlen_attn=False, scale_by_sigma=False, zero_linear_init=zero_linear_init,
                   force_optimized_native_attn=False, image_vocab_size=image_vocab_size,
                   force_text_vocab_size=text_vocab_size - 1, force_argmax_valid_indices=True, use_attention_mask=False,
                   image_model=True, unified_model=True),
        trainer=dict(precision="bf16", multimodal_batches=True, image_mode="discrete", antithetic_sampling=True,
                     importance_sampling=False, change_of_variables=False, sampling_eps=1e-3, text_loss_weight=text_loss_weight,
                     img_loss_weight=img_loss_weight, softmin_snr=softmin_snr, mask_entire_modality=mask_entire_modality,
                     force_null_sigma=True, allow_null_sigma=True, interleaved=False, compile=False,
                     use_gradient_checkpointing=False, log_seperate_modal_losses=True),
        sampling=dict(predictor=predictor, steps=sampling_steps, noise_removal=True),
        eval=dict(cfg=cfg),
        loader=dict(eval_batch_size=1),
    )
    for k, v in extra.items():
        sec, key = k.split("__", 1)
        cfgd[sec][key] = v
    return _wrap_cfg(cfgd)
