"""TEST INFRASTRUCTURE — loader for the *unmodified* reference (`/root/reference`).

Only usable in the build container (the GPU box has no /root/reference).  It is used by
`oracle/gen_golden.py` to (a) pin the pure-torch restatement in `oracle/restated.py`
against the reference's own code and (b) write the committed fixtures in `tests/golden/`.

Two shims are installed:
  * `omegaconf`  — absent here; `models/dit.py:17,1098-1099` only needs `OmegaConf.create`.
  * `diffusers.models.embeddings.get_2d_rotary_pos_embed_lumina` — absent here; restated from
    diffusers 0.32.2 (pinned in the reference's uv.lock).  Parity for the 2-D RoPE table is
    therefore pinned by this restatement only ("parity unpinned" for that one table).

`Diffusion` (model.py) cannot be imported (accelerate/tensordict/hydra/... missing), so the
pure-tensor methods on the hot path are pulled out of the reference *source files* with `ast`
and exec'd against a minimal fake `self` — the code that runs is the reference's, verbatim,
read from where it lies; nothing is copied into this repo.
"""
from __future__ import annotations

import ast
import math
import os
import random
import sys
import types
from contextlib import ExitStack, nullcontext
from types import SimpleNamespace

import numpy as np
import torch

def _default_root():
    """/root/reference in the build container; on the GPU box (where it does not exist) the verbatim copy of the backbone's
    import closure that baseline/install_reference.py staged under baseline/_ref/ (backbone only: model.py & co are not there)."""
    if os.path.isfile("/root/reference/models/dit.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


REFERENCE_ROOT = os.environ.get("UNIDISC_REFERENCE_ROOT") or _default_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "dit.py"))


# --------------------------------------------------------------------------------------
# shims
# --------------------------------------------------------------------------------------
class AttrDict(dict):
    """Attribute-access dict standing in for an OmegaConf DictConfig (getattr with defaults works)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def get(self, k, default=None):
        return dict.get(self, k, default)


def to_attrdict(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attrdict(v) for k, v in d.items()})
    return d


def lumina_1d_rotary(dim: int, pos: int, theta: float = 10000.0, linear_factor: float = 1.0, ntk_factor: float = 1.0):
    """diffusers 0.32.2 `get_1d_rotary_pos_embed(..., use_real=False)` restated."""
    p = torch.arange(pos)
    theta = theta * ntk_factor
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: (dim // 2)] / dim)) / linear_factor
    freqs = torch.outer(p, freqs).float()
    return torch.polar(torch.ones_like(freqs), freqs)


def lumina_2d_rotary(embed_dim: int, len_h: int, len_w: int, linear_factor: float = 1.0, ntk_factor: float = 1.0):
    """diffusers 0.32.2 `get_2d_rotary_pos_embed_lumina` restated: (H, W, embed_dim/2) complex64."""
    assert embed_dim % 4 == 0
    emb_h = lumina_1d_rotary(embed_dim // 2, len_h, linear_factor=linear_factor, ntk_factor=ntk_factor)
    emb_w = lumina_1d_rotary(embed_dim // 2, len_w, linear_factor=linear_factor, ntk_factor=ntk_factor)
    emb_h = emb_h.view(len_h, 1, embed_dim // 4, 1).repeat(1, len_w, 1, 1)
    emb_w = emb_w.view(1, len_w, embed_dim // 4, 1).repeat(len_h, 1, 1, 1)
    return torch.cat([emb_h, emb_w], dim=-1).flatten(2)


def _install_shims():
    if "omegaconf" not in sys.modules:
        try:
            import omegaconf  # noqa: F401
        except Exception:
            m = types.ModuleType("omegaconf")

            class OmegaConf:  # noqa: D401
                @staticmethod
                def create(d):
                    return to_attrdict(d)

            m.OmegaConf = OmegaConf
            m.DictConfig = AttrDict
            sys.modules["omegaconf"] = m
    try:
        import diffusers.models.embeddings  # noqa: F401
    except Exception:
        d = types.ModuleType("diffusers")
        dm = types.ModuleType("diffusers.models")
        de = types.ModuleType("diffusers.models.embeddings")
        de.get_2d_rotary_pos_embed_lumina = lumina_2d_rotary
        d.models = dm
        dm.embeddings = de
        sys.modules["diffusers"] = d
        sys.modules["diffusers.models"] = dm
        sys.modules["diffusers.models.embeddings"] = de


_REF_DIT = None


def load_reference_dit_module():
    """Import `/root/reference/models/dit.py` (unmodified) and return the module."""
    global _REF_DIT
    if _REF_DIT is not None:
        return _REF_DIT
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    _REF_DIT = importlib.import_module("models.dit")
    return _REF_DIT


def make_ref_config(hidden_size, n_heads, n_blocks, txt_length, img_length, *, dropout=0.0, time_conditioning=False,
                    cond_dim=128, norm_type="rms", qk_norm=True, sandwich=True, require_sample_ids=False,
                    zero_linear_init=False, linear_factor=1.0):
    """Minimal config tree `DIT.__init__` reads (SURVEY.md §8c key list)."""
    return to_attrdict(dict(
        time_conditioning=time_conditioning,
        parameterization="subs",
        trainer=dict(image_mode="discrete", multimodal_batches=True, use_gradient_checkpointing=False,
                     compile=False, compile_flag_pos_emb=True),
        data=dict(require_sample_ids=require_sample_ids),
        model=dict(hidden_size=hidden_size, cond_dim=cond_dim, n_blocks=n_blocks, n_heads=n_heads, dropout=dropout,
                   scale_by_sigma=False, length=txt_length + img_length, txt_length=txt_length, img_length=img_length,
                   attn_type="flash", force_varlen_attn=False, norm_type=norm_type, qk_norm=qk_norm,
                   full_attention=True, rope_2d=True, modality_embed=True, zero_linear_init=zero_linear_init,
                   force_optimized_native_attn=False, use_spda_attn=True, sandwich_normalization=sandwich,
                   linear_factor=linear_factor),
    ))


def build_reference_dit(cfg, vocab_size, text_vocab_size, mask_index, dtype=torch.float32):
    mod = load_reference_dit_module()
    return mod.DIT(cfg, vocab_size=vocab_size, text_vocab_size=text_vocab_size, mask_index=mask_index, dtype=dtype)


# --------------------------------------------------------------------------------------
# pulling pure-tensor methods out of the reference source without importing the module
# --------------------------------------------------------------------------------------
def _extract_functions(path: str, names):
    src = open(path).read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            node.decorator_list = []  # drop @try_except / @torch.inference_mode wrappers
            found[node.name] = ast.get_source_segment(src, node) if False else ast.unparse(node)
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError(f"{path}: functions not found: {missing}")
    return found


def _exec_functions(found: dict, extra_globals: dict):
    g = dict(torch=torch, math=math, random=random, np=np, F=torch.nn.functional, Tensor=torch.Tensor,
             ExitStack=ExitStack, nullcontext=nullcontext, is_xla_available=False, rprint=lambda *a, **k: None,
             gprint=lambda *a, **k: None, print=lambda *a, **k: None, empty_device_cache=lambda: None,
             clear_gpu_memory_if_needed=lambda: None, typing=__import__("typing"))
    g.update(extra_globals)
    for name, code in found.items():
        exec(compile(code, f"<reference:{name}>", "exec"), g)
    return g


def load_reference_diffusion_methods():
    """Return a namespace of the reference's own hot-path functions (unbound; first arg is `self`).

    model.py: q_xt, _sample_t, _subs_parameterization, _process_sigma
    model_utils.py: _sample_categorical
    model_eval.py: _sample_prior, get_cfg_weight, _ddpm_forward, _ddpm_update, _ddpm_caching_update,
                   adap_sche, _maskgit_update, _first_hitting_update
    """
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from unidisc.utils.tensor_utils import get_contiguous_blocks_per_sample  # importable as-is

    g_utils = _exec_functions(_extract_functions(os.path.join(REFERENCE_ROOT, "model_utils.py"), ["_sample_categorical"]), {})
    g_model = _exec_functions(
        _extract_functions(os.path.join(REFERENCE_ROOT, "model.py"),
                           ["q_xt", "_sample_t", "_subs_parameterization", "_process_sigma"]),
        dict(get_contiguous_blocks_per_sample=get_contiguous_blocks_per_sample,
             linear_warmup=lambda **k: k.get("final_value")))
    g_eval = _exec_functions(
        _extract_functions(os.path.join(REFERENCE_ROOT, "model_eval.py"),
                           ["_sample_prior", "get_cfg_weight", "_ddpm_forward", "_ddpm_update", "_ddpm_caching_update",
                            "adap_sche", "_maskgit_update", "_first_hitting_update"]),
        dict(_sample_categorical=g_utils["_sample_categorical"]))
    ns = SimpleNamespace()
    for k in ["q_xt", "_sample_t", "_subs_parameterization", "_process_sigma"]:
        setattr(ns, k, g_model[k])
    for k in ["_sample_prior", "get_cfg_weight", "_ddpm_forward", "_ddpm_update", "_ddpm_caching_update", "adap_sche",
              "_maskgit_update", "_first_hitting_update"]:
        setattr(ns, k, g_eval[k])
    ns._sample_categorical = g_utils["_sample_categorical"]
    return ns


def load_reference_noise():
    """`models/noise_schedule.py::LogLinearNoise` imports as-is."""
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    return importlib.import_module("models.noise_schedule")
