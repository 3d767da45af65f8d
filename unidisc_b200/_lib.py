"""ctypes binding of libunidisc_b200.so (the C ABI declared in include/unidisc_b200.h).

The product path has NO fallback: if the CUDA library is missing or a launch fails this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libunidisc_b200.so")

_lib = None

_vp, _i, _ll, _f, _u64, _i64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint64, C.c_int64

_SIGS = {
    "ud_abi_version": [],
    "ud_device_sm_count": [],
    "ud_gemm_bf16": [_i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _i, _vp, _vp, _ll, _i, _vp],
    "ud_embed_rmsnorm_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp],
    "ud_embed_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _vp, _vp, _vp],
    "ud_interleaved_prep": [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "ud_norm_residual_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _u64, _u64, _vp, _vp],
    "ud_norm_residual_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _u64, _u64, _vp, _vp],
    "ud_dropout_scales": [_vp, _i, _i, _f, _u64, _u64, _vp],
    "ud_rmsnorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp],
    "ud_qk_ln_rope_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp],
    "ud_qk_ln_rope_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "ud_attn_fwd": [_vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "ud_attn_fwd_kv": [_vp, _ll, _ll, _vp, _ll, _ll, _vp, _ll, _ll, _vp, _ll, _ll, _vp, _i, _i, _i, _i, _i, _f, _vp],
    "ud_attn_bwd": [_vp, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _vp, _ll, _vp, _i, _i, _i, _i, _f, _vp],
    "ud_colsum_bf16": [_vp, _ll, _vp, _i, _i, _vp],
    "ud_subs_nll_fwd": [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "ud_subs_nll_bwd": [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "ud_subs_logprobs": [_vp, _ll, _vp, _vp, _vp, _i, _ll, _i, _i, _i, _i, _vp],
    "ud_q_xt": [_vp, _vp, _vp, _u64, _u64, _i64, _vp, _vp, _i, _i, _vp],
    "ud_sample_categorical": [_vp, _ll, _vp, _u64, _u64, _vp, _i, _i, _vp],
    "ud_ddpm_update_probs": [_vp, _vp, _ll, _vp, _u64, _u64, _vp, _vp, _i64, _vp, _i, _i, _i, _vp],
    "ud_ddpm_update_logits": [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _i64, _i, _vp, _i, _i, _i, _vp],
    "ud_maskgit_update": [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _u64, _u64, _vp, _f, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _i, _vp],
    "ud_subs_argmax": [_vp, _ll, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "ud_adamw_step": [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _vp, _i, _vp],
    "ud_cast_f32_to_bf16": [_vp, _vp, _ll, _vp],
    "ud_sumsq_f32": [_vp, _ll, _vp, _i, _vp],
    "ud_grad_pack_bf16": [_vp, _vp, _ll, _f, _i, _vp],
    "ud_grad_unpack_bf16": [_vp, _vp, _ll, _i, _vp],
    "ud_grad_unpack_bf16_sumsq": [_vp, _vp, _ll, _i, _vp, _vp],
}

EXPORTED_SYMBOLS = tuple(_SIGS.keys())


class AdaLN(C.Structure):
    """`ud_adaln` of include/unidisc_b200.h (time conditioning; NULL in the default configuration)."""
    _fields_ = [("sel", _vp), ("img", _vp), ("shift", _vp), ("scale", _vp), ("gate", _vp), ("ld", _ll), ("tokens_per_sample", _i),
                ("d_shift", _vp), ("d_scale", _vp), ("d_gate", _vp), ("ld_d", _ll)]


def adaln(sel, img, shift=None, scale=None, gate=None, ld=0, tokens_per_sample=1, d_shift=None, d_scale=None, d_gate=None, ld_d=0):
    """Builds the struct from tensors (views into the adaLN output / its fp32 gradient buffer); returns (byref pointer, keepalive)."""
    s = AdaLN(P(sel), P(img), P(shift), P(scale), P(gate), ld, tokens_per_sample, P(d_shift), P(d_scale), P(d_gate), ld_d)
    return C.cast(C.pointer(s), _vp), s

EPI_BF16, EPI_BF16_GELU, EPI_BF16_DGELU, EPI_F32, EPI_F32_ACC, EPI_BF16_SCALED = 0, 1, 2, 3, 4, 5


class UnidiscB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UnidiscB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU / eager fallback.")
        _lib = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(_lib, name)
            fn.argtypes = args
            fn.restype = _i
    return _lib


def P(t):
    """device pointer of a tensor (None -> NULL)"""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise UnidiscB200Error(f"{name} failed with code {rc}")
