"""Thin tensor-level wrappers over the C ABI (shape checks + allocation only; all maths is in the CUDA library)."""
from __future__ import annotations

import torch

from . import _lib as L
from ._lib import P, call, stream

bf16 = torch.bfloat16


def _chk(cond, msg):
    if not cond:
        raise ValueError(msg)


def gemm(a, b, *, ta=False, tb=False, M=None, N=None, K=None, out=None, epi=L.EPI_BF16, bias=None, aux=None, bn=0):
    """C[M,N] = sum_k A(m,k) B(n,k).  a: [M,K] (ta=False) or [K,M] (ta=True); b: [N,K] (tb=False) or [K,N] (tb=True).
    Row strides may exceed the logical width (views of padded buffers)."""
    _chk(a.dtype == bf16 and b.dtype == bf16, "gemm operands must be bf16")
    _chk(a.stride(-1) == 1 and b.stride(-1) == 1, "gemm operands must be row-major")
    if M is None:
        M = a.shape[1] if ta else a.shape[0]
    if K is None:
        K = a.shape[0] if ta else a.shape[1]
    if N is None:
        N = b.shape[1] if tb else b.shape[0]
    fp32_out = epi in (L.EPI_F32, L.EPI_F32_ACC)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if fp32_out else bf16)
    _chk(out.stride(-1) == 1, "gemm output must be row-major")
    _chk(out.dtype == (torch.float32 if fp32_out else bf16), "gemm output dtype mismatch")
    if epi == L.EPI_BF16_SCALED:
        _chk(aux is not None and aux.dtype == torch.float32 and ta and tb, "EPI_BF16_SCALED: wgrad layout + device fp32 scale in aux")
    if epi == L.EPI_BF16_GELU and aux is None:
        aux = torch.empty((M, N), device=a.device, dtype=bf16)
    call("ud_gemm_bf16", int(ta), int(tb), M, N, K, P(a), a.stride(0), P(b), b.stride(0), P(out), out.stride(0), epi,
         P(bias), P(aux), aux.stride(0) if aux is not None else 0, bn, stream())
    if epi == L.EPI_BF16_GELU:
        return out, aux
    return out


def _tc(tc):
    return tc[0] if tc is not None else None


def embed_rmsnorm_fwd(ids, modality, E, Emod, w, eps=1e-6, ordinal=None, Ecount=None, tc=None):
    rows, D = ids.numel(), E.shape[1]
    x = torch.empty((rows, D), device=E.device, dtype=torch.float32)
    h = torch.empty((rows, D), device=E.device, dtype=bf16)
    rstd = torch.empty((rows,), device=E.device, dtype=torch.float32)
    call("ud_embed_rmsnorm_fwd", P(ids), P(modality), P(E), P(Emod), P(w), P(x), P(h), P(rstd), rows, D, eps, P(ordinal),
         P(Ecount), _tc(tc), stream())
    return x, h, rstd


def embed_bwd(ids, modality, g, dE, dEmod, hot_id=-1, ordinal=None, dEcount=None):
    rows, D = g.shape
    call("ud_embed_bwd", P(ids), P(modality), P(g), P(dE), P(dEmod), rows, D, hot_id, P(ordinal), P(dEcount), stream())


def interleaved_prep(modality, sample_ids, cos_tab, sin_tab, offsets):
    """modality, sample_ids: int64 [B,N]; cos_tab/sin_tab fp32 [rows,hd2]; offsets = dict(txt=, s256=, s1024=, s2304=, s4096=)
    (row offsets of the tables inside cos_tab, -1 = table absent).  Returns cos, sin [B*N,hd2] fp32 and ordinal int32 [B*N]."""
    B, N = modality.shape
    hd2 = cos_tab.shape[1]
    dev = modality.device
    cos = torch.empty((B * N, hd2), device=dev, dtype=torch.float32)
    sin = torch.empty((B * N, hd2), device=dev, dtype=torch.float32)
    ordinal = torch.empty((B * N,), device=dev, dtype=torch.int32)
    scratch = torch.empty((B, 4 * N), device=dev, dtype=torch.int32)
    call("ud_interleaved_prep", P(modality.contiguous()), P(sample_ids.contiguous()), B, N, P(cos_tab), P(sin_tab), hd2,
         offsets["txt"], offsets["s256"], offsets["s1024"], offsets["s2304"], offsets["s4096"], P(cos), P(sin), P(ordinal),
         P(scratch), stream())
    return cos, sin, ordinal


def norm_residual_fwd(a, x_in, w_a, w_n, eps=1e-6, x_out=None, h=None, p_drop=0.0, seed=0, offset=0, tc=None):
    rows, D = x_in.shape
    if x_out is None:
        x_out = torch.empty_like(x_in)
    if h is None:
        h = torch.empty((rows, D), device=x_in.device, dtype=bf16)
    ra = torch.empty((rows,), device=x_in.device, dtype=torch.float32)
    rx = torch.empty((rows,), device=x_in.device, dtype=torch.float32)
    call("ud_norm_residual_fwd", P(a), P(x_in), P(w_a), P(w_n), P(x_out), P(h), P(ra), P(rx), rows, D, eps, float(p_drop),
         seed, offset, _tc(tc), stream())
    return x_out, h, ra, rx


def dropout_scales(rows, D, p_drop, seed, offset, device):
    """The keep-scales (0 or 1/(1-p)) the fused norm kernels apply for (p_drop, seed, offset) — used by parity tests."""
    out = torch.empty((rows, D), device=device, dtype=torch.float32)
    call("ud_dropout_scales", P(out), rows, D, float(p_drop), seed, offset, stream())
    return out


def norm_residual_bwd(g_out, dh, x_out, rstd_x, w_n, a, rstd_a, w_a, dw_n, dw_a, g_in=None, da=None, db_a=None, p_drop=0.0,
                      seed=0, offset=0, tc=None):
    rows, D = x_out.shape
    if g_in is None:
        g_in = torch.empty_like(x_out)
    if da is None:
        da = torch.empty((rows, D), device=x_out.device, dtype=bf16)
    call("ud_norm_residual_bwd", P(g_out), P(dh), P(x_out), P(rstd_x), P(w_n), P(a), P(rstd_a), P(w_a), P(g_in), P(da),
         P(dw_n), P(dw_a), P(db_a), rows, D, float(p_drop), seed, offset, _tc(tc), stream())
    return g_in, da


def rmsnorm_bwd(g_out, dh, x, rstd, w, dw, g_in=None, tc=None):
    rows, D = x.shape
    if g_in is None:
        g_in = torch.empty_like(x)
    call("ud_rmsnorm_bwd", P(g_out), P(dh), P(x), P(rstd), P(w), P(g_in), P(dw), rows, D, _tc(tc), stream())
    return g_in


def qk_ln_rope_fwd(qkv, gq, bq, gk, bk, cos, sin, head_dim, eps=1e-5):
    rows, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty((rows, 2 * D), device=qkv.device, dtype=bf16)
    stats = torch.empty((rows, 4), device=qkv.device, dtype=torch.float32)
    call("ud_qk_ln_rope_fwd", P(qkv), P(gq), P(bq), P(gk), P(bk), P(cos), P(sin), P(out), P(stats), rows, D, head_dim, eps,
         stream())
    return out, stats


def qk_ln_rope_bwd(dqk, qkv, stats, gq, gk, cos, sin, dqkv, dgq, dbq, dgk, dbk, head_dim):
    rows, D3 = qkv.shape
    D = D3 // 3
    call("ud_qk_ln_rope_bwd", P(dqk), P(qkv), P(stats), P(gq), P(gk), P(cos), P(sin), P(dqkv), P(dgq), P(dbq), P(dgk),
         P(dbk), rows, D, head_dim, stream())
    return dqkv


def attn_fwd(q, k, v, B, N, H, head_dim, scale, sample_ids=None, o=None):
    """q,k,v: 2-D bf16 views [B*N, >=H*hd] (arbitrary row stride). Returns o [B*N, H*hd] bf16, lse [B,H,N] fp32."""
    D = H * head_dim
    _chk(q.stride(0) == k.stride(0), "q and k must share a row stride")
    if o is None:
        o = torch.empty((B * N, D), device=q.device, dtype=bf16)
    lse = torch.empty((B, H, N), device=q.device, dtype=torch.float32)
    call("ud_attn_fwd", P(q), P(k), q.stride(0), P(v), v.stride(0), P(o), o.stride(0), P(lse), P(sample_ids), B, N, H,
         head_dim, scale, stream())
    return o, lse


def attn_fwd_kv(q, k, v, B, Nq, Nk, H, head_dim, scale, o=None, q_bs=0, k_bs=0, v_bs=0, o_bs=0):
    """partial-query attention: Nq query tokens per sample against Nk key/value tokens, no mask.  q / k / v / o are 2-D bf16
    views whose row 0 is the first token of sample 0; `*_bs` = elements between consecutive samples (0 = dense rows * stride),
    which lets sub-ranges of a sequence (text queries, cached image keys) be used in place.  Returns o, lse [B,H,Nq]."""
    D = H * head_dim
    if o is None:
        o = torch.empty((B * Nq, D), device=q.device, dtype=bf16)
    lse = torch.empty((B, H, Nq), device=q.device, dtype=torch.float32)
    call("ud_attn_fwd_kv", P(q), q.stride(0), q_bs, P(k), k.stride(0), k_bs, P(v), v.stride(0), v_bs, P(o), o.stride(0), o_bs, P(lse),
         B, Nq, Nk, H, head_dim, scale, stream())
    return o, lse


def attn_bwd(q, k, v, o, do, lse, dq, dk, dv, B, N, H, head_dim, scale, sample_ids=None):
    _chk(dq.stride(0) == dk.stride(0), "dq and dk must share a row stride")
    _chk(o.stride(0) == do.stride(0), "o and do must share a row stride")
    delta = torch.empty((2, B, H, N), device=q.device, dtype=torch.float32)   # [0] delta, [1] lse * log2(e) (kernel scratch)
    call("ud_attn_bwd", P(q), P(k), q.stride(0), P(v), v.stride(0), P(o), P(do), o.stride(0), P(lse), P(delta), P(dq), P(dk),
         dq.stride(0), P(dv), dv.stride(0), P(sample_ids), B, N, H, head_dim, scale, stream())
    return dq, dk, dv


def colsum(dY, db, M=None, N=None):
    if M is None:
        M, N = dY.shape
    call("ud_colsum_bf16", P(dY), dY.stride(0), P(db), M, N, stream())
    return db


def subs_nll_fwd(logits2d, xt, x0, modality, V, text_vocab, mask_index):
    rows = xt.numel()
    logp = torch.empty((rows,), device=xt.device, dtype=torch.float32)
    lse = torch.empty((rows,), device=xt.device, dtype=torch.float32)
    call("ud_subs_nll_fwd", P(logits2d), logits2d.stride(0), P(xt), P(x0), P(modality), P(logp), P(lse), rows, V, text_vocab,
         mask_index, stream())
    return logp, lse


def subs_nll_bwd_(logits2d, xt, x0, modality, lse, dlogp, V, text_vocab, mask_index):
    rows = xt.numel()
    call("ud_subs_nll_bwd", P(logits2d), logits2d.stride(0), P(xt), P(x0), P(modality), P(lse), P(dlogp), rows, V, text_vocab,
         mask_index, stream())
    return logits2d


def subs_logprobs(logits2d, xt, modality, V, text_vocab, mask_index, out_dtype=torch.float32):
    rows = modality.numel()
    out = torch.empty((rows, V), device=logits2d.device, dtype=out_dtype)
    call("ud_subs_logprobs", P(logits2d), logits2d.stride(0), P(xt), P(modality), P(out), int(out_dtype == bf16), V, rows, V,
         text_vocab, mask_index, stream())
    return out


def q_xt(x, move_chance, mask_index, rand=None, seed=0, offset=0, return_move=False):
    B, N = x.shape
    xt = torch.empty_like(x)
    move = torch.empty((B, N), device=x.device, dtype=torch.uint8) if return_move else None
    mc = move_chance.reshape(B).contiguous().float()
    call("ud_q_xt", P(x), P(mc), P(rand), seed, offset, mask_index, P(xt), P(move), B, N, stream())
    return (xt, move.bool()) if return_move else xt


def sample_categorical(probs, u=None, seed=0, offset=0):
    V = probs.shape[-1]
    p2 = probs.reshape(-1, V)
    _chk(p2.dtype == torch.float32 and p2.stride(-1) == 1, "probs must be fp32 row-major")
    out = torch.empty((p2.shape[0],), device=probs.device, dtype=torch.int64)
    call("ud_sample_categorical", P(p2), p2.stride(0), P(u), seed, offset, P(out), p2.shape[0], V, stream())
    return out.view(probs.shape[:-1])


def ddpm_update_probs(x, p_x0, mc_t, mc_s, mask_index, u=None, seed=0, offset=0):
    B, N = x.shape
    V = p_x0.shape[-1]
    p2 = p_x0.reshape(-1, V)
    out = torch.empty_like(x)
    call("ud_ddpm_update_probs", P(x), P(p2), p2.stride(0), P(u), seed, offset, P(mc_t), P(mc_s), mask_index, P(out), B, N, V,
         stream())
    return out


def ddpm_update_logits(x, logits2d, modality, mc_t, mc_s, mask_index, text_vocab, V, logits_uncond=None, cfg_w=None, u=None,
                       seed=0, offset=0):
    B, N = x.shape
    out = torch.empty_like(x)
    call("ud_ddpm_update_logits", P(x), P(logits2d), P(logits_uncond), logits2d.stride(0), P(cfg_w), P(modality), P(u), seed,
         offset, P(mc_t), P(mc_s), mask_index, text_vocab, P(out), B, N, V, stream())
    return out


def maskgit_update(x, logits2d, modality, t, num_unmask, mask_index, text_vocab, V, r_temp=10.0, logits_uncond=None, cfg_w=None,
                   e_noise=None, gumbel=None, seed=0, offset=0):
    """One MaskGIT step (reference model_eval.py:3045-3114) from raw bf16 logits.  t: fp32 [B]; num_unmask: int32 [B];
    e_noise fp32 [B*N,V] (Exp(1), the draw inside torch.multinomial) + gumbel fp64 [B,N] (np.random.gumbel) = parity mode.
    Returns (x_next, pred_code, conf)."""
    B, N = x.shape
    _chk(num_unmask.dtype == torch.int32 and t.dtype == torch.float32, "maskgit: t fp32 [B], num_unmask int32 [B]")
    _chk((e_noise is None) == (gumbel is None), "maskgit: supply both noise tensors or neither")
    if gumbel is not None:
        _chk(gumbel.dtype == torch.float64 and e_noise.dtype == torch.float32, "maskgit: gumbel fp64, e_noise fp32")
        gumbel, e_noise = gumbel.contiguous(), e_noise.contiguous()
    out = torch.empty_like(x)
    pred = torch.empty_like(x)
    conf = torch.empty((B, N), device=x.device, dtype=torch.float64)
    call("ud_maskgit_update", P(x), P(logits2d), P(logits_uncond), logits2d.stride(0), P(cfg_w), P(modality), P(e_noise), P(gumbel),
         seed, offset, P(t), float(r_temp), P(num_unmask), mask_index, text_vocab, P(pred), P(conf), P(out), B, N, V, stream())
    return out, pred, conf


def subs_argmax(logits2d, xt, modality, V, text_vocab, mask_index):
    rows = xt.numel()
    out = torch.empty((rows,), device=xt.device, dtype=torch.int64)
    call("ud_subs_argmax", P(logits2d), logits2d.stride(0), P(xt), P(modality), P(out), rows, V, text_vocab, mask_index, stream())
    return out


def adamw_step(p, g, m, v, p_bf16, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None, max_ctas=0):
    call("ud_adamw_step", P(p), P(g), P(m), P(v), P(p_bf16), p.numel(), lr, beta1, beta2, eps, weight_decay, step,
         P(grad_scale), max_ctas, stream())


def cast_bf16(src, dst):
    call("ud_cast_f32_to_bf16", P(src), P(dst), src.numel(), stream())
    return dst


def sumsq(g, out, max_ctas=0):
    call("ud_sumsq_f32", P(g), g.numel(), P(out), max_ctas, stream())
    return out


def grad_pack(g, dst, inv_world, max_ctas=0):
    call("ud_grad_pack_bf16", P(g), P(dst), g.numel(), inv_world, max_ctas, stream())


def grad_unpack(src, g, max_ctas=0, sumsq=None):
    if sumsq is not None:
        call("ud_grad_unpack_bf16_sumsq", P(src), P(g), g.numel(), max_ctas, P(sumsq), stream())
        return
    call("ud_grad_unpack_bf16", P(src), P(g), g.numel(), max_ctas, stream())
