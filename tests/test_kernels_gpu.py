"""GPU parity tests of the individual CUDA kernels (called through the C ABI via ctypes) against the oracle
(oracle/restated.py) / plain fp32 torch restatements on the same seeded inputs.

Tolerances: bf16 outputs are compared after both sides round to bf16 -> allow 1 bf16 ulp (rtol 2^-7) on a tiny
fraction of elements plus rtol 1e-3/atol 1e-4 (north_star) on fp32 quantities; integer outputs must be bit-exact.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from unidisc_b200 import ops as _ops
    return _ops


def dev():
    return torch.device("cuda", 0)


def rnd(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(dev())


def close_bf16(out, ref, name, frac_bad=1e-3):
    """bf16 tensors: equal up to one bf16 rounding step (2^-8 relative) except a tiny fraction (rounding ties)."""
    out, ref = out.float(), ref.float()
    err = (out - ref).abs()
    tol = 1e-4 + ref.abs() * (2.0 ** -7)
    bad = (err > tol).float().mean().item()
    assert bad <= frac_bad, f"{name}: {bad*100:.4f}% elements off by more than 1 bf16 ulp; max err {err.max().item():.4e}"
    # and nothing is wildly off
    assert (err <= 1e-3 + ref.abs() * 0.02).all(), f"{name}: max err {err.max().item():.4e}"


# --------------------------------------------------------------------------------------------------
# GEMM family
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (200, 328, 136), (1024, 768, 2048), (130, 48385 // 8 * 8 + 8, 128)])
@pytest.mark.parametrize("bn", [0, 128, 256, 1024 + 128, 1024 + 256])     # +1024 = force the single-CTA kernel
def test_gemm_nt(ops, M, N, K, bn):
    a, b = rnd(M, K, seed=1, dtype=bf16), rnd(N, K, seed=2, dtype=bf16)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b, bn=bn)
    torch.cuda.synchronize()
    close_bf16(out, ref.to(bf16), f"gemm_nt {M}x{N}x{K}")


def test_gemm_odd_n_with_padded_ld(ops):
    # head GEMM shape class: N = 48385-like (odd), output rows padded to a multiple of 64
    M, N, K, ld = 256, 1001, 192, 1024
    a, b = rnd(M, K, seed=1, dtype=bf16), rnd(N, K, seed=2, dtype=bf16)
    bias = rnd(N, seed=3, dtype=bf16)
    buf = torch.full((M, ld), 7.0, device=dev(), dtype=bf16)
    out = ops.gemm(a, b, N=N, out=buf[:, :N], bias=bias)
    torch.cuda.synchronize()
    ref = (a.float() @ b.float().t() + bias.float()).to(bf16)
    close_bf16(out, ref, "gemm odd N")
    assert (buf[:, N:] == 7.0).all(), "padding columns must not be touched"


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (384, 200, 264), (1024, 2048, 768), (2560, 512, 4096)])
@pytest.mark.parametrize("bn", [0, 128, 1024])
def test_gemm_dgrad_layout(ops, M, N, K, bn):
    # dx[M,N] = dy[M,K] @ W[K,N]   (tb=1: B given as [K,N] row-major)
    dy, w = rnd(M, K, seed=3, dtype=bf16), rnd(K, N, seed=4, dtype=bf16)
    out = ops.gemm(dy, w, tb=True, bn=bn)
    torch.cuda.synchronize()
    close_bf16(out, (dy.float() @ w.float()).to(bf16), "gemm dgrad")


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (200, 328, 520), (768, 3072, 1024), (1000, 512, 2500)])
def test_gemm_wgrad_layout(ops, M, N, K):
    # dW[M,N] = dy[K,M]^T @ x[K,N]   (ta=1,tb=1), fp32 output, then accumulate
    dy, x = rnd(K, M, seed=5, dtype=bf16), rnd(K, N, seed=6, dtype=bf16)
    from unidisc_b200._lib import EPI_F32, EPI_F32_ACC
    out = ops.gemm(dy, x, ta=True, tb=True, epi=EPI_F32)
    torch.cuda.synchronize()
    ref = dy.float().t() @ x.float()
    assert torch.allclose(out, ref, rtol=1e-3, atol=1e-3 * math.sqrt(K)), (out - ref).abs().max()
    ops.gemm(dy, x, ta=True, tb=True, epi=EPI_F32_ACC, out=out)
    torch.cuda.synchronize()
    assert torch.allclose(out, 2 * ref, rtol=1e-3, atol=2e-3 * math.sqrt(K))


@pytest.mark.parametrize("M,N,K", [(200, 328, 520), (1000, 520, 2500), (2048, 2048, 10240), (6144, 2048, 10240)])
def test_gemm_wgrad_fused_sum_of_squares(ops, M, N, K):
    """UD_EPI_F32 with aux = fp32 device scalar: the epilogue adds sum(C^2) of what it stores (gradient-norm fusion), on ragged
    edges (zero-filled rows / columns contribute nothing) and on stream-K split tiles; the accumulate epilogue rejects it."""
    from unidisc_b200._lib import EPI_F32, EPI_F32_ACC, UnidiscB200Error
    dy, x = rnd(K, M, seed=5, scale=0.1, dtype=bf16), rnd(K, N, seed=6, scale=0.1, dtype=bf16)
    acc = torch.full((1,), 3.0, device=dev())
    out = torch.empty((M, (N + 3) // 4 * 4), device=dev(), dtype=torch.float32)[:, :N]
    ops.gemm(dy, x, ta=True, tb=True, epi=EPI_F32, out=out, aux=acc)
    ops.gemm(dy, x, ta=True, tb=True, epi=EPI_F32, out=out, aux=acc)
    torch.cuda.synchronize()
    want = 3.0 + 2 * out.double().pow(2).sum().item()
    assert abs(acc.item() - want) <= 1e-4 * want, (acc.item(), want)
    with pytest.raises(UnidiscB200Error):
        ops.gemm(dy, x, ta=True, tb=True, epi=EPI_F32_ACC, out=out, aux=acc)


@pytest.mark.parametrize("case", ["fwd_n2048", "dgrad_n2048", "wgrad_64tiles", "wgrad_192tiles", "gelu_tail", "ragged"])
def test_gemm_streamk_tail(ops, case):
    """Shapes whose last wave of 256x256 tiles is partial: the tail tiles' k-blocks are shared between all clusters
    (stream-K) through fp32 partials in a workspace.  Launched several times back to back (the arrival counters re-arm
    themselves) and compared with fp32 torch matmuls."""
    from unidisc_b200._lib import EPI_BF16_GELU, EPI_F32, EPI_F32_ACC
    if case == "fwd_n2048":            # attn_out / mlp.2 forward of the 1.4B model: 320 tiles on 74 clusters
        M, N, K = 10240, 2048, 2048
        a, b = rnd(M, K, seed=1, dtype=bf16), rnd(N, K, seed=2, scale=0.05, dtype=bf16)
        ref = (a.float() @ b.float().t()).to(bf16)
        for _ in range(3):
            out = ops.gemm(a, b)
        torch.cuda.synchronize()
        close_bf16(out, ref, case)
    elif case == "dgrad_n2048":
        M, N, K = 10240, 2048, 6144
        dy, w = rnd(M, K, seed=3, dtype=bf16), rnd(K, N, seed=4, scale=0.05, dtype=bf16)
        ref = (dy.float() @ w.float()).to(bf16)
        for _ in range(2):
            out = ops.gemm(dy, w, tb=True)
        torch.cuda.synchronize()
        close_bf16(out, ref, case)
    elif case in ("wgrad_64tiles", "wgrad_192tiles"):
        M, N, K = (2048, 2048, 10240) if case == "wgrad_64tiles" else (6144, 2048, 10240)
        dy, x = rnd(K, M, seed=5, scale=0.1, dtype=bf16), rnd(K, N, seed=6, scale=0.1, dtype=bf16)
        ref = dy.float().t() @ x.float()
        out = ops.gemm(dy, x, ta=True, tb=True, epi=EPI_F32)
        torch.cuda.synchronize()
        assert torch.allclose(out, ref, rtol=1e-3, atol=2e-3), (out - ref).abs().max()
        ops.gemm(dy, x, ta=True, tb=True, epi=EPI_F32_ACC, out=out)
        ops.gemm(dy, x, ta=True, tb=True, epi=EPI_F32_ACC, out=out)
        torch.cuda.synchronize()
        assert torch.allclose(out, 3 * ref, rtol=1e-3, atol=6e-3), (out - 3 * ref).abs().max()
    elif case == "gelu_tail":          # fused epilogue on the owner of a split tile
        M, N, K = 2560, 2304, 1024
        a, b = rnd(M, K, seed=1, dtype=bf16), rnd(N, K, seed=2, scale=0.05, dtype=bf16)
        bias = rnd(N, seed=3, dtype=bf16)
        u, g = ops.gemm(a, b, epi=EPI_BF16_GELU, bias=bias)
        torch.cuda.synchronize()
        close_bf16(u, (a.float() @ b.float().t() + bias.float()).to(bf16), "u")
        close_bf16(g, torch.nn.functional.gelu(u.float(), approximate="tanh").to(bf16), "g")
    else:                              # ragged M / N / K edges together with split tiles
        M, N, K = 1000, 1336, 1992
        a, b = rnd(M, K, seed=1, dtype=bf16), rnd(N, K, seed=2, scale=0.05, dtype=bf16)
        for _ in range(2):
            out = ops.gemm(a, b)
        torch.cuda.synchronize()
        close_bf16(out, (a.float() @ b.float().t()).to(bf16), case)


def test_gemm_bias_gelu_and_dgelu(ops):
    from unidisc_b200._lib import EPI_BF16_DGELU, EPI_BF16_GELU
    M, N, K = 256, 512, 128
    a, b = rnd(M, K, seed=1, dtype=bf16), rnd(N, K, seed=2, scale=0.2, dtype=bf16)
    bias = rnd(N, seed=3, dtype=bf16)
    u, g = ops.gemm(a, b, epi=EPI_BF16_GELU, bias=bias)
    torch.cuda.synchronize()
    u_ref = (a.float() @ b.float().t() + bias.float()).to(bf16)
    close_bf16(u, u_ref, "u")
    g_ref = torch.nn.functional.gelu(u.float(), approximate="tanh").to(bf16)
    close_bf16(g, g_ref, "gelu(u)")
    # dgelu: out = (dy @ W) * gelu'(u)
    dy, w = rnd(M, K, seed=7, dtype=bf16), rnd(K, N, seed=8, scale=0.2, dtype=bf16)
    out = ops.gemm(dy, w, tb=True, epi=EPI_BF16_DGELU, aux=u)
    torch.cuda.synchronize()
    uu = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uu, approximate="tanh").backward(dy.float() @ w.float())
    close_bf16(out, uu.grad.to(bf16), "dgelu", frac_bad=5e-3)


# --------------------------------------------------------------------------------------------------
# row kernels
# --------------------------------------------------------------------------------------------------
def _rms(x, eps=1e-6):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)


@pytest.mark.parametrize("D", [128, 384, 768, 2048])
def test_embed_rmsnorm(ops, D):
    V, rows = 500, 300
    E, Emod, w = rnd(V, D, seed=1), rnd(2, D, seed=2), 1 + 0.1 * rnd(D, seed=3)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, V, (rows,), generator=g).to(dev())
    mod = torch.randint(0, 2, (rows,), generator=g).to(dev())
    x, h, rstd = ops.embed_rmsnorm_fwd(ids, mod, E, Emod, w)
    torch.cuda.synchronize()
    x_ref = E[ids] + Emod[mod]
    assert torch.equal(x, x_ref)
    close_bf16(h, (_rms(x_ref) * w).to(bf16), "h")
    # backward scatter
    gr = rnd(rows, D, seed=4)
    dE, dEm = torch.zeros_like(E), torch.zeros_like(Emod)
    ops.embed_bwd(ids, mod, gr, dE, dEm, hot_id=int(ids[0]))
    torch.cuda.synchronize()
    dE_ref = torch.zeros_like(E).index_add_(0, ids, gr)
    dEm_ref = torch.zeros_like(Emod).index_add_(0, mod, gr)
    assert torch.allclose(dE, dE_ref, rtol=1e-4, atol=1e-4) and torch.allclose(dEm, dEm_ref, rtol=1e-4, atol=1e-3)


@pytest.fixture(params=["staged", "registers"])
def row_kernel_variant(request, monkeypatch):
    """the two heaviest row kernels exist twice: operands staged in shared memory by bulk async copies (default) and the
    register-prefetch versions (UD_NORM_BWD=1 / UD_QKLN_BWD=1; also the fallback for D > 2048).  The launcher reads the switch per call."""
    if request.param == "registers":
        monkeypatch.setenv("UD_NORM_BWD", "1")
        monkeypatch.setenv("UD_QKLN_BWD", "1")
    return request.param


@pytest.mark.parametrize("D", [128, 768, 2048])
def test_norm_residual_fwd_bwd(ops, D, row_kernel_variant):
    rows = 257
    a = rnd(rows, D, seed=1, scale=2.0, dtype=bf16)
    x_in = rnd(rows, D, seed=2)
    w_a, w_n = 1 + 0.2 * rnd(D, seed=3), 1 + 0.2 * rnd(D, seed=4)
    x_out, h, ra, rx = ops.norm_residual_fwd(a, x_in, w_a, w_n)
    torch.cuda.synchronize()

    def ref_fwd(a_, x_, wa_, wn_):
        n = _rms(a_.float())
        n = n + (n.to(bf16).float() - n).detach()      # bf16 rounding point, straight-through gradient
        xo = x_ + n * wa_
        return xo, _rms(xo) * wn_

    xo_ref, h_ref = ref_fwd(a, x_in, w_a, w_n)
    assert torch.allclose(x_out, xo_ref, rtol=1e-5, atol=1e-5)
    close_bf16(h, h_ref.to(bf16), "h")
    # backward
    g_out = rnd(rows, D, seed=5)
    dh = rnd(rows, D, seed=6, dtype=bf16)
    dw_n, dw_a = torch.zeros(D, device=dev()), torch.zeros(D, device=dev())
    db = torch.zeros(D, device=dev())
    g_in, da = ops.norm_residual_bwd(g_out, dh, x_out, rx, w_n, a, ra, w_a, dw_n, dw_a, db_a=db)
    torch.cuda.synchronize()
    a32 = a.float().requires_grad_(True)
    xi = x_in.clone().requires_grad_(True)
    wa_, wn_ = w_a.clone().requires_grad_(True), w_n.clone().requires_grad_(True)
    xo, hh = ref_fwd(a32, xi, wa_, wn_)
    (xo * g_out).sum().backward(retain_graph=True)
    (hh * dh.float()).sum().backward()
    assert torch.allclose(g_in, xi.grad, rtol=1e-3, atol=1e-4), (g_in - xi.grad).abs().max()
    close_bf16(da, a32.grad.to(bf16), "da", frac_bad=5e-3)
    assert torch.allclose(dw_n, wn_.grad, rtol=2e-3, atol=2e-3), (dw_n - wn_.grad).abs().max()
    assert torch.allclose(dw_a, wa_.grad, rtol=2e-3, atol=2e-3), (dw_a - wa_.grad).abs().max()
    assert torch.allclose(db, a32.grad.sum(0), rtol=2e-3, atol=2e-3), (db - a32.grad.sum(0)).abs().max()
    # plain rmsnorm backward
    dw = torch.zeros(D, device=dev())
    g2 = ops.rmsnorm_bwd(g_out, dh, x_out, rx, w_n, dw)
    torch.cuda.synchronize()
    x2 = x_out.clone().requires_grad_(True)
    w2 = w_n.clone().requires_grad_(True)
    ((_rms(x2) * w2) * dh.float()).sum().backward()
    assert torch.allclose(g2, g_out + x2.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(dw, w2.grad, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("D,p", [(256, 0.1), (2048, 0.1), (768, 0.5)])
def test_norm_residual_dropout_fwd_bwd(ops, D, p, row_kernel_variant):
    """training-mode dropout of the branch (reference dit.py:229-253,1024-1031): the in-kernel Philox keep-scales are
    materialised with ud_dropout_scales and fed to the torch reference."""
    rows, seed, off = 515, 1234, 77
    a = rnd(rows, D, seed=1, scale=2.0, dtype=bf16)
    x_in = rnd(rows, D, seed=2)
    w_a, w_n = 1 + 0.2 * rnd(D, seed=3), 1 + 0.2 * rnd(D, seed=4)
    ks = ops.dropout_scales(rows, D, p, seed, off, dev())
    torch.cuda.synchronize()
    vals = torch.unique(ks)
    assert vals.numel() == 2 and vals[0] == 0 and abs(float(vals[1]) - 1 / (1 - p)) < 1e-6
    assert abs(float((ks == 0).float().mean()) - p) < 0.01
    # different offsets / seeds give different masks; same triple is reproducible
    assert not torch.equal(ks, ops.dropout_scales(rows, D, p, seed, off + 1, dev()))
    assert torch.equal(ks, ops.dropout_scales(rows, D, p, seed, off, dev()))
    x_out, h, ra, rx = ops.norm_residual_fwd(a, x_in, w_a, w_n, p_drop=p, seed=seed, offset=off)

    def ref_fwd(a_, x_, wa_, wn_):
        n = _rms(a_.float())
        n = n + (n.to(bf16).float() - n).detach()
        xo = x_ + (n * wa_) * ks
        return xo, _rms(xo) * wn_

    xo_ref, h_ref = ref_fwd(a, x_in, w_a, w_n)
    assert torch.allclose(x_out, xo_ref, rtol=1e-5, atol=1e-5)
    close_bf16(h, h_ref.to(bf16), "h")
    g_out = rnd(rows, D, seed=5)
    dh = rnd(rows, D, seed=6, dtype=bf16)
    dw_n, dw_a, db = (torch.zeros(D, device=dev()) for _ in range(3))
    g_in, da = ops.norm_residual_bwd(g_out, dh, x_out, rx, w_n, a, ra, w_a, dw_n, dw_a, db_a=db, p_drop=p, seed=seed, offset=off)
    torch.cuda.synchronize()
    a32 = a.float().requires_grad_(True)
    xi = x_in.clone().requires_grad_(True)
    wa_, wn_ = w_a.clone().requires_grad_(True), w_n.clone().requires_grad_(True)
    xo, hh = ref_fwd(a32, xi, wa_, wn_)
    (xo * g_out).sum().backward(retain_graph=True)
    (hh * dh.float()).sum().backward()
    assert torch.allclose(g_in, xi.grad, rtol=1e-3, atol=1e-4), (g_in - xi.grad).abs().max()
    close_bf16(da, a32.grad.to(bf16), "da", frac_bad=5e-3)
    assert torch.allclose(dw_n, wn_.grad, rtol=2e-3, atol=2e-3), (dw_n - wn_.grad).abs().max()
    assert torch.allclose(dw_a, wa_.grad, rtol=2e-3, atol=2e-3), (dw_a - wa_.grad).abs().max()
    assert torch.allclose(db, a32.grad.sum(0), rtol=2e-3, atol=3e-3), (db - a32.grad.sum(0)).abs().max()


@pytest.mark.parametrize("D,hd", [(128, 64), (768, 64), (2048, 128), (128, 32)])
def test_qk_ln_rope_fwd_bwd(ops, D, hd, row_kernel_variant):
    from oracle import restated as R
    rows, H = 200, D // hd
    qkv = rnd(rows, 3 * D, seed=1, scale=1.5, dtype=bf16)
    gq, bq, gk, bk = 1 + 0.2 * rnd(D, seed=2), 0.1 * rnd(D, seed=3), 1 + 0.2 * rnd(D, seed=4), 0.1 * rnd(D, seed=5)
    ang = rnd(rows, hd // 2, seed=6, scale=3.0)
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    out, stats = ops.qk_ln_rope_fwd(qkv, gq, bq, gk, bk, cos, sin, hd)
    torch.cuda.synchronize()

    def ref(qkv32, gq_, bq_, gk_, bk_, ste=True):
        q, k = qkv32[:, :D], qkv32[:, D:2 * D]
        ln = lambda t, w, b: torch.nn.functional.layer_norm(t, (D,), w, b, 1e-5)
        q, k = ln(q, gq_, bq_), ln(k, gk_, bk_)
        q = q + (q.to(bf16).float() - q).detach()
        k = k + (k.to(bf16).float() - k).detach()
        q = R._rope(q.view(1, rows, H, hd), cos[None], sin[None]).reshape(rows, D)
        k = R._rope(k.view(1, rows, H, hd), cos[None], sin[None]).reshape(rows, D)
        return torch.cat([q, k], -1)

    # RoPE output = y*cos + y'*sin can cancel: the absolute error scale is one bf16 ulp of the O(4) LayerNorm outputs
    o_ref = ref(qkv.float(), gq, bq, gk, bk)
    err = (out.float() - o_ref).abs()
    assert err.max() < 3.2e-2 and (err > 1e-4 + o_ref.abs() * 2.0 ** -7).float().mean() < 2e-2, err.max()
    # backward
    dqk = rnd(rows, 2 * D, seed=7, dtype=bf16)
    dqkv = torch.zeros(rows, 3 * D, device=dev(), dtype=bf16)
    grads = [torch.zeros(D, device=dev()) for _ in range(4)]
    ops.qk_ln_rope_bwd(dqk, qkv, stats, gq, gk, cos, sin, dqkv, *grads, hd)
    torch.cuda.synchronize()
    x = qkv.float().requires_grad_(True)
    ps = [t.clone().requires_grad_(True) for t in (gq, bq, gk, bk)]
    (ref(x, *ps) * dqk.float()).sum().backward()
    close_bf16(dqkv[:, :2 * D], x.grad[:, :2 * D].to(bf16), "dqkv", frac_bad=5e-3)
    for got, p, nm in zip(grads, ps, ("dgq", "dbq", "dgk", "dbk")):
        assert torch.allclose(got, p.grad, rtol=2e-3, atol=3e-3), (nm, (got - p.grad).abs().max())


@pytest.mark.parametrize("N,ld", [(777, 784), (777, 779), (2048, 2048), (8, 8)])
def test_colsum(ops, N, ld):
    # ld % 8 == 0 takes the 16-byte kernel (the last vector's padding columns must not be summed), otherwise the 4-byte one
    M = 1000
    buf = rnd(M, ld, seed=1, dtype=bf16)
    db = torch.zeros(N, device=dev())
    ops.colsum(buf[:, :N], db, M, N)
    torch.cuda.synchronize()
    assert torch.allclose(db, buf[:, :N].float().sum(0), rtol=1e-4, atol=1e-3)


# --------------------------------------------------------------------------------------------------
# SUBS NLL
# --------------------------------------------------------------------------------------------------
def test_subs_nll_fwd_bwd_and_logprobs(ops):
    from oracle import restated as R
    B, N, V, tv, mi, ldv = 3, 40, 1001, 601, 600, 1024
    g = torch.Generator().manual_seed(0)
    logits = torch.zeros(B * N, ldv, dtype=bf16)
    logits[:, :V] = (torch.randn(B * N, V, generator=g) * 3).to(bf16)
    modality = torch.cat([torch.zeros(B, 16, dtype=torch.int64), torch.ones(B, N - 16, dtype=torch.int64)], 1)
    x0 = torch.where(modality == 0, torch.randint(0, tv - 1, (B, N), generator=g), torch.randint(tv, V, (B, N), generator=g))
    xt = torch.where(torch.rand(B, N, generator=g) < 0.6, torch.full_like(x0, mi), x0)
    lg, md, x0d, xtd = logits.to(dev()), modality.to(dev()), x0.to(dev()), xt.to(dev())
    logp, lse = ops.subs_nll_fwd(lg, xtd.view(-1), x0d.view(-1), md.view(-1), V, tv, mi)
    torch.cuda.synchronize()
    l32 = logits[:, :V].float().view(B, N, V).requires_grad_(True)
    ref = R.subs_parameterization(l32, xt, modality, mi, tv)
    ref_lp = torch.gather(ref, -1, x0[..., None]).squeeze(-1)
    assert torch.allclose(logp.cpu().view(B, N), ref_lp.detach(), rtol=1e-4, atol=1e-4), (logp.cpu().view(B, N) - ref_lp).abs().max()
    full = ops.subs_logprobs(lg, xtd.view(-1), md.view(-1), V, tv, mi)
    torch.cuda.synchronize()
    assert torch.allclose(full.cpu().view(B, N, V), ref.detach(), rtol=1e-4, atol=2e-3)
    full_noxt = ops.subs_logprobs(lg, None, md.view(-1), V, tv, mi)
    ref2 = R.subs_parameterization(logits[:, :V].float().view(B, N, V), None, modality, mi, tv)
    assert torch.allclose(full_noxt.cpu().view(B, N, V), ref2, rtol=1e-4, atol=2e-3)
    # backward
    dlogp = torch.randn(B * N, generator=g)
    ref_lp.backward(dlogp.view(B, N))
    d = ops.subs_nll_bwd_(lg.clone(), xtd.view(-1), x0d.view(-1), md.view(-1), lse, dlogp.to(dev()), V, tv, mi)
    torch.cuda.synchronize()
    close_bf16(d[:, :V].cpu(), l32.grad.view(B * N, V).to(bf16), "dlogits", frac_bad=5e-3)
    assert (d[:, V:] == 0).all()


# --------------------------------------------------------------------------------------------------
# q_xt / samplers: integer outputs, bit-exact given the same noise
# --------------------------------------------------------------------------------------------------
def test_q_xt_bit_exact_golden(ops, golden_fns, golden_dit):
    g = golden_fns
    mi = int(golden_dit["cfg"][7])
    x0 = torch.from_numpy(g["qxt_x0"]).to(dev())
    xt, mv = ops.q_xt(x0, torch.from_numpy(g["qxt_mc"]).to(dev()), mi, rand=torch.from_numpy(g["qxt_rand"]).to(dev()), return_move=True)
    torch.cuda.synchronize()
    assert np.array_equal(xt.cpu().numpy(), g["qxt_ref"]) and np.array_equal(mv.cpu().numpy(), g["qxt_move_ref"])


def test_q_xt_philox_statistics(ops):
    B, N = 8, 4096
    x = torch.zeros(B, N, dtype=torch.int64, device=dev())
    mc = torch.linspace(0.1, 0.9, B, device=dev())
    xt = ops.q_xt(x, mc, 7, seed=1234, offset=5)
    xt2 = ops.q_xt(x, mc, 7, seed=1234, offset=5)
    xt3 = ops.q_xt(x, mc, 7, seed=1235, offset=5)
    torch.cuda.synchronize()
    assert torch.equal(xt, xt2) and not torch.equal(xt, xt3)
    frac = (xt == 7).float().mean(1)
    assert torch.allclose(frac, mc, atol=0.03)


def test_sample_categorical_bit_exact(ops, golden_fns, golden_dit):
    from oracle import restated as R
    g = golden_fns
    mi = int(golden_dit["cfg"][7])
    probs, u = torch.from_numpy(g["sc_probs"]).to(dev()), torch.from_numpy(g["sc_u"]).to(dev())
    out = ops.sample_categorical(probs, u.reshape(-1, u.shape[-1]))
    torch.cuda.synchronize()
    assert torch.equal(out, R.sample_categorical(probs, u)), "vs oracle evaluated on the GPU"
    assert np.array_equal(out.cpu().numpy(), g["sc_ref"]), "vs reference (CPU) golden"
    # absorbing updates
    x, t, dt = torch.from_numpy(g["ddpm_x"]).to(dev()), torch.from_numpy(g["ddpm_t"]).to(dev()), float(g["ddpm_dt"])
    tt = t.squeeze(-1)
    u2 = torch.from_numpy(g["ddpm_u"]).to(dev())
    out2 = ops.ddpm_update_probs(x, probs, tt.contiguous(), (tt - dt).contiguous(), mi, u=u2.reshape(-1, u2.shape[-1]))
    torch.cuda.synchronize()
    assert np.array_equal(out2.cpu().numpy(), g["ddpm_cache_ref"])
    sig_t, _ = R.loglinear_noise(tt)
    sig_s, _ = R.loglinear_noise(tt - dt)
    u3 = torch.from_numpy(g["ddpm_u3"]).to(dev())
    out3 = ops.ddpm_update_probs(x, probs, (1 - torch.exp(-sig_t)).contiguous(), (1 - torch.exp(-sig_s)).contiguous(), mi,
                                 u=u3.reshape(-1, u3.shape[-1]))
    torch.cuda.synchronize()
    assert np.array_equal(out3.cpu().numpy(), g["ddpm_ref"])


def test_sample_categorical_large_vocab(ops):
    from oracle import restated as R
    R_, V = 64, 48385
    g = torch.Generator().manual_seed(3)
    probs = torch.softmax(torch.randn(R_, V, generator=g) * 4, -1).to(dev())
    u = torch.rand(R_, V, generator=g).to(dev())
    out = ops.sample_categorical(probs, u)
    torch.cuda.synchronize()
    assert torch.equal(out, R.sample_categorical(probs, u))
    # philox mode is a valid sampler: empirical distribution over a 4-way categorical
    p4 = torch.tensor([[0.1, 0.2, 0.3, 0.4]], device=dev()).repeat(20000, 1).contiguous()
    s = ops.sample_categorical(p4, None, seed=11, offset=0)
    torch.cuda.synchronize()
    freq = torch.bincount(s, minlength=4).float() / s.numel()
    assert torch.allclose(freq, p4[0], atol=0.02), freq


def test_ddpm_update_from_logits(ops):
    from oracle import restated as R
    B, N, V, tv, mi, ldv = 2, 24, 1001, 601, 600, 1024
    g = torch.Generator().manual_seed(5)
    lc = torch.zeros(B * N, ldv, dtype=bf16)
    lu = torch.zeros(B * N, ldv, dtype=bf16)
    lc[:, :V] = (torch.randn(B * N, V, generator=g) * 3).to(bf16)
    lu[:, :V] = (torch.randn(B * N, V, generator=g) * 3).to(bf16)
    modality = torch.cat([torch.zeros(B, 8, dtype=torch.int64), torch.ones(B, N - 8, dtype=torch.int64)], 1)
    x = torch.where(modality == 0, torch.randint(0, tv - 1, (B, N), generator=g), torch.randint(tv, V, (B, N), generator=g))
    x[:, ::2] = mi
    u = torch.rand(B, N, V, generator=g)
    t = torch.tensor([0.8, 0.35])
    dt = 0.05
    w = 1.5 * (1 - t)
    for cfg in (False, True):
        lg = lc[:, :V].float().view(B, N, V)
        if cfg:
            lg = (1 + w)[:, None, None] * lg - w[:, None, None] * lu[:, :V].float().view(B, N, V)
        # oracle chain evaluated on the GPU in fp32: SUBS(xt=x) -> exp -> absorbing update
        p = R.subs_parameterization(lg.to(dev()), x.to(dev()), modality.to(dev()), mi, tv).exp()
        ref = R.ddpm_caching_update(x.to(dev()), t.to(dev()), dt, p, u.to(dev()), mi)
        out = ops.ddpm_update_logits(x.to(dev()), lc.to(dev()), modality.to(dev()).view(-1), t.to(dev()), (t - dt).to(dev()), mi, tv, V,
                                     logits_uncond=lu.to(dev()) if cfg else None, cfg_w=w.to(dev()) if cfg else None,
                                     u=u.to(dev()).view(-1, V))
        torch.cuda.synchronize()
        # bit-exact: the kernel follows the oracle chain's fp32 operation order (exp(l - lse) * (mc_t - mc_s), then the division
        # by 1e-10 - log(u + 1e-10)); only the row's log-sum-exp is reduced in a different order (<= a few fp32 ulp), which
        # cannot move an arg-max unless two candidates tie to ~1e-6 relative
        assert torch.equal(out, ref), f"cfg={cfg}: {(out != ref).float().mean().item()*100:.3f}% tokens differ from the oracle chain"
        assert torch.equal(out[x.to(dev()) != mi], x.to(dev())[x.to(dev()) != mi])


def _bf16_padded(logits3d, ldv):
    B, N, V = logits3d.shape
    buf = torch.zeros(B * N, ldv, dtype=bf16)
    buf[:, :V] = logits3d.reshape(B * N, V).to(bf16)
    return buf.to(dev())


def test_ddpm_update_from_logits_reference_golden(ops, golden_sampler_logits):
    """fused absorbing update from bf16 logits with SUPPLIED uniforms == the reference's `_subs_parameterization(...).exp()` ->
    `_ddpm_caching_update` / `_ddpm_update` outputs (tests/golden/sampler_logits.npz: reference run on its own draws)."""
    from oracle import restated as R
    g = golden_sampler_logits
    V, tv, mi = [int(v) for v in g["cfg"]]
    lg = torch.from_numpy(g["logits"])
    B, N, _ = lg.shape
    l2 = _bf16_padded(lg, 192)
    assert torch.equal(l2[:, :V].float().cpu().view(B, N, V), lg), "fixture logits are bf16-representable"
    mod, x = torch.from_numpy(g["modality"]).to(dev()), torch.from_numpy(g["x"]).to(dev())
    t, dt = torch.from_numpy(g["t"]).to(dev()).squeeze(-1), float(g["dt"])
    out = ops.ddpm_update_logits(x, l2, mod.view(-1), t.contiguous(), (t - dt).contiguous(), mi, tv, V,
                                 u=torch.from_numpy(g["cache_u"]).to(dev()).view(-1, V))
    sig_t, _ = R.loglinear_noise(t)
    sig_s, _ = R.loglinear_noise(t - dt)
    out2 = ops.ddpm_update_logits(x, l2, mod.view(-1), (1 - torch.exp(-sig_t)).contiguous(), (1 - torch.exp(-sig_s)).contiguous(), mi, tv, V,
                                  u=torch.from_numpy(g["ddpm_u"]).to(dev()).view(-1, V))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), g["cache_ref"]), "ddpm_cache vs reference golden"
    assert np.array_equal(out2.cpu().numpy(), g["ddpm_ref"]), "ddpm vs reference golden"


@pytest.mark.parametrize("step", [0, 4, 7])
def test_maskgit_update_reference_golden(ops, golden_sampler_logits, step):
    """MaskGIT step kernels (vocabulary pass + per-sample k-th-largest selection) with the reference's own noise draws (the
    Exp(1) tensor torch.multinomial draws, np.random.gumbel) == the reference `_maskgit_update` output, bit for bit."""
    from oracle import restated as R
    g = golden_sampler_logits
    V, tv, mi = [int(v) for v in g["cfg"]]
    lg = torch.from_numpy(g["logits"])
    B, N, _ = lg.shape
    l2 = _bf16_padded(lg, 192)
    mod, x = torch.from_numpy(g["modality"]).to(dev()), torch.from_numpy(g["x"]).to(dev())
    t = torch.from_numpy(g["t"]).to(dev())
    sched = torch.from_numpy(g["schedule"]).to(dev())
    e, gum = torch.from_numpy(g[f"mg_e_{step}"]).to(dev()), torch.from_numpy(g[f"mg_gumbel_{step}"]).to(dev())
    out, pred, conf = ops.maskgit_update(x, l2, mod.view(-1), t.view(-1).contiguous(), sched[:, step].to(torch.int32).contiguous(), mi, tv, V,
                                         r_temp=10.0, e_noise=e, gumbel=gum)
    torch.cuda.synchronize()
    masked = (x == mi).cpu().numpy()
    assert np.array_equal(pred.cpu().numpy()[masked], g[f"mg_pred_{step}"][masked]), "multinomial draw vs reference"
    assert np.array_equal(out.cpu().numpy(), g[f"mg_ref_{step}"]), "maskgit update vs reference golden"
    # and against the oracle chain evaluated on the GPU
    p = R.subs_parameterization(lg.to(dev()), x, mod, mi, tv).exp()
    ref = R.maskgit_update_from_noise(x, t, p, e, gum, sched[:, step], mi, r_temp=10)
    assert torch.equal(out, ref)


def test_maskgit_update_large_vocab_and_cfg(ops):
    """real vocabulary (V=48385), CFG pair, N beyond one CTA stride, ties in num_unmask: kernel == oracle chain on the GPU."""
    from oracle import restated as R
    B, N, V, tv, mi = 2, 96, 48385, 32001, 32000
    ldv = (V + 63) // 64 * 64
    g = torch.Generator().manual_seed(8)
    lc = (torch.randn(B, N, V, generator=g) * 3).to(bf16).float()
    lu = (torch.randn(B, N, V, generator=g) * 3).to(bf16).float()
    modality = torch.cat([torch.zeros(B, 32, dtype=torch.int64), torch.ones(B, N - 32, dtype=torch.int64)], 1)
    x = torch.where(modality == 0, torch.randint(0, tv - 1, (B, N), generator=g), torch.randint(tv, V, (B, N), generator=g))
    x[:, 1::2] = mi
    t = torch.tensor([[0.9], [0.4]])
    w = 2.0 * (1 - t.squeeze(-1))
    e = torch.empty(B * N, V).exponential_(1, generator=g)
    gum = torch.from_numpy(np.random.default_rng(3).gumbel(size=(B, N)))
    num = torch.tensor([7, 100], dtype=torch.int32)                   # second sample: more than its 48 masked tokens
    for cfg in (False, True):
        lgf = lc.to(dev())
        if cfg:
            wd = w.to(dev())
            lgf = (1 + wd)[:, None, None] * lc.to(dev()) - wd[:, None, None] * lu.to(dev())
        p = R.subs_parameterization(lgf, None if cfg else x.to(dev()), modality.to(dev()), mi, tv).exp()
        ref = R.maskgit_update_from_noise(x.to(dev()), t.to(dev()), p, e.to(dev()), gum.to(dev()), num.to(dev()), mi, r_temp=4.5)
        out, _, _ = ops.maskgit_update(x.to(dev()), _bf16_padded(lc, ldv), modality.to(dev()).view(-1), t.view(-1).to(dev()), num.to(dev()), mi, tv, V,
                                       r_temp=4.5, logits_uncond=_bf16_padded(lu, ldv) if cfg else None, cfg_w=w.to(dev()) if cfg else None,
                                       e_noise=e.to(dev()), gumbel=gum.to(dev()))
        torch.cuda.synchronize()
        assert torch.equal(out, ref), f"cfg={cfg}: {(out != ref).sum().item()} tokens differ"
        assert int((out[0] != x.to(dev())[0]).sum()) == 7 and int((out[1] == mi).sum()) == 0


def test_maskgit_update_philox_statistics(ops):
    """in-kernel Philox mode: exactly num_unmask tokens are revealed per sample, draws follow the SUBS softmax, the call is
    deterministic in (seed, offset)."""
    V, tv, mi, ldv = 24, 9, 8, 24
    B, N = 2000, 4
    g = torch.Generator().manual_seed(0)
    row_t = (torch.randn(V, generator=g) * 1.5).to(bf16)
    logits = row_t[None].repeat(B * N, 1).contiguous().to(dev())
    modality = torch.tensor([[0, 1, 1, 1]]).repeat(B, 1).to(dev())
    x = torch.full((B, N), mi, dtype=torch.int64, device=dev())
    t = torch.full((B,), 0.5, device=dev())
    num = torch.full((B,), 2, dtype=torch.int32, device=dev())
    out, pred, conf = ops.maskgit_update(x, logits, modality.view(-1), t, num, mi, tv, V, r_temp=10.0, seed=5, offset=3)
    out_b, _, _ = ops.maskgit_update(x, logits, modality.view(-1), t, num, mi, tv, V, r_temp=10.0, seed=5, offset=3)
    torch.cuda.synchronize()
    assert torch.equal(out, out_b)
    assert torch.equal((out != mi).sum(-1), torch.full((B,), 2, device=dev()))
    lf = row_t.float()
    pi = torch.softmax(lf[tv:V], 0)
    freq = torch.bincount(pred[:, 1:].reshape(-1).cpu() - tv, minlength=V - tv).float() / (3 * B)
    assert torch.allclose(freq, pi, atol=0.02), (freq - pi).abs().max()
    assert (pred[:, 0] < tv - 1).all() and torch.isfinite(conf).all()


def test_subs_argmax_matches_materialised(ops):
    """noise-removal arg-max kernel == argmax of the materialised SUBS log-probs (model_eval.py:2440-2446)."""
    B, N, V, tv, mi, ldv = 3, 40, 1001, 601, 600, 1024
    g = torch.Generator().manual_seed(2)
    lg = torch.zeros(B * N, ldv, dtype=bf16)
    lg[:, :V] = (torch.randn(B * N, V, generator=g) * 2).to(bf16)
    lg[5, 700:705] = lg[5, 700:].max() + 1                           # an exact tie: first index wins
    modality = torch.cat([torch.zeros(B, 8, dtype=torch.int64), torch.ones(B, N - 8, dtype=torch.int64)], 1)
    x = torch.where(modality == 0, torch.randint(0, tv - 1, (B, N), generator=g), torch.randint(tv, V, (B, N), generator=g))
    x[:, ::2] = mi
    lgd, xd, md = lg.to(dev()), x.to(dev()), modality.to(dev())
    ref = ops.subs_logprobs(lgd, xd.view(-1), md.view(-1), V, tv, mi).argmax(-1)
    out = ops.subs_argmax(lgd, xd.view(-1), md.view(-1), V, tv, mi)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    assert torch.equal(out.view(B, N)[xd != mi], xd[xd != mi])


# --------------------------------------------------------------------------------------------------
# optimizer / flat-buffer helpers
# --------------------------------------------------------------------------------------------------
def test_adamw_matches_torch(ops):
    n = 10007
    p0, g = rnd(n, seed=1), rnd(n, seed=2)
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    pb = torch.empty(n, device=dev(), dtype=bf16)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    for step in range(1, 4):
        ref_p.grad = g.clone()
        opt.step()
        ops.adamw_step(p, g, m, v, pb, 1e-3, 0.9, 0.999, 1e-8, 0.05, step)
    torch.cuda.synchronize()
    assert torch.allclose(p, ref_p.data, rtol=1e-5, atol=1e-6), (p - ref_p.data).abs().max()
    assert torch.equal(pb, p.to(bf16))


def test_flat_helpers(ops):
    n = 12345
    g = rnd(n, seed=1)
    out = torch.zeros(1, device=dev())
    ops.sumsq(g, out)
    dst = torch.empty(n, device=dev(), dtype=bf16)
    ops.grad_pack(g, dst, 0.125)
    back = torch.empty_like(g)
    ops.grad_unpack(dst, back)
    torch.cuda.synchronize()
    assert torch.allclose(out, (g * g).sum(), rtol=1e-4)
    assert torch.equal(dst, (g.to(bf16) / 8))
    assert torch.equal(back, dst.float())


def test_ddpm_update_from_logits_philox_distribution(ops):
    """fast (in-kernel Philox, log-domain Gumbel-max) mode: empirical distribution == the absorbing posterior
    P(v) = (mc_t-mc_s) p_v / mc_t, P(stay masked) = mc_s / mc_t   (model_eval.py:2064-2067)."""
    V, tv, mi, ldv = 24, 9, 8, 24
    R_, N = 40000, 2
    g = torch.Generator().manual_seed(0)
    row_t = (torch.randn(V, generator=g) * 1.5).to(bf16)
    logits = row_t[None].repeat(R_ * N, 1).contiguous().to(dev())
    modality = torch.tensor([[0, 1]]).repeat(R_, 1).to(dev())
    x = torch.full((R_, N), mi, dtype=torch.int64, device=dev())
    mc_t = torch.full((R_,), 0.8, device=dev())
    mc_s = torch.full((R_,), 0.6, device=dev())
    out = ops.ddpm_update_logits(x, logits, modality.view(-1), mc_t, mc_s, mi, tv, V, seed=123, offset=7)
    out2 = ops.ddpm_update_logits(x, logits, modality.view(-1), mc_t, mc_s, mi, tv, V, seed=123, offset=7)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    lf = row_t.float()
    for pos, (lo, hi) in enumerate([(0, tv - 1), (tv, V)]):       # text row: cols 0..7 (mask col 8 excluded); image row: 9..23
        p = torch.zeros(V + 1)
        p[lo:hi] = torch.softmax(lf[lo:hi], 0) * (0.2 / 0.8)
        p[V] = 0.6 / 0.8                                            # slot V counts "stay masked"
        o = out[:, pos].cpu()
        o = torch.where(o == mi, torch.full_like(o, V), o)
        freq = torch.bincount(o, minlength=V + 1).float() / R_
        assert (freq[:lo].sum() + freq[hi:V].sum()) == 0, "sampled outside the valid vocabulary range"
        assert torch.allclose(freq, p, atol=0.012), (freq - p).abs().max()
