#!/usr/bin/env python
"""Per-kernel timing at the unidisc-1.4B shapes (B=8, N=1280, D=2048, H=16) with CUDA events. Dev tool (GPU box)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unidisc_b200 import _lib as L
from unidisc_b200 import ops

bf16 = torch.bfloat16
dev = torch.device("cuda", 0)
B, N, D, H = 8, 1280, 2048, 16
hd = D // H
M = B * N


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rb(*s):
    return (torch.randn(*s, device=dev) * 0.5).to(bf16)


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["attn", "rows", "gemm"]
    if "attn" in which:
        qk, qkv = rb(M, 2 * D), rb(M, 3 * D)
        q, k, v = qk[:, :D], qk[:, D:], qkv[:, 2 * D:]
        sc = 1 / math.sqrt(hd)
        o, lse = ops.attn_fwd(q, k, v, B, N, H, hd, sc)
        t = timeit(lambda: ops.attn_fwd(q, k, v, B, N, H, hd, sc, o=o))
        fl = 4 * B * H * N * N * hd
        print(f"attn_fwd  {t*1e3:8.1f} us  {fl/t/1e9:7.1f} TFLOP/s")
        do = rb(M, D)
        dqk, dqkv = torch.empty_like(qk), torch.empty_like(qkv)
        t = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, dqk[:, :D], dqk[:, D:], dqkv[:, 2 * D:], B, N, H, hd, sc))
        print(f"attn_bwd  {t*1e3:8.1f} us  {2.5*fl/t/1e9:7.1f} TFLOP/s (5-GEMM flops)   [UD_ATTN_BWD_V1={os.environ.get('UD_ATTN_BWD_V1')}]")
    if "attn" in which and "--vs-cudnn" in sys.argv:
        # the bar (reference models/dit.py:816-829: SDPA forced onto the cuDNN backend): same B8 H16 N1280 hd128 tensors, fwd and fwd+bwd
        import torch.nn.functional as F
        from torch.nn.attention import SDPBackend, sdpa_kernel
        for backend in (SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION):
            try:
                qh, kh, vh = (t.reshape(B, N, H, hd).permute(0, 2, 1, 3).detach().clone().requires_grad_(True) for t in (q, k, v))
                with sdpa_kernel(backends=[backend]):
                    t_f = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh))
                    oo = F.scaled_dot_product_attention(qh, kh, vh)
                    go = torch.randn_like(oo)

                    def fb():
                        o2 = F.scaled_dot_product_attention(qh, kh, vh)
                        o2.backward(go)
                        qh.grad = kh.grad = vh.grad = None
                    t_fb = timeit(fb)
                print(f"torch SDPA {backend.name:18s} fwd {t_f*1e3:8.1f} us {fl/t_f/1e9:7.1f} TFLOP/s | fwd+bwd {t_fb*1e3:8.1f} us -> bwd {(t_fb-t_f)*1e3:8.1f} us "
                      f"{2.5*fl/(t_fb-t_f)/1e9:7.1f} TFLOP/s (5-GEMM flops)")
            except Exception as e:  # noqa: BLE001
                print(f"torch SDPA {backend.name}: unavailable ({type(e).__name__}: {str(e)[:120]})")
    if "rows" in which:
        a, x = rb(M, D), torch.randn(M, D, device=dev)
        w = torch.ones(D, device=dev)
        xo, h, ra, rx = ops.norm_residual_fwd(a, x, w, w)
        t = timeit(lambda: ops.norm_residual_fwd(a, x, w, w, x_out=xo, h=h))
        print(f"norm_residual_fwd {t*1e3:8.1f} us  {(M*D*(2+4+4+2))/t/1e6:7.1f} GB/s")
        g, dh = torch.randn(M, D, device=dev), rb(M, D)
        dwn, dwa = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        gi, da = ops.norm_residual_bwd(g, dh, xo, rx, w, a, ra, w, dwn, dwa)
        t = timeit(lambda: ops.norm_residual_bwd(g, dh, xo, rx, w, a, ra, w, dwn, dwa, g_in=gi, da=da))
        print(f"norm_residual_bwd {t*1e3:8.1f} us  {(M*D*(4+2+4+2+4+2))/t/1e6:7.1f} GB/s")
        t = timeit(lambda: ops.norm_residual_bwd(g, dh, xo, rx, w, a, ra, w, dwn, dwa, g_in=gi, da=da, p_drop=0.1, seed=1, offset=3))
        print(f"norm_residual_bwd (dropout 0.1) {t*1e3:8.1f} us  {(M*D*(4+2+4+2+4+2))/t/1e6:7.1f} GB/s")
        qkv = rb(M, 3 * D)
        cos, sin = torch.rand(M, hd // 2, device=dev), torch.rand(M, hd // 2, device=dev)
        t = timeit(lambda: ops.qk_ln_rope_fwd(qkv, w, w, w, w, cos, sin, hd))
        print(f"qk_ln_rope_fwd    {t*1e3:8.1f} us  {(M*D*8)/t/1e6:7.1f} GB/s")
        out, stats = ops.qk_ln_rope_fwd(qkv, w, w, w, w, cos, sin, hd)
        dqk, dqkv = rb(M, 2 * D), torch.empty(M, 3 * D, device=dev, dtype=bf16)
        gz = [torch.zeros(D, device=dev) for _ in range(4)]
        t = timeit(lambda: ops.qk_ln_rope_bwd(dqk, qkv, stats, w, w, cos, sin, dqkv, *gz, hd))
        print(f"qk_ln_rope_bwd    {t*1e3:8.1f} us  {(M*D*12)/t/1e6:7.1f} GB/s")
        dy = rb(M, 4 * D)
        db = torch.zeros(4 * D, device=dev)
        t = timeit(lambda: ops.colsum(dy, db))
        print(f"colsum [M,4D]     {t*1e3:8.1f} us  {(M*4*D*2)/t/1e6:7.1f} GB/s")
    if "sampler" in which:
        # vocabulary-pass kernels of BASELINE configs[3] (B=64, N=1280, all rows masked, bf16 logits resident): fused absorbing
        # update and MaskGIT draw + selection, in-kernel Philox noise
        Bs, V, tv, mi = 64, 48385, 32001, 32000
        Vp = (V + 63) // 64 * 64
        lg = torch.empty(Bs * N, Vp, device=dev, dtype=bf16)
        for c in range(0, Bs * N, 8192):
            lg[c:c + 8192].normal_(0, 3)
        xs = torch.full((Bs, N), mi, dtype=torch.int64, device=dev)
        mods = torch.cat([torch.zeros(Bs, 256, dtype=torch.int64), torch.ones(Bs, 1024, dtype=torch.int64)], 1).to(dev).view(-1)
        tt = torch.full((Bs,), 0.7, device=dev)
        num = torch.full((Bs,), 20, dtype=torch.int32, device=dev)
        alg = Bs * (256 * tv + 1024 * (V - tv)) * 2 + Bs * N * 16
        t = timeit(lambda: ops.ddpm_update_logits(xs, lg, mods, tt, tt - 0.01, mi, tv, V, seed=1, offset=1), iters=10)
        print(f"ddpm_update_logits (fused, Philox) {t*1e3:8.1f} us  {alg/t/1e6:7.1f} GB/s")
        t = timeit(lambda: ops.maskgit_update(xs, lg, mods, tt, num, mi, tv, V, seed=1, offset=1), iters=10)
        print(f"maskgit_update (draw + select)     {t*1e3:8.1f} us  {alg/t/1e6:7.1f} GB/s")
        t = timeit(lambda: ops.subs_argmax(lg, xs.view(-1), mods, V, tv, mi), iters=10)
        print(f"subs_argmax (noise removal)        {t*1e3:8.1f} us  {alg/t/1e6:7.1f} GB/s")
        del lg
    if "roof" in which:
        # exactly bench.py's roofline kernel: the mlp.0 GEMM with bias + GELU epilogue (for the ncu traffic capture)
        a, b, c, g = rb(M, D), rb(4 * D, D), torch.empty(M, 4 * D, device=dev, dtype=bf16), torch.empty(M, 4 * D, device=dev, dtype=bf16)
        bias = rb(4 * D)
        t = timeit(lambda: ops.gemm(a, b, out=c, epi=L.EPI_BF16_GELU, bias=bias, aux=g), iters=10)
        print(f"gemm roofline {M}x{4*D}x{D} {t*1e3:8.1f} us  {2*M*4*D*D/t/1e9:7.1f} TFLOP/s")
    if "gemm" in which:
        shapes = [("qkv fwd", 0, 0, M, 3 * D, D, L.EPI_BF16), ("out fwd", 0, 0, M, D, D, L.EPI_BF16), ("mlp1 gelu", 0, 0, M, 4 * D, D, L.EPI_BF16_GELU),
                  ("mlp2 fwd", 0, 0, M, D, 4 * D, L.EPI_BF16), ("head fwd", 0, 0, M, 48385, D, L.EPI_BF16),
                  ("dgrad mlp2->dgelu", 0, 1, M, 4 * D, D, L.EPI_BF16_DGELU), ("dgrad mlp1", 0, 1, M, D, 4 * D, L.EPI_BF16), ("dgrad qkv", 0, 1, M, D, 3 * D, L.EPI_BF16),
                  ("dgrad out", 0, 1, M, D, D, L.EPI_BF16), ("wgrad qkv", 1, 1, 3 * D, D, M, L.EPI_F32), ("wgrad out", 1, 1, D, D, M, L.EPI_F32),
                  ("wgrad mlp1", 1, 1, 4 * D, D, M, L.EPI_F32), ("wgrad mlp2", 1, 1, D, 4 * D, M, L.EPI_F32), ("wgrad head", 1, 1, 48385, D, M, L.EPI_F32_ACC)]
        for name, ta, tb, m, n, k, epi in shapes:
            ldn = (n + 63) // 64 * 64
            A = rb(k, (m + 63) // 64 * 64)[:, :m] if ta else rb(m, k)
            Bm = rb(k, ldn)[:, :n] if tb else rb(n, k)
            fp = epi in (L.EPI_F32, L.EPI_F32_ACC)
            C = torch.zeros(m, ldn, device=dev, dtype=torch.float32 if fp else bf16)[:, :n]
            aux = rb(m, ldn)[:, :n] if epi in (L.EPI_BF16_GELU, L.EPI_BF16_DGELU) else None
            bias = rb(n) if epi == L.EPI_BF16_GELU else None
            for bn in ([0] if len(sys.argv) < 3 else [0, 128, 1024 + 256]):
                t = timeit(lambda: ops.gemm(A, Bm, ta=bool(ta), tb=bool(tb), M=m, N=n, K=k, out=C, epi=epi, aux=aux, bias=bias, bn=bn), iters=10)
                print(f"gemm {name:18s} bn={bn:4d} {m:6d}x{n:6d}x{k:6d} {t*1e3:8.1f} us  {2*m*n*k/t/1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
