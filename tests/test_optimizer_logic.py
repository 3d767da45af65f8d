"""CPU: control flow of the streamed optimizer (FusedAdamW, overlap mode) with the CUDA pieces mocked out — streams / events
become no-ops and the two kernels (ud_sumsq_f32, ud_adamw_step) are replaced by torch restatements.  What is checked is the host
logic the GPU tests cannot isolate: bucket planning (every flat element updated exactly once, in forward order), the per-bucket
gradient-norm partial sums (with and without the wgrad-epilogue fusion), the fallback to a full norm pass after gradient
accumulation, and the chaining behind ThinDDP's gradient hook.  The kernels themselves are covered by tests/test_*_gpu.py."""
import contextlib

import pytest
import torch

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_ddp_gloo import FakeFlat  # noqa: E402


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def wait_event(self, e):
        pass

    def record(self, *a):
        pass


@pytest.fixture()
def mocked(monkeypatch):
    from unidisc_b200 import ddp as D
    calls = dict(sumsq=0, adamw=[])
    monkeypatch.setattr(torch.cuda, "Stream", _Dummy)
    monkeypatch.setattr(torch.cuda, "Event", _Dummy)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: _Dummy())

    def sumsq(g, out, max_ctas=0):
        calls["sumsq"] += g.numel()
        out += (g.double() ** 2).sum().float()

    def adamw_step(p, g, m, v, pb, lr, b1, b2, eps, wd, step, grad_scale=None, max_ctas=0):
        calls["adamw"].append((p.data_ptr(), p.numel()))
        gg = g * (grad_scale if grad_scale is not None else 1.0)
        p.mul_(1 - lr * wd)
        m.mul_(b1).add_(gg, alpha=1 - b1)
        v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
        denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + eps
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
        pb.copy_(p.to(torch.bfloat16))

    monkeypatch.setattr(D.ops, "sumsq", sumsq)
    monkeypatch.setattr(D.ops, "adamw_step", adamw_step)
    return D, calls


def _module():
    m = FakeFlat(n_blocks=3)
    g = torch.Generator().manual_seed(0)
    m.flat_params.copy_(torch.randn(m.flat_params.numel(), generator=g))
    m._flat_p, m._flat_g = m.flat_params, m.flat_grads
    m._flat_bf16 = torch.zeros_like(m.flat_params, dtype=torch.bfloat16)
    m._last_bwd_fused_sumsq, m.grad_sumsq_acc, m._param_events = False, None, None
    return m


def _backward(m, seed, accumulate=False):
    """what DIT._backward_impl does to the hooks: head first, blocks last to first, then the rest"""
    g = torch.randn(m.flat_grads.numel(), generator=torch.Generator().manual_seed(seed)) * 0.01
    if accumulate:
        m.flat_grads.add_(g)
    else:
        m.flat_grads.copy_(g)
    if m._last_bwd_fused_sumsq and m.grad_sumsq_acc is not None:        # the wgrad GEMM epilogues add sum(dW^2) of the GEMM weights
        m.grad_sumsq_acc += (m.flat_grads[: m._big_end].double() ** 2).sum().float()
    for b in [m.n_blocks] + list(range(m.n_blocks - 1, -1, -1)) + [-1]:
        if m.grad_ready_hook is not None:
            m.grad_ready_hook(b)


def _reference(p0, grads_per_step, lr, wd, clip):
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=lr, weight_decay=wd)
    norms = []
    for g in grads_per_step:
        p.grad = g.clone()
        norms.append(torch.nn.utils.clip_grad_norm_([p], clip))
        opt.step()
    return p.detach(), norms


@pytest.mark.parametrize("fused_epilogue", [False, True])
def test_streamed_optimizer_single_gpu_logic(mocked, fused_epilogue):
    D, calls = mocked
    m = _module()
    p0 = m.flat_params.clone()
    opt = D.FusedAdamW(m, lr=1e-2, weight_decay=0.1, max_grad_norm=0.05, overlap=True)
    assert opt.overlap and m.grad_ready_hook is not None and m.grad_sumsq_acc is opt._sumsq
    m._last_bwd_fused_sumsq = fused_epilogue
    # the stages cover every flat element exactly once, in forward order: rest, block 0..L-1, head
    cover = torch.zeros(m.flat_params.numel(), dtype=torch.int32)
    for _, ranges in opt._stages:
        for lo, hi in ranges:
            cover[lo:hi] += 1
    assert bool((cover == 1).all())
    assert [n for n, _ in opt._stages] == ["pre", 0, 1, 2, "head"]
    grads = []
    for step in range(3):
        calls["sumsq"], calls["adamw"] = 0, []
        _backward(m, seed=10 + step)
        grads.append(m.flat_grads.clone())
        assert opt._buckets_seen == m.n_blocks + 2
        opt.step()
        # with the epilogue fusion only the small parameters are re-read for the norm
        assert calls["sumsq"] == (m.flat_grads.numel() - m._big_end if fused_epilogue else m.flat_grads.numel())
        assert sum(n for _, n in calls["adamw"]) == m.flat_params.numel()
        assert set(m._param_events) == {"pre", 0, 1, 2, "head"} and float(opt._sumsq) == 0.0
    ref, norms = _reference(p0, grads, 1e-2, 0.1, 0.05)
    assert torch.allclose(opt.last_grad_norm, norms[-1], rtol=1e-5)
    assert torch.allclose(m.flat_params, ref, rtol=1e-5, atol=1e-7)
    assert torch.equal(m._flat_bf16, m.flat_params.to(torch.bfloat16))
    # gradient accumulation: two backwards before the step -> the partial sums are stale, one full pass instead
    m._last_bwd_fused_sumsq = False
    _backward(m, seed=20)
    _backward(m, seed=21, accumulate=True)
    calls["sumsq"] = 0
    total = m.flat_grads.double().pow(2).sum().sqrt().float()
    opt.step()
    assert torch.allclose(opt.last_grad_norm, total, rtol=1e-5)
    assert calls["sumsq"] == m.flat_grads.numel()


def test_streamed_optimizer_chains_behind_thin_ddp(mocked):
    D, calls = mocked
    m = _module()
    ddp = D.ThinDDP(m)                                   # torch.distributed not initialised: world 1, same hook plumbing
    hook = m.grad_ready_hook
    assert hook.__self__ is ddp
    p0 = m.flat_params.clone()
    for handle in (m, ddp):                              # the bare module (hook owner detected) or the wrapper
        m.flat_params.copy_(p0)
        opt = D.FusedAdamW(handle, lr=1e-2, weight_decay=0.0, max_grad_norm=0.05, overlap=True)
        assert opt.ddp is ddp and m.grad_ready_hook is hook, "the all-reduce hook must stay installed"
        assert ddp.post_bucket_hook == opt._on_bucket_final and m.grad_sumsq_acc is None
        _backward(m, seed=30)
        assert opt._buckets_seen == m.n_blocks + 2
        g = m.flat_grads.clone()
        opt.step()
        ref, norms = _reference(p0, [g], 1e-2, 0.0, 0.05)
        assert torch.allclose(opt.last_grad_norm, norms[0], rtol=1e-5) and torch.allclose(m.flat_params, ref, rtol=1e-5, atol=1e-7)
        # accumulation micro-step under no_sync(): the hook is not counted, the step falls back to a full norm pass
        with ddp.no_sync():
            _backward(m, seed=31)
        assert opt._buckets_seen == 0
    # world > 1: ThinDDP calls the hook inside its communication-stream context right after grad_unpack
    opt._sumsq.zero_()
    opt._buckets_seen = 0
    calls["sumsq"] = 0
    for b in [m.n_blocks] + list(range(m.n_blocks - 1, -1, -1)) + [-1]:
        opt._on_bucket_final(b, ddp._ranges_by_block[b], True)
    assert opt._buckets_seen == m.n_blocks + 2 and calls["sumsq"] == m.flat_grads.numel()
    assert torch.allclose(opt._sumsq, m.flat_grads.double().pow(2).sum().float().reshape(1), rtol=1e-5)
    # a foreign hook is never silently replaced
    m2 = _module()
    m2.grad_ready_hook = lambda b: None
    with pytest.raises(RuntimeError):
        D.FusedAdamW(m2, max_grad_norm=1.0, overlap=True)
