"""Thin data-parallel wrapper + fused optimizer for the flat-parameter DIT.

ThinDDP replaces torch DDP + `DDPCommunicationHookType.BF16` (reference main.py:641-656, model_setup.py:703):
one process per GPU, gradients of each DiT block are all-reduced as soon as that block's backward has retired,
on a side stream, overlapped with the backward of the blocks below it.  Wire format and arithmetic follow torch's
bf16 compress hook (`buffer.to(bf16).div_(world)` -> all-reduce SUM -> copy back to the fp32 bucket).

FusedAdamW replaces `torch.optim.AdamW(fused=True)` + `clip_grad_norm_` (reference model_setup.py:385-424,
model.py:1518-1537): one pass over the flat buffers that also emits the bf16 shadow weights the tensor cores read.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops

bf16 = torch.bfloat16


def nccl_options(max_ctas=None):
    """ProcessGroupNCCL options that cap the SMs NCCL takes (`ncclConfig_t.maxCTAs`); 0 = NCCL's default (what bench.py uses).
    Measured on 2xB200 (profiles/r01_ddp_overlap.md): capping NCCL to 2/4/8/16 CTAs makes the step SLOWER (175.9 / 128.9 /
    121.8 / 118.5 ms vs 116.0 ms uncapped) — a longer-running all-reduce keeps a few SMs slow for longer and the persistent
    GEMMs' statically assigned tiles wait for them — so the knob stays off by default."""
    if max_ctas is None:
        max_ctas = int(os.environ.get("UD_NCCL_MAX_CTAS", 0))
    if max_ctas <= 0:
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = max_ctas
    opts.config.min_ctas = min(max_ctas, 4)
    return opts


class ThinDDP(nn.Module):
    def __init__(self, module, process_group=None, bf16_compress: bool = True, _pack=None, _unpack=None, side_ctas=None):
        super().__init__()
        self.module = module
        self.pg = process_group
        # (de)compression kernels; injectable so the bucket planning / collective sequencing can be exercised with a
        # gloo process group on CPU in tests (the product path always uses the CUDA kernels)
        # optional grid cap of the (de)compression kernels on the side stream (0 = uncapped, the measured optimum: a short
        # wide kernel disturbs the overlapped backward less than a long narrow one, see nccl_options)
        self.side_ctas = int(os.environ.get("UD_DDP_SIDE_CTAS", 0)) if side_ctas is None else int(side_ctas)
        self._pack = _pack or (lambda g, d, w: ops.grad_pack(g, d, w, self.side_ctas))
        self._unpack = _unpack or (lambda s, g: ops.grad_unpack(s, g, self.side_ctas))
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bf16_compress = bf16_compress
        self._sync = True
        self._is_accelerate_prepared = True     # keep accelerate from wrapping this in torch DDP
        module._ensure_ready()
        if self.world > 1:
            dist.broadcast(module.flat_params, 0, group=self.pg)
            module.mark_weights_updated()
        self._ranges_by_block = {}
        named = dict(module.named_parameters())
        slot = lambda n: (module._offs[n], module._offs[n] + (named[n].numel() + 63) // 64 * 64)
        claimed = []
        for i in range(module.n_blocks):
            (blo, bhi), (slo, shi) = module.block_grad_range(i)
            self._ranges_by_block[i] = [(blo, bhi), (slo, shi)]
            claimed += [(blo, bhi), (slo, shi)]
        head = slot("output_layer.linear.weight")
        self._ranges_by_block[module.n_blocks] = [head]
        claimed.append(head)
        # everything not owned by a block or the head weight goes out at the very end
        rest, cur = [], 0
        for lo, hi in sorted(claimed):
            if lo > cur:
                rest.append((cur, lo))
            cur = max(cur, hi)
        total = module.flat_grads.numel()
        if cur < total:
            rest.append((cur, total))
        self._ranges_by_block[-1] = rest
        self._max_range = max(hi - lo for rs in self._ranges_by_block.values() for lo, hi in rs)
        self._stage = None
        self._comm_stream = None
        self._done_event = None
        module.grad_ready_hook = self._on_grads_ready
        self.bytes_on_wire_per_step = 0
        self.debug_timing = bool(int(os.environ.get("UD_DDP_DEBUG", "0")))
        self._dbg_events = []

    def debug_report(self):
        """(UD_DDP_DEBUG=1) per-bucket timeline of the last step, in ms relative to the first gradient-ready event:
        ready = when backward produced the bucket, start/end = the pack -> all-reduce -> unpack chain on the comm stream."""
        torch.cuda.synchronize()
        if not self._dbg_events:
            return []
        t0 = self._dbg_events[0][1]
        rows = [(b, t0.elapsed_time(ev), t0.elapsed_time(c0), t0.elapsed_time(c1)) for b, ev, c0, c1 in self._dbg_events]
        self._dbg_events = []
        return rows

    def forward(self, *a, **k):
        return self.module(*a, **k)

    @contextlib.contextmanager
    def no_sync(self):
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    def _on_grads_ready(self, block_idx):
        if self.world == 1 or not self._sync:
            return
        g = self.module.flat_grads
        cuda = g.is_cuda
        if self._stage is None:
            self._stage = torch.empty(self._max_range, device=g.device, dtype=bf16 if self.bf16_compress else torch.float32)
            if cuda:
                self._comm_stream = torch.cuda.Stream(priority=-1)
        dbg = cuda and self.debug_timing
        if cuda:
            ev = torch.cuda.Event(enable_timing=dbg)
            ev.record(torch.cuda.current_stream())
            ctx = torch.cuda.stream(self._comm_stream)
        else:
            ctx = contextlib.nullcontext()
        with ctx:
            if cuda:
                self._comm_stream.wait_event(ev)
            if dbg:
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(self._comm_stream)
            for lo, hi in self._ranges_by_block.get(block_idx, []):
                seg = g[lo:hi]
                if self.bf16_compress:
                    st = self._stage[: hi - lo]
                    self._pack(seg, st, 1.0 / self.world)
                    dist.all_reduce(st, group=self.pg)
                    self._unpack(st, seg)
                    self.bytes_on_wire_per_step += 2 * (hi - lo)
                else:
                    seg.div_(self.world)
                    dist.all_reduce(seg, group=self.pg)
                    self.bytes_on_wire_per_step += 4 * (hi - lo)
            if dbg:
                c1.record(self._comm_stream)
                self._dbg_events.append((block_idx, ev, c0, c1))
            if block_idx == -1 and cuda:
                self._done_event = torch.cuda.Event(enable_timing=dbg)
                self._done_event.record(self._comm_stream)
        if block_idx == -1 and cuda and self._done_event is not None:
            torch.cuda.current_stream().wait_event(self._done_event)   # optimizer waits on ONE event


class FusedAdamW:
    """AdamW over the DIT's flat fp32 buffers (torch.optim.AdamW semantics, single param group like the reference)."""

    def __init__(self, module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None):
        self.module = module.module if isinstance(module, ThinDDP) else module
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        p = self.module.flat_params
        self.exp_avg = torch.zeros_like(p)
        self.exp_avg_sq = torch.zeros_like(p)
        self._sumsq = torch.zeros(1, device=p.device)
        self._scale = torch.ones(1, device=p.device)
        self.last_grad_norm = None

    def zero_grad(self, set_to_none: bool = True):
        self.module._force_fresh_grads = True     # next backward overwrites the GEMM grads and re-zeroes the rest

    @torch.no_grad()
    def step(self):
        m = self.module
        p, g = m._flat_p, m._flat_g              # (not the properties: they would refresh the bf16 shadow we are about to rewrite)
        self.step_count += 1
        scale = None
        if self.max_grad_norm is not None:
            self._sumsq.zero_()
            ops.sumsq(g, self._sumsq)
            norm = self._sumsq.sqrt()
            self.last_grad_norm = norm
            torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0, out=self._scale)    # torch clip_grad_norm_ coefficient
            scale = self._scale
        ops.adamw_step(p, g, self.exp_avg, self.exp_avg_sq, m._flat_bf16, self.lr, self.betas[0], self.betas[1], self.eps,
                       self.weight_decay, self.step_count, grad_scale=scale)
        m.mark_weights_updated(shadow_is_current=True)

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, lr=self.lr, betas=self.betas,
                    eps=self.eps, weight_decay=self.weight_decay)

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
