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
        # `sumsq_target` (FusedAdamW, UD_DDP_FUSED_SUMSQ=1, experimental): the decompression also accumulates the gradient norm
        self.sumsq_target = None
        self._unpack = _unpack or (lambda s, g: ops.grad_unpack(s, g, self.side_ctas, sumsq=self.sumsq_target))
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
        # GEMM-weight gradients never exist in fp32 before the all-reduce: the wgrad GEMM epilogue writes bf16(bf16(dW) / world)
        # into this buffer (DIT.attach_grad_stage), which is all-reduced in place and decompressed into the fp32 gradient
        self._wstage = None
        if self.world > 1 and bf16_compress and module.flat_grads.is_cuda and hasattr(module, "attach_grad_stage") \
                and not bool(int(os.environ.get("UD_DDP_NO_WGRAD_STAGE", "0"))):
            self._wstage = torch.zeros(module._big_end, device=module.flat_grads.device, dtype=bf16)
            self._alpha = torch.full((1,), 1.0 / self.world, device=module.flat_grads.device, dtype=torch.float32)
            module.attach_grad_stage(self._wstage, self._alpha, lambda: self._sync)
        module.grad_ready_hook = self._on_grads_ready
        self.post_bucket_hook = None       # FusedAdamW: hook(block_idx, ranges, on_side_stream) once a bucket's gradients are final
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

    def __getattr__(self, name):
        # like accelerate's unwrapped access: attributes the wrapper does not have (require_sample_ids, Vp, supports_head_rows,
        # set_flex_attention_cache, ...) are those of the wrapped backbone, so `self.backbone = ThinDDP(self.backbone)` is enough
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name == "module":
                raise
            return getattr(super().__getattr__("module"), name)

    @contextlib.contextmanager
    def no_sync(self):
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    def _on_grads_ready(self, block_idx):
        if self.world == 1 or not self._sync:
            if self._sync and self.post_bucket_hook is not None:
                self.post_bucket_hook(block_idx, self._ranges_by_block.get(block_idx, []), False)
            return
        # (not the `flat_grads` property: it refreshes the bf16 weight shadow when the backward has just marked it stale — an
        #  8.4 GB cast per step that the optimizer overwrites right after)
        g = getattr(self.module, "_flat_g", None)
        if g is None:
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
            staged = self._wstage is not None and getattr(self.module, "_last_bwd_staged", False)
            for lo, hi in self._ranges_by_block.get(block_idx, []):
                seg = g[lo:hi]
                if self.bf16_compress:
                    if staged and hi <= self._wstage.numel():
                        st = self._wstage[lo:hi]                 # already in wire format (wgrad GEMM epilogue)
                    else:
                        st = self._stage[: hi - lo]
                        self._pack(seg, st, 1.0 / self.world)
                    dist.all_reduce(st, group=self.pg)
                    self._unpack(st, seg)
                    self.bytes_on_wire_per_step += 2 * (hi - lo)
                else:
                    seg.div_(self.world)
                    dist.all_reduce(seg, group=self.pg)
                    self.bytes_on_wire_per_step += 4 * (hi - lo)
            if self.post_bucket_hook is not None:
                self.post_bucket_hook(block_idx, self._ranges_by_block.get(block_idx, []), True)
            if dbg:
                c1.record(self._comm_stream)
                self._dbg_events.append((block_idx, ev, c0, c1))
            if block_idx == -1 and cuda:
                self._done_event = torch.cuda.Event(enable_timing=dbg)
                self._done_event.record(self._comm_stream)
        if block_idx == -1 and cuda and self._done_event is not None:
            torch.cuda.current_stream().wait_event(self._done_event)   # optimizer waits on ONE event


class FusedAdamW(torch.optim.Optimizer):
    """AdamW over the DIT's flat fp32 buffers (torch.optim.AdamW semantics, single param group like the reference).

    It IS a `torch.optim.Optimizer` (one param group holding the module's parameters), so the reference's
    `hydra.utils.instantiate(config.lr_scheduler, optimizer=optimizer)` (model_setup.py:426, LambdaLR warm-up) drives it:
    lr / betas / eps / weight_decay are read from `param_groups[0]` at every step.  The moments live in two flat buffers
    (`exp_avg`, `exp_avg_sq`, the layout of DIT._flat_p), exposed per parameter as views through `state[p]`.

    `overlap=True` (default on CUDA) streams the optimizer around the compute-bound GEMM phases instead of running it as
    one exposed HBM-bound block between backward and forward:
      * the gradient-norm partial sums of a bucket (DiT block / head / rest) are taken on a side stream as soon as the
        backward (and, under ThinDDP, the all-reduce) has finalised that bucket, i.e. next to the remaining backward GEMMs;
      * `step()` returns after ENQUEUEING the update on the side stream, bucket by bucket in forward order (embeddings,
        block 0 ... block L-1, head); the next forward waits on one event per bucket right before it first reads that
        bucket's weights, so AdamW of block i+1.. runs under the forward GEMMs of blocks ..i.
    Arithmetic is unchanged (same kernels, same clip coefficient); only the schedule differs.  Anything that reads the
    parameters outside DIT.forward must call `join()` (DIT.state_dict / flat_params do)."""

    def __init__(self, module, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None, overlap=None):
        self.ddp = module if isinstance(module, ThinDDP) else None
        self.module = module.module if isinstance(module, ThinDDP) else module
        self.module._ensure_ready()           # parameters become views of the flat buffer BEFORE the param group captures them
        super().__init__([p for _, p in self.module.named_parameters()],
                         dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if self.ddp is None:
            # handed the bare DIT although a ThinDDP wraps it: its grad_ready_hook must stay in place (replacing it would
            # silently switch the gradient all-reduce off), so the optimizer chains behind it
            owner = getattr(getattr(self.module, "grad_ready_hook", None), "__self__", None)
            if isinstance(owner, ThinDDP):
                self.ddp = owner
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        p = self.module.flat_params
        self.exp_avg = torch.zeros_like(p)
        self.exp_avg_sq = torch.zeros_like(p)
        offs = getattr(self.module, "_offs", None)
        if offs is not None:                  # per-parameter views of the flat moments (torch.optim state layout)
            for n, q in self.module.named_parameters():
                o, k = offs[n], q.numel()
                self.state[q] = dict(step=torch.zeros(()), exp_avg=self.exp_avg[o:o + k].view(q.shape),
                                     exp_avg_sq=self.exp_avg_sq[o:o + k].view(q.shape))
        self._sumsq = torch.zeros(1, device=p.device)
        self._scale = torch.ones(1, device=p.device)
        self._norm = None
        if overlap is None:
            overlap = p.is_cuda and not bool(int(os.environ.get("UD_OPT_SERIAL", "0")))
        self.overlap = bool(overlap)
        self._stream = torch.cuda.Stream() if self.overlap else None
        # grid cap of the side-stream kernels that run next to tensor-core GEMMs (0 = uncapped): a bandwidth-saturating burst
        # starves the GEMMs' TMA loads, a throttled stream hides under them (measured)
        self.side_ctas = int(os.environ.get("UD_OPT_CTAS", 0))
        self.sumsq_ctas = int(os.environ.get("UD_OPT_SUMSQ_CTAS", 0))
        self.eager_buckets = int(os.environ.get("UD_OPT_EAGER", 3))      # first buckets the forward needs at once: uncapped
        self.debug_timing = bool(int(os.environ.get("UD_OPT_DEBUG", "0")))   # timing-enabled bucket events (bench timeline)
        self._last_event = None
        self._buckets_seen = 0
        self._stages = self._plan_stages() if self.overlap else None
        if self.overlap and self.max_grad_norm is not None:
            if self.ddp is not None:
                self.ddp.post_bucket_hook = self._on_bucket_final
                if bool(int(os.environ.get("UD_DDP_FUSED_SUMSQ", "1"))) and self.ddp.bf16_compress and self.ddp.world > 1 \
                        and self._sumsq.is_cuda:
                    # the decompression kernel that writes the all-reduced gradient also sums its squares: no separate pass
                    self.ddp.sumsq_target = self._sumsq        # the hook below then only counts the bucket
            else:
                prev = self.module.grad_ready_hook
                if prev is not None and not getattr(prev, "_ud_fused_adamw", False):
                    raise RuntimeError("FusedAdamW(overlap): DIT.grad_ready_hook is already taken by something that is not a ThinDDP")
                by_block = {b: self._ranges_of(b) for b in range(-1, self.module.n_blocks + 1)}      # planned once (host cost)
                big_end = self.module._big_end
                small_only = {b: [r for r in rs if r[0] >= big_end] for b, rs in by_block.items()}
                # single GPU: the wgrad GEMM epilogues add sum(dW^2) of the GEMM weights straight into the accumulator
                # (DIT.grad_sumsq_acc); only the small parameters still need a pass
                if not bool(int(os.environ.get("UD_NO_FUSED_SUMSQ", "0"))):
                    self.module.grad_sumsq_acc = self._sumsq
                hook = lambda b: self._on_bucket_final(
                    b, small_only[b] if self.module._last_bwd_fused_sumsq else by_block[b], False)
                hook._ud_fused_adamw = True          # a later FusedAdamW on the same module may replace this hook
                self.module.grad_ready_hook = hook

    # ---- bucket plan: (name, [ranges]) in the order the forward first reads the weights ----
    def _ranges_of(self, block_idx):
        m = self.module
        if 0 <= block_idx < m.n_blocks:
            return list(m.block_grad_range(block_idx))
        named = dict(m.named_parameters())
        n = "output_layer.linear.weight"
        head = (m._offs[n], m._offs[n] + (named[n].numel() + 63) // 64 * 64)
        if block_idx == m.n_blocks:
            return [head]
        claimed = sorted([r for i in range(m.n_blocks) for r in m.block_grad_range(i)] + [head])
        rest, cur = [], 0
        for lo, hi in claimed:
            if lo > cur:
                rest.append((cur, lo))
            cur = max(cur, hi)
        if cur < m._flat_g.numel():
            rest.append((cur, m._flat_g.numel()))
        return rest

    def _plan_stages(self):
        m = self.module
        return [("pre", self._ranges_of(-1))] + [(i, self._ranges_of(i)) for i in range(m.n_blocks)] + [("head", self._ranges_of(m.n_blocks))]

    def _on_bucket_final(self, block_idx, ranges, on_side_stream):
        """Partial sum of squares of a finalised gradient bucket (overlapped with the rest of the backward)."""
        g = self.module._flat_g
        if on_side_stream and self.ddp is not None and self.ddp.sumsq_target is not None:
            pass                                 # the squares were summed by grad_unpack on the same stream
        elif on_side_stream:                     # ThinDDP's comm stream, already ordered after the bucket's all-reduce
            for lo, hi in ranges:
                ops.sumsq(g[lo:hi], self._sumsq, self.sumsq_ctas)
        else:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                for lo, hi in ranges:
                    ops.sumsq(g[lo:hi], self._sumsq, 0 if block_idx == -1 else self.sumsq_ctas)
        self._buckets_seen += 1

    # hyper-parameters live in param_groups[0] (what LR schedulers mutate)
    lr = property(lambda self: self.param_groups[0]["lr"], lambda self, v: self.param_groups[0].__setitem__("lr", v))
    betas = property(lambda self: self.param_groups[0]["betas"])
    eps = property(lambda self: self.param_groups[0]["eps"])
    weight_decay = property(lambda self: self.param_groups[0]["weight_decay"])

    def zero_grad(self, set_to_none: bool = True):
        self.module._force_fresh_grads = True     # next backward overwrites the GEMM grads and re-zeroes the rest

    def join(self):
        """Make the current stream wait for the enqueued parameter update."""
        if self._last_event is not None:
            torch.cuda.current_stream().wait_event(self._last_event)

    @property
    def last_grad_norm(self):
        self.join()
        return self._norm

    def _clip_scale(self, g, partial_ok):
        """Clip coefficient of torch clip_grad_norm_ from the total gradient norm (device scalar)."""
        if not partial_ok:
            self._sumsq.zero_()
            ops.sumsq(g, self._sumsq)
        if self.ddp is not None and self.ddp.world > 1:
            # every rank sums the SAME all-reduced gradient, but with float atomics in a different order: the last bits of the
            # norm (and with them the clip coefficient and, from then on, the replicas) could differ.  One 4-byte all-reduce
            # (MAX) pins the value, so parameters stay bit-identical across ranks like under torch DDP + clip_grad_norm_.
            dist.all_reduce(self._sumsq, op=dist.ReduceOp.MAX, group=self.ddp.pg)
        self._norm = self._sumsq.sqrt()
        torch.clamp(self.max_grad_norm / (self._norm + 1e-6), max=1.0, out=self._scale)
        return self._scale

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._step_impl()
        return loss

    def _step_impl(self):
        m = self.module
        p, g = m._flat_p, m._flat_g              # (not the properties: they would refresh the bf16 shadow we are about to rewrite)
        self.step_count += 1
        args = (self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count)
        if not self.overlap:
            scale = self._clip_scale(g, False) if self.max_grad_norm is not None else None
            ops.adamw_step(p, g, self.exp_avg, self.exp_avg_sq, m._flat_bf16, *args, grad_scale=scale)
            m.mark_weights_updated(shadow_is_current=True)
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(enable_timing=self.debug_timing)
        ev.record(main)                            # backward (and the DDP tail, which the main stream already waited on) done
        self._dbg_step_start = ev
        events = {}
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ev)
            scale = None
            if self.max_grad_norm is not None:
                # per-bucket partial sums are valid iff every bucket was finalised exactly once since the last step
                scale = self._clip_scale(g, self._buckets_seen == len(self._stages))
                self._sumsq.zero_()                # ready for the next backward's partial sums (ordered before every event below)
            self._buckets_seen = 0
            for k, (name, ranges) in enumerate(self._stages):
                for lo, hi in ranges:
                    ops.adamw_step(p[lo:hi], g[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], m._flat_bf16[lo:hi], *args,
                                   grad_scale=scale, max_ctas=0 if k < self.eager_buckets else self.side_ctas)
                e = torch.cuda.Event(enable_timing=self.debug_timing)
                e.record(self._stream)
                events[name] = e
            self._last_event = events["head"]
        self._dbg_events = events
        m._param_events = dict(events)             # DIT.forward waits per bucket; DIT.wait_param_events() for other readers
        m.mark_weights_updated(shadow_is_current=True)

    def state_dict(self):
        """Flat layout (step, exp_avg, exp_avg_sq over DIT._flat_p) + the param group's hyper-parameters.  `to_torch_state_dict`
        gives the per-parameter torch.optim.AdamW layout (reference optimizer checkpoints)."""
        self.join()                                # the moments may still be in flight on the optimizer stream
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, lr=self.lr, betas=self.betas,
                    eps=self.eps, weight_decay=self.weight_decay)

    def load_state_dict(self, sd):
        self.join()
        if "state" in sd and "param_groups" in sd:                 # a torch.optim.AdamW checkpoint (per-parameter state)
            params = self.param_groups[0]["params"]
            for idx, st in sd["state"].items():
                mine = self.state[params[int(idx)]]
                mine["exp_avg"].copy_(st["exp_avg"])
                mine["exp_avg_sq"].copy_(st["exp_avg_sq"])
                self.step_count = int(st["step"])
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in sd["param_groups"][0]:
                    self.param_groups[0][k] = sd["param_groups"][0][k]
            return
        self.step_count = sd["step"]
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in sd:
                self.param_groups[0][k] = sd[k]

    def to_torch_state_dict(self):
        """The same state in torch.optim.AdamW's state_dict layout (loadable by the reference's optimizer)."""
        self.join()
        params = self.param_groups[0]["params"]
        state = {i: dict(step=torch.tensor(float(self.step_count)), exp_avg=self.state[p]["exp_avg"].clone(),
                         exp_avg_sq=self.state[p]["exp_avg_sq"].clone()) for i, p in enumerate(params)}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(params)))
        return dict(state=state, param_groups=[group])
