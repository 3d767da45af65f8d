"""Throughput / MFU monitor and a host-sync-free training step for the B200 path.

Reference: `unidisc/utils/throughput_monitor.py` — `ThroughputMonitor.on_train_step_end` (:255-302) reads
`unit.global_step / num_tokens_per_sample / step_batch_size / gradient_accumulation_steps`, calls `torch.cuda.synchronize()` and
divides rolling FLOPs by `get_available_flops(device, dtype)` (:549-633), whose table ends at H100 — on a B200 it returns
None and the MFU is silently dropped.  The reference's train loop additionally syncs the host every micro-batch
(`loss.detach().cpu().item()` model.py:1491, `torch.isfinite(loss).all()` model.py:1496) and gathers a Python object per step
(model_setup.py:996).

Here:
  * `available_flops("B200", dtype)`: the missing table row (dense, no 2:4 sparsity) + the MEASURED cuBLAS peak of this pool when
    /root/repo/MEASURED_PEAKS.json is present;
  * `flops_per_sample`: the reference's estimate `6 * model.length * non_embedding_params` (model_setup.py:821-823) and the exact
    count used by bench.py (`L (24 D^2 + 4 N D) + 2 D V`, x3 for training);
  * `ThroughputMonitor`: same metric names as the reference callback (`items_per_sec`, `device/mfu`, ...), timed with CUDA events
    recorded on the training stream — the host is only synchronised when a value is actually read (every `log_every_n_steps`);
  * `TrainStep`: one optimizer step around `Diffusion.compute_loss` without any per-micro-batch host sync; the finite-loss guard
    of model.py:1496 becomes a device-side flag that zeroes the update (`check_finite="device"`) or is skipped (`None`).
"""
from __future__ import annotations

import json
import os
import time
from collections import deque
from typing import Optional

import torch

# dense tensor-core peaks per GPU in FLOP/s (NVIDIA's headline numbers are with 2:4 sparsity: these are the dense halves)
_PEAK_FLOPS = {
    "b200": {torch.bfloat16: 2.25e15, torch.float16: 2.25e15, "fp8": 4.5e15, "fp4": 9e15, torch.float32: 1.1e15},   # tf32 for fp32
    "h100 sxm": {torch.bfloat16: 0.989e15, torch.float16: 0.989e15, torch.float32: 0.494e15},
    "h100 pcie": {torch.bfloat16: 0.756e15, torch.float16: 0.756e15, torch.float32: 0.378e15},
    "a100": {torch.bfloat16: 0.312e15, torch.float16: 0.312e15, torch.float32: 0.156e15},
    "l40s": {torch.bfloat16: 0.362e15, torch.float16: 0.362e15, torch.float32: 0.183e15},
}


def available_flops(device_name: str, dtype=torch.bfloat16, measured: bool = False) -> Optional[float]:
    """Peak FLOP/s of one GPU (reference throughput_monitor.py:549-633, plus the B200 row it lacks).  measured=True returns
    this pool's measured cuBLAS bf16 peak (MEASURED_PEAKS.json, sustained figure) when the file is present."""
    name = device_name.lower()
    if measured and "b200" in name:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        if os.path.exists(p):
            j = json.load(open(p))
            return float(j.get("bf16_tflops_sustained", j["bf16_tflops"])) * 1e12
    for key, row in _PEAK_FLOPS.items():
        if all(tok in name for tok in key.split()):
            return row.get(dtype)
    return None


def flops_per_sample(config, vocab_size: int, non_embedding_params: Optional[int] = None, exact: bool = True) -> float:
    """training FLOPs of one sample (sequence).  exact=False: the reference's `6 * length * non_embedding_params`."""
    m = config.model
    N, D, L = m.length, m.hidden_size, m.n_blocks
    if not exact:
        if non_embedding_params is None:
            raise ValueError("the reference estimate needs non_embedding_params")
        return 6.0 * N * non_embedding_params
    return 3.0 * N * (L * (24 * D * D + 4 * N * D) + 2 * D * vocab_size)


class ThroughputMonitor:
    """Rolling tokens/s, samples/s and MFU.  `on_train_step_end(unit)` follows the reference callback's contract
    (`unit.global_step`, `.num_tokens_per_sample`, `.step_batch_size`, `.gradient_accumulation_steps`)."""

    def __init__(self, flops_per_sample: Optional[float] = None, world_size: int = 1, log_every_n_steps: int = 50, window_size: int = 10,
                 device: Optional[torch.device] = None, dtype=torch.bfloat16, measured_peak: bool = False):
        self.flops_per_sample = flops_per_sample
        self.world_size = world_size
        self.log_every_n_steps = log_every_n_steps
        self.device = device
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        name = torch.cuda.get_device_name(device) if self.cuda else "cpu"
        self.available_flops = available_flops(name, dtype, measured=measured_peak)
        self._marks = deque(maxlen=window_size + 1)     # (event or perf_counter, global_step, samples, tokens)

    def _mark(self):
        if self.cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.device))      # no host sync: the event is read later
            return e
        return time.perf_counter()

    def on_train_step_end(self, unit, state=None):
        step = unit.global_step
        if step % self.log_every_n_steps != 0:
            return None
        bs = unit.step_batch_size
        self._marks.append((self._mark(), step, step * bs, step * bs * unit.num_tokens_per_sample))
        return self.compute()

    def compute(self):
        if len(self._marks) < 2:
            return {}
        (m0, s0, n0, t0), (m1, s1, n1, t1) = self._marks[0], self._marks[-1]
        if self.cuda:
            m1.synchronize()                                      # the only host sync, once per log interval
            elapsed = m0.elapsed_time(m1) * 1e-3
        else:
            elapsed = m1 - m0
        if elapsed <= 0:
            return {}
        w = self.world_size
        out = {"batches_per_sec": (s1 - s0) / elapsed, "samples_per_sec": (n1 - n0) / elapsed, "items_per_sec": (t1 - t0) / elapsed,
               "device/samples_per_sec": (n1 - n0) / elapsed / w, "device/items_per_sec": (t1 - t0) / elapsed / w, "time": elapsed}
        if self.flops_per_sample is not None:
            fps = self.flops_per_sample * (n1 - n0) / elapsed
            out["flops_per_sec"] = fps
            out["device/flops_per_sec"] = fps / w
            if self.available_flops:
                out["device/mfu"] = fps / w / self.available_flops
        return out


class TrainStep:
    """One optimizer step (gradient accumulation included) around `Diffusion.compute_loss` with NO per-micro-batch host sync.

    reference model.py:1400-1540: `loss.detach().cpu().item()` (:1491) and `torch.isfinite(loss).all()` (:1496) stall the host
    every micro-batch.  Here the loss stays on the device (read it when you log), and the finite guard is a device-side select:
    a non-finite loss contributes a zero gradient instead of being skipped by a host branch."""

    def __init__(self, model, optimizer, ddp=None, gradient_accumulation_steps: int = 1, check_finite: Optional[str] = "device",
                 monitor: Optional[ThroughputMonitor] = None):
        self.model, self.optimizer, self.ddp = model, optimizer, ddp
        self.gradient_accumulation_steps = gradient_accumulation_steps
        self.check_finite = check_finite
        self.monitor = monitor
        self.global_step = 0
        self.num_tokens_per_sample = model.config.model.length
        self.step_batch_size = None
        self.last_loss = None                       # device tensor; float(...) it only when logging

    def __call__(self, micro_batches):
        """micro_batches: a batch dict, or a list of `gradient_accumulation_steps` of them."""
        if isinstance(micro_batches, dict):
            micro_batches = [micro_batches]
        n = len(micro_batches)
        total = None
        for i, batch in enumerate(micro_batches):
            ctx = self.ddp.no_sync() if (self.ddp is not None and i + 1 < n) else _null()
            with ctx:
                loss = self.model.compute_loss(batch).loss / n
                if self.check_finite == "device":
                    loss = torch.where(torch.isfinite(loss.detach()), loss, loss.detach() * 0)    # no host branch
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        self.optimizer.step()
        self.optimizer.zero_grad()
        self.global_step += 1
        self.step_batch_size = sum(b["input_ids"].shape[0] for b in micro_batches) * (self.ddp.world if self.ddp is not None else 1)
        self.last_loss = total
        return self.monitor.on_train_step_end(self) if self.monitor is not None else None


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
