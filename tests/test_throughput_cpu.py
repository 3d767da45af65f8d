"""CPU: the throughput / MFU monitor — B200 table row, the reference's FLOPs estimate, rolling metrics."""
import time
from types import SimpleNamespace

import torch

from unidisc_b200 import throughput as T
from unidisc_b200.config import make_config


def test_b200_row_and_flops_formulas():
    assert T.available_flops("NVIDIA B200", torch.bfloat16) == 2.25e15          # dense bf16 (the reference table returns None here)
    assert T.available_flops("NVIDIA H100 80GB HBM3 SXM", torch.bfloat16) == 0.989e15
    assert T.available_flops("Some Unknown GPU", torch.bfloat16) is None
    cfg = make_config("extra_large")
    f = T.flops_per_sample(cfg, 48385)
    assert abs(f / 1280 - 8.597e9) < 2e6                                         # 8.597 GFLOP per token fwd+bwd
    assert T.flops_per_sample(cfg, 48385, non_embedding_params=1_200_000_000, exact=False) == 6.0 * 1280 * 1_200_000_000


def test_monitor_rolling_metrics():
    mon = T.ThroughputMonitor(flops_per_sample=1e9, world_size=2, log_every_n_steps=1, device=torch.device("cpu"))
    mon.available_flops = 1e12
    unit = SimpleNamespace(global_step=0, num_tokens_per_sample=100, step_batch_size=8, gradient_accumulation_steps=1)
    out = {}
    for s in range(1, 5):
        unit.global_step = s
        time.sleep(0.01)
        out = mon.on_train_step_end(unit)
    assert out["samples_per_sec"] > 0 and abs(out["items_per_sec"] / out["samples_per_sec"] - 100) < 1e-6
    assert abs(out["device/samples_per_sec"] * 2 - out["samples_per_sec"]) < 1e-9
    assert abs(out["device/mfu"] - out["samples_per_sec"] * 1e9 / 2 / 1e12) < 1e-12


def test_select_head_rows_host_logic():
    """unidisc_b200.model.select_head_rows: which token rows the output projection runs on (text rows first, split point)."""
    import torch
    from unidisc_b200.model import select_head_rows
    g = torch.Generator().manual_seed(0)
    n = 97
    is_img = torch.arange(n) >= 30
    sel = torch.rand(n, generator=g) < 0.4
    sel_t, sel_i = sel & ~is_img, sel & is_img
    n_t, n_i = int(sel_t.sum()), int(sel_i.sum())
    rows, split = select_head_rows(sel_t, sel_i, n_t, n_i, True)
    assert split == n_t and rows.dtype == torch.int64
    assert torch.equal(rows[:split], sel_t.nonzero().squeeze(1)) and torch.equal(rows[split:], sel_i.nonzero().squeeze(1))
    rows2, split2 = select_head_rows(sel_t, sel_i, n_t, n_i, False)                 # no per-modality vocabulary restriction
    assert split2 is None and torch.equal(rows2, sel.nonzero().squeeze(1))
    z = torch.zeros(n, dtype=torch.bool)
    assert select_head_rows(z, z, 0, 0, True) == (None, None)                        # nothing masked: plain head
    assert select_head_rows(~is_img, is_img, 30, n - 30, True) == (None, None)      # everything masked: plain head
    rows3, split3 = select_head_rows(z, sel_i, 0, n_i, True)                         # one modality only: rows, no split
    assert split3 is None and torch.equal(rows3, sel_i.nonzero().squeeze(1))
