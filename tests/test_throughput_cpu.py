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
