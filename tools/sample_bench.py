#!/usr/bin/env python
"""BASELINE.json configs[3]: UniDisc-1.4B inference, 64-step absorbing denoising, seq_len 1280, batch 64, one B200.
Times `Diffusion._sample` end to end (all-mask prior -> 64 x [backbone forward -> fused SUBS-softmax / absorbing update]
-> noise-removal pass) with CUDA events, and the sampler kernel alone.  Prints one JSON line.
  python tools/sample_bench.py [--batch 64] [--steps 64] [--predictor ddpm_cache|ddpm|maskgit]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--predictor", default="ddpm_cache")
    ap.add_argument("--preset", default="extra_large")
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    from unidisc_b200.config import make_config
    from unidisc_b200.model import Diffusion
    dev = torch.device("cuda", 0)
    txt, img = 256, 1024
    cfg = make_config(a.preset, txt_length=txt, img_length=img, predictor=a.predictor, sampling_steps=a.steps)
    torch.manual_seed(0)
    model = Diffusion(cfg, device=dev)
    model.eval()
    B, N = a.batch, txt + img
    modality = torch.cat([torch.zeros(B, txt, dtype=torch.int64), torch.ones(B, img, dtype=torch.int64)], 1).to(dev)
    times = []
    for r in range(a.repeat + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x, nfe = model._sample(num_steps=a.steps, batch_size_per_gpu=B, sample_modality=modality, return_nfe=True)
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            times.append(e0.elapsed_time(e1))
    assert x.shape == (B, N) and int((x == model.mask_index).sum()) == 0
    assert bool((x[:, :txt] < model.text_vocab_size).all()) and bool((x[:, txt:] >= model.text_vocab_size).all()), \
        "tokens left their modality's vocabulary"
    ms = min(times)
    fwd_flops = B * N * (cfg.model.n_blocks * (24 * cfg.model.hidden_size ** 2 + 4 * N * cfg.model.hidden_size) + 2 * cfg.model.hidden_size * model.vocab_size)
    print(json.dumps(dict(workload=f"{a.preset} 64-step absorbing sampling", predictor=a.predictor, batch=B, seq_len=N, steps=a.steps, nfe=nfe,
                          ms_total=ms, ms_per_step=ms / (nfe + 1), samples_per_s=B / (ms * 1e-3), tokens_per_s=B * N / (ms * 1e-3),
                          backbone_tflops=(nfe + 1) * fwd_flops / (ms * 1e-3) / 1e12, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)))


if __name__ == "__main__":
    main()
