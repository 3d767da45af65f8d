#!/usr/bin/env python
"""bench.py — joint-token tokens/sec of one full UniDisc training step (q_xt -> DiT fwd -> SUBS-NLL -> bwd ->
bf16 gradient all-reduce -> clip + AdamW) on synthetic seq_len=1280 (256 text + 1024 image) batches.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (reference arm: the CPU restatement of the reference path)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of value / e2e / roofline / cpu_baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (preset, per-GPU batch, txt, img)   — BASELINE.json configs[2] / configs[1]
    "unidisc-1.4B": ("extra_large", 8, 256, 1024),
    "dit-b": ("small", 32, 256, 1024),
    "tiny": ("small", 2, 64, 64),     # CI-sized sanity run, never a bench line
    # BASELINE.json configs[4]: packed / interleaved documents, seq_len 4096, bs 2 per GPU, document-masked attention
    "unidisc-1.4B-interleaved": ("extra_large", 2, 3072, 1024),
    # BASELINE.json configs[3]: 64-step absorbing denoising, batch 64 (a "step" = one whole sampling run of the batch)
    "unidisc-1.4B-sample": ("extra_large", 64, 256, 1024),
}
TEXT_VOCAB, IMAGE_VOCAB = 32001, 16384


def flops_per_token_fwd(D, L, N, V):
    return L * (24 * D * D + 4 * N * D) + 2 * D * V


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(source="measured", hbm_gbs=j["hbm_gbs"], bf16_tflops=j["bf16_tflops"],
                    bf16_tflops_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]))
    return dict(source="fallback", hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(n)
        mx = None
        for r in self.rows:
            try:
                mx = float(r[1])
                break
            except Exception:
                pass
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (q_xt -> DIT -> SUBS -> weighted NLL, fwd+bwd), host cores
# --------------------------------------------------------------------------------------------------------------
CPU_THREADS = 16          # fixed, so the CPU figure is comparable between boxes (hosts of this pool have 16-32 cores)


def cpu_reference_forward(workload):
    """Forward pass of the UNMODIFIED reference backbone (`models.dit.DIT` staged in baseline/_ref by
    baseline/install_reference.py) on the host cores: real depth, B=1, the reference's CPU default (bf16 autocast,
    SURVEY.md section 0 row 10), eval mode.  The reference's eager BACKWARD raises on CPU (in-place q/k-norm write), so the
    training-step figure comes from the port (cpu_reference_tokens_per_sec) and this is reported beside it."""
    from oracle import ref_loader as RL
    from unidisc_b200.config import MODEL_PRESETS
    if not RL.reference_available():
        return dict(value=None, unit="tokens/s (forward only)", sample=f"reference backbone not staged at {RL.REFERENCE_ROOT}")
    preset, _, txt, img = WORKLOADS[workload]
    D, L, H = MODEL_PRESETS[preset]
    V, tv, mi = TEXT_VOCAB + IMAGE_VOCAB, TEXT_VOCAB, TEXT_VOCAB - 1
    N = txt + img
    torch.set_num_threads(min(CPU_THREADS, os.cpu_count() or 1))
    rcfg = RL.make_ref_config(D, H, L, txt, img, dropout=0.0)
    model = RL.build_reference_dit(rcfg, V, tv, mi, dtype=None)
    model.eval()
    g = torch.Generator().manual_seed(42)
    ids = torch.cat([torch.randint(0, tv - 1, (1, txt), generator=g), torch.randint(tv, V, (1, img), generator=g)], 1)
    mod = torch.cat([torch.zeros(1, txt, dtype=torch.int64), torch.ones(1, img, dtype=torch.int64)], 1)
    best = None
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        for _ in range(2):
            t0 = time.perf_counter()
            model(ids, None, modality=mod)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    del model
    return dict(value=N / best, unit="tokens/s (forward only)", cores=torch.get_num_threads(), kind="reference",
                sample=f"unmodified models.dit.DIT ({RL.REFERENCE_ROOT}), depth {L}, B=1 N={N} V={V}, CPU bf16 autocast, eval, best of 2 ({best:.2f}s)")


def cpu_reference_tokens_per_sec(workload, budget_s=25.0, with_reference_forward=True):
    """Times the oracle restatement (oracle/restated.py, eager torch fp32 on the host cores) of the reference training
    path on a bounded sample: per-GPU batch 1 at the workload's D/N/V, depth 1 and 3 blocks, fwd+bwd; the per-block and
    the embed+head+loss costs are separated and extrapolated linearly to the workload's depth."""
    from oracle import restated as R

    preset, _, txt, img = WORKLOADS[workload]
    from unidisc_b200.config import MODEL_PRESETS
    D, L, H = MODEL_PRESETS[preset]
    cores = min(CPU_THREADS, os.cpu_count() or 1)
    torch.set_num_threads(cores)
    V, tv, mi = TEXT_VOCAB + IMAGE_VOCAB, TEXT_VOCAB, TEXT_VOCAB - 1
    N = txt + img
    ids, modality = R.synthetic_batch(1, txt, img, tv, V, seed=42)
    am = torch.ones_like(ids, dtype=torch.bool)

    def run(depth):
        cfg = R.OracleConfig(D, H, depth, txt, img, V, tv, mi)
        P = {k: v.requires_grad_(True) for k, v in R.init_params(cfg, seed=0).items()}
        g = torch.Generator().manual_seed(1)
        best = None
        for it in range(2):
            t0 = time.perf_counter()
            out = R.training_loss(cfg, P, ids, modality, am, torch.rand(1, generator=g), torch.rand(1, N, generator=g), mode="fp32")
            out["loss"].backward()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            for p in P.values():
                p.grad = None
        return best

    t1 = run(1)
    t3 = run(3)
    per_block = max((t3 - t1) / 2.0, 1e-6)
    fixed = max(t1 - per_block, 0.0)
    t_full = fixed + L * per_block
    out = dict(value=N / t_full, unit="tokens/s", cores=cores, kind="port",
               sample=f"oracle/restated.py eager-torch fp32 fwd+bwd, B=1 N={N} D={D} V={V}; measured depth 1 ({t1:.2f}s) and 3 ({t3:.2f}s), "
                      f"extrapolated linearly to depth {L} ({t_full:.1f}s/step); no optimizer step on the CPU arm; {cores} threads (fixed)")
    if with_reference_forward:
        try:
            out["reference_forward"] = cpu_reference_forward(workload)
        except Exception as e:  # noqa: BLE001
            out["reference_forward"] = dict(value=None, sample=f"failed: {type(e).__name__}: {e}")
    return out


def run_reference_gpu_arm(args, rank, world, local):
    """`--impl reference --ref-device cuda`: the reference's own torch path on the GPU, the denominator of BASELINE.json's
    ">= 1.5x the reference's torch-sdpa GPU path" target (the driver's reference arm stays the CPU one).

    --ref-kind stock (default): the UNMODIFIED `models.dit.DIT` staged under baseline/_ref/ (baseline/install_reference.py),
      instantiated like model_setup.py:149-161 (dtype=None, fp32 parameters) and run under the outer
      `torch.autocast(bf16)` of model.py:693-696 with UNIDISC_FORCE_CUDNN_SPDA_CONTEXT=1 (cuDNN SDPA, dit.py:816-829),
      torch.optim.AdamW(fused) + clip_grad_norm_(1.0) (model_setup.py:385-424, model.py:1518), torch DDP with
      gradient_as_bucket_view / static_graph and the BF16 compress hook for N > 1 (main.py:641-656), optionally
      torch.compile(mode=max-autotune-no-cudagraphs) with the reference's inductor settings (utils.py:502-527,
      configs/config.yaml:227).  The loss around it (q_xt, SUBS on the materialised [B,N,V] tensor, weighted NLL) is the
      library-op restatement of model.py::compute_loss in oracle/torch_eager.py (model.py itself needs accelerate / hydra /
      tensordict, absent from this image).
    --ref-kind restated: oracle/torch_eager.py's restatement of the backbone as well (round-1 cross-check)."""
    if args.ref_sdpa == "cudnn":
        os.environ["UNIDISC_FORCE_CUDNN_SPDA_CONTEXT"] = "1"          # read at import time by models/dit.py:21
    import torch.distributed as dist
    from oracle import restated as R
    from oracle import torch_eager as TE
    from unidisc_b200.config import MODEL_PRESETS

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    preset, bpg, txt, img = WORKLOADS[args.workload]
    D, L, H = MODEL_PRESETS[preset]
    V, tv, mi = TEXT_VOCAB + IMAGE_VOCAB, TEXT_VOCAB, TEXT_VOCAB - 1
    N, B = txt + img, (args.batch or bpg)
    torch.manual_seed(0)
    # reference utils.py:425-438 (called from main.py:550).  It also sets the legacy `cuda.matmul.allow_tf32 = True`; on this
    # image's torch 2.11 mixing the legacy and the new matmul-precision API makes inductor's max-autotune raise, and "medium"
    # already implies TF32 matmuls, so only the new API is used.
    torch.set_float32_matmul_precision("medium")
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    if args.ref_kind == "stock":
        from oracle import ref_loader as RL
        rcfg = RL.make_ref_config(D, H, L, txt, img, dropout=args.dropout)
        rcfg.model.force_optimized_native_attn = args.ref_sdpa == "cudnn"
        model = RL.build_reference_dit(rcfg, V, tv, mi, dtype=None).to(dev)
        src = f"unmodified models.dit.DIT from {RL.REFERENCE_ROOT}"
        call_backbone = lambda net, xt, modality: net(xt, None, modality=modality)
    else:
        ocfg = R.OracleConfig(D, H, L, txt, img, V, tv, mi)
        model = TE.EagerDIT(ocfg, dropout=args.dropout).to(dev)
        src = "oracle/torch_eager.py (restated backbone)"
        call_backbone = lambda net, xt, modality: net(xt, modality)
    model.train()
    net = model
    if world > 1:
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True, static_graph=True)
        net.register_comm_hook(None, default_hooks.bf16_compress_hook)
    fwd = net
    if args.ref_compile:
        if args.ref_sd3_config:                                       # reference utils.py:511-516 (trainer.sd3_compile_config, default on)
            import torch._inductor.config as ic
            ic.conv_1x1_as_mm = True
            ic.coordinate_descent_tuning = True
            ic.epilogue_fusion = False
            ic.coordinate_descent_check_all_directions = True
        fwd = torch.compile(net, mode=args.ref_compile_mode)          # reference utils.py:524: the (DDP-wrapped) backbone is compiled
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, fused=True)
    g = torch.Generator().manual_seed(42 + rank)
    ids_h = torch.cat([torch.randint(0, tv - 1, (B, txt), generator=g), torch.randint(tv, V, (B, img), generator=g)], 1).pin_memory()
    mod_h = torch.cat([torch.zeros(B, txt, dtype=torch.int64), torch.ones(B, img, dtype=torch.int64)], 1).pin_memory()
    am_h = torch.ones(B, N, dtype=torch.bool).pin_memory()

    class _W(torch.nn.Module):                     # lets DDP / compile wrap the backbone while the loss code calls model(xt, modality)
        def forward(self, xt, modality):
            return call_backbone(fwd, xt, modality)

    def step():
        ids, mod, am = ids_h.to(dev, non_blocking=True), mod_h.to(dev, non_blocking=True), am_h.to(dev, non_blocking=True)
        loss = TE.reference_style_loss(_W(), ids, mod, am, mi, tv)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss.detach())

    t_w0 = time.perf_counter()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    warm_s = time.perf_counter() - t_w0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        last = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        v = world * B * N / (ms.item() / args.steps * 1e-3)
        kind = "eager" + (f" + torch.compile({args.ref_compile_mode}{', sd3 inductor config' if args.ref_sd3_config else ''})" if args.ref_compile else "")
        print(json.dumps(dict(
            impl="reference", device="cuda", metric="joint_token_tokens_per_sec", value=v, unit="tokens/s", n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms.item() / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
            data="synthetic", config=dict(workload=args.workload, seq_len=N, per_gpu_batch=B, global_batch=world * B, vocab=V,
                                          parallelism=f"dp{world}", optimizer="torch AdamW(fused)+clip_grad_norm_(1.0)", dropout=args.dropout,
                                          backbone=src, mode=kind, sdpa=args.ref_sdpa,
                                          ddp="torch DDP + bf16_compress_hook, static_graph" if world > 1 else None),
            e2e=dict(value=v, unit="tokens/s", h2d_bytes_per_step=ids_h.numel() * 16 + am_h.numel(), d2h_bytes_per_step=4, loss=last),
            gpu_launches=0, clocks=clocks, peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30, warmup_wall_s=warm_s)))
    if world > 1:
        dist.destroy_process_group()


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    vals = []
    for i in range(max(1, min(args.steps, 2))):
        vals.append(cpu_reference_tokens_per_sec(args.workload, with_reference_forward=(i == 0)))
    cb = dict(vals[-1], reference_forward=vals[0].get("reference_forward"))
    v = max(x["value"] for x in vals)
    preset, bpg, txt, img = WORKLOADS[args.workload]
    line = dict(impl="reference", metric="joint_token_tokens_per_sec", value=v, unit="tokens/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=None, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp32",
                data="synthetic", config=dict(workload=args.workload, seq_len=txt + img, per_gpu_batch=1, note="CPU arm, host cores"),
                cpu_baseline=dict(cb, value=v), e2e=dict(value=v, unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: UniDisc-1.4B inference, 64-step absorbing denoising, seq_len 1280, bs 64, one B200
# --------------------------------------------------------------------------------------------------------------
def run_sampling_workload(args, rank, world, local):
    """A "step" is one complete `Diffusion._sample` call on a batch of 64 all-mask sequences: 64 x [backbone forward ->
    fused vocabulary pass (SUBS softmax -> absorbing update / MaskGIT draw + selection)] + the noise-removal forward.
    N > 1 = independent replicas (no collective).  value = sampled tokens per second over all ranks."""
    import torch.distributed as dist
    from unidisc_b200 import _lib as Lb
    from unidisc_b200 import ops
    from unidisc_b200.config import make_config
    from unidisc_b200.model import Diffusion

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    preset, B, txt, img = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    cfg = make_config(preset, txt_length=txt, img_length=img, predictor=args.predictor, sampling_steps=args.sampling_steps, seed=42 + rank)
    torch.manual_seed(0)
    model = Diffusion(cfg, device=dev)
    model.eval()
    N = txt + img
    D, Lyr, V = cfg.model.hidden_size, cfg.model.n_blocks, model.vocab_size
    mod_h = torch.cat([torch.zeros(B, txt, dtype=torch.int64), torch.ones(B, img, dtype=torch.int64)], 1).pin_memory()
    mod_d = mod_h.to(dev)
    out_h = torch.empty(B, N, dtype=torch.int64).pin_memory()
    launches = [0]
    orig_call = Lb.call

    def counting_call(name, *a):
        launches[0] += 1
        return orig_call(name, *a)

    ops.call = counting_call

    def run(e2e):
        modality = mod_h.to(dev, non_blocking=True) if e2e else mod_d
        x, nfe = model._sample(num_steps=args.sampling_steps, batch_size_per_gpu=B, sample_modality=modality, return_nfe=True)
        if e2e:
            out_h.copy_(x, non_blocking=True)            # device -> host read of the sampled tokens
        return x, nfe

    def timed(e2e, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches[0] = 0
        e0.record()
        for _ in range(steps):
            x, nfe = run(e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), x, nfe

    for _ in range(max(1, args.warmup)):
        run(True)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, x, nfe = timed(False, args.steps)
    n_launch = launches[0]
    ms_e2e, x, nfe = timed(True, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    assert x.shape == (B, N) and int((x == model.mask_index).sum()) == 0, "sampling left masked tokens"
    assert bool((x[:, :txt] < model.text_vocab_size).all()) and bool((x[:, txt:] >= model.text_vocab_size).all()), \
        "tokens left their modality's vocabulary"
    # ---- the vocabulary-pass kernel alone (HBM roofline): this predictor's kernel on a resident [B*N, Vp] bf16 logits buffer ----
    pk = peaks()
    tv = model.text_vocab_size
    lg = torch.empty(B * N, model.backbone.Vp, device=dev, dtype=torch.bfloat16)
    for c in range(0, B * N, 8192):
        lg[c:c + 8192].normal_(0, 3)
    xs = torch.full((B, N), model.mask_index, dtype=torch.int64, device=dev)
    tt = torch.full((B,), 0.7, device=dev)
    num = torch.full((B,), 20, dtype=torch.int32, device=dev)

    def vocab_pass(i):
        if args.predictor == "maskgit":
            ops.maskgit_update(xs, lg, mod_d.view(-1), tt, num, model.mask_index, tv, V, r_temp=10.0, seed=1, offset=i)
        else:
            ops.ddpm_update_logits(xs, lg, mod_d.view(-1), tt, tt - 0.01, model.mask_index, tv, V, seed=1, offset=i)

    for i in range(3):
        vocab_pass(i)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(8):
        vocab_pass(3 + i)
    s1.record()
    torch.cuda.synchronize()
    sms = s0.elapsed_time(s1) / 8
    alg_bytes = B * (txt * tv + img * (V - tv)) * 2 + B * N * 8 * 2
    del lg
    fwd_flops = B * N * flops_per_token_fwd(D, Lyr, N, V)
    per_batch_ms = ms_dev / args.steps
    value = world * B * N / (per_batch_ms * 1e-3)
    if rank == 0:
        line = dict(
            metric="sampled_tokens_per_sec", value=value, unit="tokens/s", n_gpus=world, steps=args.steps, warmup=max(1, args.warmup),
            ms_per_step=per_batch_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
            config=dict(workload=args.workload, model=f"DiT D={D} L={Lyr} H={cfg.model.n_heads}", seq_len=N, per_gpu_batch=B, vocab=V,
                        predictor=args.predictor, denoising_steps=args.sampling_steps, nfe=nfe, noise_removal=True,
                        parallelism=f"replicas x{world}", step="one 64-step sampling run of the whole batch (warm-up in the same unit)",
                        l2="logits 7.9 GB and activations per forward >> 126 MB L2 (no flush needed)"),
            e2e=dict(value=world * B * N / (ms_e2e / args.steps * 1e-3), unit="tokens/s", ms_per_step=ms_e2e / args.steps,
                     h2d_bytes_per_step=mod_h.numel() * 8, d2h_bytes_per_step=out_h.numel() * 8),
            gpu_launches=n_launch // max(args.steps, 1), clocks=clocks,
            ms_per_denoising_step=per_batch_ms / (nfe + 1), samples_per_s=world * B / (per_batch_ms * 1e-3),
            backbone_tflops_per_gpu=(nfe + 1) * fwd_flops / (per_batch_ms * 1e-3) / 1e12,
            backbone_frac_of_sustained_peak=(nfe + 1) * fwd_flops / (per_batch_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
            roofline=dict(bound="hbm", kernel=("maskgit draw + selection" if args.predictor == "maskgit" else "ddpm_update_logits_fast_kernel")
                          + " (vocabulary pass, Philox noise, all rows masked)", achieved=alg_bytes / (sms * 1e-3) / 1e9, peak=pk["hbm_gbs"],
                          unit="GB/s", frac=alg_bytes / (sms * 1e-3) / 1e9 / pk["hbm_gbs"], peak_source=pk["source"], traffic=None,
                          algorithmic_bytes=alg_bytes, ms_per_launch=sms),
            peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, cpu_baseline=None)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="unidisc-1.4B", choices=list(WORKLOADS))
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="reference arm: host cores (driver contract) or the "
                    "torch-eager restatement on the GPU (the >=1.5x target's denominator)")
    ap.add_argument("--ref-compile", action="store_true", help="with --ref-device cuda: torch.compile the backbone like the reference")
    ap.add_argument("--ref-compile-mode", default="max-autotune-no-cudagraphs")
    ap.add_argument("--ref-kind", default="stock", choices=["stock", "restated"], help="with --ref-device cuda: the unmodified "
                    "models.dit.DIT staged in baseline/_ref (default) or the oracle/torch_eager.py restatement")
    ap.add_argument("--ref-sdpa", default="cudnn", choices=["cudnn", "default"], help="UNIDISC_FORCE_CUDNN_SPDA_CONTEXT=1 (reference "
                    "setting for optimized native attention) or torch's default SDPA backend choice")
    ap.add_argument("--ref-sd3-config", type=int, default=1, help="with --ref-compile: the reference's inductor settings (utils.py:511-516)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1, help="model.dropout (reference configs/model/extra_large.yaml:9 = 0.1)")
    ap.add_argument("--predictor", default="ddpm_cache", choices=["ddpm_cache", "ddpm", "maskgit"], help="sampling workload only")
    ap.add_argument("--sampling-steps", type=int, default=64, help="sampling workload only")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's per-GPU batch")
    ap.add_argument("--no-split-head", action="store_true", help="training workloads: trainer.b200_split_head=false (masked rows x the "
                    "whole vocabulary instead of text rows x text vocabulary + image rows x image vocabulary)")
    ap.add_argument("--full-head", action="store_true", help="training workloads: trainer.b200_masked_head=false, i.e. project ALL "
                    "token rows onto the vocabulary like the reference (A/B; the default projects only the masked rows the loss reads)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.ref_device == "cuda":
            run_reference_gpu_arm(args, rank, world, local)
        else:
            run_reference_arm(args, rank, world)
        return
    if args.workload == "unidisc-1.4B-sample":
        run_sampling_workload(args, rank, world, local)
        return
    assert args.warmup >= 3 or args.workload == "tiny", "timing rules: W >= 3"

    import torch.distributed as dist
    from unidisc_b200 import _lib as Lb
    from unidisc_b200 import ops
    from unidisc_b200.config import make_config
    from unidisc_b200.ddp import FusedAdamW, ThinDDP
    from unidisc_b200.model import Diffusion

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from unidisc_b200.ddp import nccl_options
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev, pg_options=nccl_options())
    preset, bpg, txt, img = WORKLOADS[args.workload]
    if args.batch:
        bpg = args.batch
    small = args.workload == "tiny"
    packed = args.workload.endswith("-interleaved")
    extra = dict(hidden_size=256, n_blocks=2, n_heads=4) if small else {}
    if packed:       # reference configs/experiments/*interleaved*: sample ids from the packing collate + FlexAttention document mask
        extra.update(data__require_sample_ids=True, trainer__interleaved=True, trainer__interleaved_training_flex_attention=True)
    extra.update(trainer__b200_masked_head=not args.full_head, trainer__b200_split_head=not args.no_split_head)
    cfg = make_config(preset, txt_length=txt, img_length=img, dropout=args.dropout, image_vocab_size=IMAGE_VOCAB if not small else 255,
                      text_vocab_size=TEXT_VOCAB if not small else 257, **extra)
    torch.manual_seed(0)
    model = Diffusion(cfg, device=dev)
    model.train()
    net = model.backbone
    ddp = ThinDDP(net, bf16_compress=bool(int(os.environ.get('UD_DDP_COMPRESS', '1')))) if world > 1 else None
    def make_opt(overlap=None):
        return FusedAdamW(ddp if ddp is not None else net, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0,
                          overlap=overlap)

    opt = make_opt()
    assert world == 1 or net.grad_ready_hook.__self__ is ddp, "the gradient all-reduce hook must stay installed"
    B, N = bpg, txt + img
    V, tv = model.vocab_size, model.text_vocab_size
    D, Lyr = cfg.model.hidden_size, cfg.model.n_blocks

    g = torch.Generator().manual_seed(42 + rank)                      # mirrors reference main.py:1062 (seed + rank)
    sid_h = sid_d = None
    attn_pairs = float(N)                                             # mean number of keys a query attends to (dense: N)
    if packed:
        from unidisc_b200.synth import packed_batch
        ids_h, mod_h, sid_h, am_h, doc_lens = packed_batch(B, N, tv, V, seed=42 + rank)
        ids_h, mod_h, sid_h, am_h = ids_h.pin_memory(), mod_h.pin_memory(), sid_h.pin_memory(), am_h.pin_memory()
        sid_d = sid_h.to(dev)
        attn_pairs = sum(x * x for lens in doc_lens for x in lens) / float(B * N)     # cfg5: 4 N D -> 4 D sum(len_i^2) / N
    else:
        ids_h = torch.cat([torch.randint(0, tv - 1, (B, txt), generator=g), torch.randint(tv, V, (B, img), generator=g)], 1).pin_memory()
        mod_h = torch.cat([torch.zeros(B, txt, dtype=torch.int64), torch.ones(B, img, dtype=torch.int64)], 1).pin_memory()
        am_h = torch.ones(B, N, dtype=torch.bool).pin_memory()
    ids_d, mod_d, am_d = ids_h.to(dev), mod_h.to(dev), am_h.to(dev)
    h2d = ids_h.numel() * 8 + mod_h.numel() * 8 + am_h.numel() + (sid_h.numel() * 8 if packed else 0)

    launches = [0]
    orig_call = Lb.call

    def counting_call(name, *a):
        launches[0] += 1
        return orig_call(name, *a)

    ops.call = counting_call
    import unidisc_b200.ops as _o
    _o.call = counting_call

    head_rows_seen = []          # token rows the output projection ran on, per step (B*N with --full-head)

    def step(e2e):
        if e2e:
            batch = dict(input_ids=ids_h.to(dev, non_blocking=True), modality=mod_h.to(dev, non_blocking=True),
                         attention_mask=am_h.to(dev, non_blocking=True))
            if packed:
                batch["sample_ids"] = sid_h.to(dev, non_blocking=True)
        else:
            batch = dict(input_ids=ids_d, modality=mod_d, attention_mask=am_d)
            if packed:
                batch["sample_ids"] = sid_d
        losses = model.compute_loss(batch)
        head_rows_seen.append((model._last_head_rows if model._last_head_rows is not None else B * N, model._last_head_split))
        losses.loss.backward()
        opt.step()
        opt.zero_grad()
        if e2e:
            return float(losses.loss.detach())          # device -> host read of the step's result
        return losses.loss.detach()

    def timed(e2e, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches[0] = 0
        e0.record()
        last = None
        for _ in range(steps):
            last = step(e2e)
        opt.join()                      # the streamed optimizer's last update belongs to the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), float(last)

    for _ in range(args.warmup):          # (a failure here fails the run: no silent change of the optimizer schedule)
        step(True)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, loss_dev = timed(False, args.steps)
    n_launch = launches[0]
    ms_e2e, loss_e2e = timed(True, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # The same step with the output projection on EVERY token row (the reference's dataflow), measured in the same process right
    # after the product's default: the line carries both, so the sparse head's share of the result is visible in one run.
    full_head_info = None
    if not args.full_head and not small:
        n_seen = len(head_rows_seen)
        model.config.trainer["b200_masked_head"] = False
        for _ in range(2):
            step(False)
        ms_full, _ = timed(False, args.steps)
        model.config.trainer["b200_masked_head"] = True
        del head_rows_seen[n_seen:]
        full_head_info = dict(value=world * B * N / (ms_full / args.steps * 1e-3), unit="tokens/s", ms_per_step=ms_full / args.steps,
                              note="trainer.b200_masked_head=false (logits for all B*N rows x all V columns, like the reference), same process")
    if os.environ.get("UD_PHASE_TIMING"):
        # debug: CUDA-event split of one step into forward+loss / backward / optimizer (rank 0, stderr)
        for _ in range(3):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            batch = dict(input_ids=ids_d, modality=mod_d, attention_mask=am_d)
            evs[0].record()
            losses = model.compute_loss(batch)
            evs[1].record()
            losses.loss.backward()
            evs[2].record()
            opt.step()
            opt.zero_grad()
            opt.join()
            evs[3].record()
            torch.cuda.synchronize()
            if rank == 0:
                print(f"[phase] fwd+loss {evs[0].elapsed_time(evs[1]):.2f} ms  bwd {evs[1].elapsed_time(evs[2]):.2f} ms  "
                      f"opt(+allreduce tail) {evs[2].elapsed_time(evs[3]):.2f} ms", file=sys.stderr)
    if os.environ.get("UD_KERNEL_TIMES") and rank == 0:
        # debug: in-situ (warm, real clocks) duration of every C-ABI launch of one step, CUDA events around each call
        for rep in range(2):
            recs = []

            def timing_call(name, *a):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = orig_call(name, *a)
                e1.record()
                tag = name
                if name == "ud_gemm_bf16":
                    tag = f"gemm ta{a[0]} tb{a[1]} {a[2]}x{a[3]}x{a[4]} epi{a[11]}"
                recs.append((tag, e0, e1))
                return r

            _o.call = timing_call
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s0.record()
            step(False)
            opt.join()
            s1.record()
            torch.cuda.synchronize()
            _o.call = counting_call
            if rep == 0:
                continue
            agg = {}
            for tag, e0, e1 in recs:
                c, t = agg.get(tag, (0, 0.0))
                agg[tag] = (c + 1, t + e0.elapsed_time(e1))
            tot = sum(t for _, t in agg.values())
            print(f"[kernel times] step {s0.elapsed_time(s1):.2f} ms, sum of C-ABI launches {tot:.2f} ms ({len(recs)} launches)", file=sys.stderr)
            for tag, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                print(f"[kernel times] {tag:48s} n={c:4d} total {t:8.3f} ms  avg {t / c * 1e3:8.1f} us", file=sys.stderr)
    # debug (UD_OPT_DEBUG=1): when each bucket's AdamW finished vs when the next forward reached the block that needs it
    if getattr(opt, "debug_timing", False) and opt.overlap and rank == 0:
        step(False)                                    # step A: its opt.step() records the bucket events ...
        t0, evA = opt._dbg_step_start, dict(opt._dbg_events)
        net._dbg_fwd_events = []
        step(False)                                    # ... that step B's forward waits on
        fwd_ev, net._dbg_fwd_events = net._dbg_fwd_events, None
        opt.join()
        torch.cuda.synchronize()
        print("[opt timeline] ms after backward end: bucket AdamW done | forward block start", file=sys.stderr)
        names = ["pre"] + list(range(net.n_blocks)) + ["head"]
        for k, nme in enumerate(names):
            f_ms = t0.elapsed_time(fwd_ev[nme]) if isinstance(nme, int) and nme < len(fwd_ev) else float("nan")
            print(f"[opt timeline] {str(nme):>5s}  adamw_done {t0.elapsed_time(evA[nme]):8.2f}   fwd_block_start {f_ms:8.2f}", file=sys.stderr)
    if ddp is not None and ddp.debug_timing:
        ddp._dbg_events = []
        step(False)
        rows = ddp.debug_report()
        if rank == 0:
            print("[ddp timeline] bucket ready_ms start_ms end_ms chain_ms", file=sys.stderr)
            for b_, r_, s_, e_ in rows:
                print(f"[ddp timeline] {b_:3d} {r_:8.2f} {s_:8.2f} {e_:8.2f} {e_ - s_:7.2f}", file=sys.stderr)

    # ---- roofline of the dominant kernel family: the tcgen05 GEMM exactly as the model launches it for the MLP up-projection
    #      (mlp.0: bias + GELU-tanh epilogue writing u and gelu(u), reference dit.py:917-919), timed live over rotated operands ----
    pk = peaks()
    M = B * N
    sets = []
    for i in range(4):                                                # rotate operand sets (> L2) between launches
        sets.append((torch.randn(M, D, device=dev).to(torch.bfloat16), torch.randn(4 * D, D, device=dev).to(torch.bfloat16),
                     torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16), torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16)))
    bias_r = torch.randn(4 * D, device=dev).to(torch.bfloat16)
    for a, b_, c, g_ in sets:
        ops.gemm(a, b_, out=c, epi=Lb.EPI_BF16_GELU, bias=bias_r, aux=g_)
    torch.cuda.synchronize()
    iters = 24
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(iters):
        a, b_, c, g_ = sets[i % 4]
        ops.gemm(a, b_, out=c, epi=Lb.EPI_BF16_GELU, bias=bias_r, aux=g_)
    g1.record()
    torch.cuda.synchronize()
    gemm_ms = g0.elapsed_time(g1) / iters
    gemm_tflops = 2.0 * M * D * 4 * D / (gemm_ms * 1e-3) / 1e12
    del sets
    # DRAM traffic of that kernel (dram__bytes_read + dram__bytes_write of one launch) from the committed `ncu --set full` capture
    gemm_traffic, gemm_traffic_src = None, None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_roofline_ncu.json")))
        if [M, 4 * D, D] == list(cap.get("shape", [])):
            gemm_traffic, gemm_traffic_src = cap["traffic_bytes"], "profiles/r02_gemm_roofline_ncu.json (ncu --set full, one launch)"
    except Exception:  # noqa: BLE001
        pass

    # ---- secondary: the HBM-bound fused absorbing sampler (BASELINE.json configs[3]: B=64, N=1280, all rows masked) ----
    sampler_info = None
    if rank == 0 and not small:
        try:
            Bs = 64
            Vp = net.Vp
            lg = torch.empty(Bs * N, Vp, device=dev, dtype=torch.bfloat16)
            for c in range(0, Bs * N, 8192):
                lg[c:c + 8192].normal_(0, 3)
            xs = torch.full((Bs, N), model.mask_index, dtype=torch.int64, device=dev)
            mods = mod_d[:1].repeat(Bs, 1).contiguous()
            tt = torch.full((Bs,), 0.7, device=dev)
            for _ in range(2):
                ops.ddpm_update_logits(xs, lg, mods.view(-1), tt, tt - 0.01, model.mask_index, tv, V, seed=1, offset=1)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(5):
                ops.ddpm_update_logits(xs, lg, mods.view(-1), tt, tt - 0.01, model.mask_index, tv, V, seed=1, offset=2 + i)
            s1.record()
            torch.cuda.synchronize()
            sms = s0.elapsed_time(s1) / 5
            alg_bytes = Bs * (txt * tv + img * (V - tv)) * 2 + Bs * N * 8 * 2
            sampler_info = dict(kernel="ddpm_update_logits_kernel<no-cfg> (Philox noise)", workload=f"B={Bs} N={N} V={V} bf16 logits, all masked",
                                ms_per_launch=sms, algorithmic_bytes=alg_bytes, achieved_gbs=alg_bytes / (sms * 1e-3) / 1e9,
                                peak_gbs=pk["hbm_gbs"], frac=alg_bytes / (sms * 1e-3) / 1e9 / pk["hbm_gbs"])
            del lg
        except Exception as e:  # noqa: BLE001
            sampler_info = dict(error=str(e))

    tok_step = world * B * N
    fpt = 3 * flops_per_token_fwd(D, Lyr, attn_pairs, V)
    value = tok_step / (ms_dev / args.steps * 1e-3)
    e2e_v = tok_step / (ms_e2e / args.steps * 1e-3)
    # FLOPs actually executed: the output projection (2*D*V per token, x3 for forward + two backward GEMMs) runs on the masked
    # rows only unless --full-head; utilisation is reported on the EXECUTED count, the nominal model count is given beside it
    recent = head_rows_seen[-2 * args.steps:] or [(B * N, None)]
    head_frac = sum(r for r, _ in recent) / (len(recent) * B * N)
    # vocabulary columns per projected row: all V, or (split head) the text block for text rows and the image block for image rows
    head_cols = sum((mt * tv + (r - mt) * (V - tv // 8 * 8)) if mt else r * V for r, mt in recent) / max(sum(r for r, _ in recent), 1)
    fpt_exec = fpt - 3 * 2 * D * (V - head_frac * head_cols)
    model_tflops_per_gpu = value / world * fpt_exec / 1e12

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1 and not small:
            try:
                cpu = cpu_reference_tokens_per_sec(args.workload)
            except Exception as e:  # noqa: BLE001
                cpu = dict(value=None, unit="tokens/s", cores=os.cpu_count(), kind="port", sample=f"failed: {e}")
        line = dict(
            metric="joint_token_tokens_per_sec", value=value, unit="tokens/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
            config=dict(workload=args.workload, model=f"DiT D={D} L={Lyr} H={cfg.model.n_heads}", seq_len=N, per_gpu_batch=B,
                        global_batch=world * B, vocab=V, parallelism=f"dp{world}", optimizer="AdamW+clip(1.0)", optimizer_schedule="streamed" if opt.overlap else "blocking",
                        dropout=args.dropout,
                        head=("output projection + SUBS NLL on the masked, attended token rows only (exact: an unmasked token's log p is 0 "
                              f"under SUBS, reference model.py:621-658); mean {head_frac:.3f} of the rows x {head_cols:.0f} of {V} vocabulary columns "
                              "(text rows x text block, image rows x image block: force_argmax_valid_indices)" if head_frac < 1.0
                              else "output projection on all token rows (--full-head, the reference's dataflow)"),
                        **(dict(packing="documents = text U[32,512] + image {256,1024} tokens, tail padding; "
                                                              f"mean attended keys per query {attn_pairs:.0f} of {N}") if packed else {}),
                        l2="activations and weights per step >> 126 MB L2 (no flush needed)"),
            e2e=dict(value=e2e_v, unit="tokens/s", ms_per_step=ms_e2e / args.steps, h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                     loss=loss_e2e),
            gpu_launches=n_launch // max(args.steps, 1),
            clocks=clocks,
            roofline=dict(bound="tensor", kernel=f"gemm2_kernel<K-major,K-major,256,EPI_BF16_GELU> M={M} N={4*D} K={D} (mlp.0: bias + GELU epilogue, "
                                                 "as launched by the model)",
                          achieved=gemm_tflops, peak=pk["bf16_tflops"], unit="TFLOP/s", frac=gemm_tflops / pk["bf16_tflops"],
                          peak_source=pk["source"] + " (burst: kernel timed alone)", traffic=gemm_traffic, traffic_source=gemm_traffic_src,
                          algorithmic_bytes=2 * (M * D + 4 * D * D + 2 * M * 4 * D), ms_per_launch=gemm_ms),
            step_model_tflops_per_gpu=model_tflops_per_gpu,
            step_frac_of_sustained_peak=model_tflops_per_gpu / pk["bf16_tflops_sustained"],
            full_head=full_head_info,
            flops_per_token_fwd_bwd=fpt_exec, flops_per_token_fwd_bwd_nominal=fpt, head_rows_fraction=head_frac, head_cols_per_row=head_cols,
            cpu_baseline=cpu,
            sampler=sampler_info,
        )
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
