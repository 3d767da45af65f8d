"""GPU, world_size=2, NCCL (skipped on a single-GPU box): ThinDDP + FusedAdamW on hardware.
  * after a synchronised backward every rank holds the bf16-compress-hook mean of the per-rank gradients
    (torch default_hooks.bf16_compress_hook arithmetic: bf16(g)/world, summed on the wire, copied back to fp32);
  * with clipping ACTIVE the clip coefficient is identical on both ranks, so after 3 optimizer steps the fp32 parameters are
    bit-identical across ranks (what torch DDP + clip_grad_norm_ guarantees, reference main.py:641-656 / model.py:1518)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bf16 = torch.bfloat16


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from unidisc_b200.config import make_config
        from unidisc_b200.ddp import FusedAdamW, ThinDDP
        from unidisc_b200.model import Diffusion
        from unidisc_b200.synth import joint_batch
        cfg = make_config("small", hidden_size=256, n_blocks=3, n_heads=4, txt_length=64, img_length=64, image_vocab_size=255,
                          text_vocab_size=257, dropout=0.0)
        torch.manual_seed(rank)                      # different initial weights per rank: ThinDDP must broadcast rank 0's
        model = Diffusion(cfg, device=dev)
        model.train()
        net = model.backbone
        ddp = ThinDDP(net)
        opt = FusedAdamW(ddp, lr=1e-3, max_grad_norm=0.05)          # tiny threshold: the clip coefficient is < 1 every step
        ids, mod = joint_batch(4, 64, 64, model.text_vocab_size, model.vocab_size, seed=100 + rank)
        batch = dict(input_ids=ids.to(dev), modality=mod.to(dev), attention_mask=torch.ones_like(ids, dtype=torch.bool).to(dev))

        def backward(seed):
            torch.manual_seed(seed)
            model.compute_loss(batch).loss.backward()

        # (1) local gradients (no_sync) vs the synchronised ones
        opt.zero_grad()
        with ddp.no_sync():
            backward(7)
        torch.cuda.synchronize()
        g_local = net.flat_grads.clone()
        opt.zero_grad()
        backward(7)
        torch.cuda.synchronize()
        g_sync = net.flat_grads.clone()
        gathered = [torch.empty_like(g_local) for _ in range(world)]
        dist.all_gather(gathered, g_local)
        expect = gathered[0].to(bf16) / world
        for g in gathered[1:]:
            expect = expect + g.to(bf16) / world                      # bf16 sum, like the all-reduce on the bf16 wire buffers
        # Bit-equal up to the run-to-run non-determinism of the fp32 atomics in the bias / norm-weight / embedding gradient
        # reductions (the two backward passes above are separate runs: an fp32 last-bit difference occasionally moves a value
        # across a bf16 rounding boundary).  Allowed: a handful of elements, each off by at most one bf16 ulp.
        diff = (g_sync - expect.float()).abs()
        bad = diff > 0
        n_bad = int(bad.sum())
        rel = float((diff[bad] / expect.float().abs()[bad].clamp_min(1e-30)).max()) if n_bad else 0.0
        ok_mean = n_bad <= max(4, int(1e-5 * g_sync.numel())) and rel <= 2.0 ** -7
        where = ""
        if n_bad:
            offs = net._offs
            idx = bad.nonzero().squeeze(1)[:8].tolist()
            names = sorted(offs, key=lambda k: offs[k])
            where = ", ".join(f"{[n for n in names if offs[n] <= i][-1]}+{i - max(offs[n] for n in names if offs[n] <= i)}" for i in idx)
        err_mean = f"{float(diff.max()):.3e} ({n_bad} of {g_sync.numel()} elements differ, max rel {rel:.2e}: {where})"
        # (2) three clipped optimizer steps: parameters must stay bit-identical across ranks
        opt._buckets_seen = 0
        opt._sumsq.zero_()
        norms = []
        for it in range(3):
            opt.zero_grad()
            backward(20 + it)
            opt.step()
            norms.append(float(opt.last_grad_norm))
        opt.join()
        torch.cuda.synchronize()
        p_all = [torch.empty_like(net.flat_params) for _ in range(world)]
        dist.all_gather(p_all, net.flat_params.contiguous())
        ok_params = all(torch.equal(p_all[0], p) for p in p_all[1:])
        sh = [torch.empty_like(net.flat_params_bf16) for _ in range(world)]
        dist.all_gather(sh, net.flat_params_bf16.contiguous())
        ok_shadow = all(torch.equal(sh[0], s) for s in sh[1:]) and torch.equal(sh[0], p_all[0].to(bf16))
        q.put((rank, ok_mean, err_mean, ok_params, ok_shadow, norms, ddp.sumsq_target is not None, None))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, False, "nan", False, False, [], False, traceback.format_exc()))


def test_thin_ddp_two_ranks_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, ok_mean, err_mean, ok_params, ok_shadow, norms, fused, tb in sorted(res):
        assert tb is None, tb
        assert fused, "the all-reduced gradient's norm is summed by the decompression kernel (ud_grad_unpack_bf16_sumsq)"
        assert ok_mean, f"rank {rank}: synchronised gradient != bf16-compress-hook mean (max err {err_mean})"
        assert all(n > 0.05 for n in norms), f"clipping must be active in this test: norms {norms}"
        assert ok_params, f"rank {rank}: fp32 parameters differ between ranks after 3 clipped steps"
        assert ok_shadow, f"rank {rank}: bf16 shadow weights differ between ranks / from the masters"
    assert res[0][5] == res[1][5], f"gradient norms differ between ranks: {res[0][5]} vs {res[1][5]}"
