"""CPU, world_size=2, gloo: bucket planning and collective sequencing of ThinDDP (the N>1 path) on a fake flat-parameter
module.  The (de)compression kernels are injected as torch restatements of torch's bf16 compress hook
(`buffer.to(bf16).div_(world)` ... `buffer.copy_(result)`), which is also what the CUDA kernels are tested against."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeFlat(torch.nn.Module):
    """same surface ThinDDP relies on: flat buffers, per-block ranges, named parameters with slot offsets"""

    def __init__(self, n_blocks=3):
        super().__init__()
        self.n_blocks = n_blocks
        sizes = {}
        for i in range(n_blocks):
            sizes[f"blocks.{i}.attention.attn_qkv.weight"] = 3 * 64 * 64
            sizes[f"blocks.{i}.mlp.0.weight"] = 4 * 64 * 64
        sizes["output_layer.linear.weight"] = 100 * 64
        self._big = list(sizes)
        small = {"vocab_embed.embedding": 100 * 64, "modality_embed.embedding": 2 * 64}
        for i in range(n_blocks):
            small[f"blocks.{i}.norm1.weight"] = 64
            small[f"blocks.{i}.mlp.0.bias"] = 256
        small["output_layer.linear.bias"] = 100
        sizes.update(small)
        self._offs, off = {}, 0
        for n, s in sizes.items():
            self._offs[n] = off
            off += (s + 63) // 64 * 64
        self._big_end = self._offs["vocab_embed.embedding"]
        self._sizes = sizes
        self.flat_params = torch.zeros(off)
        self.flat_grads = torch.zeros(off)
        self.grad_ready_hook = None
        self._params = {n: torch.nn.Parameter(self.flat_params[o:o + sizes[n]]) for n, o in self._offs.items()}

    def named_parameters(self, *a, **k):
        return iter(self._params.items())

    def _ensure_ready(self):
        pass

    def mark_weights_updated(self, shadow_is_current=False):
        pass

    def block_grad_range(self, i):
        names = [n for n in self._offs if n.startswith(f"blocks.{i}.")]
        big = [n for n in names if self._offs[n] < self._big_end]
        small = [n for n in names if self._offs[n] >= self._big_end]
        rng = lambda ns: (min(self._offs[n] for n in ns), max(self._offs[n] + (self._sizes[n] + 63) // 64 * 64 for n in ns))
        return rng(big), rng(small)


def _pack(g, dst, inv_world):
    dst.copy_((g.to(torch.bfloat16).float() * inv_world).to(torch.bfloat16))


def _unpack(src, g):
    g.copy_(src.float())


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unidisc_b200.ddp import ThinDDP
    m = FakeFlat()
    m.flat_params.fill_(float(rank + 1))
    ddp = ThinDDP(m, _pack=_pack, _unpack=_unpack)
    ok_bcast = bool((m.flat_params == 1.0).all())           # rank-0 weights everywhere
    gen = torch.Generator().manual_seed(100 + rank)
    grads = torch.randn(m.flat_grads.numel(), generator=gen)
    # expected: bf16 compress hook semantics over both ranks
    all_g = [torch.randn(m.flat_grads.numel(), generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    expect = sum((g.to(torch.bfloat16) / world) for g in all_g)          # bf16 sum, like NCCL/gloo on bf16 buffers
    m.flat_grads.copy_(grads)
    with ddp.no_sync():                                      # accumulation micro-step: nothing may be communicated
        for i in [m.n_blocks] + list(range(m.n_blocks - 1, -1, -1)) + [-1]:
            m.grad_ready_hook(i)
    ok_nosync = torch.equal(m.flat_grads, grads)
    # post-bucket hook (the streamed optimizer's gradient-norm partial sums): once per bucket, AFTER that bucket's all-reduce
    seen = []

    def post(block_idx, ranges, on_side_stream):
        ok = all(torch.equal(m.flat_grads[lo:hi], expect.float()[lo:hi]) for lo, hi in ranges)
        seen.append((block_idx, on_side_stream, ok, sum(hi - lo for lo, hi in ranges)))

    ddp.post_bucket_hook = post
    for i in [m.n_blocks] + list(range(m.n_blocks - 1, -1, -1)) + [-1]:     # order used by DIT._backward_impl
        m.grad_ready_hook(i)
    got = m.flat_grads.clone()
    ok_post = ([b for b, _, _, _ in seen] == [m.n_blocks] + list(range(m.n_blocks - 1, -1, -1)) + [-1]
               and all(side and ok for _, side, ok, _ in seen) and sum(n for _, _, _, n in seen) == got.numel())
    # an optimizer handed the bare module must chain behind the DDP hook, never replace it
    from unidisc_b200.ddp import FusedAdamW
    opt = FusedAdamW(m, max_grad_norm=1.0)
    ok_chain = opt.ddp is ddp and getattr(m.grad_ready_hook, "__self__", None) is ddp and not opt.overlap
    opt2 = FusedAdamW(ddp, max_grad_norm=1.0)
    ok_chain = ok_chain and opt2.ddp is ddp and opt2.module is m
    # every element reduced exactly once
    covered = torch.zeros_like(got, dtype=torch.int32)
    for rs in ddp._ranges_by_block.values():
        for lo, hi in rs:
            covered[lo:hi] += 1
    q.put((rank, ok_bcast, ok_nosync, bool((covered == 1).all()), float((got - expect.float()).abs().max()),
           ddp.bytes_on_wire_per_step, 2 * got.numel(), ok_post, ok_chain))
    dist.destroy_process_group()


def test_thin_ddp_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_bcast, ok_nosync, covered_once, err, wire, expect_wire, ok_post, ok_chain in res:
        assert ok_post, "post_bucket_hook must fire once per bucket, on the communication path, after the bucket's all-reduce"
        assert ok_chain, "FusedAdamW must chain behind ThinDDP's gradient hook (replacing it disables the all-reduce)"
        assert ok_bcast, "weights must be broadcast from rank 0"
        assert ok_nosync, "no_sync() must suppress the all-reduce"
        assert covered_once, "every gradient element must be all-reduced exactly once"
        assert err <= 2e-2, err          # bf16 wire format
        assert wire == expect_wire
