"""Debug tool: in-kernel timeline of ONE CTA of the attention-backward kernels (clock64 stamps of its pipeline events).

Builds its own copy of csrc/attention.cu with -DUD_ATTN_TRACE into tools/_build/libattn_trace.so (the product library never
carries the stamps), runs ud_attn_bwd on the bench shape and prints, per streamed sub-tile, when each hand-off happened
(cycles since the CTA started).  Usage:  python tools/attn_trace.py [--build-only] [--B 8 --N 1280 --H 16]
"""
from __future__ import annotations

import argparse
import ctypes
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tools", "_build")
LIB = os.path.join(OUT, "libattn_trace.so")

EVENTS = {
    17: "cta start / fixed tiles landed",
    13: "TMA: stage free",
    0: "MMA: scores issue start",
    1: "MMA: scores issued",
    14: "WG: loop top",
    5: "WG: S ready (woke)",
    6: "WG: S loaded",
    7: "WG: P computed",
    8: "WG: P stored+arrived",
    2: "MMA: p_rdy seen",
    9: "WG: dP ready (woke)",
    10: "WG: dP loaded",
    11: "WG: dS computed",
    12: "WG: dS stored+arrived",
    3: "MMA: ds_rdy seen",
    4: "MMA: acc MMAs issued",
    18: "MMA: dP issue start",
    19: "MMA: dP issued",
}
COLS = [13, 0, 1, 18, 19, 14, 5, 6, 7, 8, 2, 9, 10, 11, 12, 3, 4]


def build():
    os.makedirs(OUT, exist_ok=True)
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "unidisc_b200", "csrc"), "--expt-relaxed-constexpr",
           "-DUD_ATTN_TRACE", "-shared", "-o", LIB, os.path.join(ROOT, "unidisc_b200", "csrc", "attention.cu"),
           os.path.join(ROOT, "unidisc_b200", "csrc", "gemm.cu")]       # gemm.cu holds the tensor-map encoder
    subprocess.run(cmd, check=True)
    return LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build-only", action="store_true")
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=1280)
    ap.add_argument("--H", type=int, default=16)
    ap.add_argument("--iters", type=int, default=24)
    args = ap.parse_args()
    if args.build_only or not os.path.exists(LIB):
        build()
        if args.build_only:
            return
    import torch

    lib = ctypes.CDLL(LIB)
    B, N, H, hd = args.B, args.N, args.H, 128
    D = H * hd
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = (torch.randn(B * N, 3 * D, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    o = torch.randn(B * N, D, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(B * N, D, device=dev, generator=g).to(torch.bfloat16)
    # a plausible lse: log-sum-exp of the scaled scores
    scale = 1.0 / math.sqrt(hd)
    qh = q.view(B, N, H, hd).permute(0, 2, 1, 3).float()
    kh = k.view(B, N, H, hd).permute(0, 2, 1, 3).float()
    lse = torch.logsumexp(qh @ kh.transpose(-1, -2) * scale, dim=-1).contiguous()
    delta = torch.empty(2, B, H, N, device=dev)
    dqkv = torch.empty(B * N, 3 * D, device=dev, dtype=torch.bfloat16)
    dq, dk, dv = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
    trace = torch.zeros(32 * 64, device=dev, dtype=torch.int64)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    LL = ctypes.c_longlong
    assert lib.ud_attn_set_trace(P(trace)) == 0

    def run():
        rc = lib.ud_attn_bwd(P(q), P(k), LL(q.stride(0)), P(v), LL(v.stride(0)), P(o), P(do), LL(o.stride(0)), P(lse), P(delta),
                             P(dq), P(dk), LL(dq.stride(0)), P(dv), LL(dv.stride(0)), None, B, N, H, hd, ctypes.c_float(scale),
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    # both kernels stamp the same table (the dQ kernel runs second and overwrites): trace them one at a time through the
    # environment switch of the launcher if present, else report the last writer.
    for which in ("UD_ATTN_TRACE_ONLY=dkv", "UD_ATTN_TRACE_ONLY=dq"):
        key, val = which.split("=")
        os.environ[key] = val
        trace.zero_()
        run()
        torch.cuda.synchronize()
        t = trace.view(32, 64).cpu()
        t0 = int(t[17, 0])
        print(f"\n=== {val} kernel, CTA (3,5,2): cycles since CTA start; fixed tiles landed at {int(t[17, 1]) - t0}")
        print("sub " + " ".join(f"{c:>7d}" for c in COLS))
        for i in range(min(args.iters, 16)):
            if int(t[0, i]) == 0 and int(t[5, i]) == 0:
                break
            print(f"{i:3d} " + " ".join(f"{(int(t[c, i]) - t0) if int(t[c, i]) else 0:7d}" for c in COLS))
        for g in (1, 2, 3):          # v3 kernels: warpgroup g stamps its own rows at offset 16 * g (v2: both share offset 0)
            if int(t[5, 16 * g]) == 0:
                continue
            print(f"warpgroup {g} (warpgroup events only):")
            for i in range(16 * g, 16 * g + 16):
                if int(t[5, i]) == 0:
                    break
                print(f"{i - 16 * g:3d} " + " ".join(f"{(int(t[c, i]) - t0) if int(t[c, i]) else 0:7d}" for c in COLS))
        print(f"epilogue: all_done seen wg0 {int(t[15, 0]) - t0} wg1 {int(t[15, 1]) - t0}; stores done wg0 {int(t[16, 0]) - t0} "
              f"wg1 {int(t[16, 1]) - t0}")
        print("legend: " + "; ".join(f"{c}={EVENTS[c]}" for c in COLS))


if __name__ == "__main__":
    main()
