#!/usr/bin/env python
"""Micro-benchmarks of the individual kernels at FLUX.1-dev shapes (CUDA events, L2 flushed by
rotating through buffers > L2).  Used for tuning and as the short command under ncu."""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflectionflow_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_attention(n_tok=4608, heads=24, iters=20):
    lib = L.load()
    nbuf = 3
    qkvs = [torch.randn(n_tok, 3 * heads * 128, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    out = torch.empty(n_tok, heads * 128, dtype=torch.bfloat16, device=dev)
    inner = heads * 128
    i = [0]

    def fn():
        qkv = qkvs[i[0] % nbuf]
        i[0] += 1
        L.check(lib.rf_op_attention(L.ptr(qkv), L.ptr(qkv[:, inner:]), L.ptr(qkv[:, 2 * inner:]),
                                    qkv.stride(0), L.ptr(out), out.stride(0), n_tok, heads, 1, n_tok, 0,
                                    0.0, L.cur_stream()))
    ms = timeit(fn, iters)
    fl = 4.0 * n_tok * n_tok * 128 * heads
    print(f"attention n={n_tok} h={heads}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")


def bench_gemm(M, N, K, epi, iters=20):
    lib = L.load()
    nbuf = max(2, int(300e6 // (2 * (M * K + N * K))) + 1)
    nbuf = min(nbuf, 6)
    xs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    Ws = [(torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16) for _ in range(nbuf)]
    b = torch.randn(N, device=dev).to(torch.bfloat16)
    y = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    gate = torch.randn(N, device=dev).to(torch.bfloat16)
    cos = torch.rand(M, 64, device=dev)
    sin = torch.rand(M, 64, device=dev)
    nq = torch.ones(128, dtype=torch.bfloat16, device=dev)
    i = [0]

    def fn():
        x, W = xs[i[0] % nbuf], Ws[i[0] % nbuf]
        i[0] += 1
        L.check(lib.rf_op_linear(epi, M, N, K, L.ptr(x), K, L.ptr(W), L.ptr(b), L.ptr(y), N, None, 0,
                                 L.ptr(y) if epi == 2 else None, N, L.ptr(gate) if epi == 2 else None,
                                 L.ptr(cos) if epi == 3 else None, L.ptr(sin) if epi == 3 else None,
                                 L.ptr(nq) if epi == 3 else None, L.ptr(nq) if epi == 3 else None,
                                 L.cur_stream()))
    ms = timeit(fn, iters)
    fl = 2.0 * M * N * K
    names = {0: "bias", 1: "gelu", 2: "gate_res", 3: "qkv"}
    print(f"gemm {names[epi]:8s} M={M} N={N} K={K}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")


def bench_ln(rows=4608, dim=3072, iters=50):
    lib = L.load()
    xs = [torch.randn(rows, dim, device=dev).to(torch.bfloat16) for _ in range(6)]
    out = torch.empty_like(xs[0])
    sc = torch.randn(dim, device=dev).to(torch.bfloat16)
    i = [0]

    def fn():
        x = xs[i[0] % 6]
        i[0] += 1
        L.check(lib.rf_op_ln_modulate(L.ptr(x), dim, L.ptr(out), dim, rows, dim, L.ptr(sc), L.ptr(sc),
                                      rows, 0, L.cur_stream()))
    ms = timeit(fn, iters)
    print(f"ln_modulate rows={rows}: {ms*1e3:.1f} us  {4.0*rows*dim/ms/1e6:.1f} GB/s")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="all")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    if a.what in ("all", "attn"):
        bench_attention(4608, 24, a.iters)
        if a.what == "all":
            bench_attention(5632, 24, a.iters)
    if a.what in ("all", "gemm"):
        for (M, N, K, e) in [(4608, 9216, 3072, 3), (4608, 12288, 3072, 1), (4608, 3072, 15360, 2),
                             (4608, 3072, 3072, 2), (4608, 3072, 12288, 2), (4608, 3072, 3072, 0)]:
            bench_gemm(M, N, K, e, a.iters)
    if a.what in ("all", "ln"):
        bench_ln(iters=a.iters)
