#!/usr/bin/env python
"""dev tool: one gate+residual GEMM per shape (run under ncu with RF_GEMM_BAND=n to compare rasters)"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflectionflow_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
for (M, N, K, epi) in [(4608, 3072, 12288, 2), (4608, 3072, 15360, 2), (4608, 12288, 3072, 1), (4608, 9216, 3072, 0)]:
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device=dev).to(torch.bfloat16)
    y = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    gate = torch.randn(N, device=dev).to(torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for it in range(3):
        flush.zero_()
        L.check(lib.rf_op_linear(epi, M, N, K, L.ptr(x), K, L.ptr(W), L.ptr(b), L.ptr(y), N, None, 0,
                                 L.ptr(y) if epi == 2 else None, N, L.ptr(gate) if epi == 2 else None,
                                 None, None, None, None, L.cur_stream()))
    torch.cuda.synchronize()
