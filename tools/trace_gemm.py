#!/usr/bin/env python
"""dev tool: per-tile clock64 timeline of CTA 0 of the GEMM kernel."""
import os, sys, ctypes, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflectionflow_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
def run_case(M, N, K, epi):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    b = torch.randn(N, device=dev).to(torch.bfloat16)
    y = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    gate = torch.randn(N, device=dev).to(torch.bfloat16)
    cos = torch.rand(M, 64, device=dev); sin = torch.rand(M, 64, device=dev)
    nq = torch.ones(128, dtype=torch.bfloat16, device=dev)
    tr = torch.zeros(16 * 8, dtype=torch.int64, device=dev)
    def run():
        L.check(lib.rf_op_linear(epi, M, N, K, L.ptr(x), K, L.ptr(W), L.ptr(b), L.ptr(y), N, None, 0,
                                 L.ptr(y) if epi == 2 else None, N, L.ptr(gate) if epi == 2 else None,
                                 L.ptr(cos) if epi == 3 else None, L.ptr(sin) if epi == 3 else None,
                                 L.ptr(nq) if epi == 3 else None, L.ptr(nq) if epi == 3 else None, L.cur_stream()))
    run(); torch.cuda.synchronize()
    lib.rf_dbg_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
    run(); torch.cuda.synchronize()
    lib.rf_dbg_set_gemm_trace(None)
    t = tr.cpu().view(16, 8)
    t0 = int(t[0][0])
    ntile = int((t[:, 3] > 0).sum())
    print(f"== M={M} N={N} K={K} epi={epi}: tiles on CTA0 = {ntile}; ideal mainloop/tile = {K//64*512} cycles")
    print("tile  mma_start  tempty_wait  fullbar_stall  mma_issue_end  epi_start  epi_end  epi_len  epi_wait_from")
    for i in range(ntile):
        r = [int(v) for v in t[i]]
        print(f"{i:3d} {r[0]-t0:10d} {r[1]-r[0]:11d} {r[2]:13d} {r[3]-t0:13d} {r[4]-t0:10d} {r[5]-t0:8d} {r[5]-r[4]:7d} {r[7]-t0:10d}")
for c in [(4608, 3072, 3072, 0), (4608, 3072, 3072, 2), (4608, 12288, 3072, 1), (4608, 9216, 3072, 3), (4608, 3072, 15360, 2)]:
    run_case(*c)
