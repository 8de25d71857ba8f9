#!/usr/bin/env python
"""dev tool: clock64 timeline of CTA 0 of the attention kernel (MMA thread + one lane per softmax WG).
Needs a DEV build (`make -C reflectionflow_b200/csrc clean all DEV=1`): the probes are compiled out of production
builds (they cost ~12 % of the softmax loop's instructions)."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflectionflow_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
n_tok, heads = 4608, 24
qkv = torch.randn(n_tok, 3 * heads * 128, device=dev).to(torch.bfloat16)
out = torch.empty(n_tok, heads * 128, dtype=torch.bfloat16, device=dev)
inner = heads * 128
tr = torch.zeros(24 * 16, dtype=torch.int64, device=dev)
def run():
    L.check(lib.rf_op_attention(L.ptr(qkv), L.ptr(qkv[:, inner:]), L.ptr(qkv[:, 2 * inner:]), qkv.stride(0),
                                L.ptr(out), out.stride(0), n_tok, heads, 1, n_tok, 0, 0.0, L.cur_stream()))
run(); torch.cuda.synchronize()
lib.rf_dbg_set_attn_trace(ctypes.c_void_p(tr.data_ptr()))
run(); torch.cuda.synchronize()
lib.rf_dbg_set_attn_trace(None)
t = tr.cpu().view(24, 16)
t0 = int(t[0][t[0] > 0].min())
names = ["mma:pA", "mma:issuedA", "mma:pB", "mma:issuedB", "A:s_full", "A:ld", "A:exp", "A:arrive",
         "B:s_full", "B:ld", "B:exp", "B:arrive", "A:max", "A:c1st", "A:plo", "A:c3st"]
print("j  " + " ".join(f"{n:>8s}" for n in names))
for j in range(4, 14):
    print(f"{j:2d} " + " ".join(f"{int(t[j][i]) - t0:8d}" for i in range(16)))
print("per-iteration period (mma:pA):", [int(t[j + 1][0] - t[j][0]) for j in range(4, 14)])
