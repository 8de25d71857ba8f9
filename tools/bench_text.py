"""time the native T5-v1.1-XXL + CLIP-L encoders (random weights, full size) — dev tool"""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reflectionflow_b200.text import B200TextEncoders
from reflectionflow_b200 import _lib as L

enc = B200TextEncoders().init_synthetic_weights(0)
res = {}
for B in (1, 4):
    ids = torch.randint(0, 32000, (B, 512))
    cids = torch.randint(3, 49000, (B, 77)); cids[:, -1] = 49407
    for _ in range(2):
        enc.t5_encode(ids); enc.clip_encode(cids)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 5
    e[0].record()
    for _ in range(n): enc.t5_encode(ids)
    e[1].record()
    for _ in range(n): enc.clip_encode(cids)
    e[2].record()
    torch.cuda.synchronize()
    res[f"B{B}"] = {"t5_ms": e[0].elapsed_time(e[1]) / n, "clip_ms": e[1].elapsed_time(e[2]) / n}
L.profile_start()
enc.t5_encode(torch.randint(0, 32000, (1, 512)))
res["t5_profile_B1"] = L.profile_stop()
print(json.dumps(res, indent=1))
