#!/usr/bin/env python
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflectionflow_b200 import _lib as L
from reflectionflow_b200.vae import B200AutoencoderKL
vae = B200AutoencoderKL().init_synthetic_weights(0)
H = W = 1024
lat = torch.randn(1, 4096, 64).to(torch.bfloat16).cuda()
for _ in range(2):
    vae.decode_packed(lat, H, W, "u8")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = vae.decode_packed(lat, H, W, "u8")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
res = {"vae_decode_ms": ms, "tflops": 10.47 / ms * 1e3, "mean_u8": float(out.float().mean()),
       "graph": os.environ.get("RF_VAE_GRAPH", "1")}
# encode of the 512x512 condition image (the tree's parent -> condition path)
img = torch.randint(0, 256, (1, 512, 512, 3), dtype=torch.uint8).cuda()
for _ in range(2):
    vae.encode_packed(img)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    vae.encode_packed(img)
e1.record(); torch.cuda.synchronize()
res["vae_encode_512_ms"] = e0.elapsed_time(e1) / 5
print(json.dumps(res))
L.profile_start()
vae.decode_packed(lat, H, W, "u8")
prof = L.profile_stop()
tot = sum(v["ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    tf = v["flops"] / v["ms"] / 1e9 if v["flops"] else 0
    gb = v["bytes"] / v["ms"] / 1e6
    print(f"  {k:16s} launches {v['launches']:3d}  ms {v['ms']:7.3f}  share {v['ms']/tot:.3f}  {tf:7.1f} TF/s  {gb:7.1f} GB/s")
