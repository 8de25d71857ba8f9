#!/usr/bin/env python
"""dev tool: in-process A/B of the LoRA arithmetic / launch forms on the headline step (5632 tokens):
exact with 256-wide (single accumulator stage) / 128-wide (double-buffered) condition tiles, merged — alternated
several times in ONE process so that box / clock state is shared."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from reflectionflow_b200.pipeline import Condition, flow_match_schedule
from reflectionflow_b200.tts import reflectionflow as RF
from reflectionflow_b200.tts.dist import DistCtx

dev = torch.device("cuda:0")
ctx = DistCtx(0, 1, dev)
cfgj = json.load(open(os.path.join(os.path.dirname(RF.__file__), "configs", "headline_tree_flux_dev.json")))


class A:
    synthetic, layers, text_encoders, lora_mode = True, "19,38", "hash", "exact"


pipe = RF.build_pipeline(cfgj, A, ctx)
m = pipe.transformer
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16
lat = torch.randn(1, 4096, 64, generator=g).to(bf).to(dev)
txt = torch.randn(1, 512, 4096, generator=g).to(bf).to(dev)
pool = torch.randn(1, 768, generator=g).to(bf).to(dev)
cond = torch.randn(1, 1024, 64, generator=g).to(bf).to(dev)
img_ids = pipe._prepare_latent_image_ids(1, 64, 64, dev, bf)
txt_ids = torch.zeros(512, 3, dtype=bf, device=dev)
_, cond_ids, _ = Condition("cot", latents=cond, position_delta=[0, -32]).encode(pipe)
K = int(os.environ.get("STEPS", "10"))
ts, sig = flow_match_schedule(K, 4096)
t_in = ts.to(bf) / 1000
lora = RF.synthetic_lora(m.cfg, seed=1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(mode, bn):
    os.environ["RF_LORA_BN"] = str(bn)  # condition-stream tile width of the fused-LoRA GEMM (read at capture time)
    m.load_lora(lora, mode=mode)   # drops the captured graph: the next denoise re-captures with this form
    m.denoise(lat, txt, pool, t_in[:3], sig[:4], 3.5, img_ids, txt_ids, cond, cond_ids, {})
    torch.cuda.synchronize()
    e0.record()
    m.denoise(lat, txt, pool, t_in, sig, 3.5, img_ids, txt_ids, cond, cond_ids, {})
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


res = {"exact_bn256": [], "exact_bn128": [], "merged": []}
for rep in range(3):
    res["exact_bn256"].append(run("exact", 256))
    res["exact_bn128"].append(run("exact", 128))
    res["merged"].append(run("merged", 256))
print(json.dumps({k: [round(x, 2) for x in v] for k, v in res.items()}))
