#!/usr/bin/env python
"""Step time of entry B (condition stream + LoRA on condition tokens): the headline tree's step."""
import os, sys, json, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflectionflow_b200 import _lib as L
from reflectionflow_b200.config import FluxDiTConfig
from reflectionflow_b200.pipeline import B200FluxPipeline, Condition, generate, flow_match_schedule
from reflectionflow_b200.tts.reflectionflow import synthetic_lora

dev = torch.device("cuda:0")
layers = os.environ.get("LAYERS", "19,38").split(",")
cfg = FluxDiTConfig(num_layers=int(layers[0]), num_single_layers=int(layers[1]))
use_lora = os.environ.get("LORA", "1") == "1"
pipe = B200FluxPipeline.from_synthetic(cfg, seed=0, device=dev, lora_rank=32 if use_lora else 0)
if use_lora:
    pipe.transformer.load_lora(synthetic_lora(cfg), mode=os.environ.get("LORA_MODE", "exact"))
g = torch.Generator().manual_seed(0)
lat = torch.randn(1, 4096, 64, generator=g).to(torch.bfloat16)
txt = torch.randn(1, 512, 4096, generator=g).to(torch.bfloat16)
pool = torch.randn(1, 768, generator=g).to(torch.bfloat16)
cond = Condition("cot", latents=torch.randn(1, 1024, 64, generator=g).to(torch.bfloat16), position_delta=[0, -32])
K = int(os.environ.get("STEPS", "8"))
def run(n):
    return generate(pipe, conditions=[cond], model_config={"union_cond_attn": True}, default_lora=True,
                    prompt_embeds=txt, pooled_prompt_embeds=pool, latents=lat, height=1024, width=1024,
                    num_inference_steps=n, output_type="latent").images
run(3); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = run(K); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
tf = (57 * 24 * 3072 * 3072 * 5632 + 57 * 4 * 5632 * 5632 * 3072) / 1e12
print(json.dumps({"entry": "B", "lora": use_lora, "lora_mode": pipe.transformer.lora_mode, "n_tok": 5632, "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                  "step_tflop": tf, "tflops": tf / ms * 1e3, "finite": bool(torch.isfinite(out.float()).all())}))
from reflectionflow_b200.transformer import tranformer_forward
m = pipe.transformer
ids_i = pipe._prepare_latent_image_ids(1, 64, 64, dev, torch.bfloat16)
ids_t = torch.zeros(512, 3, dtype=torch.bfloat16, device=dev)
ctok, cids, _ = cond.encode(pipe)
kw = dict(hidden_states=lat, encoder_hidden_states=txt, pooled_projections=pool,
          timestep=torch.tensor([0.5], dtype=torch.bfloat16), img_ids=ids_i, txt_ids=ids_t,
          guidance=torch.tensor([3.5]), return_dict=False)
tranformer_forward(m, ctok, cids, None, {}, 0, **kw)
torch.cuda.synchronize()
L.profile_start()
tranformer_forward(m, ctok, cids, None, {}, 0, **kw)
prof = L.profile_stop()
tot = sum(v["ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:20s} launches {v['launches']:4d}  ms {v['ms']:7.3f}  share {v['ms']/tot:.3f}")
