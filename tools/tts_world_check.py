#!/usr/bin/env python
"""Run the two-stage outer loop (tiny-depth synthetic model) under the current torchrun world and,
on rank 0, compare every produced latent with a reference directory made by a world-size-1 run.
usage: torchrun --nproc-per-node N tools/tts_world_check.py OUT [REF]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reflectionflow_b200.tts import noise_scaling, reflectionflow
from reflectionflow_b200.tts.dist import DistCtx

out = sys.argv[1]
ref = sys.argv[2] if len(sys.argv) > 2 else None
ctx = DistCtx.from_env()
cfg = {
    "pipeline_args": {"pretrained_model_name_or_path": "black-forest-labs/FLUX.1-dev", "cache_dir": "x",
                      "torch_dtype": "bf16", "height": 256, "width": 256, "condition_size": 128,
                      "max_sequence_length": 512, "guidance_scale": 3.5, "num_inference_steps": 4,
                      "lora_path": "LORA"},
    "verifier_args": {"name": "nvila", "model_name": "stub", "cache_dir": "x"},
    "refine_args": {"name": "openai", "choice_of_metric": "overall_score", "max_new_tokens": 64,
                    "refine_prompt_relpath": "r", "reflexion_prompt_relpath": "x", "verifier_prompt_relpath": "v"},
    "search_args": {"search_method": "random", "search_branch": 4, "search_rounds": 2},
    "model": {"add_cond_attn": False, "latent_lora": False, "union_cond_attn": True},
    "reflection_args": {"run_reflection": True, "name": "openai"},
    "prompt_refiner_args": {"run_refinement": True}, "use_low_gpu_vram": False, "batch_size_for_img_gen": 1}
if ctx.rank == 0:
    os.makedirs(out, exist_ok=True)
    json.dump(dict(cfg, search_args=dict(cfg["search_args"], search_rounds=1)), open(f"{out}/cfg0.json", "w"))
    json.dump(cfg, open(f"{out}/cfg.json", "w"))
    open(f"{out}/meta.jsonl", "w").write(json.dumps({"prompt": "a green bench left of a dog", "tag": "position"}) + "\n")
ctx.barrier()
common = ["--synthetic", "--seed", "0", "--layers", "1,1"]
noise_scaling.main(["--pipeline_config_path", f"{out}/cfg0.json", "--meta_path", f"{out}/meta.jsonl",
                    "--output_dir", f"{out}/s0"] + common, ctx=ctx)
reflectionflow.main(["--pipeline_config_path", f"{out}/cfg.json", "--imgpath", f"{out}/s0",
                     "--output_dir", f"{out}/s1"] + common, ctx=ctx)
ctx.barrier()
if ctx.rank == 0:
    files = []
    for d in ("s0/00000/samples", "s1/00000/midimg"):
        files += [os.path.join(d, f) for f in sorted(os.listdir(os.path.join(out, d))) if f.endswith(".latent.pt")]
    print(f"world={ctx.world}: {len(files)} candidate latents written")
    if ref:
        same = all(torch.equal(torch.load(os.path.join(out, f)), torch.load(os.path.join(ref, f))) for f in files)
        print("identical to world-1 run:", same)
        assert same
