"""Noise-scaling search — mirror of tts/tts_t2i_noise_scaling.py (sample :16-77, main :80-159):
per prompt x round, draw `search_branch` seeded noises and run the stock FLUX pipeline (entry A)
on each.  Candidates are sharded over the ranks; like the reference this stage only generates
(no verifier is invoked) — the reflection stage scores them."""
from __future__ import annotations

import copy
import json
import os
from typing import Dict, List, Optional

import torch

from . import search as S
from .dist import DistCtx
from .reflectionflow import _ensure_pixels, _save_candidate, build_pipeline, flush_saves
from .utils import TORCH_DTYPE_MAP, get_latent_prep_fn, get_noises, parse_cli_args
from .verifiers import Candidate

MAX_SEED = S.MAX_SEED


def sample(noises: Dict[int, torch.Tensor], prompts: List[str], search_round: int, pipe, config: dict,
           original_prompt: str, midimg_path: str, *, ctx: Optional[DistCtx] = None) -> dict:
    ctx = ctx or DistCtx()
    config_cp = copy.deepcopy(config)
    bs = config_cp.get("batch_size_for_img_gen", 1)
    pa = config_cp["pipeline_args"]
    noise_items = list(noises.items())
    names = [os.path.join(midimg_path, f"{search_round}_round@{seed}.png") for seed, _ in noise_items]
    mine = ctx.my_candidates(len(noise_items))
    local = []
    for b0 in range(0, len(mine), bs):
        idxs = mine[b0:b0 + bs]
        seeds_batch = [noise_items[i][0] for i in idxs]
        if ctx.rank == 0:
            print(f"Generating images for batch with seeds: {seeds_batch}.")
        batched_latents = torch.stack([noise_items[i][1] for i in idxs]).squeeze(dim=1)
        batched_prompts = [prompts[i] for i in idxs]
        res = pipe(prompt=batched_prompts, latents=batched_latents,
                   guidance_scale=pa["guidance_scale"], num_inference_steps=pa["num_inference_steps"],
                   height=pa["height"], width=pa["width"], output_type="latent")
        for j, i in enumerate(idxs):
            c = Candidate(names[i], noise_items[i][0], latents=res.images[j:j + 1])
            _ensure_pixels(pipe, c, pa["height"], pa["width"])
            _save_candidate(c, names[i])
            local.append((i, c))
    return {"prompt": original_prompt, "search_round": search_round, "num_noises": len(noises),
            "generated_img": names, "local": local}


@torch.no_grad()
def main(argv=None, ctx: Optional[DistCtx] = None):
    args = parse_cli_args(argv)
    with open(args.pipeline_config_path, "r") as f:
        config = json.load(f)
    config.update(vars(args))
    config.setdefault("use_low_gpu_vram", False)
    ctx = ctx or DistCtx.from_env()
    if args.seed is not None:
        torch.manual_seed(args.seed)
    else:
        torch.manual_seed(int(ctx.broadcast_object(int(torch.seed() % (2 ** 31)))))
    search_rounds = config["search_args"]["search_rounds"]
    search_branch = config["search_args"]["search_branch"]
    pipeline_name = config["pipeline_args"].get("pretrained_model_name_or_path")
    root_dir = config["output_dir"]
    os.makedirs(root_dir, exist_ok=True)
    torch_dtype = TORCH_DTYPE_MAP[config["pipeline_args"].get("torch_dtype")]
    cfg_noload = dict(config)
    cfg_noload["pipeline_args"] = dict(config["pipeline_args"], lora_path=None)  # entry A: no LoRA
    pipe = build_pipeline(cfg_noload, args, ctx)
    with open(args.meta_path) as fp:
        metadatas = [json.loads(line) for line in fp]
    metadatas = metadatas[args.start_index:] if args.end_index == -1 else \
        metadatas[args.start_index:args.end_index]
    for index, metadata in enumerate(metadatas):
        original_prompt = metadata["prompt"]
        current_prompts = [original_prompt] * search_branch
        outpath = os.path.join(root_dir, f"{index + args.start_index:0>5}")
        midimg_path = os.path.join(outpath, "samples")
        if ctx.rank == 0:
            os.makedirs(midimg_path, exist_ok=True)
            with open(os.path.join(outpath, "metadata.jsonl"), "w") as fp:
                json.dump(metadata, fp)
        ctx.barrier()
        for rnd in range(1, search_rounds + 1):
            if ctx.rank == 0:
                print(f"\n=== Round: {rnd} ===")
            noises = get_noises(max_seed=MAX_SEED, num_samples=search_branch,
                                height=config["pipeline_args"]["height"],
                                width=config["pipeline_args"]["width"], dtype=torch_dtype,
                                fn=get_latent_prep_fn(pipeline_name))
            sample(noises=noises, prompts=current_prompts, search_round=rnd, pipe=pipe, config=config,
                   original_prompt=original_prompt, midimg_path=midimg_path, ctx=ctx)
    ctx.barrier()
    flush_saves()
    ctx.barrier()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
