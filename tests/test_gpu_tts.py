"""-m gpu: the re-hosted outer loop end to end on one GPU with a tiny-depth (1+1 block, full width)
synthetic model: noise-scaling stage (entry A) -> reflection rounds (entry B: condition stream, merged
LoRA, VAE decode -> PIL-exact resize -> VAE encode for the parent condition), stub verifier/reflector.
Checks the artefact layout of the reference and run-to-run determinism (seeded)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from reflectionflow_b200.tts import noise_scaling, reflectionflow  # noqa: E402
from reflectionflow_b200.tts.dist import DistCtx  # noqa: E402

CONFIG = {
    "pipeline_args": {"pretrained_model_name_or_path": "black-forest-labs/FLUX.1-dev", "cache_dir": "x",
                      "torch_dtype": "bf16", "height": 256, "width": 256, "condition_size": 128,
                      "max_sequence_length": 512, "guidance_scale": 3.5, "num_inference_steps": 4,
                      "lora_path": "LORA"},
    "verifier_args": {"name": "nvila", "model_name": "stub", "cache_dir": "x"},
    "refine_args": {"name": "openai", "choice_of_metric": "overall_score", "max_new_tokens": 64,
                    "refine_prompt_relpath": "r.txt", "reflexion_prompt_relpath": "x.txt",
                    "verifier_prompt_relpath": "v.json"},
    "search_args": {"search_method": "random", "search_branch": 2, "search_rounds": 2},
    "model": {"add_cond_attn": False, "latent_lora": False, "union_cond_attn": True},
    "reflection_args": {"run_reflection": True, "name": "openai"},
    "prompt_refiner_args": {"run_refinement": True},
    "use_low_gpu_vram": False, "batch_size_for_img_gen": 1,
}


def _run(base):
    os.makedirs(base, exist_ok=True)
    cfg = os.path.join(base, "cfg.json")
    cfg0 = dict(CONFIG, search_args=dict(CONFIG["search_args"], search_rounds=1))
    json.dump(cfg0, open(os.path.join(base, "cfg0.json"), "w"))
    json.dump(CONFIG, open(cfg, "w"))
    meta = os.path.join(base, "meta.jsonl")
    with open(meta, "w") as f:
        f.write(json.dumps({"prompt": "a photo of a red cube on a blue sphere", "tag": "position"}) + "\n")
    ctx = DistCtx(0, 1, "cuda:0")
    common = ["--synthetic", "--seed", "0", "--layers", "1,1"]
    noise_scaling.main(["--pipeline_config_path", os.path.join(base, "cfg0.json"), "--meta_path", meta,
                        "--output_dir", os.path.join(base, "s0")] + common, ctx=ctx)
    reflectionflow.main(["--pipeline_config_path", cfg, "--imgpath", os.path.join(base, "s0"),
                         "--output_dir", os.path.join(base, "s1")] + common, ctx=ctx)
    torch.cuda.synchronize()
    return os.path.join(base, "s0", "00000"), os.path.join(base, "s1", "00000")


def test_outer_loop_end_to_end(tmp_path):
    s0, s1 = _run(str(tmp_path / "a"))
    samples = sorted(os.listdir(os.path.join(s0, "samples")))
    assert sum(f.endswith(".png") for f in samples) == 2 and sum(f.endswith(".latent.pt") for f in samples) == 2
    assert all(f.startswith("1_round@") for f in samples)
    mid = sorted(os.listdir(os.path.join(s1, "midimg")))
    assert sum(f.endswith(".png") for f in mid) == 4
    assert {f.split("_")[0] for f in mid} == {"1", "2"}
    for d in ("samples_lastround", "samples_best", "samples_path_bestround"):
        assert any(f.endswith(".png") for f in os.listdir(os.path.join(s1, d))), d
    meta = open(os.path.join(s1, "best_img_meta.jsonl")).read()
    assert "reflections1:" in meta and "refined_prompt2:" in meta and "filenames_batch2:" in meta
    from PIL import Image
    im = Image.open(os.path.join(s1, "midimg", mid[0] if mid[0].endswith(".png") else mid[1]))
    assert im.size == (256, 256) and im.mode == "RGB"
    # determinism: a second seeded run reproduces every latent bit for bit
    t0, t1 = _run(str(tmp_path / "b"))
    for f in mid:
        if f.endswith(".latent.pt"):
            a = torch.load(os.path.join(s1, "midimg", f))
            b = torch.load(os.path.join(t1, "midimg", f))
            assert torch.equal(a, b), f
