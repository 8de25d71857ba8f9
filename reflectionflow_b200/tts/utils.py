"""Mirror of the reference's tts/utils.py for the FLUX path: CLI, dtype map, seeded noise factory.

`get_noises` keeps the reference's exact RNG protocol (tts/utils.py:131-155): seeds come from the
GLOBAL torch RNG, and each noise is drawn on the global CPU generator re-seeded with that seed
(`torch.manual_seed(seed)` — a side effect the next round's seeds depend on),
directly in bf16, then packed.  Same inputs => bit-identical seeds and noises as the reference."""
from __future__ import annotations

import argparse
import hashlib
import json
import re
from typing import Callable, Dict

import torch

TORCH_DTYPE_MAP = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
MODEL_NAME_MAP = {"black-forest-labs/FLUX.1-dev": "flux.1-dev"}


def parse_cli_args(argv=None):
    """Same flags as tts/utils.py:24-67, plus --seed / --synthetic for offline runs."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--pipeline_config_path", type=str, default="configs/headline_tree_flux_dev.json")
    parser.add_argument("--start_index", type=int, default=0)
    parser.add_argument("--end_index", type=int, default=-1)
    parser.add_argument("--imgpath", type=str, default="")
    parser.add_argument("--output_dir", type=str, default="output")
    parser.add_argument("--meta_path", type=str, default="meta.jsonl")
    parser.add_argument("--seed", type=int, default=None,
                        help="seed the global RNG (the reference leaves it unseeded)")
    parser.add_argument("--synthetic", action="store_true",
                        help="random-init weights + hash text embeddings + stub verifier (no network)")
    parser.add_argument("--text_encoders", choices=("hash", "native"), default="hash",
                        help="--synthetic only: hash embeddings, or random-init T5-XXL + CLIP-L run on the device")
    parser.add_argument("--layers", type=str, default=None, help="debug: 'double,single' layer counts")
    parser.add_argument("--lora_mode", choices=("exact", "merged"), default="exact",
                        help="exact = peft's unfused low-rank arithmetic (the reference's); merged = fuse_lora")
    return parser.parse_args(argv)


def _pack_latents(latents, batch_size, num_channels_latents, height, width):
    latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
    latents = latents.permute(0, 2, 4, 1, 3, 5)
    return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)


def prepare_latents_for_flux(batch_size: int, height: int, width: int, generator: torch.Generator,
                             device: str, dtype: torch.dtype) -> torch.Tensor:
    """tts/utils.py:71-87."""
    num_latent_channels = 16
    vae_scale_factor = 8
    height = 2 * (int(height) // (vae_scale_factor * 2))
    width = 2 * (int(width) // (vae_scale_factor * 2))
    shape = (batch_size, num_latent_channels, height, width)
    # diffusers randn_tensor: CPU generator => draw on the CPU in `dtype`, then move
    gdev = generator.device.type if generator is not None else "cpu"
    latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
    return _pack_latents(latents, batch_size, num_latent_channels, height, width)


def get_latent_prep_fn(pretrained_model_name_or_path: str) -> Callable:
    if pretrained_model_name_or_path != "black-forest-labs/FLUX.1-dev":
        raise KeyError(f"{pretrained_model_name_or_path}: only the FLUX.1-dev path is in scope")
    return prepare_latents_for_flux


def get_noises(max_seed: int, num_samples: int, height: int, width: int, device="cpu",
               dtype: torch.dtype = torch.bfloat16, fn: Callable = prepare_latents_for_flux
               ) -> Dict[int, torch.Tensor]:
    """tts/utils.py:131-155 (device defaults to "cpu": values are identical either way because the
    draw happens on the CPU generator; callers move what they need to the GPU)."""
    seeds = torch.randint(0, high=max_seed, size=(num_samples,))
    noises = {}
    for noise_seed in seeds:
        latents = fn(batch_size=1, height=height, width=width,
                     generator=torch.manual_seed(int(noise_seed)), device=device, dtype=dtype)
        noises.update({int(noise_seed): latents})
    assert len(noises) == len(seeds)
    return noises


def load_verifier_prompt(path: str):
    if path.endswith(".txt"):
        with open(path, "r") as f:
            return f.read().replace('"""', "")
    if path.endswith(".json"):
        with open(path, "r") as f:
            return json.load(f)
    raise ValueError("Unsupported file type. Please provide a .txt or .json file.")


def prompt_to_filename(prompt, max_length=100):
    filename = re.sub(r"[^a-zA-Z0-9]", "_", prompt.strip())
    filename = re.sub(r"_+", "_", filename)
    hash_digest = hashlib.sha256(prompt.encode()).hexdigest()[:8]
    base_filename = f"prompt@{filename}_hash@{hash_digest}"
    if len(base_filename) > max_length:
        base_length = max_length - len(hash_digest) - 7
        base_filename = f"prompt@{filename[:base_length]}_hash@{hash_digest}"
    return base_filename


def recover_json_from_output(output: str):
    start = output.find("{")
    end = output.rfind("}") + 1
    return json.loads(output[start:end])


def get_batches(items, batch_size):
    return [items[i:i + batch_size] for i in range(0, len(items), batch_size)]


def load_image(path_or_url):
    """tts/utils.py:188-200: a PIL image is returned as is, "http..." is fetched (needs network), else a local path."""
    from PIL import Image
    if isinstance(path_or_url, Image.Image):
        return path_or_url
    if str(path_or_url).startswith("http"):
        import io
        import requests
        response = requests.get(path_or_url, stream=True)
        response.raise_for_status()
        return Image.open(io.BytesIO(response.content))
    return Image.open(path_or_url)


def convert_to_bytes(path_or_url) -> bytes:
    """tts/utils.py:203-208: RGB PNG bytes of an image given by object, path or URL."""
    import io
    buf = io.BytesIO()
    load_image(path_or_url).convert("RGB").save(buf, format="PNG")
    return buf.getvalue()
