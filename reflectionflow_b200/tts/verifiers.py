"""Verifier / reflection / refinement hooks of the outer loop.

The reference wires three external models into tts/tts_reflectionflow.py:
  * a verifier — OpenAIVerifier.score (verifiers/openai_verifier.py:122-164, GPT-4o over HTTP) or
    NVILA-Lite-2B-Verifier (verifiers/nvila_verifier.py:4-10, remote code from the Hub), used at
    tts_reflectionflow.py:144-170 and :337-354;
  * a reflection writer (GPT-4o or a local Qwen server, :196-242);
  * a prompt refiner (GPT-4o, :245-257).
All three need the network or Hub checkpoints, so offline they are replaced by deterministic stubs
with the same return shapes; real models plug in through the same small interfaces."""
from __future__ import annotations

import hashlib
import math
import os
import threading
from typing import Any, Dict, List, Optional, Sequence

import torch


class Candidate:
    """One generated sample: identifier (the reference's PNG path "<dir>/<round>_round@<seed>.png"),
    final latent (always) and decoded image (when a VAE is attached)."""

    def __init__(self, name: str, seed: int, latents: Optional[torch.Tensor] = None, image=None,
                 stub_score: Optional[float] = None, image_u8: Optional[torch.Tensor] = None):
        self.name, self.seed, self.latents, self.image = name, int(seed), latents, image
        self.stub_score = stub_score
        self.image_u8 = image_u8  # decoded uint8 [H, W, 3] on the device (when a VAE is attached)

    def pil(self):
        if self.image is None and self.image_u8 is not None:
            from PIL import Image
            self.image = Image.fromarray(self.image_u8.cpu().numpy())
        return self.image

    def png_bytes(self) -> Optional[bytes]:
        """PNG encoding of the image, computed once per candidate (the round's artefact layout stores
        the same image under up to four names); thread-safe, callable from the save workers once
        `pil()` has been materialised on the calling thread."""
        if self.image is None:
            return None
        lock = self.__dict__.setdefault("_png_lock", threading.Lock())
        with lock:
            if self.__dict__.get("_png") is None:
                import io
                buf = io.BytesIO()
                self.image.save(buf, format="PNG")
                self._png = buf.getvalue()
        return self._png


def latent_functional(latents: torch.Tensor) -> float:
    """Fixed linear functional of a latent in fp64 (see search.stub_verifier_score)."""
    from .search import stub_verifier_score
    return float(stub_verifier_score(latents.reshape(1, -1, latents.shape[-1]))[0].item())


class StubVerifier:
    """Deterministic stand-in with the NVILA or OpenAI output shape.

    nvila : {"image_name", "label": "yes"|"no", "score": float}   (tts_reflectionflow.py:160-164)
    openai: {<aspect>: {"score": int 0-10, "explanation": str}, ..., "overall_score": {...}}
            (verifiers/openai_verifier.py:23-69) — `choice_of_metric` indexes it."""

    def __init__(self, name: str = "nvila", choice_of_metric: str = "overall_score"):
        assert name in ("nvila", "openai")
        self.name = name
        self.choice_of_metric = choice_of_metric

    def value(self, cand: Candidate) -> float:
        if cand.stub_score is not None:
            return float(cand.stub_score)
        return latent_functional(cand.latents)

    def score_one(self, cand: Candidate, prompt: str) -> Dict[str, Any]:
        v = self.value(cand)
        if self.name == "nvila":
            p_yes = 1.0 / (1.0 + math.exp(-40.0 * v))
            if p_yes >= 0.5:
                return {"image_name": cand.name, "label": "yes", "score": p_yes}
            return {"image_name": cand.name, "label": "no", "score": 1.0 - p_yes}
        s = max(0, min(10, int(round(5 + 60.0 * v))))
        return {self.choice_of_metric: {"score": s, "explanation": "stub"},
                "image_name": cand.name}

    def score(self, cands: Sequence[Candidate], prompts: Sequence[str]) -> List[Dict[str, Any]]:
        return [self.score_one(c, p) for c, p in zip(cands, prompts)]


class NvilaVerifier:
    """Adapter for the real NVILA verifier returned by the reference's load_model():
    `model.generate_content([PIL, prompt]) -> (label, scores)`; score = scores[0][0, yes|no id]
    (tts_reflectionflow.py:160-164).  Needs decoded images (VAE) and Hub weights."""

    name = "nvila"

    def __init__(self, model, yes_id: int, no_id: int):
        self.model, self.yes_id, self.no_id = model, yes_id, no_id

    def score(self, cands: Sequence[Candidate], prompts: Sequence[str]):
        out = []
        for c, p in zip(cands, prompts):
            if c.image is None:
                raise RuntimeError("NVILA needs decoded images: attach a VAE to the pipeline")
            r1, scores1 = self.model.generate_content([c.image, p])
            tok = self.yes_id if r1 == "yes" else self.no_id
            out.append({"image_name": c.name, "label": "yes" if r1 == "yes" else "no",
                        "score": scores1[0][0, tok].detach().cpu().float().item()})
        return out


def load_verifier(verifier_args: dict, synthetic: bool, choice_of_metric: str = "overall_score"):
    """tts_reflectionflow.py:515-522: only "openai" and "nvila" are accepted."""
    name = verifier_args.get("name", "openai")
    if name not in ("openai", "nvila"):
        raise ValueError(f"Verifier {name} not supported")
    if synthetic:
        return StubVerifier(name, choice_of_metric)
    if name == "nvila":
        from transformers import AutoModel
        model = AutoModel.from_pretrained(verifier_args["model_name"], trust_remote_code=True,
                                          device_map="auto", cache_dir=verifier_args.get("cache_dir"))
        yes_id = model.tokenizer.encode("yes", add_special_tokens=False)[0]
        no_id = model.tokenizer.encode("no", add_special_tokens=False)[0]
        return NvilaVerifier(model, yes_id, no_id)
    raise RuntimeError("the OpenAI verifier needs network access; run with --synthetic offline")


class StubReflector:
    """Deterministic reflection writer + prompt refiner (same list-in / list-out shapes as
    OpenAIVerifier.generate_reflections / refine_prompt, openai_verifier.py:241-317)."""

    def generate_reflections(self, cands, original_prompt, current_prompts, reflections, evaluations):
        out = []
        for c, ev in zip(cands, evaluations):
            ev_ = ev.replace(c.name, os.path.basename(c.name))  # independent of the output directory
            h = hashlib.sha256((os.path.basename(c.name) + ev_).encode()).hexdigest()[:6]
            out.append(f"Make the subject match the prompt more closely ({h}).")
        return out

    def refine_prompt(self, cands, original_prompt, current_prompts, reflections, evaluations=None):
        return [f"{original_prompt}, detailed, faithful to the description" for _ in cands]


class HashTextEncoder:
    """Offline stand-in for T5-XXL + CLIP-L (the `text_encoder_hook` of B200FluxPipeline): prompt ->
    seeded N(0,1) embeddings, a pure function of the prompt strings.  Synthetic data, not a model."""

    def __init__(self, joint_dim: int = 4096, pooled_dim: int = 768):
        self.joint_dim, self.pooled_dim = joint_dim, pooled_dim

    @staticmethod
    def _seed(s: str) -> int:
        return int.from_bytes(hashlib.sha256(s.encode()).digest()[:4], "little") & 0x7FFFFFFF

    def __call__(self, prompt: List[str], prompt_2: List[str], max_sequence_length: int = 512):
        embs, pooled = [], []
        for p1, p2 in zip(prompt, prompt_2):
            g2 = torch.Generator().manual_seed(self._seed("t5:" + p2))
            g1 = torch.Generator().manual_seed(self._seed("clip:" + p1))
            embs.append(torch.randn(max_sequence_length, self.joint_dim, generator=g2))
            pooled.append(torch.randn(self.pooled_dim, generator=g1))
        return torch.stack(embs).to(torch.bfloat16), torch.stack(pooled).to(torch.bfloat16)


class ByteTokenizer:
    """Offline stand-in for the CLIP BPE / T5 sentencepiece tokenizers (their vocabularies are
    checkpoint assets): UTF-8 bytes offset into the vocabulary, EOS, padding.  Same call shape as
    the tokenizer arguments of B200TextEncoders.as_hook: (prompts, max_len) -> LongTensor ids."""

    def __init__(self, vocab: int, eos: int, pad: int, bos: Optional[int] = None):
        self.vocab, self.eos, self.pad, self.bos = vocab, eos, pad, bos

    def __call__(self, prompts: List[str], max_len: int) -> torch.Tensor:
        out = torch.full((len(prompts), max_len), self.pad, dtype=torch.long)
        for i, p in enumerate(prompts):
            toks = ([self.bos] if self.bos is not None else []) + [3 + b % (self.vocab - 4) for b in p.encode()]
            toks = toks[: max_len - 1] + [self.eos]
            out[i, : len(toks)] = torch.tensor(toks)
        return out
