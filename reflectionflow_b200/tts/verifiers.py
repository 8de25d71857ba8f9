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

    needs_images = False

    def __init__(self, name: str = "nvila", choice_of_metric: str = "overall_score"):
        assert name in ("nvila", "openai", "ours")
        self.name = name
        self.choice_of_metric = choice_of_metric

    def value(self, cand: Candidate) -> float:
        if cand.stub_score is not None:
            return float(cand.stub_score)
        if cand.latents is None:  # pixel-only candidate (a stage-0 PNG written by the reference): same kind of
            if cand.image_u8 is None:  # fixed functional, of the image
                raise RuntimeError(f"{cand.name}: neither latents nor pixels to score")
            px = cand.image_u8.to(torch.float64)
            return float((px.mean() / 255.0 - 0.5) * 0.1 + (px[::7, ::5].mean() - px[::5, ::7].mean()) / 2550.0)
        return latent_functional(cand.latents)

    def score_one(self, cand: Candidate, prompt: str) -> Dict[str, Any]:
        v = self.value(cand)
        if self.name == "nvila":
            p_yes = 1.0 / (1.0 + math.exp(-40.0 * v))
            if p_yes >= 0.5:
                return {"image_name": cand.name, "label": "yes", "score": p_yes}
            return {"image_name": cand.name, "label": "no", "score": 1.0 - p_yes}
        if self.name == "ours":  # scalar reward under choice_of_metric (ImageVerifierOurs shape)
            return {self.choice_of_metric: float(v), "VQ": float(v), "image_name": cand.name}
        s = max(0, min(10, int(round(5 + 60.0 * v))))
        # no "image_name": the reference's OpenAI outputs are the bare pydantic dumps (openai_verifier.py:147), and
        # `outputs.index` in the top-k rule tells candidates apart by dict equality (App. B.7)
        return {self.choice_of_metric: {"score": s, "explanation": "stub"}}

    def score(self, cands: Sequence[Candidate], prompts: Sequence[str], tag=None) -> List[Dict[str, Any]]:
        return [self.score_one(c, p) for c, p in zip(cands, prompts)]


def _images_of(cands: Sequence[Candidate]):
    imgs = []
    for c in cands:
        im = c.pil()
        if im is None:
            raise RuntimeError(f"{c.name}: the verifier needs pixels — decode the candidate first "
                               "(pipeline built with with_vae=True)")
        imgs.append(im)
    return imgs


class NvilaVerifier:
    """Adapter for the real NVILA verifier returned by the reference's load_model()
    (verifiers/nvila_verifier.py:4-10): `model.generate_content([PIL, prompt]) -> (label, scores)`;
    score = scores[0][0, yes_id | no_id] (tts_reflectionflow.py:157-164, 345-352).  Uniform hook:
    score(candidates, prompts) -> [{"image_name", "label", "score"}], images taken from HBM-decoded
    candidates instead of re-opened PNGs."""

    name = "nvila"
    needs_images = True

    def __init__(self, model, yes_id: int, no_id: int):
        self.model, self.yes_id, self.no_id = model, yes_id, no_id

    def score(self, cands: Sequence[Candidate], prompts: Sequence[str], tag=None):
        out = []
        for c, im, p in zip(cands, _images_of(cands), prompts):
            r1, scores1 = self.model.generate_content([im, p])
            yes = r1 == "yes"
            tok = self.yes_id if yes else self.no_id
            out.append({"image_name": c.name, "label": "yes" if yes else "no",
                        "score": scores1[0][0, tok].detach().cpu().float().item()})
        return out


# ---- response schemas of the OpenAI-shaped verifier (verifiers/openai_verifier.py:23-69): the JSON
# keys are part of the on-disk format (best_img_detailedscore.jsonl) and of `choice_of_metric`.
GRADING_ASPECTS = {
    None: ("accuracy_to_prompt", "creativity_and_originality", "visual_quality_and_realism",
           "consistency_and_cohesion", "emotional_or_thematic_resonance", "overall_score"),
    "single_object": ("object_completeness", "detectability", "occlusion_handling", "overall_score"),
    "two_object": ("separation_clarity", "individual_completeness", "relationship_accuracy", "overall_score"),
    "counting": ("count_accuracy", "object_uniformity", "spatial_legibility", "overall_score"),
    "colors": ("color_fidelity", "contrast_effectiveness", "multi_object_consistency", "overall_score"),
    "position": ("position_accuracy", "occlusion_management", "perspective_consistency", "overall_score"),
    "color_attr": ("attribute_binding", "contrast_effectiveness", "material_consistency", "overall_score"),
}
_GRADING_MODELS: Dict[Any, Any] = {}


def grading_model(tag: Optional[str]):
    """pydantic response_format for a GenEval tag (None = the general rubric)."""
    if tag not in GRADING_ASPECTS:
        raise KeyError(f"unknown verifier tag {tag!r}")
    if tag not in _GRADING_MODELS:
        from pydantic import BaseModel, create_model

        class Score(BaseModel):
            score: int
            explanation: str
        name = "Grading" if tag is None else f"Grading_{tag}"
        _GRADING_MODELS[tag] = create_model(name, **{a: (Score, ...) for a in GRADING_ASPECTS[tag]})
    return _GRADING_MODELS[tag]


def _jpeg_data_url(image) -> str:
    import base64
    import io
    if isinstance(image, str):
        with open(image, "rb") as f:
            raw = f.read()
    else:
        buf = io.BytesIO()
        image.save(buf, format="JPEG")
        raw = buf.getvalue()
    return "data:image/jpeg;base64," + base64.b64encode(raw).decode("utf-8")


def _map_in_order(fn, items, max_workers: int = 4):
    """The reference fans the HTTP calls out over <= 4 threads and appends results in COMPLETION
    order (openai_verifier.py:153-164), so outputs[i] need not belong to image i; here results keep
    the input order (a candidate is always scored by its own response)."""
    if not items:
        return []
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(items), max_workers)) as ex:
        return list(ex.map(fn, items))


class OpenAIShapedVerifier:
    """Accept-a-client adapter with the schema of verifiers/openai_verifier.py: any object exposing
    `client.beta.chat.completions.parse(model=, messages=, temperature=, response_format=)` (the
    OpenAI SDK, or a local server with the same surface) scores images against the rubric prompts.
    score(candidates, prompts, tag) -> [{<aspect>: {"score": int, "explanation": str}, ...}]."""

    name = "openai"
    needs_images = True

    def __init__(self, client, system_instruction, model_name: str = "gpt-4o-2024-11-20",
                 choice_of_metric: str = "overall_score"):
        self.client, self.system_instruction = client, system_instruction
        self.model_name, self.choice_of_metric = model_name, choice_of_metric

    def prepare_inputs(self, images, prompts):
        images = images if isinstance(images, list) else [images]
        prompts = prompts if isinstance(prompts, list) else [prompts]
        return [{"role": "user", "content": [{"type": "text", "text": p},
                                              {"type": "image_url", "image_url": {"url": _jpeg_data_url(im)}}]}
                for p, im in zip(prompts, images)]

    def score_inputs(self, inputs, tag=None, **_):
        instr = self.system_instruction if tag is None else self.system_instruction[tag]
        system_message = {"role": "system", "content": instr}
        fmt = grading_model(tag)

        def call(parts):
            r = self.client.beta.chat.completions.parse(model=self.model_name, messages=[system_message, parts],
                                                        temperature=1, response_format=fmt)
            return r.choices[0].message.parsed.model_dump()
        return _map_in_order(call, list(inputs))

    def score(self, cands: Sequence[Candidate], prompts: Sequence[str], tag=None):
        return self.score_inputs(self.prepare_inputs(_images_of(cands), list(prompts)), tag=tag)


class ImageVerifierOurs:
    """The "ours" branch the reference's loader lacks (tts_reflectionflow.py:515-522 accepts only
    openai / nvila): the repo's own Image-Verifier — a Qwen2.5-VL backbone with a scalar `rm_head`
    (reward_modeling/trainer.py:59-172) driven through reward_modeling/inference.py:155-180's
    `reward(images, prompts, use_norm) -> [{"VQ": r, "Overall": r}]`.  One batched in-process call per
    round share; output dicts carry the scalar under `choice_of_metric` ("Overall"), so the
    metric-ordered selection rules apply unchanged (descending reward)."""

    name = "ours"
    needs_images = True

    def __init__(self, inferencer, choice_of_metric: str = "Overall", use_norm: bool = True,
                 batch_size: int = 8):
        self.inferencer, self.choice_of_metric = inferencer, choice_of_metric
        self.use_norm, self.batch_size = use_norm, batch_size

    def score(self, cands: Sequence[Candidate], prompts: Sequence[str], tag=None):
        imgs, out = _images_of(cands), []
        for o in range(0, len(imgs), self.batch_size):
            rs = self.inferencer.reward(imgs[o:o + self.batch_size], list(prompts[o:o + self.batch_size]),
                                        use_norm=self.use_norm)
            out.extend(rs)
        res = []
        for c, r in zip(cands, out):
            d = {k: float(v) for k, v in r.items()}
            if self.choice_of_metric not in d:
                raise KeyError(f"reward model returned {sorted(d)}, not {self.choice_of_metric!r}")
            d["image_name"] = c.name
            res.append(d)
        return res


def load_verifier(verifier_args: dict, synthetic: bool, choice_of_metric: str = "overall_score",
                  client=None, inferencer=None):
    """tts_reflectionflow.py:515-522 accepts "openai" and "nvila"; "ours" (the Image-Verifier reward
    model) is the branch SURVEY §8f-3 adds.  `client` / `inferencer` inject the external model."""
    name = verifier_args.get("name", "openai")
    if name not in ("openai", "nvila", "ours"):
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
    if name == "ours":
        if inferencer is None:
            raise RuntimeError("verifier 'ours' needs the Image-Verifier inferencer (an object with "
                               ".reward(images, prompts, use_norm)); its checkpoint is not reachable offline")
        return ImageVerifierOurs(inferencer, choice_of_metric)
    if client is None:
        raise RuntimeError("the OpenAI verifier needs a client (network access); run with --synthetic offline")
    from .utils import load_verifier_prompt
    return OpenAIShapedVerifier(client, load_verifier_prompt(verifier_args["verifier_prompt_relpath"]),
                                verifier_args.get("model_name", "gpt-4o-2024-11-20"), choice_of_metric)


class OpenAIShapedReflector:
    """Accept-a-client reflection writer + prompt refiner with the message layout of
    openai_verifier.py:166-317 (`client.chat.completions.create(model=, messages=, temperature=)`).
    Same list-in / list-out shapes as StubReflector; results keep the input order."""

    def __init__(self, client, reflexion_instruction: str, refine_instruction: str,
                 model_name: str = "gpt-4o-2024-11-20"):
        self.client, self.model_name = client, model_name
        self.system_message_reflexion = {"role": "system", "content": reflexion_instruction}
        self.system_message_refine = {"role": "system", "content": refine_instruction}

    def _chat(self, system_message, inputs):
        def call(parts):
            r = self.client.chat.completions.create(model=self.model_name, messages=[system_message, parts],
                                                    temperature=1)
            return r.choices[0].message.content
        return _map_in_order(call, list(inputs))

    def prepare_reflexion_prompt_inputs(self, images, original_prompt, current_prompt, reflections, evaluations):
        inputs = []
        for im, op, cp, rf, ev in zip(images, original_prompt, current_prompt, reflections, evaluations):
            inputs.append({"role": "user", "content": [
                {"type": "text", "text": "Original prompt: " + op},
                {"type": "text", "text": f"The updated prompt to generate the image is: {cp}[Reflexion]: {rf}"},
                {"type": "text", "text": f"Evaluation of the generated image: {ev}"},
                {"type": "text", "text": "Generated images:"},
                {"type": "image_url", "image_url": {"url": _jpeg_data_url(im)}},
                {"type": "text", "text": "Please generate instructions following the defined rules."}]})
        return inputs

    def prepare_refine_prompt_inputs(self, original_prompt, images=None, evaluations=None, current_prompt=None,
                                     reflections=None):
        n = len(original_prompt)
        fill = lambda v: [None] * n if v is None else list(v)[:n]
        inputs = []
        for op, im, ev, cp, rf in zip(original_prompt, fill(images), fill(evaluations), fill(current_prompt),
                                      fill(reflections)):
            content = [{"type": "text", "text": f"Original prompt: {op}"}]
            if cp:
                content.append({"type": "text", "text": f"Current prompt: {cp}"})
            if rf:
                content.append({"type": "text", "text": f"Reflection prompt: {rf}"})
            if im is not None:
                content.append({"type": "image_url", "image_url": {"url": _jpeg_data_url(im)}})
            if ev:
                content.append({"type": "text", "text": f"Evaluation of the generated images: {ev}"})
            content.append({"type": "text", "text": "Please refine the current prompt to improve the overall "
                                                    "quality of the future generated images."})
            inputs.append({"role": "user", "content": content})
        return inputs

    # ---- the two hooks sample() calls (same signatures as StubReflector)
    def generate_reflections(self, cands, original_prompt, current_prompts, reflections, evaluations):
        n = len(cands)
        inputs = self.prepare_reflexion_prompt_inputs(_images_of(cands), [original_prompt] * n,
                                                      list(current_prompts), list(reflections), list(evaluations))
        out = self._chat(self.system_message_reflexion, inputs)
        if len(out) != n:
            raise RuntimeError("reflection writer returned a short list")
        return out

    def refine_prompt(self, cands, original_prompt, current_prompts, reflections, evaluations=None):
        n = len(cands)
        inputs = self.prepare_refine_prompt_inputs([original_prompt] * n, images=_images_of(cands),
                                                   evaluations=evaluations, current_prompt=list(current_prompts),
                                                   reflections=reflections)
        return self._chat(self.system_message_refine, inputs)


class ReflectionGeneratorOurs:
    """`reflection_args.name` other than "openai" (tts_reflectionflow.py:26-44, 220-238): the fine-tuned
    Qwen2.5-VL reflection generator behind an OpenAI-compatible server — one
    `client.chat.completions.create(messages=, model=)` request per selected image, each retried up to
    `max_retries` times (the reference then silently drops the entry; here the last error is raised, so the list
    cannot come back short).  Prompt refinement stays with the OpenAI-shaped refiner (:247-254): `refiner`, or
    the current prompts unchanged when there is none."""

    def __init__(self, client, refiner=None, model_name: str = "Qwen/Qwen2.5-VL-7B-Instruct",
                 max_retries: int = 5, retry_delay: float = 2.0):
        self.client, self.refiner, self.model_name = client, refiner, model_name
        self.max_retries, self.retry_delay = max_retries, retry_delay

    @staticmethod
    def generate_messages(bad_image: str, prompt: str):
        """tts_reflectionflow.py:27-41 (`bad_image`: a path or URL the server can read, or a data URL)"""
        return [{"role": "system", "content": "You are a helpful assistant."},
                {"role": "user", "content": [
                    {"type": "image_url", "image_url": {"url": bad_image}},
                    {"type": "text", "text": "Generate reflections to improve the input image according to the "
                                             f"prompt. The prompt is: \"{prompt}\""}]}]

    def generate_reflections(self, cands, original_prompt, current_prompts, reflections, evaluations):
        import time
        out = []
        for c in cands:
            url = c.name if os.path.exists(c.name) else _jpeg_data_url(_images_of([c])[0])
            messages = self.generate_messages(url, original_prompt)
            for attempt in range(self.max_retries):
                try:
                    res = self.client.chat.completions.create(messages=messages, model=self.model_name)
                    out.append(res.choices[0].message.content)
                    break
                except Exception as e:  # noqa: BLE001 - the reference retries on anything (:231-238)
                    if attempt + 1 == self.max_retries:
                        raise
                    print(f"Error generating reflection: {e}. Retrying in {self.retry_delay} seconds...")
                    time.sleep(self.retry_delay)
        return out

    def refine_prompt(self, cands, original_prompt, current_prompts, reflections, evaluations=None):
        if self.refiner is None:
            return list(current_prompts)
        return self.refiner.refine_prompt(cands, original_prompt, current_prompts, reflections, evaluations)


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
