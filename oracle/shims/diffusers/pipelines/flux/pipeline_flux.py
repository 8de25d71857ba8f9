"""Restated pieces of diffusers.pipelines.flux.pipeline_flux that generate.py:3,8-13 and
pipeline_tools.py use: calculate_shift, retrieve_timesteps, FluxPipelineOutput and a minimal
FluxPipeline (latent-space only: no VAE / text encoders — callers pass prompt_embeds and
output_type="latent")."""
import logging as _logging
import math
from contextlib import contextmanager
from dataclasses import dataclass

import numpy as np  # noqa: F401  (generate.py imports np from here)
import torch

from oracle import flux_oracle as fo

logger = _logging.getLogger("diffusers.shim.pipeline_flux")
calculate_shift = fo.calculate_shift


@dataclass
class FluxPipelineOutput:
    images: object


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class FlowMatchEulerDiscreteScheduler:
    """diffusers FlowMatchEulerDiscreteScheduler with the FLUX.1-dev scheduler_config.json."""
    order = 1

    def __init__(self):
        self.config = _Cfg(num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True,
                           base_shift=0.5, max_shift=1.15, base_image_seq_len=256,
                           max_image_seq_len=4096)
        self._step_index = None

    def time_shift(self, mu, sigma, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None):
        sigmas = np.array(sigmas).astype(np.float32)
        sigmas = self.time_shift(mu, 1.0, sigmas)
        sigmas = torch.from_numpy(np.asarray(sigmas, dtype=np.float32)).to(dtype=torch.float32, device=device)
        self.timesteps = sigmas * self.config.num_train_timesteps
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self._step_index = None

    def step(self, model_output, timestep, sample, return_dict=True):
        if self._step_index is None:
            self._step_index = 0
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self._step_index]
        sigma_next = self.sigmas[self._step_index + 1]
        prev = sample + (sigma_next - sigma) * model_output
        prev = prev.to(model_output.dtype)
        self._step_index += 1
        return (prev,)


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None,
                       **kwargs):
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed.")
    if sigmas is not None:
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    else:
        raise ValueError("shim supports the sigmas path only (what generate.py uses)")
    return timesteps, num_inference_steps


class FluxPipeline:
    """Latent-space subset of diffusers FluxPipeline: the attributes generate.py:114-310 touches."""

    vae_scale_factor = 8
    default_sample_size = 128

    def __init__(self, transformer, dtype=torch.bfloat16):
        self.transformer = transformer
        self.scheduler = FlowMatchEulerDiscreteScheduler()
        self.dtype = dtype
        self.device = torch.device("cpu")
        self._execution_device = torch.device("cpu")
        self._joint_attention_kwargs = None
        self._interrupt = False

    @property
    def joint_attention_kwargs(self):
        return self._joint_attention_kwargs

    @property
    def interrupt(self):
        return self._interrupt

    def check_inputs(self, prompt, prompt_2, height, width, prompt_embeds=None,
                     pooled_prompt_embeds=None, callback_on_step_end_tensor_inputs=None,
                     max_sequence_length=None):
        if height % 16 != 0 or width % 16 != 0:
            raise ValueError("`height` and `width` have to be divisible by 16")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError("Cannot forward both `prompt` and `prompt_embeds`")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`")
        if prompt_embeds is not None and pooled_prompt_embeds is None:
            raise ValueError("If `prompt_embeds` are provided, `pooled_prompt_embeds` also have to be passed")

    def encode_prompt(self, prompt=None, prompt_2=None, prompt_embeds=None, pooled_prompt_embeds=None,
                      device=None, num_images_per_prompt=1, max_sequence_length=512, lora_scale=None):
        if prompt_embeds is None:
            raise RuntimeError("shim pipeline has no text encoders: pass prompt_embeds")
        text_ids = torch.zeros(prompt_embeds.shape[1], 3).to(device=device, dtype=prompt_embeds.dtype)
        return prompt_embeds, pooled_prompt_embeds, text_ids

    _pack_latents = staticmethod(fo.pack_latents)
    _unpack_latents = staticmethod(fo.unpack_latents)

    @staticmethod
    def _prepare_latent_image_ids(batch_size, height, width, device, dtype):
        return fo.prepare_latent_image_ids(height, width, dtype).to(device)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator,
                        latents=None):
        height = 2 * (int(height) // (self.vae_scale_factor * 2))
        width = 2 * (int(width) // (self.vae_scale_factor * 2))
        ids = self._prepare_latent_image_ids(batch_size, height // 2, width // 2, device, dtype)
        if latents is not None:
            return latents.to(device=device, dtype=dtype), ids
        from diffusers.utils.torch_utils import randn_tensor
        latents = randn_tensor((batch_size, num_channels_latents, height, width), generator=generator,
                               device=device, dtype=dtype)
        return self._pack_latents(latents, batch_size, num_channels_latents, height, width), ids

    @contextmanager
    def progress_bar(self, total=None):
        class _PB:
            def update(self):
                pass
        yield _PB()

    def maybe_free_model_hooks(self):
        pass

    def set_adapters(self, *a, **k):
        pass
